// Path-compatible entry point: code written against HJReachability/ilqgames includes
// <ilqgames/cost/quadratic_difference_cost.h>; the B200 host classes live in <ilqgames/b200/costs.h>.
#ifndef ILQGAMES_B200_FWD_COST_QUADRATIC_DIFFERENCE_COST_H
#define ILQGAMES_B200_FWD_COST_QUADRATIC_DIFFERENCE_COST_H
#include <ilqgames/b200/costs.h>
#endif
