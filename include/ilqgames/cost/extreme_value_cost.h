// Path-compatible entry point: code written against HJReachability/ilqgames includes
// <ilqgames/cost/extreme_value_cost.h>; the B200 host classes live in <ilqgames/b200/costs.h>.
#ifndef ILQGAMES_B200_FWD_COST_EXTREME_VALUE_COST_H
#define ILQGAMES_B200_FWD_COST_EXTREME_VALUE_COST_H
#include <ilqgames/b200/costs.h>
#endif
