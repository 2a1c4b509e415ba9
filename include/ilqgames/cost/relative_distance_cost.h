// Path-compatible entry point: code written against HJReachability/ilqgames includes
// <ilqgames/cost/relative_distance_cost.h>; the B200 host classes live in <ilqgames/b200/costs.h>
// (RelativeDistanceCost itself is not among them yet: no in-tree example constructs one).
#ifndef ILQGAMES_B200_FWD_COST_RELATIVE_DISTANCE_COST_H
#define ILQGAMES_B200_FWD_COST_RELATIVE_DISTANCE_COST_H
#include <ilqgames/b200/costs.h>
#endif
