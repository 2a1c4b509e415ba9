// Path-compatible entry point: code written against HJReachability/ilqgames includes
// <ilqgames/utils/linear_dynamics_approximation.h>; the B200 host classes live in <ilqgames/b200/core.h>.
#ifndef ILQGAMES_B200_FWD_UTILS_LINEAR_DYNAMICS_APPROXIMATION_H
#define ILQGAMES_B200_FWD_UTILS_LINEAR_DYNAMICS_APPROXIMATION_H
#include <ilqgames/b200/core.h>
#endif
