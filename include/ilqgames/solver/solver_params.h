// Path-compatible entry point: code written against HJReachability/ilqgames includes
// <ilqgames/solver/solver_params.h>; the B200 host classes live in <ilqgames/b200/core.h>.
#ifndef ILQGAMES_B200_FWD_SOLVER_SOLVER_PARAMS_H
#define ILQGAMES_B200_FWD_SOLVER_SOLVER_PARAMS_H
#include <ilqgames/b200/core.h>
#endif
