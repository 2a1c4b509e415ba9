// Path-compatible entry point: code written against HJReachability/ilqgames includes
// <ilqgames/utils/solver_log.h>; the B200 host classes live in <ilqgames/b200/core.h>.
#ifndef ILQGAMES_B200_FWD_UTILS_SOLVER_LOG_H
#define ILQGAMES_B200_FWD_UTILS_SOLVER_LOG_H
#include <ilqgames/b200/core.h>
#endif
