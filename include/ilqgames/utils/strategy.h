// Path-compatible entry point: code written against HJReachability/ilqgames includes
// <ilqgames/utils/strategy.h>; the B200 host classes live in <ilqgames/b200/core.h>.
#ifndef ILQGAMES_B200_FWD_UTILS_STRATEGY_H
#define ILQGAMES_B200_FWD_UTILS_STRATEGY_H
#include <ilqgames/b200/core.h>
#endif
