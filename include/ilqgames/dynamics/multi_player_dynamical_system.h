// Path-compatible entry point: code written against HJReachability/ilqgames includes
// <ilqgames/dynamics/multi_player_dynamical_system.h>; the B200 host classes live in <ilqgames/b200/dynamics.h>.
#ifndef ILQGAMES_B200_FWD_DYNAMICS_MULTI_PLAYER_DYNAMICAL_SYSTEM_H
#define ILQGAMES_B200_FWD_DYNAMICS_MULTI_PLAYER_DYNAMICAL_SYSTEM_H
#include <ilqgames/b200/dynamics.h>
#endif
