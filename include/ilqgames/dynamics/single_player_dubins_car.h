// Path-compatible entry point: code written against HJReachability/ilqgames includes
// <ilqgames/dynamics/single_player_dubins_car.h>; the B200 host classes live in
// <ilqgames/b200/dynamics.h>.
#ifndef ILQGAMES_B200_FWD_DYNAMICS_SINGLE_PLAYER_DUBINS_CAR_H
#define ILQGAMES_B200_FWD_DYNAMICS_SINGLE_PLAYER_DUBINS_CAR_H
#include <ilqgames/b200/dynamics.h>
#endif
