// Path-compatible entry point: code written against HJReachability/ilqgames includes
// <ilqgames/dynamics/two_player_unicycle_4d.h>; the B200 host classes live in
// <ilqgames/b200/dynamics.h>.
#ifndef ILQGAMES_B200_FWD_DYNAMICS_TWO_PLAYER_UNICYCLE_4D_H
#define ILQGAMES_B200_FWD_DYNAMICS_TWO_PLAYER_UNICYCLE_4D_H
#include <ilqgames/b200/dynamics.h>
#endif
