// Path-compatible entry point: code written against HJReachability/ilqgames includes
// <ilqgames/dynamics/single_player_point_mass_2d.h>; the B200 host classes live in
// <ilqgames/b200/dynamics.h>.
#ifndef ILQGAMES_B200_FWD_DYNAMICS_SINGLE_PLAYER_POINT_MASS_2D_H
#define ILQGAMES_B200_FWD_DYNAMICS_SINGLE_PLAYER_POINT_MASS_2D_H
#include <ilqgames/b200/dynamics.h>
#endif
