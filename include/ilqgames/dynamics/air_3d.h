// Path-compatible entry point: code written against HJReachability/ilqgames includes
// <ilqgames/dynamics/air_3d.h>; the B200 host classes live in <ilqgames/b200/dynamics.h>.
#ifndef ILQGAMES_B200_FWD_DYNAMICS_AIR_3D_H
#define ILQGAMES_B200_FWD_DYNAMICS_AIR_3D_H
#include <ilqgames/b200/dynamics.h>
#endif
