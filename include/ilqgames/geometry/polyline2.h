// Path-compatible entry point: code written against HJReachability/ilqgames includes
// <ilqgames/geometry/polyline2.h>; the B200 host classes live in <ilqgames/b200/costs.h>.
#ifndef ILQGAMES_B200_FWD_GEOMETRY_POLYLINE2_H
#define ILQGAMES_B200_FWD_GEOMETRY_POLYLINE2_H
#include <ilqgames/b200/costs.h>
#endif
