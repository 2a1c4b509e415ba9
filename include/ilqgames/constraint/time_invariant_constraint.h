// Path-compatible entry point: code written against HJReachability/ilqgames includes
// <ilqgames/constraint/time_invariant_constraint.h>; the B200 host classes live in <ilqgames/b200/costs.h>.
#ifndef ILQGAMES_B200_FWD_CONSTRAINT_TIME_INVARIANT_CONSTRAINT_H
#define ILQGAMES_B200_FWD_CONSTRAINT_TIME_INVARIANT_CONSTRAINT_H
#include <ilqgames/b200/costs.h>
#endif
