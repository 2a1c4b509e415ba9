// Path-compatible entry point: code written against HJReachability/ilqgames includes
// <ilqgames/solver/lq_open_loop_solver.h>; the B200 host classes live in <ilqgames/b200/solvers.h>.
#ifndef ILQGAMES_B200_FWD_SOLVER_LQ_OPEN_LOOP_SOLVER_H
#define ILQGAMES_B200_FWD_SOLVER_LQ_OPEN_LOOP_SOLVER_H
#include <ilqgames/b200/solvers.h>
#endif
