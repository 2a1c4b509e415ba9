// Path-compatible entry point: code written against HJReachability/ilqgames includes
// <ilqgames/solver/top_down_renderable_problem.h>; the B200 host classes live in <ilqgames/b200/solvers.h>.
#ifndef ILQGAMES_B200_FWD_SOLVER_TOP_DOWN_RENDERABLE_PROBLEM_H
#define ILQGAMES_B200_FWD_SOLVER_TOP_DOWN_RENDERABLE_PROBLEM_H
#include <ilqgames/b200/solvers.h>
#endif
