"""ilqgames_b200 -- B200-native batched iLQ-games inner solver.

Scope: the single hot path of HJReachability/ilqgames named in BASELINE.json
(rollout -> linearize -> quadraticize -> backward Riccati Nash recursion, plus the
Armijo linesearch glue), as hand-written sm_100a CUDA behind the C ABI declared in
include/ilqg.h.  This Python package is binding glue for tests and bench.py:

  _abi.py      ctypes mirror of include/ilqg.h, handle wrapper
  problems.py  POD problem descriptors restating the reference example classes
  build.py     nvcc build recipe for csrc/ -> lib/libilqg_b200.so
  csrc/        CUDA kernels + the C ABI implementation (the product)
"""
from . import _abi as abi  # noqa: F401
from ._abi import Handle, IlqgError, Library, SolverParams, product_library  # noqa: F401
