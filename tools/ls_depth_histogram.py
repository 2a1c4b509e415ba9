"""How deep do linesearches backtrack?  The oracle (CPU) on samples of the three headline batches, one iLQ
iteration at a time: histogram of rollouts per linesearch (j + 1; max_backtracking_steps + 1 = failed).
The tier split of the continued linesearch (ilqg_abi.cu: ILQG_LS_TIERS) was chosen from this table
(profiles/r02_summary.md).  usage: python tools/ls_depth_histogram.py"""
import sys, numpy as np, collections
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ilqgames_b200 import _abi as abi
lib = abi.Library(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle', '_build', 'libilqg_oracle.so'))
for cfg, B in (("metric", 192), ("c3", 48), ("c4", 256)):
    desc, params, x0 = bench.workload(cfg, 4096 if cfg != "c4" else 16384, 4096)
    idx = np.random.default_rng(0).choice(len(x0), B, replace=False)
    x0 = x0[idx]
    h = abi.Handle(lib, desc, params, B, 0)
    h.upload_x0(x0); h.solve_begin()
    prev = h.download(abi.BACKTRACKS).copy()
    hist = collections.Counter()
    for it in range(10):
        st0 = h.download(abi.STATUS).copy()
        h.iterate(1)
        bt = h.download(abi.BACKTRACKS); st = h.download(abi.STATUS)
        d = (bt - prev)[st0 == abi.STATUS_RUNNING]; prev = bt.copy()
        failed = (st == abi.STATUS_LINESEARCH_FAILED) & (st0 == abi.STATUS_RUNNING)
        for v in d: hist[int(v)] += 1
    h.close()
    tot = sum(hist.values())
    print(cfg, "linesearches", tot, "rollouts histogram (j+1; 101 = failed):")
    cum = 0
    for k in sorted(hist):
        cum += hist[k]
        print(f"   {k}: {hist[k]}  cum {cum/tot:.3f}")
