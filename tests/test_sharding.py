"""The N > 1 path of bench.py on CPU: two gloo ranks shard a batch of independent games, each
solves its slice (with the oracle standing in for the device), and the plumbing (slice bounds,
MAX-time / SUM-work reduction, final all-gather) must reproduce the single-process result."""
import os
import subprocess
import sys
import textwrap

import numpy as np

from ilqgames_b200 import sharding

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_partition_the_batch():
    for B in (1, 7, 64, 4096):
        for W in (1, 2, 3, 8):
            cuts = [sharding.shard_bounds(B, W, r) for r in range(W)]
            assert cuts[0][0] == 0 and cuts[-1][1] == B
            assert all(cuts[r][1] == cuts[r + 1][0] for r in range(W - 1))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1


WORKER = textwrap.dedent("""
    import os, sys, json
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, {repo!r})
    from ilqgames_b200 import _abi as abi, problems, sharding
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    B = 6
    desc, _ = problems.three_player_intersection()
    params = problems.three_player_intersection_params(max_solver_iters=2)
    x0 = problems.three_player_intersection_x0_batch(B, 11)
    lo, hi = sharding.shard_bounds(B, world, rank)
    lib = abi.Library(os.path.join({repo!r}, "oracle", "_build", "libilqg_oracle.so"))
    h = abi.Handle(lib, desc, params, hi - lo)
    h.upload_x0(x0[lo:hi]); h.solve_begin(); h.iterate(2)
    xs = h.download(abi.XS)
    done = int(h.download(abi.ITERS).sum())
    ms, total = sharding.reduce_metrics(10.0 * (rank + 1), done)
    allxs = sharding.gather_trajectories(xs)
    if rank == 0:
        np.save({out!r}, allxs.reshape((-1,) + xs.shape[1:]))
        print(json.dumps({{"ms": ms, "total": total}}))
    dist.destroy_process_group()
""")


def test_two_rank_gloo_run_matches_single_process(tmp_path, oracle):
    from ilqgames_b200 import _abi as abi, problems
    out = str(tmp_path / "xs.npy")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(repo=REPO, out=out))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, timeout=300, env=env)
    assert res.returncode == 0, res.stderr[-2000:]
    import json
    line = [l for l in res.stdout.splitlines() if l.startswith("{")][-1]
    agg = json.loads(line)
    assert agg["ms"] == 20.0  # MAX over ranks
    desc, _ = problems.three_player_intersection()
    params = problems.three_player_intersection_params(max_solver_iters=2)
    x0 = problems.three_player_intersection_x0_batch(6, 11)
    h = abi.Handle(oracle, desc, params, 6)
    h.upload_x0(x0)
    h.solve_begin()
    h.iterate(2)
    assert agg["total"] == int(h.download(abi.ITERS).sum())
    np.testing.assert_array_equal(np.load(out), h.download(abi.XS))


def test_multi_device_handle_shards_one_batch(oracle):
    """ilqg_create_multi through the ABI (the oracle ignores the device list): one handle, the global
    batch addressed as a whole -- the results equal a single-device handle's, row for row."""
    import numpy as np
    from ilqgames_b200 import _abi as abi, problems
    desc, _ = problems.three_player_intersection(num_time_steps=20)
    params = problems.three_player_intersection_params(max_solver_iters=2)
    x0 = problems.three_player_intersection_x0_batch(5, 3)
    out = []
    for dev in (0, [0, 1]):
        h = abi.Handle(oracle, desc, params, 5, dev)
        h.upload_x0(x0)
        h.solve_begin()
        h.solve(chunk=2)
        out.append((h.download(abi.XS), h.download(abi.STATUS)))
        h.close()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    import ctypes as C
    hh = C.c_void_p()
    devs = (C.c_int * 2)(0, 1)
    assert oracle.lib.ilqg_create_multi(C.byref(desc), C.byref(params), 1, devs, 2, C.byref(hh)) == -1  # batch < devices
