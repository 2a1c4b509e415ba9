#!/usr/bin/env python
"""bench.py -- iLQ iterations/s on a batch of ThreePlayerIntersection games (T = 100); with
`--config c1..c5` the same measurement on each configuration BASELINE.json lists.

Contract (see the task prompt): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON
line; for N > 1 it is launched under torch.distributed.run, one rank per GPU.

  step        one benchmark solve of the whole per-GPU batch: Solve() prologue + 10 iLQ
              iterations (linearize+quadraticize, backward Riccati, linesearch) from the zero
              warm start, lambda = 0, mu = 10, convergence exit disabled
  value       completed instance-iterations / second, all ranks, inputs resident in HBM,
              CUDA-event timed on the launch stream, max over ranks
  e2e         same metric through the C ABI with HOST buffers: per step the initial states go
              pinned-host -> device and trajectories + status come back device -> pinned-host
  roofline    dominant kernel: algorithmic bytes per launch / mean launch duration (CUDA events)
  cpu_baseline  the CPU oracle (restatement of the reference's Eigen path) on a bounded sample
  config      the workload, stated identically by both arms (nothing measured, nothing host-dependent)
  workload_stats  what the run measured about the workload: iterations completed per step, status
              histogram, how the linesearches ended (first try / backtracked / failed)

`--impl reference` times the reference algorithm's CPU path instead (the oracle port, all host
cores) on the same workload/metric; no GPU code runs in that arm.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

ITERS_PER_SOLVE = 10
ORACLE_LIB = os.path.join(REPO, "oracle", "_build", "libilqg_oracle.so")
REF_BUILD_LIB = os.path.join(REPO, "oracle", "_ref", "libilqg_ref.so")


# BASELINE.json's configurations.  `metric` (the default) is the one the metric string is quoted on;
# c1..c5 are BASELINE.json `configs[0..4]` at their stated batch and horizon (reference-true state
# dimensions: SURVEY.md section 8 "Shapes at BASELINE configs").  batch = per GPU.
CONFIGS = {
    "metric": dict(problem="three_player_intersection", batch=4096, T=100,
                   label="ThreePlayerIntersection (2x car6d + unicycle4d, n=16, m=(2,2,2))"),
    "c1": dict(problem="three_player_intersection", batch=1, T=100,
               label="ThreePlayerIntersection single instance (the reference's own exec; plumbing)"),
    "c2": dict(problem="three_player_intersection", batch=1024, T=100,
               label="ThreePlayerIntersection (reference-true n=16; BASELINE.json states n=12), randomized x0"),
    "c3": dict(problem="roundabout_merging", batch=4096, T=150, label="RoundaboutMerging (4x car6d, n=24, m=(2,2,2,2))"),
    "c4": dict(problem="air_3d", batch=16384, T=50, label="Air3D two-player zero-sum (n=3, m=(1,1)), x0 grid sweep"),
    "c5": dict(problem="three_player_intersection", batch=8192, T=100,
               label="ThreePlayerIntersection sharded, 8192 games per GPU (65536 on 8)"),
}


def workload(config: str, batch: int, seed: int, lo: int = 0, hi: int = None):
    """Descriptor, solver parameters and rows [lo, hi) of the GLOBAL batch of `batch` initial states."""
    from ilqgames_b200 import problems
    c = CONFIGS[config]
    kw = dict(max_solver_iters=ITERS_PER_SOLVE, disable_convergence_exit=1)
    if c["problem"] == "three_player_intersection":
        desc, _ = problems.three_player_intersection(num_time_steps=c["T"])
        params = problems.three_player_intersection_params(**kw)
        x0 = problems.three_player_intersection_x0_batch(batch, seed)
    elif c["problem"] == "roundabout_merging":
        desc, _ = problems.roundabout_merging(num_time_steps=c["T"])
        params = problems.roundabout_params(**kw)
        x0 = problems.roundabout_x0_batch(batch, seed)
    else:
        desc, _ = problems.air_3d(num_time_steps=c["T"])
        params = problems.air_3d_params(**kw)
        side = int(np.ceil(np.sqrt(batch)))
        x0 = problems.air_3d_x0_grid(side)[:batch]
    return desc, params, np.ascontiguousarray(x0[lo:hi])


def usable_cores() -> int:
    """Host cores this process may really use: the affinity mask, capped by a cgroup CPU quota."""
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            cores = max(1, min(cores, int(float(quota) / float(period))))
    except (OSError, ValueError):
        pass
    return cores


def l2_policy(batch: int, layout):
    """(need_flush, note): the strategy buffers alone that one iteration rewrites and re-reads
    (2 x T x M x (n + 1) floats per game) against the 126 MB L2 -- a criterion both arms can evaluate
    from the problem's shape, so that their `config` objects are the same."""
    ws_mb = batch * layout.num_time_steps * 4 * 2 * layout.total_udim * (layout.xdim + 1) / 1e6
    need_flush = ws_mb <= 126
    note = (f"one iteration streams {ws_mb:.0f} MB of strategies (plus the LQ records) per {batch} instances: "
            + ("fits the 126 MB L2, so a 256 MB write flushes L2 before every timed step" if need_flush
               else "exceeds the 126 MB L2, no explicit flush"))
    return need_flush, note


def bench_config(config: str, batch: int, world: int, seed: int, layout) -> dict:
    """The workload description, IDENTICAL in both arms (`--impl reference` prints the same object;
    everything measured goes to `workload_stats`, what a CPU step really solves to `cpu_baseline.sample`)."""
    c = CONFIGS[config]
    return {"workload": f"batch {batch} {c['label']}, T={c['T']}, "
                        f"{ITERS_PER_SOLVE} iLQ iterations/solve, lambda=0 mu=10, convergence exit disabled",
            "config": config, "batch_per_gpu": batch, "global_batch": batch * world,
            "iterations_per_step": ITERS_PER_SOLVE, "seed": seed,
            "l2": l2_policy(batch, layout)[1],
            "parallelism": f"dp{world} (independent games, no hot-path collective)",
            "reference_arm": "--impl reference times the CPU path on a bounded SAMPLE of this workload: each step "
                             "solves the first min(batch, 16 x host cores) instances of the same seeded batch, one "
                             "process per core (throughput per instance-iteration does not depend on the batch on "
                             "the CPU); the exact sample is stated in its cpu_baseline.sample"}


def algorithmic_bytes(layout) -> dict:
    """Per instance-iteration algorithmic bytes of each kernel.  `dense`: SURVEY.md section 8d's byte
    model (lin / quad materialised once in HBM) -- the figure `roofline.achieved` is defined on.
    `compact`: what the round-2 kernels are designed to move (compact records: item values + g_k)."""
    T, n, M, N = layout.num_time_steps, layout.xdim, layout.total_udim, layout.num_players
    op = T * (n + M)
    AB = T * (n * n + n * M)
    QR = T * (N * (n * n + n) + layout.R_floats + layout.r_floats)
    Pa = T * (M * n + M)
    crec = T * layout.compact_record_floats
    out = {
        "linearize_quadraticize": 4 * (op + AB + QR),
        "lq_backward": 4 * (AB + QR + Pa),  # records read once, strategy written once
        "linesearch_per_rollout": 4 * (2 * op + Pa),
    }
    if crec:
        out["compact"] = {"linearize_quadraticize": 4 * (op + crec), "lq_backward": 4 * (crec + Pa)}
    return out


# ------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML in-process, every
    20 ms; nvidia-smi as a fallback).  Rows: [sm_mhz, sm_max_mhz, power_w, hw_slowdown,
    hw_thermal_slowdown, sw_thermal_slowdown, sw_power_cap]."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices; map through CUDA_VISIBLE_DEVICES when it is a plain list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = index
            if vis and all(tok.strip().isdigit() for tok in vis.split(",")):
                phys = int(vis.split(",")[index])
            self._dev = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self._nvml = pynvml
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        n = self._nvml
        sm = n.nvmlDeviceGetClockInfo(self._dev, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self._dev, n.NVML_CLOCK_SM)
        try:
            pw = n.nvmlDeviceGetPowerUsage(self._dev) / 1000.0
        except Exception:
            pw = 0.0
        r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self._dev)
        act = lambda bit: "Active" if (r & bit) else "Not Active"
        return [str(sm), str(mx), str(pw), act(n.nvmlClocksThrottleReasonHwSlowdown),
                act(n.nvmlClocksThrottleReasonHwThermalSlowdown), act(n.nvmlClocksThrottleReasonSwThermalSlowdown),
                act(n.nvmlClocksThrottleReasonSwPowerCap)]

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._nvml is not None:
                    self.rows.append(self._sample_nvml())
                    self._stop.wait(0.02)
                    continue
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5)
                if out.returncode == 0 and out.stdout.strip():
                    self.rows.append([c.strip() for c in out.stdout.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for k, nm in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# ------------------------------------------------------------------------- CPU arm
_CFG = "metric"   # the configuration of this run (set in main, inherited by forked workers)


def _oracle_worker(args):
    x0, iters = args
    from ilqgames_b200 import _abi as abi
    desc, params, _ = workload(_CFG, 1, 0)
    lib = abi.Library(ORACLE_LIB)
    h = abi.Handle(lib, desc, params, x0.shape[0])
    h.upload_x0(x0)
    t = time.perf_counter()
    h.reset(h.RESET_SOLVER)
    h.solve_begin()
    h.solve(chunk=iters)
    dt = time.perf_counter() - t
    done = int(h.download(abi.ITERS).sum())
    rolls = int(h.download(abi.BACKTRACKS).sum())
    h.close()
    return dt, done, rolls


def cpu_single_core(x0_sample):
    dt, done, rolls = _oracle_worker((x0_sample, ITERS_PER_SOLVE))
    return done / dt, done, rolls, dt


def cpu_reference_build(x0_sample):
    """The reference's OWN sources (oracle/_ref/libilqg_ref.so: /root/reference/src compiled against
    the Eigen/glog/gflags stand-ins of oracle/ref_shim) on the same games, one host core.  Reported
    next to the port, not instead of it: the stand-in's dense kernels are plain loops, so this is a
    lower bound on what the reference built on real Eigen would do.  None when the library was not
    built (it needs /root/reference at build time)."""
    if not os.path.exists(REF_BUILD_LIB) or CONFIGS[_CFG]["T"] != 100:
        return None  # (the reference's example Problems are built with T = 100)
    from tests.golden import ref_lib
    ref = ref_lib.RefLibrary(REF_BUILD_LIB)
    _, params, _ = workload(_CFG, 1, 0)
    which = {"three_player_intersection": ref_lib.INTERSECTION, "roundabout_merging": ref_lib.ROUNDABOUT,
             "air_3d": ref_lib.AIR3D}[CONFIGS[_CFG]["problem"]]
    rp = ref_lib.RefParams.from_abi(params)
    rp.convergence_tolerance = 0.0   # HasConverged needs |delta| < tolerance: never, like the GPU arm
    t = time.perf_counter()
    done = 0
    for x in x0_sample:
        done += ref.solve(which, ref_lib.ILQ, x, rp, mu0=10.0, max_log=1)["iterates"] - 1
    dt = time.perf_counter() - t
    return {"value": done / dt, "unit": "instance-iterations/s", "cores": 1,
            "sample": f"{len(x0_sample)} instances, {dt:.1f} s; reference sources on oracle/ref_shim (no Eigen3 in the image)"}


_worker_state = {}


def _worker_init():
    from ilqgames_b200 import _abi as abi
    desc, params, _ = workload(_CFG, 1, 0)
    _worker_state["abi"] = abi
    _worker_state["lib"] = abi.Library(ORACLE_LIB)
    _worker_state["desc"], _worker_state["params"] = desc, params
    _worker_state["handles"] = {}


def _worker_solve(x0):
    abi = _worker_state["abi"]
    key = x0.shape[0]
    h = _worker_state["handles"].get(key)
    if h is None:
        h = abi.Handle(_worker_state["lib"], _worker_state["desc"], _worker_state["params"], key)
        _worker_state["handles"][key] = h
    h.upload_x0(x0)
    h.reset(h.RESET_SOLVER)
    h.solve_begin()
    h.solve(chunk=ITERS_PER_SOLVE)
    return int(h.download(abi.ITERS).sum())


def run_reference(args):
    """--impl reference: the CPU path (oracle port of the reference's single-threaded Eigen
    algorithm) on all host cores, one process per core; each step is a bounded sample of the same
    workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = usable_cores()
    per_worker = 16
    _, _, x0 = workload(args.config, args.batch, args.seed, 0, min(args.batch, cores * per_worker))
    sample = x0
    chunks = [c for c in np.array_split(sample, cores) if len(c)]
    from ilqgames_b200 import _abi as abi
    desc_, params_, _ = workload(args.config, 1, 0)
    h_ = abi.Handle(abi.Library(ORACLE_LIB), desc_, params_, 1)
    ref_layout = h_.layout
    h_.close()
    ctx = mp.get_context("fork")
    times = []
    done_total = 0
    with ctx.Pool(len(chunks), initializer=_worker_init) as pool:
        for step in range(args.warmup + args.steps):
            t = time.perf_counter()
            res = pool.map(_worker_solve, chunks, chunksize=1)
            dt = time.perf_counter() - t
            if step >= args.warmup:
                times.append(dt)
                done_total += sum(res)
    total = sum(times)
    value = done_total / total
    line = {
        "impl": "reference", "metric": "ilq_instance_iterations_per_second", "value": value,
        "unit": "instance-iterations/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # same workload description as the repo arm; the CPU arm's step is a bounded SAMPLE of it
        # the repo arm's `config`, key for key; this arm's step is a bounded SAMPLE of that workload
        "config": bench_config(args.config, args.batch, max(1, args.gpus), args.seed, ref_layout),
        "workload_stats": {"reference_sample_instances": len(sample),
                           "reference_sample": f"each step solves the first {len(sample)} of the {args.batch} instances "
                                               f"({len(chunks)} processes x {per_worker})"},
        "cpu_baseline": {"value": value, "unit": "instance-iterations/s", "cores": len(chunks), "kind": "port",
                         "sample": f"first {len(sample)} instances of the batch x {ITERS_PER_SOLVE} iterations per step, "
                                   f"{len(chunks)} processes; oracle port of the reference's Eigen path: faster than "
                                   "the reference's own sources built on the Eigen stand-ins of oracle/ref_shim "
                                   "(reference_build), to which it is bit-identical (tests/test_ref_pins.py)",
                         "reference_build": cpu_reference_build(sample[:16])},
        "e2e": {"value": value, "unit": "instance-iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from ilqgames_b200 import _abi as abi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference "
                         "for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        # one JSON line only on stdout: whatever NCCL logs (its version banner at NCCL_DEBUG >= VERSION)
        # goes to a file
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/ilqg_nccl_%h_%p.log")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ONE global batch of world x batch games; each rank owns a contiguous slice of it (weak scaling:
    # fixed games per GPU): independent games, no hot-path collective
    from ilqgames_b200 import sharding
    lo_row, hi_row = sharding.shard_bounds(args.batch * world, world, rank)
    desc, params, x0 = workload(args.config, args.batch * world, args.seed, lo_row, hi_row)
    assert x0.shape[0] == args.batch
    lib = abi.product_library()
    h = abi.Handle(lib, desc, params, args.batch, local)
    # a non-default torch stream so torch.cuda.Event and the library's launches share it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    h.set_stream(stream.cuda_stream)
    h.upload_x0(x0)
    bytes_model = algorithmic_bytes(h.layout)

    def solve():
        h.reset(h.RESET_SOLVER)  # a fresh ILQSolver per solve (last merit = +inf)
        h.solve_begin()
        # ITERS_PER_SOLVE passes complete ITERS_PER_SOLVE iterations for every instance whose
        # linesearches resolve within their first window; stragglers (deep backtracking) finish in
        # extra passes.  count_running() is the only host sync of a solve.
        h.iterate(ITERS_PER_SOLVE)
        while h.count_running() > 0:
            h.iterate(2)

    # ---- device-resident timing ------------------------------------------------------
    for _ in range(args.warmup):
        solve()
    barrier()
    launches0 = h.kernel_launches()
    # working set of one iteration (records + both strategy buffers): above the 126 MB L2 for every
    # batched configuration; the small ones (c1) get an explicit L2 flush before each timed step
    need_flush, _ = l2_policy(args.batch, h.layout)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if need_flush else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        if not need_flush:
            ev0.record(stream)
            for _ in range(args.steps):
                solve()
            ev1.record(stream)
            barrier()
            elapsed_ms = ev0.elapsed_time(ev1)
        else:
            elapsed_ms = 0.0
            for _ in range(args.steps):
                flush_buf.zero_()
                ev0.record(stream)
                solve()
                ev1.record(stream)
                torch.cuda.synchronize()
                elapsed_ms += ev0.elapsed_time(ev1)
            barrier()
    launches = h.kernel_launches() - launches0
    iters = h.download(abi.ITERS)
    rollouts = h.download(abi.BACKTRACKS)  # cumulative over all solves since creation
    status = h.download(abi.STATUS)
    done_per_step = int(iters.sum())
    max_ms, total_done_per_step = sharding.reduce_metrics(elapsed_ms, done_per_step, device="cuda")
    value = total_done_per_step * args.steps / (max_ms * 1e-3)

    # ---- per-kernel profile pass (not part of the timed region above) ----------------
    # a second handle with a single group / stream, so each kernel is timed alone on the device
    # (the timed handle overlaps several groups' kernels on concurrent streams)
    groups_env = os.environ.get("ILQG_GROUPS")
    os.environ["ILQG_GROUPS"] = "1"
    hp = abi.Handle(lib, desc, params, args.batch, local)
    if groups_env is None:
        del os.environ["ILQG_GROUPS"]
    else:
        os.environ["ILQG_GROUPS"] = groups_env
    hp.set_stream(stream.cuda_stream)
    hp.upload_x0(x0)

    def solve_p():
        hp.reset(hp.RESET_SOLVER)
        hp.solve_begin()
        hp.iterate(ITERS_PER_SOLVE)

    solve_p()
    hp.profile(True)
    prof_steps = max(1, min(args.steps, 3))
    roll_before = int(hp.download(abi.BACKTRACKS).sum())
    for _ in range(prof_steps):
        solve_p()
    prof = hp.profile_read()
    hp.profile(False)
    roll_per_step = (int(hp.download(abi.BACKTRACKS).sum()) - roll_before) / prof_steps
    # how the linesearches of one solve end, iteration by iteration (one extra solve, stepped):
    # accepted at the first candidate / after backtracking / failed (all candidates rejected)
    hp.reset(hp.RESET_SOLVER)
    hp.solve_begin()
    prev_roll, prev_it = hp.download(abi.BACKTRACKS).astype(np.int64), hp.download(abi.ITERS).astype(np.int64)
    first_try = backtracked = failed = rolls_bt = 0
    max_bt = int(params.max_backtracking_steps)
    for _ in range(ITERS_PER_SOLVE):
        hp.iterate(1)
        roll, itn, st = (hp.download(w).astype(np.int64) for w in (abi.BACKTRACKS, abi.ITERS, abi.STATUS))
        d_roll, d_it = roll - prev_roll, itn - prev_it
        did = d_it > 0
        fail_now = did & (st == abi.STATUS_LINESEARCH_FAILED) & (d_roll == max_bt + 1)
        first_try += int((did & (d_roll == 1)).sum())
        failed += int(fail_now.sum())
        bt = did & (d_roll > 1) & ~fail_now
        backtracked += int(bt.sum())
        rolls_bt += int(d_roll[bt].sum())
        prev_roll, prev_it = roll, itn
    total_ls = max(first_try + backtracked + failed, 1)
    ls_split = {"accepted_first_try": first_try / total_ls, "backtracked": backtracked / total_ls,
                "failed": failed / total_ls, "mean_rollouts_when_backtracked": rolls_bt / max(backtracked, 1),
                "rollouts_of_a_failed_linesearch": max_bt + 1, "linesearches": total_ls}
    hp.close()
    kern = {}
    for name, (ms, n) in prof.items():
        if n:
            kern[name] = {"ms_per_launch": ms / n, "launches_per_step": n / prof_steps, "ms_per_step": ms / prof_steps}
    passes = max(kern.get("lq_backward", {}).get("launches_per_step", ITERS_PER_SOLVE), 1)
    inst_iters_per_launch = done_per_step / passes
    # a linesearch window (k_ls_rollout + k_ls_merit) runs once per window: the first window for every instance, then the
    # queued window chunk by chunk (chunks past the end of the queue exit at once) - one entry
    ev = [kern[k] for k in ("ls_eval_fresh", "ls_eval_queued") if k in kern]
    if ev:
        n_l = sum(e["launches_per_step"] for e in ev)
        ms = sum(e["ms_per_step"] for e in ev)
        kern["ls_eval"] = {"ms_per_launch": ms / n_l, "launches_per_step": n_l, "ms_per_step": ms}
    per_kernel_bytes = {
        "linearize_quadraticize": bytes_model["linearize_quadraticize"] * inst_iters_per_launch,
        "lq_backward": bytes_model["lq_backward"] * inst_iters_per_launch,
    }
    compact_bytes = {k: v * inst_iters_per_launch for k, v in bytes_model.get("compact", {}).items()}
    if "ls_eval" in kern:
        per_kernel_bytes["ls_eval"] = bytes_model["linesearch_per_rollout"] * roll_per_step / kern["ls_eval"]["launches_per_step"]
    hot = [k for k in ("linearize_quadraticize", "lq_backward", "ls_eval") if k in kern]
    dominant = max(hot, key=lambda k: kern[k]["ms_per_step"])
    peaks_path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = per_kernel_bytes[dominant] / (kern[dominant]["ms_per_launch"] * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(REPO, "profiles", "traffic.json")
    if os.path.exists(tpath) and args.config == "metric":
        traffic = json.load(open(tpath)).get(dominant)  # ncu --set full, per launch, metric configuration
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": per_kernel_bytes[dominant],
                "byte_model": "SURVEY 8(d): dense lin/quad records materialised once in HBM; the round-2 kernels move "
                              "compact records instead (designed_bytes_per_launch), so `traffic` sits far below "
                              "the algorithmic bytes",
                "designed_bytes_per_launch": compact_bytes.get(dominant),
                "kernels": {k: dict(v, algorithmic_GBps=per_kernel_bytes[k] / (v["ms_per_launch"] * 1e-3) / 1e9
                                    if k in per_kernel_bytes else None,
                                    designed_GBps=compact_bytes[k] / (v["ms_per_launch"] * 1e-3) / 1e9
                                    if k in compact_bytes else None) for k, v in kern.items()}}
    # ---- end to end through the C ABI with host buffers ------------------------------
    pin_x0 = torch.from_numpy(x0).pin_memory()
    lo = h.layout
    pin_xs = torch.empty((args.batch, lo.num_time_steps, lo.xdim), dtype=torch.float32).pin_memory()
    pin_us = torch.empty((args.batch, lo.num_time_steps, lo.total_udim), dtype=torch.float32).pin_memory()
    pin_status = torch.empty((args.batch,), dtype=torch.int32).pin_memory()
    pin_iters = torch.empty((args.batch,), dtype=torch.int32).pin_memory()
    h2d = pin_x0.numel() * 4
    d2h = (pin_xs.numel() + pin_us.numel()) * 4 + (pin_status.numel() + pin_iters.numel()) * 4

    def e2e_step():
        h.upload_x0_ptr(pin_x0.data_ptr(), h2d)
        solve()
        h.download_ptr(abi.XS, pin_xs.data_ptr(), pin_xs.numel() * 4)
        h.download_ptr(abi.US, pin_us.data_ptr(), pin_us.numel() * 4)
        h.download_ptr(abi.STATUS, pin_status.data_ptr(), pin_status.numel() * 4)
        h.download_ptr(abi.ITERS, pin_iters.data_ptr(), pin_iters.numel() * 4)

    for _ in range(max(1, min(args.warmup, 2))):
        e2e_step()
    barrier()
    e2e_steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_done = int(pin_iters.sum().item())
    e2e_max_s, e2e_total = sharding.reduce_metrics(e2e_s, e2e_done, device="cuda")
    e2e_value = e2e_total * e2e_steps / e2e_max_s

    # ---- gather of converged trajectories over NCCL (off the hot path, reported apart) ----
    gather_ms = None
    if world > 1:
        dev_xs = pin_xs.cuda(non_blocking=True)
        out = torch.empty((world,) + tuple(dev_xs.shape), dtype=dev_xs.dtype, device="cuda")
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        dist.all_gather_into_tensor(out.view(-1), dev_xs.view(-1))
        g1.record()
        torch.cuda.synchronize()
        gather_ms = g0.elapsed_time(g1)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sample = x0[: min(args.batch, args.cpu_sample)]
        v, done, rolls, dt = cpu_single_core(sample)
        cpu = {"value": v, "unit": "instance-iterations/s", "cores": 1, "kind": "port",
               "sample": f"first {len(sample)} instances of the same batch, {ITERS_PER_SOLVE} iterations each, "
                         f"{dt:.1f} s on one host core (of {os.cpu_count()}); oracle port of the reference's "
                         "Eigen path (bit-identical to the reference's own sources built on Eigen stand-ins, "
                         "tests/test_ref_pins.py; the image has no Eigen3/glog/gflags)"}
        ref_build = cpu_reference_build(sample[:16])
        if ref_build:
            cpu["reference_build"] = ref_build

    if rank == 0:
        hist = np.bincount(status, minlength=6).tolist()
        line = {
            "metric": "ilq_instance_iterations_per_second", "value": value, "unit": "instance-iterations/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": max_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(args.config, args.batch, world, args.seed, h.layout),
            "workload_stats": {"instance_iterations_per_step": total_done_per_step,
                               "mean_rollouts_per_iteration": roll_per_step / max(done_per_step, 1),
                               "status_histogram_rank0": hist,
                               "linesearch_split": ls_split},
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": "instance-iterations/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps},
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "gather_ms": gather_ms,
        }
        if args.with_al and world == 1:
            line["al_solve"] = al_solve_table(lib, args.batch, args.seed, local)
        emit(line)
    h.close()
    if world > 1:
        dist.destroy_process_group()


def al_solve_table(lib, batch: int, seed: int, device: int = 0, al_iterates_cap: int = 100, repeats: int = 2):
    """SURVEY 8(d)'s second table: the FULL AugmentedLagrangianSolver::Solve of the batch (every
    game runs its own outer loop: inner ILQSolver solves capped at unconstrained_solver_max_iters,
    multiplier sweeps, exit on constraint error or the NumIterates cap), timed on the host clock
    around a synchronised solve.  Works on any library implementing include/ilqg.h."""
    from ilqgames_b200 import _abi as abi, al, problems
    desc, _ = problems.three_player_intersection()
    p = problems.three_player_intersection_params()
    p.max_solver_iters = p.unconstrained_solver_max_iters   # the handle's cap is the inner solver's
    x0 = problems.three_player_intersection_x0_batch(batch, seed)
    h = abi.Handle(lib, desc, p, batch, device)
    h.upload_x0(x0)
    best = None
    for _ in range(repeats + 1):          # first pass = warm-up
        h.reset(h.RESET_SOLVER | h.RESET_MULTIPLIERS | h.RESET_SOLUTION)
        h.synchronize()
        t0 = time.perf_counter()
        out = al.solve_augmented_lagrangian(h, al_iterates_cap, p.constraint_error_tolerance)
        h.synchronize()
        dt = time.perf_counter() - t0
        row = {"seconds": dt, "inner_solves": out.rounds, "logged_iterates": int(out.iterates.sum()),
               "logged_iterates_per_s": int(out.iterates.sum()) / dt, "games": batch,
               "games_per_s": batch / dt, "success_fraction": float(out.success.mean()),
               "al_iterates_cap": al_iterates_cap}
        if best is None or row["seconds"] < best["seconds"]:
            best = row
    h.close()
    return best


_RESULT_FD = None


def emit(line: dict):
    """The one JSON line of the contract, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    # stdout carries exactly one JSON line: whatever libraries write to fd 1 meanwhile (NCCL prints
    # its version banner there) is sent to stderr instead
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="metric", choices=sorted(CONFIGS),
                    help="metric (default: the batch-4096 configuration BASELINE.json's metric is quoted on) or "
                         "c1..c5 = BASELINE.json configs[0..4]")
    ap.add_argument("--batch", type=int, default=None, help="instances per GPU (default: the configuration's)")
    ap.add_argument("--seed", type=int, default=4096)
    ap.add_argument("--cpu-sample", type=int, default=256)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--with-al", action="store_true",
                    help="also time the full augmented-Lagrangian solve of the batch (adds `al_solve`; N = 1 only)")
    args = ap.parse_args()
    global _CFG
    _CFG = args.config
    if args.batch is None:
        args.batch = CONFIGS[args.config]["batch"]
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
