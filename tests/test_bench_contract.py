"""bench.py's output contract on the CPU arm (`--impl reference`): exactly one JSON line on stdout
with the keys the driver reads.  The GPU arm prints the same line plus `roofline` (checked on the
GPU box by the driver's own run)."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    res = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--batch", "32"], capture_output=True, text=True, cwd=REPO, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "ilq_instance_iterations_per_second"
    assert line["unit"] == "instance-iterations/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["steps"] == 1 and line["warmup"] == 0
    cpu = line["cpu_baseline"]
    assert cpu["kind"] == "port" and cpu["cores"] >= 1 and cpu["value"] == line["value"] and cpu["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]
    # the reference arm prints the repo arm's `config`, key for key (the driver compares the two): nothing
    # measured and nothing that depends on the host lives in it
    import bench
    from ilqgames_b200 import _abi as abi
    desc, params, _ = bench.workload("metric", 1, 0)
    h = abi.Handle(abi.Library(bench.ORACLE_LIB), desc, params, 1)
    assert line["config"] == bench.bench_config("metric", 32, 1, 4096, h.layout)
    assert set(line["config"]) == {"workload", "config", "batch_per_gpu", "global_batch", "iterations_per_step", "seed",
                                   "l2", "parallelism", "reference_arm"}
    assert line["workload_stats"]["reference_sample_instances"] >= 16


def test_gpu_arm_refuses_to_run_without_a_device():
    """No CPU fallback: without CUDA the default arm exits with an error instead of timing the oracle."""
    import torch
    if torch.cuda.is_available():
        return
    res = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, cwd=REPO, timeout=600)
    assert res.returncode != 0 and "CUDA" in (res.stderr + res.stdout)
    assert not [l for l in res.stdout.splitlines() if l.strip().startswith("{")]


def test_al_solve_table_runs_through_the_abi(oracle):
    """bench.py --with-al (SURVEY 8d's second table, the full AugmentedLagrangianSolver::Solve of
    the batch): the helper is library-agnostic, so the oracle can stand in here."""
    import bench
    row = bench.al_solve_table(oracle, 3, 1024, 0, al_iterates_cap=25, repeats=0)
    assert row["games"] == 3 and row["inner_solves"] >= 2
    assert row["logged_iterates"] >= 3 * 2 and row["seconds"] > 0
    assert 0.0 <= row["success_fraction"] <= 1.0
