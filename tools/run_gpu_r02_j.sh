rm -f gpurun_out/parity_counts.jsonl
timeout 1000 python -m pytest tests -m gpu -q 2>&1 | grep "^E  \|^tests/\|^___\|passed\|failed\|^FAILED\|skipped" | grep -v "ACTUAL\|DESIRED\|^E   *\[" | head -40 > gpurun_out/pytest_r02o.txt; cat gpurun_out/pytest_r02o.txt
show() { python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$1', round(d['value']), round(d['ms_per_step'],2), d['config']['status_histogram_rank0'], round(d['config']['linesearch_split']['mean_rollouts_when_backtracked'],2), {k: round(v['ms_per_launch'],3) for k,v in d['roofline']['kernels'].items()})"; }
python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | tee gpurun_out/bench_r02o_metric.json | show metric
ILQG_PIPELINE=0 ILQG_TRACE=gpurun_out/trace_c1_p0.txt python tools/profile_target.py 4096 8 c1 > /dev/null
python tools/trace_view.py gpurun_out/trace_c1_p0.txt 4 1
ILQG_TRACE=gpurun_out/trace_c1.txt python tools/profile_target.py 4096 8 c1 > /dev/null
python tools/trace_view.py gpurun_out/trace_c1.txt 4 1
