rm -f gpurun_out/parity_counts.jsonl
timeout 1000 python -m pytest tests -m gpu -q 2>&1 | grep "^E  \|^tests/\|^___\|passed\|failed\|^FAILED\|skipped" | grep -v "ACTUAL\|DESIRED\|^E   *\[" | head -40 > gpurun_out/pytest_r02r.txt; cat gpurun_out/pytest_r02r.txt
show() { python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$1', round(d['value']), round(d['ms_per_step'],2), d['config']['status_histogram_rank0'], round(d['config']['linesearch_split']['mean_rollouts_when_backtracked'],2), {k: round(v['ms_per_launch'],3) for k,v in d['roofline']['kernels'].items()})"; }
python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | tee gpurun_out/bench_r02r_metric.json | show metric
python bench.py --config c3 --steps 5 --warmup 3 --no-cpu 2>/dev/null | show c3
ILQG_TRACE=gpurun_out/trace_c1_r.txt python tools/profile_target.py 4096 8 c1 > /dev/null
python tools/trace_view.py gpurun_out/trace_c1_r.txt 4 1
