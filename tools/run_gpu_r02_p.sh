rm -f gpurun_out/parity_counts.jsonl
timeout 1000 python -m pytest tests -m gpu -q 2>&1 | grep "^E  \|^tests/\|^___\|passed\|failed\|^FAILED\|skipped" | grep -v "ACTUAL\|DESIRED\|^E   *\[" | head -40 > gpurun_out/pytest_r02q.txt; cat gpurun_out/pytest_r02q.txt
show() { python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$1', round(d['value']), round(d['ms_per_step'],2), d['config']['status_histogram_rank0'], round(d['config']['linesearch_split']['mean_rollouts_when_backtracked'],2), {k: round(v['ms_per_launch'],3) for k,v in d['roofline']['kernels'].items()})"; }
python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | tee gpurun_out/bench_r02q_metric.json | show metric
for c in c3 c4 c2; do
python bench.py --config $c --steps 5 --warmup 3 --no-cpu 2>/dev/null | tee gpurun_out/bench_r02q_$c.json | show $c
done
ILQG_ROLLOUT=sp python bench.py --config c3 --steps 5 --warmup 3 --no-cpu 2>/dev/null | show c3:sp
ILQG_ROLLOUT=lanes python bench.py --config c3 --steps 5 --warmup 3 --no-cpu 2>/dev/null | show c3:lanes
ILQG_LS_TIERS=7,32 python bench.py --config c4 --steps 5 --warmup 3 --no-cpu 2>/dev/null | show c4:7,32
