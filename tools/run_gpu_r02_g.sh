# tiers of the continued linesearch + rollout micro-optimisations: GPU suite, then tier A/B on the three headline configs
rm -f gpurun_out/parity_counts.jsonl
timeout 1000 python -m pytest tests -m gpu -q 2>&1 | grep "^E  \|^tests/\|^___\|passed\|failed\|^FAILED\|skipped" | grep -v "ACTUAL\|DESIRED\|^E   *\[" | head -40 > gpurun_out/pytest_r02n.txt; cat gpurun_out/pytest_r02n.txt
show() { python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$1', round(d['value']), round(d['ms_per_step'],2), d['config']['status_histogram_rank0'], round(d['config']['linesearch_split']['mean_rollouts_when_backtracked'],2), {k: round(v['ms_per_launch'],3) for k,v in d['roofline']['kernels'].items()})"; }
for t in 99 39 47 31 7,32 15,32; do
ILQG_LS_TIERS=$t python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | tee gpurun_out/bench_r02n_metric_$t.json | show metric:$t
done
for c in c3 c4; do for t in 99 39 7,32; do
ILQG_LS_TIERS=$t python bench.py --config $c --steps 5 --warmup 3 --no-cpu 2>/dev/null | tee gpurun_out/bench_r02n_${c}_$t.json | show $c:$t
done; done
