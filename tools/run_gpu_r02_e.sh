# stage-parallel rollout: the bit-identity test first, then the whole GPU suite, then the bench A/B
rm -f gpurun_out/parity_counts.jsonl
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "rollout_kernels" 2>&1 | grep "^E  \|^tests/\|^___\|passed\|failed\|^FAILED\|skipped" | head -60 > gpurun_out/pytest_r02l_rollout.txt; cat gpurun_out/pytest_r02l_rollout.txt
timeout 1000 python -m pytest tests -m gpu -q 2>&1 | grep "^E  \|^tests/\|^___\|passed\|failed\|^FAILED\|skipped" | head -40 > gpurun_out/pytest_r02l.txt; cat gpurun_out/pytest_r02l.txt
for r in lanes sp; do
ILQG_ROLLOUT=$r python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | tee gpurun_out/bench_r02l_$r.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$r', round(d['value']), round(d['ms_per_step'],2), d['config']['status_histogram_rank0'], d['config']['linesearch_split']['mean_rollouts_when_backtracked'], {k: round(v['ms_per_launch'],3) for k,v in d['roofline']['kernels'].items()})"
done
