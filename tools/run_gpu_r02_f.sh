# stage-parallel rollout, second cut: bit-identity test, whole GPU suite, bench A/B, one ncu capture
rm -f gpurun_out/parity_counts.jsonl
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "rollout_kernels" 2>&1 | grep "^E  \|^tests/\|^___\|passed\|failed\|^FAILED\|skipped" | grep -v "ACTUAL\|DESIRED\|^E   *\[" | head -60 > gpurun_out/pytest_r02m_rollout.txt; cat gpurun_out/pytest_r02m_rollout.txt
timeout 1000 python -m pytest tests -m gpu -q -x 2>&1 | grep "^E  \|^tests/\|^___\|passed\|failed\|^FAILED\|skipped" | head -40 > gpurun_out/pytest_r02m.txt; cat gpurun_out/pytest_r02m.txt
for r in lanes sp; do
ILQG_ROLLOUT=$r python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | tee gpurun_out/bench_r02m_$r.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$r', round(d['value']), round(d['ms_per_step'],2), d['config']['status_histogram_rank0'], d['config']['linesearch_split']['mean_rollouts_when_backtracked'], {k: round(v['ms_per_launch'],3) for k,v in d['roofline']['kernels'].items()})"
done
export ILQG_GROUPS=1 ILQG_PIPELINE=0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ls_rollout_sp -s 3 -c 1 -f -o gpurun_out/r02e_ls_rollout_sp python tools/profile_target.py 4096 5 > gpurun_out/ncu_r02e_ls_rollout_sp.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r02e_target.csv python tools/profile_target.py 4096 3 > /dev/null 2>&1
grep -o '"k_[a-z_0-9]*[^"]*","[^"]*","[^"]*","[^"]*","[0-9.]*"$' gpurun_out/launches_r02e_target.csv | head -0
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches_r02e_target.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
for r in rows[hdr+1:]:
    if len(r)>5 and 'gpu__time_duration' in r[-3]: print(r[4][:60], r[-1])
PY
