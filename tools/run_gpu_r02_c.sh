rm -f gpurun_out/parity_counts.jsonl
timeout 800 python -m pytest tests -m gpu -q 2>&1 | tail -70 > gpurun_out/pytest_r02e.txt
tail -4 gpurun_out/pytest_r02e.txt
for cfg in "1 0 2" "2 0 2" "2 1 2" "3 1 2" "4 1 2" "2 1 1" "2 1 0" "4 1 1"; do
set -- $cfg
ILQG_GROUPS=$1 ILQG_STAGGER=$2 ILQG_PIPELINE=$3 python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('groups $1 stagger $2 pipeline $3', round(d['value']), round(d['ms_per_step'],2), d['config'].get('status_histogram_rank0'))"
done > gpurun_out/sched_r02b.txt 2>&1
cat gpurun_out/sched_r02b.txt
