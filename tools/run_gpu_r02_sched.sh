timeout 120 python tools/stage_errors.py three_player_intersection 16 > gpurun_out/stage_err_c1.txt 2>&1
timeout 120 python tools/stage_errors.py roundabout_merging 16 > gpurun_out/stage_err_c3.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_r02b.txt
for pl in 0 1 2; do for gr in 1 2; do
ILQG_PIPELINE=$pl ILQG_GROUPS=$gr python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('pipeline',$pl,'groups',$gr, round(d['value']), d['ms_per_step'], d['config'].get('status_histogram_rank0'))"
done; done > gpurun_out/sched_r02a.txt 2>&1
cat gpurun_out/sched_r02a.txt
