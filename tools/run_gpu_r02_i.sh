show() { python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$1', round(d['value']), round(d['ms_per_step'],2), d['config']['status_histogram_rank0'])"; }
for pl in 2 1 0; do
ILQG_PIPELINE=$pl python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | show pipeline=$pl
done
ILQG_PIPELINE=1 ILQG_TRACE=gpurun_out/trace_c1_p1.txt python tools/profile_target.py 4096 8 c1 > /dev/null
python tools/trace_view.py gpurun_out/trace_c1_p1.txt 3 2
