python -m pytest tests -m gpu -x -q > gpurun_out/pytest16.log 2>&1; tail -5 gpurun_out/pytest16.log
python bench.py > gpurun_out/bench16.json 2>gpurun_out/bench16.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench16.json"))
print(round(d["value"]), round(d["ms_per_step"],2), round(d["e2e"]["value"]), d["clocks"], d["roofline"]["kernel"], round(d["roofline"]["frac"],4), d["roofline"]["traffic"], {k:round(v["ms_per_launch"],3) for k,v in d["roofline"]["kernels"].items()})
print(d["cpu_baseline"])
PY
cat /sys/fs/cgroup/cpu.max 2>&1; nproc; lscpu | grep -i "model name\|socket\|thread\|core(s)" ; cat /proc/loadavg
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench16_ref.json 2>gpurun_out/bench16_ref.err; cut -c1-200 gpurun_out/bench16_ref.json; python - <<PY
import json
d=json.load(open("gpurun_out/bench16_ref.json")); print(d["value"], d["ms_per_step"], d["cpu_baseline"]["cores"])
PY
cat /proc/loadavg
