python -m pytest tests -m gpu -x -q > gpurun_out/pytest13.log 2>&1; tail -3 gpurun_out/pytest13.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench13.json 2>gpurun_out/bench13.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench13.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"], {k:round(v["ms_per_launch"],3) for k,v in d["roofline"]["kernels"].items()})
PY
