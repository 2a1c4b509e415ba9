python bench.py --steps 10 --warmup 3 > gpurun_out/bench18.json 2>gpurun_out/bench18.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench18.json"))
print(round(d["value"]), round(d["ms_per_step"],2), round(d["e2e"]["value"]), d["clocks"], {k:round(v["ms_per_launch"],3) for k,v in d["roofline"]["kernels"].items()})
PY
export ILQG_GROUPS=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ls_eval -s 1 -c 1 -f -o gpurun_out/r01c_k_ls_eval_fresh python tools/profile_target.py 4096 3 > gpurun_out/ncu_r01c.log 2>&1
