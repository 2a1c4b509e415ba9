for cfg in "99 512" "99 2048" "64 2048" "48 2048" "32 2048"; do set -- $cfg; ILQG_LS_JB=$1 ILQG_LS_CAP=$2 python bench.py --steps 10 --warmup 3 > gpurun_out/bench14.json 2>gpurun_out/bench14.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench14.json"))
print("$cfg", round(d["value"]), round(d["ms_per_step"],2), {k:round(v["ms_per_launch"],3) for k,v in d["roofline"]["kernels"].items() if k.startswith("ls_") or k=="linesearch"}, d["roofline"]["kernels"]["ls_eval_queued"]["launches_per_step"])
PY
done
