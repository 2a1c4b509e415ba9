for cfg in "1 4096" "2 4096" "3 4096" "4 4096"; do set -- $cfg; ILQG_LS_JA=$1 python bench.py --steps 10 --warmup 3 > gpurun_out/bench23.json 2>gpurun_out/bench23.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench23.json"))
print("JA $1", round(d["value"]), round(d["ms_per_step"],2), d["config"]["status_histogram_rank0"], round(d["config"]["mean_rollouts_per_iteration"],2), {k:round(v["ms_per_launch"],3) for k,v in d["roofline"]["kernels"].items() if k.startswith("ls_")})
PY
done
