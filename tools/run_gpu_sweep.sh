for cfg in "0 0 4" "1 1 4" "2 0 4" "2 1 4" "2 1 1" "2 1 2" "1 1 2"; do set -- $cfg; ILQG_PIPELINE=$1 ILQG_SIDE_PRIORITY=$2 ILQG_GROUPS=$3 python bench.py --steps 10 --warmup 3 > gpurun_out/bench22.json 2>gpurun_out/bench22.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench22.json"))
print("pipeline/prio/groups $cfg", round(d["value"]), round(d["ms_per_step"],2), d["config"]["status_histogram_rank0"])
PY
done
