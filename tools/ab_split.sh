for v in 0 1; do ILQG_LS_SPLIT=$v python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_split$v.json 2>gpurun_out/bench_split$v.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_split$v.json"))
print("split=$v", round(d["value"]), round(d["ms_per_step"],2), d["config"]["status_histogram_rank0"], round(d["config"]["mean_rollouts_per_iteration"],4), {k: round(v["ms_per_launch"],3) for k,v in d["roofline"]["kernels"].items()})
PY
done
