cp ilqgames_b200/lib/libilqg_b200.so /tmp/orig.so
for u in 2 4 8 16; do cp ilqgames_b200/lib/variants/libilqg_b200_u$u.so ilqgames_b200/lib/libilqg_b200.so
python bench.py --steps 10 --warmup 3 > gpurun_out/bench19.json 2>gpurun_out/bench19.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench19.json"))
print("u$u", round(d["value"]), round(d["ms_per_step"],2), {k:round(v["ms_per_launch"],3) for k,v in d["roofline"]["kernels"].items() if k in ("lq_backward","linearize_quadraticize","ls_eval_fresh")})
PY
done
cp /tmp/orig.so ilqgames_b200/lib/libilqg_b200.so
