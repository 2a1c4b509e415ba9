#!/bin/bash
# A/B of prebuilt library variants (build_variants/*.so, built on the CPU container) on one box:
# usage: run_gpu_variants.sh "<so> [ENV=VAL ...] [-- bench args]" ...
# each variant is copied over the product library and run through a short bench.
mkdir -p gpurun_out
cp ilqgames_b200/lib/libilqg_b200.so /tmp/keep.so
i=0
for spec in "$@"; do
  i=$((i+1))
  set -- $spec
  so=$1; shift
  envs=""; args=""
  while [ $# -gt 0 ]; do
    if [ "$1" = "--" ]; then shift; args="$*"; break; fi
    envs="$envs $1"; shift
  done
  cp build_variants/$so.so ilqgames_b200/lib/libilqg_b200.so
  env $envs python bench.py --steps 10 --warmup 3 --no-cpu $args > gpurun_out/var_$i.json 2> gpurun_out/var_$i.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/var_$i.json"))
    print("$spec |", round(d["value"]), "inst-iter/s", round(d["ms_per_step"], 2), "ms/step",
          {k: round(v["ms_per_launch"], 3) for k, v in d["roofline"]["kernels"].items() if k in ("linearize_quadraticize", "lq_backward", "ls_eval_fresh", "ls_eval_queued")})
except Exception as e:
    print("$spec FAILED", e, open("gpurun_out/var_$i.err").read()[-600:])
PY
done
cp /tmp/keep.so ilqgames_b200/lib/libilqg_b200.so
