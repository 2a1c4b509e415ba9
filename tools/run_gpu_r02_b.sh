rm -f gpurun_out/parity_counts.jsonl
timeout 800 python -m pytest tests -m gpu -q -x 2>&1 | tail -70 > gpurun_out/pytest_r02d.txt
tail -4 gpurun_out/pytest_r02d.txt
for c in metric c2 c3 c4 c5; do
  timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_r02_$c.json 2> gpurun_out/bench_r02_$c.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_r02_$c.json"))
    k = d["roofline"]["kernels"]
    print("$c", round(d["value"]), "inst-it/s", round(d["ms_per_step"], 2), "ms/step e2e", round(d["e2e"]["value"]), d["config"]["status_histogram_rank0"], {n: round(v["ms_per_launch"], 3) for n, v in k.items()}, d["config"]["linesearch_split"])
except Exception as e:
    print("$c failed", e, open("gpurun_out/bench_r02_$c.err").read()[-600:])
PY
done
