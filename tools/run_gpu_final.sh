# final pass of a round: GPU tests, smoke(), every BASELINE config, the AL solve, the reference arm
R=${1:-r02f}
rm -f gpurun_out/parity_counts.jsonl
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | grep "^E  \|^tests/\|^___\|passed\|failed\|^FAILED\|skipped" | grep -v "ACTUAL\|DESIRED\|^E   *\[" | head -40 > gpurun_out/pytest_${R}.txt; cat gpurun_out/pytest_${R}.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_${R}_metric.json 2> gpurun_out/bench_${R}_metric.err
cut -c1-300 gpurun_out/bench_${R}_metric.json
for c in c1 c2 c3 c4 c5; do
  python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/bench_${R}_$c.json 2> gpurun_out/bench_${R}_$c.err
  cut -c1-200 gpurun_out/bench_${R}_$c.json
done
python bench.py --with-al --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_${R}_al.json 2> gpurun_out/bench_${R}_al.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${R}_reference.json 2> gpurun_out/bench_${R}_reference.err
cut -c1-200 gpurun_out/bench_${R}_reference.json
ILQG_TRACE=gpurun_out/trace_${R}.txt python tools/profile_target.py 4096 8 c1 > /dev/null; python tools/trace_view.py gpurun_out/trace_${R}.txt 4 1
ILQG_PIPELINE=0 ILQG_TRACE=gpurun_out/trace_${R}_p0.txt python tools/profile_target.py 4096 8 c1 > /dev/null; python tools/trace_view.py gpurun_out/trace_${R}_p0.txt 4 1
