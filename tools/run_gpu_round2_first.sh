# First GPU call of the next round: everything that was written after round 1's GPU minutes ran out.
#   1. GPU tests (the two n = 18 tests are non-strict xfail: look for XPASS / xfailed in the summary)
#   2. compute-sanitizer on the n = 18 path and on k_integrate_plan
#   3. the K_bwd dependent-chain microbenchmark (FFMA vs mma.sync TF32 x3 / x1)
#   4. the default bench line plus the full-AL-solve table
R=${1:-r02}
timeout 600 python -m pytest tests -m gpu -q -rxX 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_${R}.txt
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool python tools/profile_target.py 64 2 overtaking finish 2>&1 | tail -4 | tee gpurun_out/sanitizer_${tool}_overtaking_${R}.txt
done
bash tools/run_gpu_mma_bench.sh
python bench.py --with-al > gpurun_out/bench_${R}.json 2> gpurun_out/bench_${R}.err
cut -c1-300 gpurun_out/bench_${R}.json
python -c "import json; d = json.load(open('gpurun_out/bench_${R}.json')); print(d.get('al_solve'))"
