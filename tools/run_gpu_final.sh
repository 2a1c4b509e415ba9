# round-end check: GPU tests, smoke, launch list of the bench command, default bench
R=${1:-r01s2}
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_${R}.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/launches_${R}_bench.log 2>&1
python bench.py > gpurun_out/bench_${R}.json 2> gpurun_out/bench_${R}.err
cut -c1-300 gpurun_out/bench_${R}.json
