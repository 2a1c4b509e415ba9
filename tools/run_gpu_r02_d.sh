rm -f gpurun_out/parity_counts.jsonl
timeout 1000 python -m pytest tests -m gpu -q 2>&1 | grep "^E  \|^tests/\|^___\|passed\|failed\|^FAILED\|skipped" | head -40 > gpurun_out/pytest_r02k.txt; cat gpurun_out/pytest_r02k.txt
python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(round(d['value']), round(d['ms_per_step'],2), {k: round(v['ms_per_launch'],3) for k,v in d['roofline']['kernels'].items()})"
for tool in memcheck racecheck; do
  for cfg in "64 2 c1 finish" "32 2 c3 finish" "64 2 overtaking finish" "64 2 c4 finish"; do
    timeout 600 compute-sanitizer --tool $tool python tools/profile_target.py $cfg 2>&1 | tail -3 | sed "s/^/[$tool $cfg] /"
  done
done > gpurun_out/sanitizer_r02.txt 2>&1
cat gpurun_out/sanitizer_r02.txt
