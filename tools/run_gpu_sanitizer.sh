# compute-sanitizer over the kernels written this session (stage-parallel rollout, K_lq v4 phase 2, decide, tiers)
for tool in memcheck racecheck initcheck; do
  for cfg in "64 2 c1 finish" "32 2 c3 finish" "64 2 overtaking finish" "64 2 c4 finish"; do
    timeout 900 compute-sanitizer --tool $tool python tools/profile_target.py $cfg 2>&1 | grep -v "^=========     \|^========= $" | tail -4 | sed "s/^/[$tool $cfg] /"
  done
done > gpurun_out/sanitizer_r02b.txt 2>&1
cat gpurun_out/sanitizer_r02b.txt
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "rollout_kernels and (two_player_reachability or dubins or modified_air or collision_avoidance)" 2>&1 | tail -4
