# part 2: every BASELINE config, the AL solve, the reference arm; captures of the tier-1 window kernels and the decide kernel
R=r02
python bench.py > gpurun_out/bench_${R}_metric.json 2> gpurun_out/bench_${R}_metric.err
cut -c1-300 gpurun_out/bench_${R}_metric.json
for c in c2 c3 c4 c5; do
  python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/bench_${R}_$c.json 2> gpurun_out/bench_${R}_$c.err
  cut -c1-200 gpurun_out/bench_${R}_$c.json
done
python bench.py --with-al --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_${R}_al.json 2> gpurun_out/bench_${R}_al.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${R}_reference.json 2> gpurun_out/bench_${R}_reference.err
export ILQG_GROUPS=1 ILQG_PIPELINE=0
for k in k_ls_rollout_sp:5:ls_rollout_tier1 k_ls_merit:5:ls_merit_tier1 k_ls_decide:2:ls_decide; do
  IFS=: read name skip tag <<< "$k"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$name -s $skip -c 1 -f -o gpurun_out/${R}_$tag python tools/profile_target.py 4096 5 > gpurun_out/ncu_${R}_$tag.log 2>&1
done
du -sh gpurun_out
