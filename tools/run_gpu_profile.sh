# profile pass: one full ncu capture of each hot kernel + launch list of the bench command
R=${1:-r01s2}
export ILQG_GROUPS=1 ILQG_PIPELINE=0
for k in k_ls_rollout:3:ls_rollout k_ls_rollout:4:ls_rollout_queued k_ls_merit:3:ls_merit k_ls_merit:4:ls_merit_queued k_lq_backward_hw:1:lq_backward k_linearize_quadraticize_v3:1:linearize_quadraticize; do
  IFS=: read name skip tag <<< "$k"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$name -s $skip -c 1 -f -o gpurun_out/${R}_$tag python tools/profile_target.py 4096 5 > gpurun_out/ncu_${R}_$tag.log 2>&1
done
unset ILQG_GROUPS ILQG_PIPELINE
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_${R}.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/launches_${R}_bench.log 2>&1
python bench.py > gpurun_out/bench_${R}.json 2> gpurun_out/bench_${R}.err
cut -c1-600 gpurun_out/bench_${R}.json
