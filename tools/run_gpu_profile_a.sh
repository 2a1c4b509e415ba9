# data for the round-2 documents, part 1: ncu --set full capture of the four hot kernels + launch list of the bench command
R=r02
export ILQG_GROUPS=1 ILQG_PIPELINE=0
for k in k_linearize_quadraticize_v4:2:linearize_quadraticize k_lq_backward_tc:2:lq_backward k_ls_rollout_sp:4:ls_rollout k_ls_merit:4:ls_merit; do
  IFS=: read name skip tag <<< "$k"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$name -s $skip -c 1 -f -o gpurun_out/${R}_$tag python tools/profile_target.py 4096 5 > gpurun_out/ncu_${R}_$tag.log 2>&1
done
unset ILQG_GROUPS ILQG_PIPELINE
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_${R}.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/launches_${R}_bench.log 2>&1
ls -la gpurun_out; du -sh gpurun_out
