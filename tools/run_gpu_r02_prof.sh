# profile pass (round 2): one full ncu capture of each hot kernel
R=${1:-r02}
export ILQG_GROUPS=1 ILQG_PIPELINE=0
for k in k_lq_backward_tc:1:lq_backward_tc k_linearize_quadraticize_v4:1:linearize_quadraticize_v4 k_ls_rollout:3:ls_rollout k_ls_merit:3:ls_merit; do
  IFS=: read name skip tag <<< "$k"
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$name -s $skip -c 1 -f -o gpurun_out/${R}_$tag python tools/profile_target.py 4096 5 > gpurun_out/ncu_${R}_$tag.log 2>&1
done
ls -la gpurun_out/${R}_*.ncu-rep
