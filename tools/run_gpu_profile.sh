# round profile pass: launch list of the bench command + one full capture of each hot kernel
R=${1:-r01}
export ILQG_GROUPS=1
for k in k_ls_eval:1 k_lq_backward_hw:0 k_linearize_quadraticize_v3:0; do
  name=${k%%:*}; skip=${k##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$name -s $skip -c 1 -f -o gpurun_out/${R}_$name python tools/profile_target.py 4096 2 > gpurun_out/ncu_${R}_$name.log 2>&1
done
unset ILQG_GROUPS
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_${R}.csv python bench.py --steps 2 --warmup 1 > gpurun_out/launches_${R}_bench.log 2>&1
ls -la gpurun_out | tail -8
