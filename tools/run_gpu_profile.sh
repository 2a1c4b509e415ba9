# round profile pass: one full ncu capture of each hot kernel + launch list of the bench command
R=${1:-r01}
export ILQG_GROUPS=1 ILQG_PIPELINE=0
for k in k_ls_eval:4:ls_eval k_ls_eval:11:ls_eval_queued k_lq_backward_hw:1:lq_backward k_linearize_quadraticize_v3:1:linearize_quadraticize; do
  IFS=: read name skip tag <<< "$k"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$name -s $skip -c 1 -f -o gpurun_out/${R}_$tag python tools/profile_target.py 4096 5 > gpurun_out/ncu_${R}_$tag.log 2>&1
done
unset ILQG_GROUPS ILQG_PIPELINE
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_${R}.csv python bench.py --steps 2 --warmup 1 > gpurun_out/launches_${R}_bench.log 2>&1
python bench.py > gpurun_out/bench_${R}.json 2> gpurun_out/bench_${R}.err
cut -c1-400 gpurun_out/bench_${R}.json
