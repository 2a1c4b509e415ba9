export ILQG_GROUPS=1
for k in k_ls_eval:1:fresh k_ls_eval:2:queued; do
  IFS=: read name skip tag <<< "$k"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$name -s $skip -c 1 -f -o gpurun_out/r01b_${name}_$tag python tools/profile_target.py 4096 3 > gpurun_out/ncu_r01b_${name}_$tag.log 2>&1
done
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
