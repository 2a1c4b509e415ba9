export ILQG_GROUPS=1 ILQG_PIPELINE=0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ls_eval -s 11 -c 1 -f -o gpurun_out/r01d_k_ls_eval_queued python tools/profile_target.py 4096 5 > gpurun_out/ncu_r01d.log 2>&1
tail -3 gpurun_out/ncu_r01d.log
