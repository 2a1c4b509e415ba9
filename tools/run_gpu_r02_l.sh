export ILQG_GROUPS=1 ILQG_PIPELINE=0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_linearize_quadraticize_v4 -s 2 -c 1 -f -o gpurun_out/r02f_linearize_quadraticize_v4 python tools/profile_target.py 4096 4 > gpurun_out/ncu_r02f_klq.log 2>&1
tail -2 gpurun_out/ncu_r02f_klq.log
