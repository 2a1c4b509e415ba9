ILQG_TRACE=gpurun_out/trace_c1.txt python tools/profile_target.py 4096 8 c1 > /dev/null
python tools/trace_view.py gpurun_out/trace_c1.txt 3 3
ILQG_TRACE=gpurun_out/trace_c1_p0.txt ILQG_PIPELINE=0 python tools/profile_target.py 4096 8 c1 > /dev/null
python tools/trace_view.py gpurun_out/trace_c1_p0.txt 3 2
