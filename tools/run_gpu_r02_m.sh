# data for the round-2 documents: one ncu --set full capture per hot kernel, the launch list of the bench command, every BASELINE config
R=r02
export ILQG_GROUPS=1 ILQG_PIPELINE=0
for k in k_linearize_quadraticize_v4:2:linearize_quadraticize k_lq_backward_tc:2:lq_backward k_ls_rollout_sp:4:ls_rollout k_ls_rollout_sp:5:ls_rollout_tier1 k_ls_merit:4:ls_merit k_ls_merit:5:ls_merit_tier1 k_ls_decide:2:ls_decide; do
  IFS=: read name skip tag <<< "$k"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$name -s $skip -c 1 -f -o gpurun_out/${R}_$tag python tools/profile_target.py 4096 5 > gpurun_out/ncu_${R}_$tag.log 2>&1
done
unset ILQG_GROUPS ILQG_PIPELINE
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_${R}.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/launches_${R}_bench.log 2>&1
python bench.py > gpurun_out/bench_${R}_metric.json 2> gpurun_out/bench_${R}_metric.err
cut -c1-300 gpurun_out/bench_${R}_metric.json
for c in c2 c3 c4 c5; do
  python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/bench_${R}_$c.json 2> gpurun_out/bench_${R}_$c.err
  cut -c1-200 gpurun_out/bench_${R}_$c.json
done
python bench.py --with-al --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_${R}_al.json 2> gpurun_out/bench_${R}_al.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_${R}_al.json').readline()); print(d.get('al_solve'))"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${R}_reference.json 2> gpurun_out/bench_${R}_reference.err
cut -c1-400 gpurun_out/bench_${R}_reference.json
