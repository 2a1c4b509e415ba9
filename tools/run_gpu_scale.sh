# weak scaling on one box: bench.py under torchrun at N = 2, 4, 8 (as many as the box has), metric config and C5
N=$(nvidia-smi -L | wc -l)
for n in 2 4 8; do
  [ $n -le $N ] || continue
  for c in metric c5; do
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --config $c --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_scale_${c}_$n.json 2> gpurun_out/bench_scale_${c}_$n.err
    python -c "
import json; d=json.loads(open('gpurun_out/bench_scale_${c}_$n.json').readline()); print('$c', $n, round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['value']), d.get('gather_ms'))"
  done
done
