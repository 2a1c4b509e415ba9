# two GPUs: the multi-device handle test, the sharded bench (torchrun), C5 at 8192 games per GPU
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_device or any_shape" 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r02_metric_2gpu.json 2> gpurun_out/bench_r02_metric_2gpu.err
cut -c1-260 gpurun_out/bench_r02_metric_2gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config c5 --steps 10 --warmup 3 > gpurun_out/bench_r02_c5_2gpu.json 2> gpurun_out/bench_r02_c5_2gpu.err
cut -c1-260 gpurun_out/bench_r02_c5_2gpu.json
