# first thing to measure in the next round: the dependent chain of
# K_bwd on FFMA vs mma.sync TF32 (x3 split / x1) at 4..64 warps per SM
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_chain_bench tools/mma_chain_bench.cu &&
  timeout 120 /tmp/mma_chain_bench 99 20 | tee gpurun_out/mma_chain_bench.txt
