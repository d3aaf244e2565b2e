#!/bin/bash
# TWO B200s, one box: N=1 baseline, then N=2 with the peer-memory transport (default) and with the NCCL transport.
mkdir -p gpurun_out
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c4_n1.log 2> gpurun_out/c4_n1.err
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/c4_n2_ce.log 2> gpurun_out/c4_n2_ce.err
B200_DP_TRANSPORT=nccl timeout 300 $TR --master-port 29542 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/c4_n2_nccl.log 2> gpurun_out/c4_n2_nccl.err
for f in c4_n1 c4_n2_ce c4_n2_nccl; do grep '^{' gpurun_out/$f.log | cut -c1-200; done
