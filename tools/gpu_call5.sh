#!/bin/bash
# TWO B200s, one box: data-parallel equivalence test, then N=2 with the peer-memory transport (unrolled reduce-scatter
# kernel) and with the NCCL transport.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dp.py -q --timeout 280 > gpurun_out/c5_pytest_dp.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c5_pytest_dp.log
tail -3 gpurun_out/c5_pytest_dp.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29551 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/c5_n2_ce.log 2> gpurun_out/c5_n2_ce.err
B200_DP_TRANSPORT=nccl timeout 300 $TR --master-port 29552 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/c5_n2_nccl.log 2> gpurun_out/c5_n2_nccl.err
for f in c5_n2_ce c5_n2_nccl; do grep '^{' gpurun_out/$f.log | cut -c1-200; done
