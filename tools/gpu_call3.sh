#!/bin/bash
# TWO B200s (gpurun --gpus 2): the whole GPU suite (the data-parallel equivalence test needs 2 devices: all-reduce, ZeRO-1
# over NCCL, ZeRO-1 over peer memory) and the N=2 bench line with the default transport.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c3_smi.txt
timeout 700 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/c3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c3_pytest.log
tail -4 gpurun_out/c3_pytest.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/c3_bench_n2.log 2> gpurun_out/c3_bench_n2.err
tail -c 1200 gpurun_out/c3_bench_n2.log; tail -5 gpurun_out/c3_bench_n2.err
