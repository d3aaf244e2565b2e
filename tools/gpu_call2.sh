#!/bin/bash
# One B200: GPU tests (incl. the staged norm kernels), A/B bandwidth of the norm kernels, the default bench line (with the
# bounded CPU arm), the reference arm, and the per-kernel table of a pi0 step.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c2_pytest.log
tail -3 gpurun_out/c2_pytest.log
timeout 200 python tools/bench_elementwise.py > gpurun_out/c2_hbm_events.log 2>&1; head -8 gpurun_out/c2_hbm_events.log
( time timeout 420 python bench.py ) > gpurun_out/c2_bench_cogact.log 2> gpurun_out/c2_bench_cogact.err; tail -c 900 gpurun_out/c2_bench_cogact.log; tail -4 gpurun_out/c2_bench_cogact.err
( time timeout 400 python bench.py --impl reference ) > gpurun_out/c2_bench_ref.log 2> gpurun_out/c2_bench_ref.err; tail -c 700 gpurun_out/c2_bench_ref.log; tail -4 gpurun_out/c2_bench_ref.err
timeout 200 python tools/profile_step.py pi0_2b > gpurun_out/c2_profile_pi0.log 2>&1
nproc > gpurun_out/c2_host.txt; lscpu | head -20 >> gpurun_out/c2_host.txt; free -g >> gpurun_out/c2_host.txt
