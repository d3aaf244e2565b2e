#!/bin/bash
# Round-2 evidence run on ONE B200 (gpurun): GPU tests, the default bench line, ncu rows for every HBM-bound kernel class,
# ncu --set full of the flash-attention kernels and of every decoder GEMM mode, the per-kernel step table, and the ncu
# launch list of the device-timed bench steps.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/c1_smi.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
tail -4 gpurun_out/c1_pytest.log
timeout 400 python bench.py > gpurun_out/c1_bench_cogact.log 2> gpurun_out/c1_bench_cogact.err; tail -c 600 gpurun_out/c1_bench_cogact.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file gpurun_out/r2_hbm_kernels.csv python tools/bench_elementwise.py --once > gpurun_out/c1_hbm_once.log 2>&1
timeout 200 python tools/bench_elementwise.py > gpurun_out/c1_hbm_events.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:flash_|attn_delta' -o gpurun_out/r2_flash_attn python tools/ncu_flash.py cogact pi0 > gpurun_out/c1_flash_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 6 -c 6 -o gpurun_out/r2_gemm_modes python tools/ncu_gemm.py > gpurun_out/c1_gemm_ncu.log 2>&1
timeout 300 python tools/profile_step.py cogact_7b > gpurun_out/c1_profile_step.log 2>&1
B200_PROFILER_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_step.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/c1_launch_bench.log 2>&1
timeout 300 python bench.py --workload pi0_2b --no-cpu-baseline > gpurun_out/c1_bench_pi0_2b.log 2>&1
timeout 200 python tools/bench_gemm.py > gpurun_out/c1_bench_gemm.log 2>&1
ls -la gpurun_out | head -40
