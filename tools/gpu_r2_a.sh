#!/bin/bash
# round-2 GPU call A: new tcgen05 tile width + fused forward kernel: parity first, then stage times
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_gpu.txt 2>&1
timeout 300 python -m pytest tests/test_tc_gemm_gpu.py -x -q > gpurun_out/a_gemm.log 2>&1; echo "gemm rc=$?" >> gpurun_out/a_gemm.log
tail -5 gpurun_out/a_gemm.log
timeout 600 python -m pytest tests/test_sac_gpu.py -x -q > gpurun_out/a_sac.log 2>&1; echo "sac rc=$?" >> gpurun_out/a_sac.log
tail -15 gpurun_out/a_sac.log
DDRL_FUSE_L1=0 timeout 300 python -m pytest tests/test_sac_gpu.py -x -q -k "one_step" > gpurun_out/a_sac_unfused.log 2>&1; echo "sac unfused rc=$?" >> gpurun_out/a_sac_unfused.log
tail -5 gpurun_out/a_sac_unfused.log
for c in C2 C1; do timeout 120 python tools/stage_times.py $c > gpurun_out/a_stage_$c.log 2>&1; cat gpurun_out/a_stage_$c.log; done
DDRL_FUSE_L1=0 timeout 120 python tools/stage_times.py C2 > gpurun_out/a_stage_C2_unfused.log 2>&1; cat gpurun_out/a_stage_C2_unfused.log
DDRL_FUSE_L1=0 DDRL_TC_BN=128 timeout 120 python tools/stage_times.py C2 > gpurun_out/a_stage_C2_unfused128.log 2>&1; cat gpurun_out/a_stage_C2_unfused128.log
timeout 200 python tools/micro_sac.py > gpurun_out/a_micro_sac.log 2>&1; cat gpurun_out/a_micro_sac.log
