#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tc_gemm_gpu.py -x -q > gpurun_out/k_gemm.log 2>&1; echo "gemm rc=$?" >> gpurun_out/k_gemm.log; tail -n 2 gpurun_out/k_gemm.log
timeout 600 python -m pytest tests/test_sac_gpu.py -x -q > gpurun_out/k_sac.log 2>&1; echo "sac rc=$?" >> gpurun_out/k_sac.log; tail -n 3 gpurun_out/k_sac.log
timeout 100 python tools/tc_trace.py C2 1 > gpurun_out/k_trace_1.log 2>&1; cat gpurun_out/k_trace_1.log
timeout 120 python tools/stage_times.py C2 > gpurun_out/k_stage_C2.log 2>&1; cat gpurun_out/k_stage_C2.log
timeout 200 python tools/micro_sac.py > gpurun_out/k_micro_sac.log 2>&1; cat gpurun_out/k_micro_sac.log
DDRL_PDL=1 timeout 200 python tools/micro_sac.py > gpurun_out/k_micro_sac_pdl.log 2>&1; cat gpurun_out/k_micro_sac_pdl.log
