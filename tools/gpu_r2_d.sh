#!/bin/bash
# round-2 GPU call D: 4 rotating TMEM accumulators: parity, traces, stage times, step time
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tc_gemm_gpu.py -x -q > gpurun_out/d_gemm.log 2>&1; echo "gemm rc=$?" >> gpurun_out/d_gemm.log; tail -3 gpurun_out/d_gemm.log
timeout 600 python -m pytest tests/test_sac_gpu.py -x -q > gpurun_out/d_sac.log 2>&1; echo "sac rc=$?" >> gpurun_out/d_sac.log; tail -4 gpurun_out/d_sac.log
for st in 1 4 6; do
  timeout 100 python tools/tc_trace.py C2 $st > gpurun_out/d_trace_$st.log 2>&1; cat gpurun_out/d_trace_$st.log
done
DDRL_TC_BN=128 timeout 100 python tools/tc_trace.py C2 4 > gpurun_out/d_trace128_4.log 2>&1; cat gpurun_out/d_trace128_4.log
timeout 120 python tools/stage_times.py C2 > gpurun_out/d_stage_C2.log 2>&1; cat gpurun_out/d_stage_C2.log
DDRL_TC_BN=128 timeout 120 python tools/stage_times.py C2 > gpurun_out/d_stage_C2_128.log 2>&1; cat gpurun_out/d_stage_C2_128.log
timeout 200 python tools/micro_sac.py > gpurun_out/d_micro_sac.log 2>&1; cat gpurun_out/d_micro_sac.log
