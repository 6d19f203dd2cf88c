#!/bin/bash
mkdir -p gpurun_out
for st in 4 5 6; do
  DDRL_TC_BN=64 timeout 100 python tools/tc_trace.py C2 $st > gpurun_out/c_trace64_$st.log 2>&1; cat gpurun_out/c_trace64_$st.log
  DDRL_TC_BN=128 timeout 100 python tools/tc_trace.py C2 $st > gpurun_out/c_trace128_$st.log 2>&1; cat gpurun_out/c_trace128_$st.log
done
