#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sac_gpu.py -x -q > gpurun_out/e_sac.log 2>&1; echo "sac rc=$?" >> gpurun_out/e_sac.log; tail -4 gpurun_out/e_sac.log
timeout 100 python tools/tc_trace.py C2 1 > gpurun_out/e_trace_1.log 2>&1; cat gpurun_out/e_trace_1.log
timeout 100 python tools/tc_trace.py C2 3 > gpurun_out/e_trace_3.log 2>&1; cat gpurun_out/e_trace_3.log
timeout 120 python tools/stage_times.py C2 > gpurun_out/e_stage_C2.log 2>&1; cat gpurun_out/e_stage_C2.log
