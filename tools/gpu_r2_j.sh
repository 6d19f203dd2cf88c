#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sac_gpu.py -x -q > gpurun_out/j_sac.log 2>&1; echo "sac rc=$?" >> gpurun_out/j_sac.log; tail -n 3 gpurun_out/j_sac.log
timeout 100 python tools/tc_trace.py C2 1 > gpurun_out/j_trace_1.log 2>&1; cat gpurun_out/j_trace_1.log
timeout 120 python tools/stage_times.py C2 > gpurun_out/j_stage_C2.log 2>&1; cat gpurun_out/j_stage_C2.log
timeout 200 python tools/micro_sac.py > gpurun_out/j_micro_sac.log 2>&1; cat gpurun_out/j_micro_sac.log
