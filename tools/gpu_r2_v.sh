#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc_gemm_gpu.py tests/test_sac_gpu.py tests/test_workers_gpu.py -x -q > gpurun_out/v_sac.log 2>&1; echo "sac rc=$?"; tail -4 gpurun_out/v_sac.log
timeout 200 python tools/stage_times.py C2 > gpurun_out/v_stage_C2.log 2>&1; cat gpurun_out/v_stage_C2.log | tail -16
timeout 200 python tools/stage_times.py C3 > gpurun_out/v_stage_C3.log 2>&1; cat gpurun_out/v_stage_C3.log | tail -16
timeout 200 python tools/micro_sac.py > gpurun_out/v_micro_sac.log 2>&1; tail -5 gpurun_out/v_micro_sac.log
timeout 200 python tools/dbg_fullsize.py > gpurun_out/v_fullsize.log 2>&1; tail -12 gpurun_out/v_fullsize.log
