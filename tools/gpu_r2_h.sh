#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sac_gpu.py -x -q > gpurun_out/h_sac.log 2>&1; echo "sac rc=$?" >> gpurun_out/h_sac.log; tail -n 12 gpurun_out/h_sac.log
timeout 120 python tools/stage_times.py C2 > gpurun_out/h_stage_C2.log 2>&1; cat gpurun_out/h_stage_C2.log
timeout 200 python tools/micro_sac.py > gpurun_out/h_micro_sac.log 2>&1; cat gpurun_out/h_micro_sac.log
DDRL_PDL=1 timeout 200 python tools/micro_sac.py > gpurun_out/h_micro_sac_pdl.log 2>&1; cat gpurun_out/h_micro_sac_pdl.log
