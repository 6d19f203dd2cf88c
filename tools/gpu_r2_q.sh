#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_compat_gpu.py tests/test_workers_gpu.py tests/test_replay_gpu.py -x -q -s > gpurun_out/q_compat.log 2>&1; echo "rc=$?" >> gpurun_out/q_compat.log; grep -E "sha256|passed|failed|rc=|Error" gpurun_out/q_compat.log | cut -c1-400
for c in C1 C2 C3; do timeout 200 python tools/replay_curve.py $c default > gpurun_out/q_curve_$c.log 2>&1; grep -E "^\{" gpurun_out/q_curve_$c.log | cut -c1-200; done
for al in 32 96 128; do DDRL_ROW_ALIGN=$al timeout 200 python tools/replay_curve.py C1 align$al > gpurun_out/q_curve_C1_a$al.log 2>&1; grep -E "sample" gpurun_out/q_curve_C1_a$al.log | cut -c1-200; done
DDRL_GATHER_MODE=2 timeout 200 python tools/replay_curve.py C1 regs > gpurun_out/q_curve_C1_regs.log 2>&1; grep -E "sample" gpurun_out/q_curve_C1_regs.log | cut -c1-200
DDRL_GATHER_MODE=1 timeout 200 python tools/replay_curve.py C1 bulk > gpurun_out/q_curve_C1_bulk.log 2>&1; grep -E "sample" gpurun_out/q_curve_C1_bulk.log | cut -c1-200
DDRL_GATHER_MODE=1 timeout 200 python tools/replay_curve.py C3 bulk > gpurun_out/q_curve_C3_bulk.log 2>&1; grep -E "sample" gpurun_out/q_curve_C3_bulk.log | cut -c1-200
