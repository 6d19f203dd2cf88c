#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_replay_gpu.py tests/test_capi.py -x -q > gpurun_out/w_replay.log 2>&1; echo "replay rc=$?"; tail -2 gpurun_out/w_replay.log
timeout 200 python tools/replay_curve.py C1 default > gpurun_out/w_curve_C1.log 2>&1; grep sample gpurun_out/w_curve_C1.log | cut -c1-200
DDRL_GATHER_FLAT=0 timeout 200 python tools/replay_curve.py C1 pow2 > gpurun_out/w_curve_C1_pow2.log 2>&1; grep sample gpurun_out/w_curve_C1_pow2.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rb_gather_flat -s 2 -c 1 -f -o gpurun_out/r02_gather_c1 python tools/prof_replay.py C1 > gpurun_out/w_ncu.log 2>&1; ls -la gpurun_out/r02_gather_c1.ncu-rep
