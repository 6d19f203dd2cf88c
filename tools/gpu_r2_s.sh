#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do timeout 300 python -m pytest tests/test_compat_gpu.py -x -q > gpurun_out/s_compat_$i.log 2>&1; echo "compat$i rc=$?"; done
timeout 900 python -m pytest tests/test_replay_gpu.py tests/test_frames_gpu.py tests/test_nstep_gpu.py tests/test_capi.py tests/test_workers_gpu.py -x -q > gpurun_out/s_replay.log 2>&1; echo "replay rc=$?"; tail -5 gpurun_out/s_replay.log
timeout 300 python tools/micro_frames.py > gpurun_out/s_micro_frames.log 2>&1; grep "^{" gpurun_out/s_micro_frames.log | cut -c1-220
for c in C3; do timeout 200 python tools/replay_curve.py $c default > gpurun_out/s_curve_$c.log 2>&1; grep -E "^\{" gpurun_out/s_curve_$c.log | grep sample | cut -c1-200; done
DDRL_GATHER_MODE=3 timeout 200 python tools/replay_curve.py C3 tma > gpurun_out/s_curve_C3_tma.log 2>&1; grep -E "sample" gpurun_out/s_curve_C3_tma.log | cut -c1-200
DDRL_GATHER_MODE=3 timeout 200 python tools/replay_curve.py C2 tma > gpurun_out/s_curve_C2_tma.log 2>&1; grep -E "sample" gpurun_out/s_curve_C2_tma.log | cut -c1-200
DDRL_GATHER_MODE=3 DDRL_NO_GRAPH=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:rb_gather_tma -s 2 -c 1 -f -o gpurun_out/r02_gather_c3_tma python tools/prof_replay.py C3 64 > gpurun_out/s_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rb_gather_tma -s 6 -c 1 -f -o gpurun_out/r02_gather_c4_dedup_tma python tools/micro_frames.py > gpurun_out/s_ncu2.log 2>&1
ls -la gpurun_out/r02_gather_c3_tma.ncu-rep gpurun_out/r02_gather_c4_dedup_tma.ncu-rep
