#!/bin/bash
# round-2 GPU call B: is a tcgen05 stage bound by L2 -> SM operand traffic?  warm-L2 ncu captures of BQ (64 / 128 tiles) and the fused stage
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --cache-control none --import-source on"
DDRL_TC_BN=64 $NCU -k regex:gemm_grouped_tc -s 3 -c 1 -o gpurun_out/b_bq64 python tools/prof_stage.py C2 4 6 > gpurun_out/b1.log 2>&1
DDRL_TC_BN=128 $NCU -k regex:gemm_grouped_tc -s 3 -c 1 -o gpurun_out/b_bq128 python tools/prof_stage.py C2 4 6 > gpurun_out/b2.log 2>&1
$NCU -k regex:fwd_fused_tc -s 3 -c 1 -o gpurun_out/b_fused python tools/prof_stage.py C2 1 6 > gpurun_out/b3.log 2>&1
$NCU -k regex:gemm_grouped_tc -s 3 -c 1 -o gpurun_out/b_bp3 python tools/prof_stage.py C2 6 6 > gpurun_out/b4.log 2>&1
$NCU -k regex:k_adam -s 3 -c 1 -o gpurun_out/b_adam python tools/prof_stage.py C2 11 6 > gpurun_out/b5.log 2>&1
$NCU -k regex:k_qheads -s 3 -c 1 -o gpurun_out/b_qheads python tools/prof_stage.py C2 9 6 > gpurun_out/b6.log 2>&1
tail -2 gpurun_out/b*.log
ls -la gpurun_out/*.ncu-rep
