# Round profile set (run on the GPU box via gpurun); outputs under gpurun_out/
set -x
R=${1:-r01}
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${R}_bench_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_bench_under_ncu.log 2>&1
# tcgen05 GEMM: the L2 stage (5 passes x 1024 x 256 x 256) and the BQ stage (dgrad + wgrad tiles) of a C2 step
DDRL_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:gemm_grouped_tc -s 15 -c 1 -o gpurun_out/${R}_gemm_tc_c2_l2 python tools/prof_sac.py C2 4 > gpurun_out/p3.log 2>&1
DDRL_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:gemm_grouped_tc -s 18 -c 1 -o gpurun_out/${R}_gemm_tc_c2_bq python tools/prof_sac.py C2 4 > gpurun_out/p4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rb_gather_bulk -s 2 -c 1 -o gpurun_out/${R}_gather_bulk_c2 python tools/prof_replay.py C2 > gpurun_out/p1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rb_gather_wide -s 2 -c 1 -o gpurun_out/${R}_gather_wide_c3 python tools/prof_replay.py C3 > gpurun_out/p2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rb_store_rows -s 2 -c 1 -o gpurun_out/${R}_store_c2 python tools/prof_replay.py C2 > gpurun_out/p5.log 2>&1
ls -la gpurun_out/*.ncu-rep
