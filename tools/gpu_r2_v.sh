#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc_gemm_gpu.py tests/test_sac_gpu.py tests/test_workers_gpu.py -x -q > gpurun_out/v_sac.log 2>&1; echo "sac rc=$?"; tail -4 gpurun_out/v_sac.log
timeout 200 python tools/stage_times.py C2 > gpurun_out/v_stage_C2.log 2>&1; cat gpurun_out/v_stage_C2.log | tail -16
timeout 200 python tools/stage_times.py C3 > gpurun_out/v_stage_C3.log 2>&1; cat gpurun_out/v_stage_C3.log | tail -16
timeout 200 python tools/micro_sac.py > gpurun_out/v_micro_sac.log 2>&1; tail -5 gpurun_out/v_micro_sac.log
timeout 300 python bench.py --only-primary --no-cpu-baseline > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err; python -c "import json; d=json.loads(open(\"gpurun_out/v_bench.json\").read().strip().splitlines()[-1]); print(\"C2 step %.1f us e2e %.1f us roofline %.4f (%.2f us graph, %.2f us stream)\" % (d[\"ms_per_step\"]*1e3, d[\"e2e\"][\"ms_per_step\"]*1e3, d[\"roofline\"][\"frac\"], d[\"roofline\"][\"us_per_launch\"], d[\"roofline\"][\"us_per_stream_launch\"]), {k: round(v[\"ms_per_step\"]*1e3,1) for k,v in d[\"e2e\"].items() if isinstance(v, dict)})"
