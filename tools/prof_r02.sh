#!/bin/bash
# Round-2 measurement set (GPU box, via gpurun): tests, default bench, launch list, ncu --set full of the dominant kernels.
# Outputs under gpurun_out/; tools/ncu_traffic.py + tools/ncu_summary.py turn the captures into profiles/ files.
R=${1:-r02}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${R}_pytest_gpu.log
tail -3 gpurun_out/${R}_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${R}_bench_reference.json 2> gpurun_out/${R}_bench_reference.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${R}_bench_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --only-primary > gpurun_out/${R}_bench_under_ncu.log 2>&1
cap() { # name, kernel regex, skip, script...
  n=$1; k=$2; s=$3; shift 3
  DDRL_NO_GRAPH=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -f -o gpurun_out/${R}_$n "$@" > gpurun_out/${R}_$n.log 2>&1
}
cap fwd_c2 fwd_fused_tc 4 python tools/prof_stage.py C2 1 6
cap fwd_c1 fwd_fused_tc 4 python tools/prof_stage.py C1 1 6
cap fwd_c3 gemm_grouped_tc 9 python tools/prof_stage.py C3 1 6
cap l1_c3 gemm_grouped_tc 9 python tools/prof_stage.py C3 0 6
cap bq_c2 gemm_grouped_tc 4 python tools/prof_stage.py C2 4 6
cap gather_c1 rb_gather 2 python tools/prof_replay.py C1
cap gather_c2 rb_gather 2 python tools/prof_replay.py C2
cap gather_c3 rb_gather 2 python tools/prof_replay.py C3 64
cap store_c2 rb_store_staged 2 python tools/prof_replay.py C2
cap gather_c4_dedup rb_gather_tma 6 python tools/micro_frames.py
cap store_c4_frames fb_store_frames 2 python tools/micro_frames.py
cap store_n3 seg_store_rows 14 python tools/micro_frames.py
DDRL_GATHER_MODE=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:rb_gather_wide -s 2 -c 1 -f -o gpurun_out/${R}_gather_c3_regs python tools/prof_replay.py C3 64 > gpurun_out/${R}_gather_c3_regs.log 2>&1
timeout 300 python tools/parity_margins.py > gpurun_out/${R}_parity_margins.log 2>&1
for c in C1 C2 C3; do timeout 200 python tools/replay_curve.py $c default > gpurun_out/${R}_curve_$c.log 2>&1; done
timeout 200 python tools/stage_times.py C2 > gpurun_out/${R}_stage_times_C2.log 2>&1
timeout 200 python tools/stage_times.py C3 > gpurun_out/${R}_stage_times_C3.log 2>&1
timeout 200 python tools/tc_trace.py C2 1 > gpurun_out/${R}_tc_trace_fused_c2.log 2>&1
timeout 200 python tools/tc_trace.py C2 4 > gpurun_out/${R}_tc_trace_bq_c2.log 2>&1
./tools/probes/tick_probe > gpurun_out/${R}_tick_probe.log 2>&1
ls -la gpurun_out/${R}_*.ncu-rep
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench.json").read().strip().splitlines()[-1])
print("C2 value %.3fM us/step %.1f e2e %.3fM (%.1f us) roofline frac %.4f us %.2f" % (d["value"]/1e6, d["ms_per_step"]*1e3, d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"]*1e3, d["roofline"]["frac"], d["roofline"]["us_per_launch"]))
for k, v in (d.get("configs") or {}).items():
    print(k, "value %.3fM" % (v["value"]/1e6), "us/step %s" % (v.get("ms_per_step") and round(v["ms_per_step"]*1e3, 1)), "e2e", v.get("e2e", {}).get("value"), "gather", [round(g["frac"], 3) for g in v.get("replay", {}).get("gather", v.get("gather", []))], "store", [round(g["frac"], 3) for g in v.get("replay", {}).get("store", [])] or v.get("store"))
print("cpu", {k: (v if not isinstance(v, dict) else v.get("value", v.get("sample_transitions_per_s"))) for k, v in d["cpu_baseline"].items() if k in ("value", "threads1", "replay_only", "cores", "kind")})
PY
