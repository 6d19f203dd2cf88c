#!/bin/bash
# final single-GPU verification of the round: whole GPU suite, default bench (what the driver runs), reference arm, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --only-primary > gpurun_out/r02_bench_under_ncu.log 2>&1
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench.json").read().strip().splitlines()[-1])
print("C2 value %.3fM us/step %.1f e2e %.3fM (%.1f us) roofline frac %.4f us %.2f launches %d" % (d["value"]/1e6, d["ms_per_step"]*1e3, d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"]*1e3, d["roofline"]["frac"], d["roofline"]["us_per_launch"], d["gpu_launches"]))
print({k: round(v["ms_per_step"]*1e3, 1) for k, v in d["e2e"].items() if isinstance(v, dict)}, "c5 %.1f" % (d["c5"]["ms_per_step"]*1e3))
for k, v in (d.get("configs") or {}).items():
    if "replay" in v:
        print(k, "value %.3fM" % (v["value"]/1e6), "us/step %.1f" % (v["ms_per_step"]*1e3), "e2e %.3fM" % (v["e2e"]["value"]/1e6), "roofline %.4f" % v["roofline"]["frac"], "gather", [round(g["frac"], 3) for g in v["replay"]["gather"]], "store", [round(g["frac"], 3) for g in v["replay"]["store"]])
    elif "gather" in v:
        print(k, [(g["batch"], round(g["kernel"]["frac"], 3), round(g["host_api"]["frac"], 3), round(g["kernel"]["us"], 1), round(g["host_api"]["us"], 1)) for g in v["gather"]], v.get("store"))
    else:
        print(k, {kk: vv for kk, vv in v.items() if kk != "workload"})
print("c4_sharded", {k: (round(v["value"]/1e6, 2), round(v["frac"], 3)) for k, v in d["c4_sharded"]["batches"].items()})
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["threads1"]["value"], d["cpu_baseline"]["replay_only"]["sample_transitions_per_s"])
PY
