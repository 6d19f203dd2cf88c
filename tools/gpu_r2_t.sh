#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_pytest.log
timeout 300 python tools/micro_frames.py > gpurun_out/t_micro_frames.log 2>&1; grep "^{" gpurun_out/t_micro_frames.log | cut -c1-220
timeout 900 python bench.py > gpurun_out/t_bench.json 2> gpurun_out/t_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/t_bench.err
timeout 200 python tools/e2e_profile.py C2 > gpurun_out/t_e2e_profile.log 2>&1; tail -14 gpurun_out/t_e2e_profile.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/t_bench.json").read().strip().splitlines()[-1])
print("C2 value %.3fM us/step %.1f e2e %.3fM (%.1f us) blocking %.1f us roofline frac %.4f; c5 %.1f us wall %.1f" % (d["value"]/1e6, d["ms_per_step"]*1e3, d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"]*1e3, d["e2e"]["blocking"]["ms_per_step"]*1e3, d["roofline"]["frac"], d["c5"]["ms_per_step"]*1e3, d["c5"]["wall_ms_per_step"]*1e3))
for k, v in (d.get("configs") or {}).items():
    if "replay" in v:
        print(k, "value %.3fM" % (v["value"]/1e6), "us/step %.1f" % (v["ms_per_step"]*1e3), "e2e %.3fM" % (v["e2e"]["value"]/1e6), "gather", [round(g["frac"], 3) for g in v["replay"]["gather"]], "store", [round(g["frac"], 3) for g in v["replay"]["store"]])
    else:
        print(k, [(g["batch"], round(g["kernel"]["frac"], 3), round(g["host_api"]["frac"], 3), round(g["kernel"]["us"], 1), round(g["host_api"]["us"], 1)) for g in v["gather"]], v.get("store"))
print("c4_sharded", d.get("c4_sharded", {}).get("batches"))
PY
