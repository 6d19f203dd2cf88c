#!/bin/bash
# 8 GPUs: data-parallel modes compared (short runs)
mkdir -p gpurun_out
run() { # name, nproc, env...
  name=$1; n=$2; shift 2
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $n --steps 600 --warmup 50 > gpurun_out/p_$name.json 2> gpurun_out/p_$name.err
}
run n8_single 8 DDRL_DP_OVERLAP=0
run n8_v1 8 DDRL_DP_V1=1
run n8_overlap 8 DDRL_DP_OVERLAP=1
run n4_single 4 DDRL_DP_OVERLAP=0
run n4_v1 4 DDRL_DP_V1=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/p_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], "value %.3fM" % (d["value"]/1e6), "us/step %.1f" % (d["ms_per_step"]*1e3), "e2e %.3fM" % (d["e2e"]["value"]/1e6), "c5 us %.1f" % (d["c5"]["ms_per_step"]*1e3))
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-800:])
PY
