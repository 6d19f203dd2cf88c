#!/bin/bash
mkdir -p gpurun_out
run() { # name, N, env...
  name=$1; N=$2; shift 2
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 1000 --warmup 50 --only-primary > gpurun_out/z_${N}_$name.json 2> gpurun_out/z_${N}_$name.err
}
run v1 8 DDRL_DP_V1=1
run one 8 DDRL_DP_V1=0
run v1 4 DDRL_DP_V1=1
run one 4 DDRL_DP_V1=0
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/z_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], "value %.3fM" % (d["value"]/1e6), "us/step %.1f" % (d["ms_per_step"]*1e3), "e2e %.1f us" % (d["e2e"]["ms_per_step"]*1e3), "c5 us %.1f" % (d["c5"]["ms_per_step"]*1e3))
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-600:])
PY
