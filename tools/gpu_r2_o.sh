#!/bin/bash
mkdir -p gpurun_out
DDRL_DP_TRACE=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/dp_trace.py > gpurun_out/o_dp_trace.log 2>&1; grep "rank" gpurun_out/o_dp_trace.log || tail -n 20 gpurun_out/o_dp_trace.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1000 --warmup 50 > gpurun_out/o_bench2.json 2> gpurun_out/o_bench2.err
python - <<'PY'
import json
for f in ("o_bench2",):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["n_gpus"], "value %.3fM" % (d["value"]/1e6), "us/step %.1f" % (d["ms_per_step"]*1e3), "e2e %.3fM" % (d["e2e"]["value"]/1e6), "c5 us %.1f" % (d["c5"]["ms_per_step"]*1e3))
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
