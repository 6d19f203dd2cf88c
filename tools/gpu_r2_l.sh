#!/bin/bash
# 2 GPUs: data-parallel parity + short scaling runs
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dist_gpu.py -x -q > gpurun_out/l_dist.log 2>&1; echo "dist rc=$?" >> gpurun_out/l_dist.log; tail -n 6 gpurun_out/l_dist.log
timeout 200 python bench.py --gpus 1 --steps 1000 --warmup 50 --no-cpu-baseline > gpurun_out/l_bench1.json 2> gpurun_out/l_bench1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1000 --warmup 50 > gpurun_out/l_bench2.json 2> gpurun_out/l_bench2.err
DDRL_DP_V1=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 1000 --warmup 50 > gpurun_out/l_bench2_v1.json 2> gpurun_out/l_bench2_v1.err
python - <<'PY'
import json
for f in ("l_bench1","l_bench2","l_bench2_v1"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["n_gpus"], "value %.3fM" % (d["value"]/1e6), "us/step %.1f" % (d["ms_per_step"]*1e3), "e2e %.3fM" % (d["e2e"]["value"]/1e6), "c5 us %.1f" % (d["c5"]["ms_per_step"]*1e3), d["config"]["parallelism"][:40])
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
