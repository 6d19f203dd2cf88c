#!/bin/bash
# N GPUs: the NVLS (multimem.ld_reduce) form of the fused data-parallel exchange against the peer-read form
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then timeout 300 python -m pytest "tests/test_dist_gpu.py::test_two_gpu_step_equals_single_gpu_on_concatenated_batch" -q -s 2>&1 | grep -E "NVLS|passed|failed|Error" | head; fi
run() { name=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $N --steps 1000 --warmup 50 --only-primary > gpurun_out/nv_${N}_$name.json 2> gpurun_out/nv_${N}_$name.err; }
run nvls DDRL_DP_NVLS=1
run peer DDRL_DP_NVLS=0
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/nv_${N}_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], "value %.3fM" % (d["value"]/1e6), "us/step %.1f" % (d["ms_per_step"]*1e3), d["config"]["parallelism"][:110])
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
