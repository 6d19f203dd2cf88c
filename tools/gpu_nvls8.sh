#!/bin/bash
mkdir -p gpurun_out
for nv in 1 0; do DDRL_DP_NVLS=$nv timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 tools/dp_replica_check.py 200 2>&1 | grep -E "differ|Error" ; done
bash tools/gpu_nvls.sh 8
