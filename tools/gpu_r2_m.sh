#!/bin/bash
mkdir -p gpurun_out
DDRL_DP_TRACE=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/dp_trace.py > gpurun_out/m_dp_trace.log 2>&1; grep "k_adam_dp" gpurun_out/m_dp_trace.log || tail -n 20 gpurun_out/m_dp_trace.log
