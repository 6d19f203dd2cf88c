#!/bin/bash
# 2 GPUs: the multi-GPU parity tests (the driver's box has one GPU and skips them) + bench at N = 2
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_dist_tests_2gpu.log
timeout 600 python -m pytest tests/test_dist_gpu.py tests/test_sac_gpu.py::test_host_block_path_and_cache_prefetch_equal_the_plain_calls -v 2>&1 | grep -E "PASSED|FAILED|SKIPPED|passed|failed|Error" >> gpurun_out/r02_dist_tests_2gpu.log; cat gpurun_out/r02_dist_tests_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 1000 --warmup 50 > gpurun_out/u_bench2.json 2> gpurun_out/u_bench2.err; echo "bench2 rc=$?"; tail -3 gpurun_out/u_bench2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/u_bench2.json").read().strip().splitlines()[-1])
print("N=2 value %.3fM us/step %.1f e2e %.3fM (%.1f us) c5 %.1f us" % (d["value"]/1e6, d["ms_per_step"]*1e3, d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"]*1e3, d["c5"]["ms_per_step"]*1e3))
print({k: round(v["ms_per_step"]*1e3, 1) for k, v in d["e2e"].items() if isinstance(v, dict)})
print("c4_sharded", d.get("c4_sharded"))
PY
