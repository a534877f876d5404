#!/bin/bash
# 2-GPU validation pass: sharded parity tests + the bench under torchrun
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02p_pytest2.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02p_pytest2.log; tail -4 gpurun_out/r02p_pytest2.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --timeline-out gpurun_out/r02p_timeline_2gpu.json \
  > gpurun_out/r02p_bench_2gpu.json 2> gpurun_out/r02p_bench_2gpu.err; echo "bench exit $?"; tail -3 gpurun_out/r02p_bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02p_bench_2gpu.json').read().strip().splitlines()[-1])
for k in ("value","ms_per_step","e2e","device_loop_ms_per_step","exchange_wait_ms_per_step_by_rank","config4_100M","config5_batch"):
    print(k, json.dumps(d.get(k))[:900])
PY
