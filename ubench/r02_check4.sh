#!/bin/bash
# 8-GPU pass: sharded parity on 8 ranks, then the bench under torchrun with the per-rank round timeline
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py \
  > gpurun_out/r02m_multi4.log 2>&1; echo "multi4 exit $?"; grep -a "multi-GPU check\|peer-memory\|mismatch\|Error" gpurun_out/r02m_multi4.log | tail -5
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --timeline-out gpurun_out/r02m_timeline_4gpu.json \
  > gpurun_out/r02m_bench_4gpu.json 2> gpurun_out/r02m_bench_4gpu.err; echo "bench exit $?"; tail -3 gpurun_out/r02m_bench_4gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02m_bench_4gpu.json').read().strip().splitlines()[-1])
for k in ("value","ms_per_step","e2e","e2e_breakdown","device_loop_ms_per_step","exchange_wait_ms_per_step_by_rank","config4_100M","config5_batch"):
    print(k, json.dumps(d.get(k))[:1100])
PY
