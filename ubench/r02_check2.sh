#!/bin/bash
# 2-GPU validation pass: sharded parity tests + the bench under torchrun
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02i_pytest2.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02i_pytest2.log; tail -4 gpurun_out/r02i_pytest2.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 \
  > gpurun_out/r02i_bench_2gpu.json 2> gpurun_out/r02i_bench_2gpu.err; echo "bench exit $?"; head -c 9000 gpurun_out/r02i_bench_2gpu.json
