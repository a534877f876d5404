#!/bin/bash
# round-2 first GPU pass: parity, score-loop unroll variants, HBM kernels at 100M, host overhead, a short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_smi.txt
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest.log
for U in 2 4 8; do PR_SCORE_UNROLL=$U timeout 300 python ubench/score_rate.py 10000000 2>&1 | sed "s/^/U=$U /" >> gpurun_out/r02_score_unroll.log; done
for U in 2 4 8; do PR_SCORE_UNROLL=$U timeout 300 python ubench/score_rate.py 2500000 2>&1 | sed "s/^/U=$U /" >> gpurun_out/r02_score_unroll.log; done
timeout 600 python ubench/hbm_rate.py > gpurun_out/r02_hbm_rate.log 2>&1
timeout 600 python ubench/host_overhead.py > gpurun_out/r02_host_overhead.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 --no-hbm-100m > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err
tail -5 gpurun_out/r02_pytest.log; cat gpurun_out/r02_score_unroll.log gpurun_out/r02_hbm_rate.log gpurun_out/r02_host_overhead.log
