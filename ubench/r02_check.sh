#!/bin/bash
# quick 1-GPU validation pass: parity, stress, round-loop overhead, short bench
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02i_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02i_pytest.log; tail -4 gpurun_out/r02i_pytest.log
timeout 900 python ubench/stress_loops.py > gpurun_out/r02i_stress.log 2>&1; echo "stress exit $?"; tail -3 gpurun_out/r02i_stress.log
timeout 600 python ubench/host_overhead.py > gpurun_out/r02i_overhead.log 2>&1; echo "overhead exit $?"; tail -30 gpurun_out/r02i_overhead.log
timeout 1500 python bench.py --no-config4 > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err; echo "bench exit $?"; cat gpurun_out/r02i_bench.json | head -c 6000
