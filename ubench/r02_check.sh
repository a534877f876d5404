#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02r_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02r_pytest.log; tail -4 gpurun_out/r02r_pytest.log
timeout 200 python ubench/batch_rate.py 2>&1 | tail -1 | tee gpurun_out/r02r_batch.log
