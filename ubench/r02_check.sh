#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "batch" 2>&1 | tail -2
timeout 200 python ubench/batch_rate.py 2>&1 | tail -1
