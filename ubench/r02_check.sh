#!/bin/bash
mkdir -p gpurun_out
timeout 600 python ubench/host_overhead.py > gpurun_out/r02q_overhead.log 2>&1; echo "overhead exit $?"
timeout 200 python ubench/batch_rate.py > gpurun_out/r02q_batch.log 2>&1; tail -1 gpurun_out/r02q_batch.log
grep "no events" gpurun_out/r02q_overhead.log
