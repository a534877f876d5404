#!/bin/bash
# quick 1-GPU validation pass: parity, stress, round-loop overhead, short bench
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02j_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02j_pytest.log; tail -15 gpurun_out/r02j_pytest.log
timeout 900 python ubench/stress_loops.py > gpurun_out/r02j_stress.log 2>&1; echo "stress exit $?"; tail -3 gpurun_out/r02j_stress.log
timeout 1500 python bench.py --no-config4 --timeline-out gpurun_out/r02j_timeline_1gpu.json > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err; echo "bench exit $?"; tail -5 gpurun_out/r02j_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02j_bench.json').read().strip().splitlines()[-1])
for k in ("value","ms_per_step","e2e","roofline","device_loop_ms_per_step","kernel_ms_per_step"):
    print(k, json.dumps(d.get(k))[:1200])
PY
