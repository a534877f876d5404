#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c_pytest.log
tail -15 gpurun_out/r02c_pytest.log
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; echo "bench exit $?"
tail -3 gpurun_out/r02c_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02c_bench.json'))
print('ms_per_step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'frac',d['roofline']['frac'])
print('kernel_ms',d['kernel_ms_per_step'])
print('config5',d.get('config5_batch'))
c4=d.get('config4_100M') or {}
print('config4 ms',c4.get('ms_per_extraction'),'frac',(c4.get('roofline') or {}).get('frac'))
PY
# C++ end to end through the shim at the headline size
python - <<'PY'
import numpy as np, sys
sys.path.insert(0,'.')
from dialog_b200 import synth
synth.indoor_scene().points(0,10_000_000)[:,:3].astype('<f4').tofile('/tmp/cloud10m.f32')
PY
timeout 600 ./examples/plane_detect_demo /tmp/cloud10m.f32 0.1 4095 500 --prob 1.0 --max-planes 20 --bench 5 > gpurun_out/r02c_cpp_e2e.log 2>&1; cat gpurun_out/r02c_cpp_e2e.log
