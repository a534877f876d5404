#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02n_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02n_pytest.log; tail -4 gpurun_out/r02n_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python bench.py --timeline-out gpurun_out/r02n_timeline_1gpu.json > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err; echo "bench exit $?"; tail -2 gpurun_out/r02n_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02n_bench.json').read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["config2_1M"], d["config5_batch"]["ms_per_batch"], d["clocks"])
PY
