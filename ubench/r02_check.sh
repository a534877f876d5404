#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02l_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02l_pytest.log; tail -5 gpurun_out/r02l_pytest.log
for w in default 1,1,1,1,1,1,1,1 1,2,4,8; do
  if [ "$w" = default ]; then unset PR_UPLOAD_WEIGHTS; else export PR_UPLOAD_WEIGHTS=$w; fi
  timeout 300 python ubench/e2e_chunks.py 2>&1 | tail -2
done | tee gpurun_out/r02l_chunks.log
unset PR_UPLOAD_WEIGHTS
timeout 300 python ubench/e2e_round0.py 2>&1 | tail -6
timeout 900 python bench.py --no-config4 --steps 3 --no-cpu-baseline > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err; echo "bench exit $?"; tail -3 gpurun_out/r02l_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02l_bench.json').read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], json.dumps(d["e2e"]), json.dumps(d["config5_batch"])[:1500])
PY
