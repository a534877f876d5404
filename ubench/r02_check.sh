#!/bin/bash
# evidence refresh after the last scoring-kernel change: bench line, launch list, score capture
mkdir -p gpurun_out
T=r02p
timeout 1500 python bench.py --timeline-out gpurun_out/${T}_timeline_1gpu.json > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench exit $?"
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-config4 --no-config5"
timeout 900 $CMD > gpurun_out/${T}_plain1.log 2>&1 && \
  timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/${T}_ncu1.log 2>&1
timeout 900 $CMD > gpurun_out/${T}_plain2.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_kernel -s 6 -c 2 -o gpurun_out/${T}_prof_score $CMD > gpurun_out/${T}_ncu2.log 2>&1
ls -la gpurun_out/${T}_*.ncu-rep
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02p_bench.json').read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["roofline"].get("frac_cuda_events"), d["config2_1M"]["ms_per_extraction"], d["config5_batch"]["ms_per_batch"], d["config5_batch"]["frac_of_fp32_peak_whole_call"], d["clocks"])
PY
