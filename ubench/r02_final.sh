#!/bin/bash
# round-2 evidence pass (1 GPU): parity, bench lines, round timeline, ncu captures of the final kernels
mkdir -p gpurun_out
T=r02m
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log; tail -4 gpurun_out/${T}_pytest.log
timeout 1500 python bench.py --timeline-out gpurun_out/${T}_timeline_1gpu.json > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo "reference exit $?"
timeout 1800 python bench.py --steps 3 --extras --no-config4 --no-config5 > gpurun_out/${T}_bench_extras.json 2> gpurun_out/${T}_bench_extras.err; echo "extras exit $?"; tail -3 gpurun_out/${T}_bench_extras.err
timeout 600 python ubench/host_overhead.py > gpurun_out/${T}_overhead.log 2>&1; echo "overhead exit $?"
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-config4 --no-config5"
timeout 900 $CMD > gpurun_out/${T}_plain1.log 2>&1 && \
  timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/${T}_ncu1.log 2>&1
timeout 900 $CMD > gpurun_out/${T}_plain2.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_kernel -s 6 -c 2 -o gpurun_out/${T}_prof_score $CMD > gpurun_out/${T}_ncu2.log 2>&1
timeout 300 python ubench/hbm_100m.py > gpurun_out/${T}_plain3.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:compact_kernel|refit_kernel" -s 2 -c 2 -o gpurun_out/${T}_prof_hbm \
    python ubench/hbm_100m.py > gpurun_out/${T}_ncu3.log 2>&1
cat gpurun_out/${T}_plain3.log; ls -la gpurun_out/${T}_*.ncu-rep
