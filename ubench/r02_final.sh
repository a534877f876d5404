#!/bin/bash
# round-2 evidence pass: parity, bench lines, ncu captures (final kernels)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02h_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02h_pytest.log; tail -4 gpurun_out/r02h_pytest.log
timeout 1500 python bench.py > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference > gpurun_out/r02h_bench_reference.json 2> gpurun_out/r02h_bench_reference.err; echo "reference exit $?"
timeout 1800 python bench.py --steps 3 --extras --no-config4 --no-config5 > gpurun_out/r02h_bench_extras.json 2> gpurun_out/r02h_bench_extras.err; echo "extras exit $?"; tail -3 gpurun_out/r02h_bench_extras.err
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-config4 --no-config5"
timeout 900 $CMD > gpurun_out/r02h_plain1.log 2>&1 && \
  timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02h_launches.csv $CMD > gpurun_out/r02h_ncu1.log 2>&1
timeout 900 $CMD > gpurun_out/r02h_plain2.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_kernel -s 6 -c 2 -o gpurun_out/r02h_prof_score $CMD > gpurun_out/r02h_ncu2.log 2>&1
timeout 300 python ubench/hbm_100m.py > gpurun_out/r02h_plain3.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:compact_kernel|refit_kernel" -s 2 -c 2 -o gpurun_out/r02h_prof_hbm \
    python ubench/hbm_100m.py > gpurun_out/r02h_ncu3.log 2>&1
cat gpurun_out/r02h_plain3.log; ls -la gpurun_out/r02h_*.ncu-rep
