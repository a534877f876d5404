#!/bin/bash
# round-2 second GPU pass: parity after the kernel fusions, unroll 16/32, host overhead, bench, ncu captures
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02b_pytest.log
tail -4 gpurun_out/r02b_pytest.log
for U in 8 16 32; do PR_SCORE_UNROLL=$U timeout 300 python ubench/score_rate.py 10000000 2>&1 | sed "s/^/U=$U /" >> gpurun_out/r02b_score_unroll.log; done
cat gpurun_out/r02b_score_unroll.log
timeout 600 python ubench/host_overhead.py > gpurun_out/r02b_host_overhead.log 2>&1; cat gpurun_out/r02b_host_overhead.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench exit $?"
tail -3 gpurun_out/r02b_bench.err
# ncu: launch list of a short bench, then full captures of the scoring kernel and of K3/K5 at 100M points
timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-config4 > gpurun_out/r02b_plain1.log 2>&1 && \
  timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02b_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-config4 > gpurun_out/r02b_ncu1.log 2>&1
timeout 300 python ubench/score_rate.py 10000000 > gpurun_out/r02b_plain2.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_kernel -s 8 -c 2 -o gpurun_out/r02b_prof_score \
    python ubench/score_rate.py 10000000 > gpurun_out/r02b_ncu2.log 2>&1
timeout 300 python ubench/hbm_100m.py > gpurun_out/r02b_plain3.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:compact_kernel|refit_kernel" -s 2 -c 2 -o gpurun_out/r02b_prof_hbm \
    python ubench/hbm_100m.py > gpurun_out/r02b_ncu3.log 2>&1
cat gpurun_out/r02b_plain3.log; ls -la gpurun_out/*.ncu-rep | tail -3
