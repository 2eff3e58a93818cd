#!/bin/bash
# Last single-GPU pass of round 2 (run on the GPU box from the repo root; every profiled target first runs WITHOUT ncu).
O=gpurun_out
timeout 400 python -m pytest tests -x -q -m gpu > $O/r02zc_tests_n1.log 2>&1; tail -3 $O/r02zc_tests_n1.log
timeout 300 python bench.py > $O/r02zc_bench_c4_n1.json 2> $O/r02zc_bench_c4_n1.err; tail -c 300 $O/r02zc_bench_c4_n1.json
timeout 120 python bench.py --workload c5 > $O/r02zc_bench_c5_n1.json 2> $O/r02zc_bench_c5_n1.err; tail -c 300 $O/r02zc_bench_c5_n1.json
timeout 120 python profiles/prof_fit.py c4 3 auto > $O/r02zc_prof_fit_plain.log 2>&1 && \
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02zc_launches_c4_fit.csv \
    python profiles/prof_fit.py c4 3 auto > $O/r02zc_ncu_list.log 2>&1
grep -c "" $O/r02zc_launches_c4_fit.csv
