#!/bin/bash
# Last single-GPU pass of round 2, final code (run on the GPU box from the repo root; profiled targets first run WITHOUT ncu).
O=gpurun_out
timeout 400 python -m pytest tests -x -q -m gpu > $O/r02zg_tests_n1.log 2>&1; tail -2 $O/r02zg_tests_n1.log
timeout 300 python bench.py > $O/r02zg_bench_c4_n1.json 2> $O/r02zg_bench_c4_n1.err; tail -c 200 $O/r02zg_bench_c4_n1.json
timeout 120 python bench.py --workload c5 > $O/r02zg_bench_c5_n1.json 2> $O/r02zg_bench_c5_n1.err; tail -c 200 $O/r02zg_bench_c5_n1.json
timeout 120 python profiles/prof_fit.py c4 3 auto > $O/r02zg_prof_fit_plain.log 2>&1 || exit 1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02zg_launches_c4_fit.csv \
    python profiles/prof_fit.py c4 3 auto > $O/r02zg_ncu_list.log 2>&1
# full captures: the first symmetric association launch (EPI_STORE = <1>) and the ring-fed cover_apply of step 1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_f4s_2sm_kernel -c 1 -o $O/r02zg_prof_assoc_f4s \
    python profiles/prof_fit.py c4 3 auto > $O/r02zg_ncu_assoc.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:cover_apply_ring_kernel -c 1 -o $O/r02zg_prof_apply_ring \
    python profiles/prof_fit.py c4 3 auto > $O/r02zg_ncu_apply.log 2>&1
for r in r02zg_prof_assoc_f4s r02zg_prof_apply_ring; do
  ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_raw.csv 2>/dev/null
done
ls -la $O/r02zg*
