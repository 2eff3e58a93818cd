#!/bin/bash
# Round-2 profiling pass (run on the GPU box from the repo root; every target first runs WITHOUT ncu).
set -x
O=gpurun_out
python profiles/prof_fit.py c4 3 auto > $O/r02_prof_fit_plain.log 2>&1 || exit 1
# launch list of a 3-step incremental fit (per-launch times are cold-cache and serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_c4_fit.csv \
    python profiles/prof_fit.py c4 3 auto > $O/r02_ncu_list.log 2>&1
# full capture: the full-pass scoring GEMM (first launch) and the two incremental re-scoring launches of step 1
ncu --set full --clock-control none --import-source on -k regex:gemm_f4s_2sm_kernel -s 3 -c 3 -o $O/r02_prof_gain_f4s \
    python profiles/prof_fit.py c4 3 auto > $O/r02_ncu_gain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cover_apply_kernel -c 1 -o $O/r02_prof_apply_compact \
    python profiles/prof_fit.py c4 3 auto > $O/r02_ncu_apply.log 2>&1
for r in r02_prof_gain_f4s r02_prof_apply_compact; do
  ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_raw.csv 2>/dev/null
done
