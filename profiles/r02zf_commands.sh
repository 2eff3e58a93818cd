O=gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu -k "basis or expand or cover_apply or matches_reference or c2_k20 or c4_k20 or assoc" > $O/r02zf_tests.log 2>&1; tail -3 $O/r02zf_tests.log
for cfg in "16 8" "32 8" "64 8" "16 4" "32 16"; do set -- $cfg
  BMF_STAGE_SLOT_MB=$1 BMF_STAGE_THREADS=$2 timeout 120 python profiles/fit_trace.py > /dev/null 2> $O/r02zf_fit_trace_slot$1_t$2.log
  echo "slot $1 MB, $2 threads:"; grep "fit trace" $O/r02zf_fit_trace_slot$1_t$2.log | tail -1 | cut -c1-140; grep "trace=0" $O/r02zf_fit_trace_slot$1_t$2.log | tail -2 | cut -c1-50
done
timeout 120 python profiles/prof_fit.py c4 3 auto > $O/r02zf_prof_fit_plain.log 2>&1 && \
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02zf_launches_c4_fit.csv \
    python profiles/prof_fit.py c4 3 auto > $O/r02zf_ncu_list.log 2>&1
grep "basis_\|min_counts\|expand_bits_f4" $O/r02zf_launches_c4_fit.csv | awk -F'","' '{print substr($5,1,40), $NF}'
