O=gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu -k "basis or expand or cover_apply or matches_reference or c2_k20 or c4_k20 or assoc" > $O/r02ze_tests.log 2>&1; tail -3 $O/r02ze_tests.log
timeout 120 python profiles/fit_trace.py > /dev/null 2> $O/r02ze_fit_trace_n1.log; tail -5 $O/r02ze_fit_trace_n1.log
timeout 120 python profiles/prof_fit.py c4 3 auto > $O/r02ze_prof_fit_plain.log 2>&1 && \
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02ze_launches_c4_fit.csv \
    python profiles/prof_fit.py c4 3 auto > $O/r02ze_ncu_list.log 2>&1
grep "basis_\|expand_bits_f4" $O/r02ze_launches_c4_fit.csv | awk -F'","' '{print substr($5,1,40), $NF}'
