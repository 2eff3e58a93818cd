O=gpurun_out
timeout 400 python -m pytest tests -x -q -m gpu > $O/r02zd_tests_n1.log 2>&1; tail -3 $O/r02zd_tests_n1.log
timeout 120 python profiles/fit_time.py c4 20 3 > $O/r02zd_fit_time_c4.log 2>&1; cut -c1-400 $O/r02zd_fit_time_c4.log
BMF_FIT_TRACE=1 timeout 120 python profiles/fit_trace.py > $O/r02zd_fit_trace_n1.log 2>&1; tail -4 $O/r02zd_fit_trace_n1.log
timeout 120 python profiles/prof_fit.py c4 3 auto > $O/r02zd_prof_fit_plain.log 2>&1 && \
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02zd_launches_c4_fit.csv \
    python profiles/prof_fit.py c4 3 auto > $O/r02zd_ncu_list.log 2>&1
grep "gemm_f4s_2sm_kernel<1>\|basis_threshold\|expand_bits_f4" $O/r02zd_launches_c4_fit.csv | awk -F'","' '{print $5, $(NF-1), $NF}' | cut -c1-120
