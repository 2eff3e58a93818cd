set -x
timeout 600 python -m pytest tests/test_asso_gpu.py -m gpu -x -q -k "general or canonical" 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 > gpurun_out/r01b_bench_c4_n1.json 2> gpurun_out/r01b_bench_c4_n1.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01b_launches_c4_bench.csv python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_list_r01b.log 2>&1
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/plain_a.json 2>/dev/null &&
ncu --set full --clock-control none --import-source on -k regex:gemm_i8_2sm_kernel -s 4 -c 1 -o gpurun_out/r01b_prof_gain python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_gain.log 2>&1
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --w-fp 0.2 > gpurun_out/plain_b.json 2>/dev/null &&
ncu --set full --clock-control none --import-source on -k regex:gemm_i8_2sm_kernel -s 4 -c 1 -o gpurun_out/r01b_prof_gain2 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --w-fp 0.2 > gpurun_out/ncu_gain2.log 2>&1
python bench.py --workload c5 --points -1 --steps 1 --warmup 1 > gpurun_out/plain_c.json 2>/dev/null &&
ncu --set full --clock-control none --import-source on -k regex:panel -c 2 -o gpurun_out/r01b_prof_c5 python bench.py --workload c5 --points -1 --steps 1 --warmup 1 > gpurun_out/ncu_c5.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
