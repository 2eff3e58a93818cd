# Round-1 final evidence pass (after the converged-warp issue fix) on 1 x B200; every ncu command follows a plain run of it.
set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/r01d_bench_c4_n1.json 2> gpurun_out/r01d_bench_c4_n1.err
python bench.py --steps 3 --warmup 3 > gpurun_out/r01d_bench_c4_n1_s3.json 2> /dev/null &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01d_launches_c4_bench.csv python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_list_r01d.log 2>&1
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/plain_a.json 2>/dev/null &&
ncu --set full --clock-control none --import-source on -k regex:gemm_f4s_2sm_kernel -s 2 -c 1 -o gpurun_out/r01d_prof_gain_f4s -f python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_gain.log 2>&1
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --w-fp 0.2 > gpurun_out/plain_b.json 2>/dev/null &&
ncu --set full --clock-control none --import-source on -k regex:gemm_f4_2sm_kernel -s 2 -c 1 -o gpurun_out/r01d_prof_gain2_f4 -f python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --w-fp 0.2 > gpurun_out/ncu_gain2.log 2>&1
python bench.py --workload c2 --steps 20 --warmup 3 > gpurun_out/r01d_bench_c2_n1.json 2> /dev/null
ls -la gpurun_out/r01d_*
