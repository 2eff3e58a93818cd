# Round-1 (final kernels) evidence pass on 1 x B200; every ncu command follows a plain run of the same command line.
set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/r01c_bench_c4_n1.json 2> gpurun_out/r01c_bench_c4_n1.err
python bench.py --steps 3 --warmup 3 > gpurun_out/r01c_bench_c4_n1_s3.json 2> /dev/null &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01c_launches_c4_bench.csv python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_list_r01c.log 2>&1
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/plain_a.json 2>/dev/null &&
ncu --set full --clock-control none --import-source on -k regex:gemm_f4_2sm_kernel -s 4 -c 1 -o gpurun_out/r01c_prof_gain_f4 -f python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_gain.log 2>&1
python bench.py --workload c5 --points -1 --steps 1 --warmup 1 > gpurun_out/plain_c.json 2>/dev/null &&
ncu --set full --clock-control none --import-source on -k regex:confusion_panel -c 1 -o gpurun_out/r01c_prof_c5_confusion -f python bench.py --workload c5 --points -1 --steps 1 --warmup 1 > gpurun_out/ncu_c5a.log 2>&1
python bench.py --workload c5 --steps 20 --warmup 3 > gpurun_out/r01c_bench_c5_n1.json 2> gpurun_out/r01c_bench_c5_n1.err
python bench.py --workload c2 --steps 20 --warmup 3 > gpurun_out/r01c_bench_c2_n1.json 2> gpurun_out/r01c_bench_c2_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01c_bench_reference_arm.json 2> /dev/null
ls -la gpurun_out/r01c_*
