#!/bin/bash
# Raster-group sweep of the FP4 scoring GEMM at c4 (N=1): step time, SM clock, power and DRAM bytes per launch.
# BMF_GROUP_M2 = candidate tiles (of 256) per raster group; 16 is the round-1 default.
O=gpurun_out
for g in 8 16 24 35 70; do
  BMF_GROUP_M2=$g python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-extras > $O/r02_group_${g}.json 2>/dev/null
  python - $g <<'PY'
import json, sys
g = sys.argv[1]
d = json.load(open("gpurun_out/r02_group_%s.json" % g))
print("group %s: %.3f ms/step  kernel %.3f ms  %.2f Pop/s  clocks %s" % (g, d["ms_per_step"], d["roofline"]["kernel_ms"], d["value"] / 1e6, d["clocks"]))
PY
done
for g in 16 35 70; do
  BMF_GROUP_M2=$g ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none \
      -k regex:gemm_f4s_2sm_kernel -s 3 -c 1 --csv --log-file $O/r02_group_${g}_dram.csv python profiles/prof_fit.py c4 1 full > /dev/null 2>&1
  grep -v "^==" $O/r02_group_${g}_dram.csv | tail -4 | cut -d, -f5,13-16 | cut -c1-200
done
