#!/bin/bash
# c5 largest point (1M x 100k, k = 64), 1 GPU: count mode of the confusion kernel x row order, store policy of the product kernel
O=gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" python bench.py --workload c5 --steps 10 --warmup 3 --points -1 > $O/r02_c5_$name.json 2>/dev/null
  python - $name <<'PY'
import json, sys
d = json.load(open("gpurun_out/r02_c5_%s.json" % sys.argv[1]))
p = d["sweep"][-1]
print("%-28s product %.3f ms %5.0f GB/s   confusion %.3f ms %5.0f GB/s  agree=%s" % (sys.argv[1], p["product_ms"], p["product_gbs"], p["confusion_ms"], p["confusion_gbs"], p["kernels_agree"]))
PY
}
run count0_rowmap0 BMF_CONFUSION_COUNT=0 BMF_PANEL_ROWMAP=0
run count1_rowmap0 BMF_CONFUSION_COUNT=1 BMF_PANEL_ROWMAP=0
run count2_rowmap0 BMF_CONFUSION_COUNT=2 BMF_PANEL_ROWMAP=0
run count0_rowmap1 BMF_CONFUSION_COUNT=0 BMF_PANEL_ROWMAP=1
run store1 BMF_PRODUCT_STORE=1
run store2 BMF_PRODUCT_STORE=2
