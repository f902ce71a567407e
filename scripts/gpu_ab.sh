#!/bin/bash
# same-box A/B of an environment switch: gpu_ab.sh "<env A>" "<env B>"  (train step + sampling, short)
O=gpurun_out; mkdir -p $O
i=0
for E in "$@" "$1"; do
  i=$((i+1))
  env $E timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu --sample-tiles 128 > $O/ab_$i.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("$O/ab_$i.json")); s=d["sample"]
print("[$E]", "train ms/step", round(d["ms_per_step"],2), "tiles/s", round(d["value"],1), "| sample tiles/s", round(s["value"],2), "ms", round(s["ms_per_step"],1), "| clk", d["clocks"]["sm_mhz"], s["clocks"]["sm_mhz"])
PY
done
