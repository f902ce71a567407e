#!/bin/bash
# same-box comparison of several environment settings on the bench (train step only): gpu_ab_env2.sh <tag> "<env1>" "<env2>" ...
TAG=$1; shift
O=gpurun_out; mkdir -p $O
i=0
for rep in 1 2; do
for E in "$@"; do
  i=$((i+1))
  env $E timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --sample-tiles 64 > $O/${TAG}_bench_$i.json 2> $O/${TAG}_bench_$i.err
  python - <<PY
import json
d=json.load(open("$O/${TAG}_bench_$i.json")); k=d["kernels"]; s=d.get("sample") or {}
print("[$E]", "train ms/step", round(d["ms_per_step"],2), "| sample", round(s.get("value",0),2), "| clk", d["clocks"]["sm_mhz"])
print("   ", {n: (round(v["ms"],2), round(v.get("frac_hbm_peak",0),2)) for n,v in k.items() if n.startswith("gn_")})
PY
done
done
