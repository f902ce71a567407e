#!/bin/bash
# same-box A/B of an environment switch on kbench gn + the bench: gpu_ab_env.sh <tag> "<env A>" "<env B>"
TAG=$1; A=$2; B=$3
O=gpurun_out; mkdir -p $O
i=0
for E in "$A" "$B"; do
  i=$((i+1))
  env $E timeout 300 python scripts/kbench.py gn --batch 64 --iters 10 --gn-shapes 128x256,256x128,256x64,512x32 --json $O/${TAG}_kb_$i.json > $O/${TAG}_kb_$i.log 2>&1; echo "kbench [$E] rc=$?"
done
python - <<PY
import json
a=json.load(open("$O/${TAG}_kb_1.json")); b=json.load(open("$O/${TAG}_kb_2.json"))
for x,y in zip(a,b):
    if "bwd_apply" in x["kernel"]:
        print(f"{x['kernel']:32s} {x['shape']:14s} A {x['ms']:.3f} ms ({x['frac_hbm']:.2f})  B {y['ms']:.3f} ms ({y['frac_hbm']:.2f})")
PY
i=0
for E in "$A" "$B" "$A" "$B"; do
  i=$((i+1))
  env $E timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-sample > $O/${TAG}_bench_$i.json 2> $O/${TAG}_bench_$i.err
  python - <<PY
import json
d=json.load(open("$O/${TAG}_bench_$i.json")); k=d["kernels"]
print("[$E]", "train ms/step", round(d["ms_per_step"],2), "tiles/s", round(d["value"],1), "| clk", d["clocks"]["sm_mhz"])
print("   ", {n: (round(v["ms"],2), round(v.get("frac_hbm_peak",0),2)) for n,v in k.items() if n.startswith("gn_bwd_apply")})
PY
done
