#!/bin/bash
# Same-box A/B of two builds of the library: gpu_ab_lib.sh <tag> <old .so> [pytest 0|1]
#   [pytest -m gpu on the new build] -> kbench gn (B = 64 shapes of the step) old / new -> bench.py old, new, old.
TAG=${1:-ab}
OLD=${2:-stain2stain_b200/lib/libs2s_b200_old.so}
PYT=${3:-1}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_gpu.csv 2>&1
if [ "$PYT" = "1" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
  tail -15 $O/${TAG}_pytest.log
fi
for V in old new; do
  if [ "$V" = "old" ]; then export S2S_LIB_PATH=$OLD; else unset S2S_LIB_PATH; fi
  timeout 300 python scripts/kbench.py gn --batch 64 --gn-shapes 128x256,256x128,256x64,512x32 --json $O/${TAG}_kbench_$V.json > $O/${TAG}_kbench_$V.log 2>&1
  echo "kbench $V rc=$?"
done
unset S2S_LIB_PATH
python - <<PY
import json
a=json.load(open("$O/${TAG}_kbench_old.json")); b=json.load(open("$O/${TAG}_kbench_new.json"))
for x,y in zip(a,b):
    print(f"{x['kernel']:32s} {x['shape']:14s} old {x['ms']:.3f} ms ({x['frac_hbm']:.2f})  new {y['ms']:.3f} ms ({y['frac_hbm']:.2f})")
PY
i=0
for V in old new old new; do
  i=$((i+1))
  if [ "$V" = "old" ]; then export S2S_LIB_PATH=$OLD; else unset S2S_LIB_PATH; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --sample-tiles 128 > $O/${TAG}_bench_${i}_$V.json 2> $O/${TAG}_bench_${i}_$V.err
  python - <<PY
import json
d=json.load(open("$O/${TAG}_bench_${i}_$V.json")); s=d["sample"]; k=d["kernels"]
print("[$V]", "train ms/step", round(d["ms_per_step"],2), "tiles/s", round(d["value"],1), "| sample tiles/s", round(s["value"],2), "| clk", d["clocks"]["sm_mhz"], s["clocks"]["sm_mhz"])
print("   ", {n: (round(v["ms"],2), round(v.get("frac_hbm_peak",0),2)) for n,v in k.items() if n.startswith("gn_")})
PY
done
unset S2S_LIB_PATH
