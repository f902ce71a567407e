#!/bin/bash
# kbench sweep of one environment knob: gpu_sweep_env.sh <tag> <VAR> "<v1 v2 ...>" "<kbench args>"
TAG=$1; VAR=$2; VALS=$3; ARGS=$4
O=gpurun_out; mkdir -p $O
for V in $VALS; do
  env $VAR=$V timeout 300 python scripts/kbench.py $ARGS --json $O/${TAG}_${VAR}_$V.json > $O/${TAG}_${VAR}_$V.log 2>&1; echo "$VAR=$V rc=$?"
done
python - <<PY
import json
vals="$VALS".split()
rows=[json.load(open("$O/${TAG}_${VAR}_%s.json" % v)) for v in vals]
print("%-32s %-14s " % ("kernel","shape") + " ".join("%10s" % v for v in vals))
for i,x in enumerate(rows[0]):
    print("%-32s %-14s " % (x["kernel"], x["shape"]) + " ".join("%6.3f(%.2f)" % (r[i]["ms"], r[i].get("frac_hbm", 0)) for r in rows))
PY
