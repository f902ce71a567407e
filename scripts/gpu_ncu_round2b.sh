#!/bin/bash
# Round-2 final profiler evidence: (1) plain runs must exit 0, (2) ncu launch list (durations only) of ~2 eager training steps,
# (3) ncu --set full (with source) of the normalisation kernels on the benchmark shapes (kbench gn --once, B = 64).
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-sample --no-graph"
timeout 600 $CMD > $O/${TAG}_plain.json 2> $O/${TAG}_plain.err; rc=$?; echo "plain rc=$rc"
[ $rc -ne 0 ] && { tail -5 $O/${TAG}_plain.err; exit 1; }
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 4000 -c 2600 --csv \
  --log-file $O/${TAG}_launches.csv $CMD > $O/${TAG}_ncu_list.log 2>&1; echo "ncu list rc=$?"
python scripts/summarize_launches.py $O/${TAG}_launches.csv 2>&1 | head -50
CMD2="python scripts/kbench.py gn --once --batch 64 --gn-shapes 128x256,256x64"
timeout 300 $CMD2 > $O/${TAG}_once.log 2>&1; rc=$?; echo "once rc=$rc"
[ $rc -ne 0 ] && { tail -5 $O/${TAG}_once.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:gn_(apply|bwd_reduce|bwd_apply)_kernel' -c 40 \
  -o /tmp/${TAG}_gn $CMD2 > $O/${TAG}_ncu_gn.log 2>&1; echo "ncu gn rc=$?"
ncu -i /tmp/${TAG}_gn.ncu-rep --page raw --csv > $O/${TAG}_ncu_gn_raw.csv 2>/dev/null
python scripts/ncu_summary.py $O/${TAG}_ncu_gn_raw.csv | cut -c1-230
SZ=$(stat -c %s /tmp/${TAG}_gn.ncu-rep 2>/dev/null || echo 0); echo "rep size $SZ"
if [ "$SZ" -lt 40000000 ]; then cp /tmp/${TAG}_gn.ncu-rep $O/; fi
