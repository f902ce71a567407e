#!/bin/bash
# Round-2 profiler evidence: (1) plain run must exit 0, (2) ncu launch list (durations only) of ~2 eager training steps,
# (3) ncu --set full of the dominant conv kernel's launches at the benchmark batch (B = 64).
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-sample --no-graph"
timeout 600 $CMD > $O/${TAG}_plain.json 2> $O/${TAG}_plain.err; rc=$?; echo "plain rc=$rc"
[ $rc -ne 0 ] && { tail -5 $O/${TAG}_plain.err; exit 1; }
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 4000 -c 2600 --csv \
  --log-file $O/${TAG}_launches.csv $CMD > $O/${TAG}_ncu_list.log 2>&1; echo "ncu list rc=$?"
python scripts/summarize_launches.py $O/${TAG}_launches.csv 2>&1 | head -60
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:conv_halo_pair --launch-skip 60 -c 24 \
  -o /tmp/${TAG}_conv $CMD > $O/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/${TAG}_conv.ncu-rep --page raw --csv > $O/${TAG}_ncu_conv_raw.csv 2>/dev/null
python scripts/ncu_summary.py $O/${TAG}_ncu_conv_raw.csv | cut -c1-230
SZ=$(stat -c %s /tmp/${TAG}_conv.ncu-rep 2>/dev/null || echo 0); echo "rep size $SZ"
if [ "$SZ" -lt 40000000 ]; then cp /tmp/${TAG}_conv.ncu-rep $O/; fi
