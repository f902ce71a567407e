#!/bin/bash
O=gpurun_out; TAG=${1:-r2s}; mkdir -p $O
timeout 200 python scripts/ncu_shortk.py 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_ --launch-skip 4 -c 2 -o /tmp/${TAG}_shortk python scripts/ncu_shortk.py > $O/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/${TAG}_shortk.ncu-rep --page raw --csv > $O/${TAG}_raw.csv 2>/dev/null
python scripts/ncu_summary.py $O/${TAG}_raw.csv | cut -c1-220
ncu -i /tmp/${TAG}_shortk.ncu-rep --page source --csv > $O/${TAG}_source.csv 2>/dev/null
ls -la $O/${TAG}_source.csv; head -c 600 $O/${TAG}_source.csv
ncu -i /tmp/${TAG}_shortk.ncu-rep --page details --csv 2>/dev/null | grep -i -E "stall|Issued Warp|No Eligible|Eligible Warps|Active Warps|One or More" | cut -c1-260 | head -60
