#!/bin/bash
# kbench timings of the norm / attention kernels + one ncu --set full capture of the norm-backward and attention kernels.
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
timeout 600 python scripts/kbench.py gn attn --batch 32 --json $O/${TAG}_kbench_gn.json > $O/${TAG}_kbench_gn.log 2>&1; echo "kbench rc=$?"
grep -E "gn_bwd|gn_apply|attn" $O/${TAG}_kbench_gn.log | cut -c1-170
CMD="python scripts/kbench.py gn attn --once --batch 32"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:gn_bwd_(reduce|apply)_kernel|attn_' -c 40 -o /tmp/${TAG}_gn $CMD > $O/${TAG}_ncu_gn.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/${TAG}_gn.ncu-rep --page raw --csv > $O/${TAG}_ncu_gn_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}_gn.ncu-rep --page details --csv > $O/${TAG}_ncu_gn_details.csv 2>/dev/null
python scripts/ncu_summary.py $O/${TAG}_ncu_gn_raw.csv | cut -c1-220
SZ=$(stat -c %s /tmp/${TAG}_gn.ncu-rep 2>/dev/null || echo 0)
if [ "$SZ" -lt 30000000 ]; then cp /tmp/${TAG}_gn.ncu-rep $O/; fi
