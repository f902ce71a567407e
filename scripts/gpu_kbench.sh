#!/bin/bash
# [tests] + per-kernel micro-benchmark + (optional) one small ncu --set full capture of the same kernels.
# Usage: gpu_kbench.sh <tag> <ncu 0|1> <pytest 0|1> [ncu kernel count]
TAG=${1:-r1}
NCU=${2:-0}
PYT=${3:-1}
CNT=${4:-28}
O=gpurun_out
mkdir -p $O
if [ "$PYT" = "1" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${TAG}_pytest.log
fi
timeout 600 python scripts/kbench.py --batch 32 --json $O/${TAG}_kbench.json > $O/${TAG}_kbench.log 2>&1; echo "kbench rc=$?"
grep -v Warning $O/${TAG}_kbench.log | cut -c1-200
if [ "$NCU" = "1" ]; then
  CMD="python scripts/kbench.py conv wgrad gn --once --batch 32"
  timeout 300 $CMD > $O/${TAG}_once.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none -k 'regex:conv_igemm|conv_wgrad|gn_|convert16' -c $CNT -o /tmp/${TAG}_kernels $CMD > $O/${TAG}_ncu.log 2>&1
  echo "ncu rc=$?"
  ncu -i /tmp/${TAG}_kernels.ncu-rep --page raw --csv > $O/${TAG}_ncu_raw.csv 2>/dev/null
  ncu -i /tmp/${TAG}_kernels.ncu-rep --page details --csv > $O/${TAG}_ncu_details.csv 2>/dev/null
  SZ=$(stat -c %s /tmp/${TAG}_kernels.ncu-rep 2>/dev/null || echo 0)
  if [ "$SZ" -lt 40000000 ]; then cp /tmp/${TAG}_kernels.ncu-rep $O/; fi
  du -sh $O
fi
