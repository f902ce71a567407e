#!/bin/bash
# ncu launch list (gpu__time_duration per launch) of the bench command, only after the same command exited 0 without ncu.
TAG=${1:-r1}
shift
CMD="python bench.py --steps 1 --warmup 3 --no-cpu $@"
O=gpurun_out
mkdir -p $O
$CMD > $O/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file $O/${TAG}_launches.csv $CMD > $O/${TAG}_ncu.log 2>&1
echo "rc=$?"
tail -3 $O/${TAG}_ncu.log
