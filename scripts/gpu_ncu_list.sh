#!/bin/bash
# ncu launch list of ONE training step of the bench command (after a plain run of the same command that exited 0).
# ncu costs ~80 ms per intercepted launch even when skipped, and a step is ~1500 launches (817 own + torch glue): the list is
# taken with one warm-up step, skipping it, and the application is killed once the step has been captured.
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
CMD="python bench.py --steps 1 --warmup 1 --no-cpu"
timeout 300 $CMD > $O/${TAG}_plain.json 2> $O/${TAG}_plain.err && \
timeout 800 ncu --metrics gpu__time_duration.sum --clock-control none -s 1550 -c 1550 --kill on --csv --log-file $O/${TAG}_launches.csv $CMD > $O/${TAG}_ncu_list.log 2>&1
echo "launch list rc=$?"
ls -la $O | grep ${TAG}
