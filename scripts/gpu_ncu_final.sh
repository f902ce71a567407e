#!/bin/bash
# Final evidence capture: (1) launch list of one whole training step of the bench command, (2) ncu --set full of the conv /
# wgrad / norm kernels on the kbench shapes.  Each ncu run follows a plain run of the same command that exited 0.
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
CMD="python bench.py --steps 1 --warmup 3 --no-cpu"
timeout 300 $CMD > $O/${TAG}_plain.json 2> $O/${TAG}_plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 9400 -c 4800 --csv --log-file $O/${TAG}_launches.csv $CMD > $O/${TAG}_ncu_list.log 2>&1
echo "launch list rc=$?"
CMD2="python scripts/kbench.py conv wgrad gn --once --batch 32"
timeout 300 $CMD2 > $O/${TAG}_once.log 2>&1 && \
timeout 900 ncu --set full --clock-control none -k 'regex:conv_halo|conv_igemm|conv_wgrad|gn_|convert16' -c 40 -o /tmp/${TAG}_kernels $CMD2 > $O/${TAG}_ncu_full.log 2>&1
echo "ncu full rc=$?"
ncu -i /tmp/${TAG}_kernels.ncu-rep --page raw --csv > $O/${TAG}_ncu_raw.csv 2>/dev/null
du -sh $O; ls -la $O | grep ${TAG}
