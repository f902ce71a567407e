#!/bin/bash
# ncu --set full of the conv kernels of the first training step (all forward + backward shapes), after a plain run.
TAG=${1:-r1}
CMD="python bench.py --steps 1 --warmup 3 --no-cpu"
O=gpurun_out
mkdir -p $O
$CMD > $O/${TAG}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_igemm -c 140 -o $O/${TAG}_conv_igemm $CMD > $O/${TAG}_ncu1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_wgrad -c 80 -o $O/${TAG}_conv_wgrad $CMD > $O/${TAG}_ncu2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gn_ -c 60 -s 100 -o $O/${TAG}_gn $CMD > $O/${TAG}_ncu3.log 2>&1
echo "rc=$?"
ls -la $O | grep ${TAG}
