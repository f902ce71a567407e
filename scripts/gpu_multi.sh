#!/bin/bash
# N-GPU bench (torchrun as the driver launches it).  Usage: gpu_multi.sh <tag> <N> [extra bench args]
TAG=${1:-r2}
N=${2:-2}
shift; shift
O=gpurun_out
mkdir -p $O
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 10 --warmup 3 "$@" > $O/${TAG}_ddp${N}.json 2> $O/${TAG}_ddp${N}.err; echo "bench N=$N rc=$?"
tail -c 1500 $O/${TAG}_ddp${N}.err | grep -v Warning
python scripts/show_bench.py $O/${TAG}_ddp${N}.json
