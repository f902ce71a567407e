#!/bin/bash
# 2-GPU multitask bench: per-rank BatchNorm (graph), SyncBatchNorm (graph: NCCL all-reduces captured), SyncBatchNorm (eager DDP)
O=gpurun_out; mkdir -p $O
run() { # tag, extra args
  TAG=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --mode multitask --steps 8 --warmup 3 "$@" > $O/syncbn_$TAG.json 2> $O/syncbn_$TAG.err; echo "$TAG rc=$?"
  tail -c 600 $O/syncbn_$TAG.err | grep -v Warning | tail -5
  python - <<PY
import json
try:
    d=json.load(open("$O/syncbn_$TAG.json")); print("$TAG", round(d["value"],1), "tiles/s", round(d["ms_per_step"],2), "ms/step loss", d["loss"], d["config"].get("batchnorm"), "|", d["config"]["step_launch"][:30])
except Exception as e: print("$TAG: no json", e)
PY
}
run plain_graph
run sync_graph --sync-bn
run sync_eager --sync-bn --no-graph
