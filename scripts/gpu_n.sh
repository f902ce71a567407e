#!/bin/bash
# N-GPU default bench line (train + sampling shards), as the driver launches it.  Usage: gpu_n.sh <tag> <N> [bench args]
TAG=${1:-r3}; N=${2:-8}; shift; shift
O=gpurun_out; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 10 --warmup 3 "$@" > $O/${TAG}_train_${N}gpu.json 2> $O/${TAG}_train_${N}gpu.err; echo "train N=$N rc=$?"
tail -c 400 $O/${TAG}_train_${N}gpu.err | grep -v -i warn
python - <<PY
import json
d=json.load(open("$O/${TAG}_train_${N}gpu.json"))
print("train", "tiles/s", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "clk", (d.get("clocks") or {}).get("sm_mhz"))
s=d.get("sample")
if s: print("  sample tiles/s", round(s["value"],2), "e2e", round(s["e2e"]["value"],2), "tiles", s["config"]["tiles_total"], "frac", round(s["roofline"]["frac"],3))
PY
