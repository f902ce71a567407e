#!/bin/bash
# 8-GPU evidence for BASELINE.md section 4: configs[1]+[2] (default line: train + 4096-tile sampling), configs[3], configs[4].
TAG=${1:-r2}
N=${2:-8}
O=gpurun_out
mkdir -p $O
run() { name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N "$@" > $O/${TAG}_${name}_${N}gpu.json 2> $O/${TAG}_${name}_${N}gpu.err; echo "$name rc=$?"
  tail -c 400 $O/${TAG}_${name}_${N}gpu.err | grep -v -i warn
  python - <<PY
import json
d=json.load(open("$O/${TAG}_${name}_${N}gpu.json"))
print("$name", "tiles/s", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "clk", (d.get("clocks") or {}).get("sm_mhz"))
s=d.get("sample")
if s: print("  sample tiles/s", round(s["value"],2), "e2e", round(s["e2e"]["value"],2), "tiles", s["config"]["tiles_total"], "frac", round(s["roofline"]["frac"],3))
PY
}
run train --steps 10 --warmup 3
run classcond --mode classcond --steps 10 --warmup 3 --no-cpu
run multitask --mode multitask --steps 10 --warmup 3
run train_ddp --steps 10 --warmup 3 --no-graph --no-sample --no-cpu
