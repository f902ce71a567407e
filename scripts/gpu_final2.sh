#!/bin/bash
# Final validation as the driver runs it: GPU tests, smoke, reference arm, default bench line (20 steps, 5 warm-up).
TAG=${1:-r2final}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $O/${TAG}_gpu.csv 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
tail -4 $O/${TAG}_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/${TAG}_smoke.log
tail -2 $O/${TAG}_smoke.log
SECONDS=0; timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/${TAG}_ref.json 2> $O/${TAG}_ref.err; echo "ref rc=$?"
python -c "import json; d=json.load(open('$O/${TAG}_ref.json')); print('reference', d['value'], d['steps'], d['ms_per_step'], d['cpu_baseline']['cores'])"
echo "reference arm wall ${SECONDS}s"; SECONDS=0; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
python scripts/show_bench.py $O/${TAG}_bench.json
echo "bench wall ${SECONDS}s"
