#!/bin/bash
# tests + default bench + multitask bench (1 GPU).  Usage: gpu_round3.sh <tag>
TAG=${1:-r3}
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
tail -4 $O/${TAG}_pytest.log
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu --sample-tiles 128 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
python scripts/show_bench.py $O/${TAG}_bench.json | grep -vE "^e2e|cpu_baseline|peak_mem|loss"
timeout 600 python bench.py --mode multitask --steps 8 --warmup 3 --no-cpu > $O/${TAG}_mt.json 2> $O/${TAG}_mt.err; echo "multitask rc=$?"
python scripts/show_bench.py $O/${TAG}_mt.json | grep -vE "^e2e|cpu_baseline|peak_mem|loss"
