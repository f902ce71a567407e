#!/bin/bash
# Last validation of the round: GPU tests, then the default driver-style bench line.
TAG=${1:-r4}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_gpu.csv 2>&1
SECONDS=0
timeout 150 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$? ${SECONDS}s" | tee -a $O/${TAG}_pytest.log
tail -3 $O/${TAG}_pytest.log
SECONDS=0
timeout 170 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$? ${SECONDS}s"
python scripts/show_bench.py $O/${TAG}_bench.json | head -12
