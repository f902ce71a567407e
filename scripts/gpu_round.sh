#!/bin/bash
# One gpurun call: parity tests, smoke, train + sample bench, then (only if the plain bench exited 0) the ncu launch list.
# Usage: scripts/gpu_round.sh <tag> [train batch]
TAG=${1:-r1}
TB=${2:-64}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $O/${TAG}_gpu.csv 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/${TAG}_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 --batch $TB > $O/${TAG}_bench_train.json 2> $O/${TAG}_bench_train.err; echo "bench train rc=$?"
tail -c 600 $O/${TAG}_bench_train.err
timeout 900 python bench.py --mode sample --steps 2 --warmup 1 > $O/${TAG}_bench_sample.json 2> $O/${TAG}_bench_sample.err; echo "bench sample rc=$?"
tail -c 600 $O/${TAG}_bench_sample.err
