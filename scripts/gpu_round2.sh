#!/bin/bash
# One gpurun call: parity tests, smoke, default bench (train line + sampling record).  Usage: gpu_round2.sh <tag> [sample tiles]
TAG=${1:-r2}
ST=${2:-128}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $O/${TAG}_gpu.csv 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
tail -25 $O/${TAG}_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/${TAG}_smoke.log
tail -2 $O/${TAG}_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 --sample-tiles $ST > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
tail -c 600 $O/${TAG}_bench.err
python scripts/show_bench.py $O/${TAG}_bench.json
