#!/bin/bash
# train + sample bench only.  Usage: gpu_bench.sh <tag> [train batch] [extra bench args]
TAG=${1:-r1}
TB=${2:-64}
O=gpurun_out
mkdir -p $O
timeout 900 python bench.py --steps 5 --warmup 3 --batch $TB > $O/${TAG}_bench_train.json 2> $O/${TAG}_bench_train.err; echo "bench train rc=$?"
tail -c 400 $O/${TAG}_bench_train.err
python scripts/show_bench.py $O/${TAG}_bench_train.json
timeout 900 python bench.py --mode sample --steps 2 --warmup 1 > $O/${TAG}_bench_sample.json 2> $O/${TAG}_bench_sample.err; echo "bench sample rc=$?"
python scripts/show_bench.py $O/${TAG}_bench_sample.json
