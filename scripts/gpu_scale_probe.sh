#!/bin/bash
# Same box, back to back: N=1 graph, N=2 graph, N=2 eager DDP (one bucket), N=2 eager DDP (25 MB buckets).
O=gpurun_out; TAG=${1:-r2}; mkdir -p $O
run() { # name N args...
  name=$1; N=$2; shift; shift
  if [ "$N" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-sample --no-cpu "$@" > $O/${TAG}_${name}.json 2> $O/${TAG}_${name}.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 10 --warmup 3 --no-sample --no-cpu "$@" > $O/${TAG}_${name}.json 2> $O/${TAG}_${name}.err
  fi
  echo "$name rc=$?"; python -c "
import json,sys
d=json.load(open('$O/${TAG}_${name}.json'))
print('$name', 'ms/step', round(d['ms_per_step'],2), 'tiles/s', round(d['value'],1), 'clk', d['clocks']['sm_mhz'], 'prof_sum', d.get('profiled_kernel_ms_per_step'))
"
}
run n1_graph 1
run n2_graph 2
run n2_ddp1b 2 --no-graph
run n2_ddp25 2 --no-graph --bucket-mb 25
run n1_graph_b 1
