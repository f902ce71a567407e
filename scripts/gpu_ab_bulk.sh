#!/bin/bash
# bulk-copy staged norm kernels: tests with the switch on, then same-box A/B (kbench gn + bench)
TAG=${1:-bulk}
O=gpurun_out; mkdir -p $O
S2S_GN_BULK=1 timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_unet.py tests/test_gpu_parity_config_a.py tests/test_gpu_graphed.py tests/test_gpu_multitask.py -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest(bulk) rc=$?"
tail -5 $O/${TAG}_pytest.log
bash scripts/gpu_ab_env.sh $TAG S2S_GN_BULK=0 S2S_GN_BULK=1
