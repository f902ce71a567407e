"""Per-kernel CUDA-event breakdown of ONE eager velocity evaluation + Euler update (the body of the sampler's CUDA graph)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from stain2stain_b200 import kernels as K  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda", 0)
lit = bench.build_lit(dev)
lit.eval()
net = lit.net
x = torch.rand(B, 3, 256, 256, device=dev) * 2 - 1
t = torch.full((B,), 0.3, device=dev)
with torch.no_grad():
    for _ in range(3):
        net.euler_step_(t, x, 0.02)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        net.euler_step_(t, x, 0.02)
    e1.record()
    torch.cuda.synchronize()
    print(f"eager evaluation: {e0.elapsed_time(e1) / 5:.3f} ms at B={B}")
    K.PROFILE = []
    net.euler_step_(t, x, 0.02)
    torch.cuda.synchronize()
    recs = list(K.PROFILE)
    prof = K.profile_summary(K.PROFILE)
    K.PROFILE = None
    print("per-launch conv records (ms, TFLOP/s algorithmic, GFLOP):")
    for i, (name, a, b, fl, by, xfl) in enumerate(recs):
        if name == "conv_igemm":
            ms = a.elapsed_time(b)
            print(f"   #{i:3d} {ms:7.3f} ms {fl / ms / 1e9:7.0f} TF/s  alg {fl / 1e9:8.1f} GF  exec {xfl / 1e9:8.1f} GF")
tot = sum(d["ms"] for d in prof.values())
print(f"profiled kernels: {tot:.3f} ms")
for k, d in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
    extra = f"{d['flops'] / d['ms'] / 1e9:8.0f} TFLOP/s" if d["flops"] else (f"{d['bytes'] / d['ms'] / 1e6:8.0f} GB/s" if d["bytes"] else "")
    print(f"  {k:22s} n={d['launches']:3d} {d['ms']:8.3f} ms  {extra}")
