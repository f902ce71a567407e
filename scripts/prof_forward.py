import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from stain2stain_b200 import kernels as K
dev = torch.device("cuda", 0)
lit = bench.build_lit(dev)
lit.eval()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
x = torch.rand(B, 3, 256, 256, device=dev) * 2 - 1
t = torch.zeros(B, device=dev)
with torch.no_grad():
    for _ in range(3):
        lit.net.euler_step_(t, x, 0.02)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        lit.net.euler_step_(t, x, 0.02)
    e1.record()
    torch.cuda.synchronize()
    print("eager ms per evaluation:", e0.elapsed_time(e1) / 5)
    K.PROFILE = []
    lit.net.euler_step_(t, x, 0.02)
    torch.cuda.synchronize()
    prof = K.profile_summary(K.PROFILE)
    K.PROFILE = None
tot = 0
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
    tot += v["ms"]
    extra = f"{v['flops'] / v['ms'] / 1e9:.0f} TF/s" if v["flops"] else f"{v['bytes'] / v['ms'] / 1e6:.0f} GB/s"
    print(f"{k:20s} n={v['launches']:3d} {v['ms']:7.3f} ms  {extra}")
print("sum", tot)
