"""Per-launch CUDA-event list of ONE eager training step: every tensor-core launch with its algorithmic TFLOP/s and the time it
loses against a 1500 TFLOP/s kernel -- finds the individual layers that are off the roofline."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from stain2stain_b200 import kernels as K  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda", 0)
lit = bench.build_lit(dev)
lit.train()
opt = lit.configure_optimizers()["optimizer"]
x0 = torch.rand(B, 3, 256, 256, device=dev) * 2 - 1
x1 = torch.rand(B, 3, 256, 256, device=dev) * 2 - 1


def step():
    opt.zero_grad(set_to_none=True)
    loss = lit.training_step((x0, x1), 0)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
K.PROFILE = []
step()
torch.cuda.synchronize()
recs = list(K.PROFILE)
K.PROFILE = None
rows = []
for i, (name, a, b, fl, by, xfl) in enumerate(recs):
    ms = a.elapsed_time(b)
    if fl > 0:
        rows.append((ms - xfl / 1.5e12, i, name, ms, fl / ms / 1e9, xfl / ms / 1e9, fl / 1e9))
tot = sum(r[3] for r in rows)
print(f"{len(rows)} tensor launches, {tot:.2f} ms; lost vs 1500 TFLOP/s (executed flops): {sum(max(r[0], 0) for r in rows):.2f} ms")
print("   idx kernel          ms   alg TF/s  exec TF/s    alg GF   lost ms")
for lost, i, name, ms, tf, xtf, gf in sorted(rows, reverse=True)[:40]:
    print(f"  {i:4d} {name:12s} {ms:7.3f} {tf:9.0f} {xtf:9.0f} {gf:9.1f} {lost:8.3f}")
