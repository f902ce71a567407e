"""One training step under torch.profiler: GPU time per kernel NAME, grouped into own kernels / torch glue / libraries.
Usage: python scripts/prof_step.py [--batch 64] [--mode train|sample]   (diagnostic; numbers under a profiler are not bench values)"""
import argparse
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--top", type=int, default=45)
ap.add_argument("--mode", default="train", choices=["train", "multitask"])
a = ap.parse_args()
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
if a.mode == "multitask":
    lit = bench.build_multitask_lit(dev)
    S = 512
    batch = (torch.rand(a.batch, 3, S, S, device=dev, generator=g) * 2 - 1, torch.rand(a.batch, 3, S, S, device=dev, generator=g) * 2 - 1,
             torch.randint(0, 5, (a.batch, 1, S, S), device=dev, generator=g).float())
else:
    lit = bench.build_lit(dev)
    batch = (torch.rand(a.batch, 3, 256, 256, device=dev, generator=g) * 2 - 1,
             torch.rand(a.batch, 3, 256, 256, device=dev, generator=g) * 2 - 1)
lit.train()
opt = lit.configure_optimizers()["optimizer"]


def step():
    opt.zero_grad(set_to_none=True)
    loss = lit.training_step(batch, 0)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
tot = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        n = ev.name
        tot[n][0] += 1
        tot[n][1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
groups = collections.defaultdict(lambda: [0, 0.0])
for n, (c, us) in tot.items():
    if "s2s::" in n:
        grp = "own"
    elif n.startswith("void at::") or "at::native" in n or "elementwise" in n:
        grp = "torch glue"
    elif "Memcpy" in n or "Memset" in n:
        grp = "memcpy/memset"
    else:
        grp = "library"
    groups[grp][0] += c
    groups[grp][1] += us
allus = sum(v[1] for v in tot.values())
print(f"total GPU kernel time {allus / 1e3:.2f} ms over {sum(v[0] for v in tot.values())} launches")
for grp, (c, us) in sorted(groups.items(), key=lambda kv: -kv[1][1]):
    print(f"  {grp:14s} launches {c:5d}  {us / 1e3:8.3f} ms")
print("top kernels that are NOT the big own kernels:")
k = 0
for n, (c, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"  {us / 1e3:8.3f} ms  n={c:4d}  {n[:150]}")
    k += 1
    if k >= a.top:
        break
