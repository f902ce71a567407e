import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import test_gpu_multitask as T
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
feats = [int(a) for a in (sys.argv[1] if len(sys.argv) > 1 else "64,128,256,512").split(",")]
H = int(sys.argv[2]) if len(sys.argv) > 2 else 64
lit, ref = T._engine_and_oracle(feats, 5, 64, 11)
ref = ref.to("cuda")
inp = {k: v.to("cuda") for k, v in T._inputs(41, 4, H, 5).items()}
lit.train(); ref.train()
total, d = lit.model_step((inp["x0"], inp["x1"], inp["mask"]), t=inp["t"])
total_ref, d_ref = ref.model_step((inp["x0"], inp["x1"], inp["mask"]), t=inp["t"])
print({k: (float(d[k]), float(d_ref[k])) for k in d_ref})
total.backward(); total_ref.backward()
rows = []
for (n, p), (_, q) in zip(lit.named_parameters(), ref.named_parameters()):
    e = float((p.grad.double() - q.grad.double()).norm()); s = float(q.grad.double().norm())
    rows.append((e / max(s, 1e-30), e, s, n))
for r in sorted(rows, reverse=True)[:25]:
    print("%.4f  err %.3e  norm %.3e  %s" % r)
num = sum(r[1] ** 2 for r in rows); den = sum(r[2] ** 2 for r in rows)
print("whole", (num / den) ** 0.5)
