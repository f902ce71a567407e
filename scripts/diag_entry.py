"""GPU diagnostic: bisect a non-finite `generate` of the small entry-point model (batch size, graph, trained weights)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stain2stain_b200 import entry, hydra_lite, neural_ode  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL_NET = dict(dim=[3, 64, 64], num_channels=64, num_res_blocks=1, attention_resolutions="16,8", dropout=0.1,
                 use_scale_shift_norm=True, num_heads=4, num_head_channels=32, channel_mult=[1, 2, 2, 4])
cfg = hydra_lite.load_yaml(os.path.join(ROOT, "configs", "model", "conditional_flow_matching.yaml"))
cfg["net"].update(SMALL_NET)
out = entry.train(cfg, steps=12, batch=4, device="cuda", ckpt_path="/tmp/diag.ckpt")
print("losses", out["losses"])
model = out["model"]
bad = [k for k, v in model.state_dict().items() if not torch.isfinite(v).all()]
print("non-finite params after training:", bad)
model.eval()
net = model.net
torch.manual_seed(0)
for B in (4, 3, 2, 1):
    x = (torch.rand(B, 3, 64, 64) * 2 - 1).cuda()
    t = torch.full((B,), 0.3, device="cuda")
    with torch.no_grad():
        v = net(t, x)
        print(f"B={B} forward finite={bool(torch.isfinite(v).all())} absmax={float(v.abs().max()):.4f}")
        xs = x.clone()
        net.euler_step_(t, xs, 0.2)
        print(f"B={B} euler_step_ finite={bool(torch.isfinite(xs).all())}")
        ts = torch.linspace(0, 1, 6, device="cuda")
        for g in (False, True):
            r = neural_ode.fused_euler(net, x, ts, use_graph=g)
            print(f"B={B} fused_euler graph={g} finite={bool(torch.isfinite(r).all())}")
gen, img = entry.infer_simple(cfg, "/tmp/diag.ckpt", torch.rand(3, 3, 64, 64) * 2 - 1, num_steps=6)
print("infer_simple finite:", bool(torch.isfinite(gen).all()))
fresh = entry.build_model(cfg).cuda().eval()
with torch.no_grad():
    for B in (3, 4):
        x = (torch.rand(B, 3, 64, 64) * 2 - 1).cuda()
        print(f"fresh B={B} generate finite:", bool(torch.isfinite(fresh.generate(x, num_steps=6)).all()))
