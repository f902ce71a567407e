"""GPU engine vs the vectors the REFERENCE'S OWN LitModule code produced (tests/golden/simple_fm_small.pt,
class_cond_small.pt; generator: oracle/make_golden.py).  Tolerances are the north star's: velocity / loss rel-L2
<= 1e-2, sampled tiles PSNR >= 40 dB.  The multitask vectors are checked in tests/test_gpu_multitask.py."""
import functools
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _inputs(seed, B, H, classes=0):
    g = torch.Generator().manual_seed(seed)
    d = dict(x0=torch.rand(B, 3, H, H, generator=g) * 2 - 1, x1=torch.rand(B, 3, H, H, generator=g) * 2 - 1,
             t=torch.rand(B, generator=g))
    if classes:
        d["y"] = torch.randint(0, classes, (B,), generator=g)
    return d


@pytest.mark.parametrize("fixture,class_cond", [("simple_fm_small.pt", False), ("class_cond_small.pt", True)])
def test_engine_reproduces_reference_litmodule_vectors(fixture, class_cond):
    from oracle import unet as ounet
    from oracle.flow import psnr, rel_l2
    from stain2stain_b200.flow_matching import ConditionalFlowMatcher
    from stain2stain_b200.lit import ClassConditionalFlowMatchingLitModule, ConditionalFlowMatchingLitModule
    from stain2stain_b200.neural_ode import NeuralODE
    from stain2stain_b200.optim import FusedAdam
    from stain2stain_b200.unet import UNetModel
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    gold = torch.load(os.path.join(GOLD, fixture), map_location="cpu", weights_only=False)
    torch.manual_seed(gold["weight_seed"])
    ref = ounet.dezero_(ounet.UNetModel(**gold["config"]), seed=gold["dezero_seed"])
    net = UNetModel(**gold["config"])
    net.load_state_dict(ref.state_dict(), strict=True)
    for k, v in gold["checksums"].items():
        assert abs(float(net.state_dict()[k[4:]].double().sum()) - v) <= 1e-9 * max(1.0, abs(v)), k
    cls = ClassConditionalFlowMatchingLitModule if class_cond else ConditionalFlowMatchingLitModule
    lit = cls(net=net, flow_matcher=ConditionalFlowMatcher(0.0),
              solver=functools.partial(NeuralODE, solver="dopri5", sensitivity="adjoint", atol=1e-4, rtol=1e-4),
              optimizer=functools.partial(FusedAdam, lr=1e-4), scheduler=None, reference_solver_defaults=True).to(DEV)
    lit.eval()
    inp = {k: v.to(DEV) for k, v in _inputs(gold["input_seed"], 2, 64, classes=3 if class_cond else 0).items()}
    with torch.no_grad():
        v = lit(inp["t"], inp["x0"], inp["y"]) if class_cond else lit(inp["t"], inp["x0"])
    r = rel_l2(v.cpu(), gold["forward"])
    assert r <= 1e-2, f"velocity rel-L2 vs the reference vector: {r}"
    ms = gold["model_step"]
    batch = (inp["x0"], inp["x1"], inp["y"]) if class_cond else (inp["x0"], inp["x1"])
    torch.manual_seed(ms["rng_seed"])  # the engine draws t exactly like torchcfm: CPU default generator
    loss = lit.model_step(batch)
    assert abs(float(loss) - float(ms["loss"])) <= 1e-2 * abs(float(ms["loss"])), (float(loss), float(ms["loss"]))
    lit.zero_grad()
    loss.backward()
    num = den = 0.0
    for k, p in lit.named_parameters():
        num += (float(p.grad.double().norm()) - ms["grad_norms"][k]) ** 2
        den += ms["grad_norms"][k] ** 2
    assert (num / den) ** 0.5 <= 2e-2, f"gradient-norm profile vs the reference: {(num / den) ** 0.5}"
    # generate(): literal reference behaviour = adaptive dopri5 @ 1e-4 over linspace(0, 1, 2)
    gen = lit.generate(inp["x0"][:1], 1, num_steps=2) if class_cond else lit.generate(inp["x0"][:1], num_steps=2)
    p = psnr(gen.cpu(), gold["generate_num_steps2"])
    assert p >= 40.0, f"dopri5 sample PSNR vs the reference vector: {p:.1f} dB"
