"""The import shims (shims/README.md): `torchdyn.core.NeuralODE`, `torchcfm.conditional_flow_matching.ConditionalFlowMatcher`,
`torchcfm.models.unet.UNetModel` and `torchcfm.models.unet.unet.UNetModel` resolve to the B200 implementations, so the
reference's LitModule files and Hydra yamls run unedited."""
import os
import subprocess
import sys
import textwrap

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIMS = os.path.join(ROOT, "shims")


def _run(code: str):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([SHIMS, ROOT]))
    r = subprocess.run([sys.executable, "-c", textwrap.dedent(code)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def test_shim_imports_resolve_to_the_engine():
    out = _run("""
        import torchdyn, torchcfm
        from torchdyn.core import NeuralODE
        from torchcfm.conditional_flow_matching import ConditionalFlowMatcher
        from torchcfm.models.unet import UNetModel
        from torchcfm.models.unet.unet import UNetModel as Raw
        import stain2stain_b200.neural_ode as n, stain2stain_b200.flow_matching as f, stain2stain_b200.unet as u
        assert NeuralODE is n.NeuralODE and ConditionalFlowMatcher is f.ConditionalFlowMatcher
        assert UNetModel is u.UNetModel and Raw is u.RawUNetModel
        print("ok")
    """)
    assert out.strip().endswith("ok")


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/models"), reason="the reference tree only exists in the build container")
def test_reference_litmodule_file_runs_unedited_over_the_shims():
    """Imports /root/reference/src/models/conditional_flow_matching.py AS IS (only `lightning` / `wandb`, absent from the
    image, are stubbed), instantiates it the way its yaml does, and checks that the module-level NeuralODE it will call in
    generate() and the net it holds are the B200 ones.  On host tensors the engine refuses to run (no CPU fallback) --
    which proves the call reached it."""
    out = _run("""
        import sys, functools, torch
        import torchdyn.core, torchcfm.conditional_flow_matching, torchcfm.models.unet   # the shims, before any stub
        from oracle import ref_bridge
        m = ref_bridge.reference_module("src.models.conditional_flow_matching")
        import stain2stain_b200.neural_ode as n, stain2stain_b200.unet as u, stain2stain_b200.flow_matching as f
        assert m.NeuralODE is n.NeuralODE and m.ConditionalFlowMatcher is f.ConditionalFlowMatcher
        assert m.__file__.startswith("/root/reference/")
        from torchcfm.models.unet import UNetModel
        net = UNetModel(dim=[3, 64, 64], num_channels=64, num_res_blocks=1, attention_resolutions="16,8", dropout=0.1,
                        use_scale_shift_norm=True, num_heads=4, num_head_channels=32, channel_mult=[1, 2, 2, 4])
        lit = m.ConditionalFlowMatchingLitModule(
            net=net, flow_matcher=m.ConditionalFlowMatcher(sigma=0.0),
            solver=functools.partial(m.NeuralODE, solver="dopri5", sensitivity="adjoint", atol=1e-4, rtol=1e-4),
            optimizer=functools.partial(torch.optim.Adam, lr=1e-4), scheduler=None)
        assert isinstance(lit.net, u.RawUNetModel)
        x = torch.zeros(1, 3, 64, 64)
        for call in (lambda: lit.model_step((x, x)), lambda: lit.generate(x, num_steps=2)):
            try:
                call()
            except RuntimeError as e:
                assert "CUDA" in str(e), e
            else:
                raise AssertionError("the engine ran on host tensors")
        print("ok")
    """)
    assert out.strip().endswith("ok")


@pytest.mark.gpu
def test_reference_call_sequence_through_the_shims_on_gpu():
    """What src/models/conditional_flow_matching.py:53-74 and :133-170 execute, spelled with the reference's imports (the
    reference tree is not on the GPU box), against the fp32 oracle."""
    out = _run("""
        import torch
        from torchcfm.conditional_flow_matching import ConditionalFlowMatcher
        from torchcfm.models.unet import UNetModel
        from torchdyn.core import NeuralODE
        from oracle import unet as ounet, flow as oflow
        torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
        cfg = dict(dim=[3, 64, 64], num_channels=64, num_res_blocks=1, attention_resolutions="16,8", dropout=0.0,
                   use_scale_shift_norm=True, num_heads=4, num_head_channels=32, channel_mult=[1, 2, 2, 4])
        torch.manual_seed(0)
        ref = ounet.UNetModel(**cfg); ounet.dezero_(ref)
        net = UNetModel(**cfg); net.load_state_dict(ref.state_dict(), strict=True)
        ref, net = ref.cuda().eval(), net.cuda().eval()
        g = torch.Generator(device="cuda").manual_seed(1)
        x0 = torch.rand(2, 3, 64, 64, device="cuda", generator=g) * 2 - 1
        x1 = torch.rand(2, 3, 64, 64, device="cuda", generator=g) * 2 - 1
        fm = ConditionalFlowMatcher(sigma=0.0)
        torch.manual_seed(5)
        t, xt, ut = fm.sample_location_and_conditional_flow(x0, x1)            # :66
        loss = torch.mean((net(t, xt) - ut) ** 2)                                # :69-72
        loss_ref = torch.mean((ref(t, xt) - ut) ** 2)
        assert abs(float(loss) - float(loss_ref)) <= 1e-2 * float(loss_ref)
        with torch.no_grad():
            node = NeuralODE(net, solver="dopri5", sensitivity="adjoint", atol=1e-4, rtol=1e-4)      # :158-163
            traj = node.trajectory(x0, t_span=torch.linspace(0, 1, 2, device="cuda"))              # :166-167
            want = oflow.generate(ref, x0, num_steps=2, solver="dopri5")
        assert traj.shape[0] == 2 and oflow.psnr(traj[-1], want) >= 35.0
        print("ok")
    """)
    assert out.strip().endswith("ok")
