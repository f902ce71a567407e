"""SyncBatchNorm of the multitask model (configs/trainer/ddp.yaml:9 `sync_batchnorm: True`): two ranks, each
with half of a batch, must reproduce ONE process normalising the whole batch -- outputs, input gradients, running statistics
-- and hold the LOCAL parameter gradients (their sum over ranks is the full-batch gradient), exactly like torch.nn.SyncBatchNorm.

Both ranks run on cuda:0 over the gloo backend (CUDA tensors are staged through the host by gloo; no kernel of one rank waits
on the other), so the test needs ONE GPU."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from stain2stain_b200 import kernels as K, ops
        dev = torch.device("cuda", 0)
        torch.manual_seed(7)
        B, C, H, W = 4, 64, 32, 32
        x_all = torch.randn(B, C, H, W, device=dev) * 1.7 + 0.3
        g_all = torch.randn(B, C, H, W, device=dev)
        gamma = torch.rand(C, device=dev) + 0.5
        beta = torch.randn(C, device=dev) * 0.1
        # fp32 reference on the whole batch (the 16-bit roundings of x and g applied first)
        xr = K.to_float(K.from_float(x_all.permute(0, 2, 3, 1).contiguous(), K.ACT), K.ACT).permute(0, 3, 1, 2).contiguous()
        gr = K.to_float(K.from_float(g_all.permute(0, 2, 3, 1).contiguous(), K.GRAD), K.GRAD).permute(0, 3, 1, 2).contiguous()
        ref_bn = torch.nn.BatchNorm2d(C).to(dev).train()
        with torch.no_grad():
            ref_bn.weight.copy_(gamma)
            ref_bn.bias.copy_(beta)
        xr.requires_grad_(True)
        torch.relu(ref_bn(xr)).backward(gr)
        # engine: SyncBatchNorm module (what convert_sync_batchnorm makes of the model's BatchNorm2d), this rank's half
        bn = torch.nn.SyncBatchNorm.convert_sync_batchnorm(torch.nn.Sequential(torch.nn.BatchNorm2d(C)))[0].to(dev).train()
        assert isinstance(bn, torch.nn.SyncBatchNorm)
        with torch.no_grad():
            bn.weight.copy_(gamma)
            bn.bias.copy_(beta)
        lo, hi = rank * B // world, (rank + 1) * B // world
        x = K.from_float(x_all[lo:hi].permute(0, 2, 3, 1).contiguous(), K.ACT).requires_grad_(True)
        y = ops.batch_norm_relu(x, bn, relu=True)
        y.backward(K.from_float(g_all[lo:hi].permute(0, 2, 3, 1).contiguous(), K.GRAD))

        def rel(a, b):
            return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))

        y_ref = torch.relu(torch.nn.functional.batch_norm(xr.detach(), None, None, gamma, beta, True, 0.0, ref_bn.eps))[lo:hi]
        res = dict(
            y=rel(K.to_float(y.detach(), K.ACT).permute(0, 3, 1, 2), y_ref),
            dx=rel(K.to_float(x.grad, K.GRAD).permute(0, 3, 1, 2), xr.grad[lo:hi]),
            rm=rel(bn.running_mean, ref_bn.running_mean), rv=rel(bn.running_var, ref_bn.running_var),
            nbt=int(bn.num_batches_tracked))
        # parameter gradients: local on each rank, their sum = the full-batch gradient
        dg, db = bn.weight.grad.clone(), bn.bias.grad.clone()
        dist.all_reduce(dg)
        dist.all_reduce(db)
        res["dgamma"] = rel(dg, ref_bn.weight.grad)
        res["dbeta"] = rel(db, ref_bn.bias.grad)
        # and an un-synchronised BatchNorm2d on the same half batch must NOT match (the test can tell the two apart)
        plain = torch.nn.BatchNorm2d(C).to(dev).train()
        with torch.no_grad():
            plain.weight.copy_(gamma)
            plain.bias.copy_(beta)
        y_plain = ops.batch_norm_relu(x.detach(), plain, relu=True)
        res["plain_differs"] = rel(K.to_float(y_plain, K.ACT).permute(0, 3, 1, 2), y_ref)
        out[rank] = res
    finally:
        dist.destroy_process_group()


def test_sync_batchnorm_two_ranks_match_full_batch():
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0, "worker failed"
    for r in range(2):
        res = out[r]
        assert res["y"] < 2e-3 and res["dx"] < 1.2e-2, res          # fp16 / bf16 storage of y and dx
        assert res["rm"] < 1e-5 and res["rv"] < 1e-4 and res["nbt"] == 1, res
        assert res["dgamma"] < 1e-3 and res["dbeta"] < 1e-3, res
        assert res["plain_differs"] > 5e-3, res  # 25x the synchronised error: the two paths are told apart
