"""N > 1 host logic on CPU: 2 ranks over gloo (the reference's own way of testing multi-rank without a cluster is
`ddp_spawn` x2 on CPU, tests/test_train.py:65-77 + configs/trainer/ddp_sim.yaml).

The training step shards by batch with one gradient all-reduce (DDP around the LitModule's training_step, exactly
what Lightning's `strategy: ddp` does); sampling shards tiles with no collective.  The CUDA engine cannot run here,
so the LitModule drives a small torch `net` through its generic (reference-shaped) path: this covers the wrapper
module bench.py uses, the per-rank batch split rule of PairedDataModule.setup (src/data/paired_data_module.py:273-278)
and the max-over-ranks timing reduction."""
import functools
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


class TinyField(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv = nn.Conv2d(3, 3, 3, padding=1)

    def forward(self, t, x, y=None):
        return self.conv(x) * (1 + t.view(-1, 1, 1, 1))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_lit():
    from stain2stain_b200.flow_matching import ConditionalFlowMatcher
    from stain2stain_b200.lit import ConditionalFlowMatchingLitModule
    from stain2stain_b200.neural_ode import NeuralODE
    torch.manual_seed(7)
    return ConditionalFlowMatchingLitModule(
        net=TinyField(), flow_matcher=ConditionalFlowMatcher(0.0),
        solver=functools.partial(NeuralODE, solver="euler"),
        optimizer=functools.partial(torch.optim.Adam, lr=1e-2), scheduler=None)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import bench
        from stain2stain_b200.parallel import per_rank_batch, shard_range
        lit = _make_lit()
        g = torch.Generator().manual_seed(11)
        gb = 8  # global batch
        x0, x1, t = torch.rand(gb, 3, 8, 8, generator=g), torch.rand(gb, 3, 8, 8, generator=g), torch.rand(gb, generator=g)
        pb = per_rank_batch(gb, world)
        lo, hi = rank * pb, (rank + 1) * pb

        class Step(nn.Module):  # bench._StepModule with an explicit t so both layouts see the same draw
            def __init__(self, lit):
                super().__init__()
                self.lit = lit

            def forward(self, a, b, tt):
                return self.lit.model_step((a, b), t=tt)

        ddp = torch.nn.parallel.DistributedDataParallel(Step(lit))
        loss = ddp(x0[lo:hi], x1[lo:hi], t[lo:hi])
        loss.backward()
        grads = [p.grad.clone() for p in lit.parameters()]
        # single-process reference on the global batch: mean over ranks of per-rank means == global mean (equal shards)
        ref = _make_lit()
        ref.model_step((x0, x1), t=t).backward()
        ok = all(torch.allclose(a, p.grad, atol=1e-6) for a, p in zip(grads, ref.parameters()))
        # the graphed step's collective (graphed.py): every gradient gathered into ONE flat buffer, one all-reduce(SUM), the
        # optimizer reads it scaled by 1 / world -- must equal the DDP mean, i.e. the global-batch gradient
        from stain2stain_b200.parallel import FlatGradients
        lit2 = _make_lit()
        lit2.model_step((x0[lo:hi], x1[lo:hi]), t=t[lo:hi]).backward()
        fg = FlatGradients(list(lit2.parameters()))
        fg.gather()
        fg.all_reduce_sum()
        ok = ok and all(torch.allclose(fg.views[p] / world, q.grad, atol=1e-6) for p, q in zip(lit2.parameters(), ref.parameters()))
        ok = ok and fg.flat.numel() == sum(p.numel() for p in lit2.parameters()) and fg.offsets[0] == (0, 81)
        # the wrapper bench.py hands to DDP calls training_step like Lightning does
        sm = bench._StepModule(lit)
        assert sm(x0[lo:hi], x1[lo:hi]).dim() == 0 and "train/loss" in lit.logged
        # sampling shards tiles, no collective: disjoint cover of the tile range
        r = shard_range(4096, rank, world)
        # timing reduction = max over ranks
        tmax = bench._max_over_ranks(10.0 + rank, world, torch.device("cpu"))
        q.put((rank, ok, float(loss), r, tmax))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_gloo_data_parallel_step_and_tile_sharding():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    out = sorted(q.get(timeout=150) for _ in range(world))
    [p.join(30) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert all(o[1] for o in out), "DDP-averaged gradients differ from the global-batch gradients"
    assert out[0][3] == (0, 2048) and out[1][3] == (2048, 4096)
    assert out[0][4] == out[1][4] == 11.0


def test_batch_split_rule_and_tile_partition():
    from stain2stain_b200.parallel import per_rank_batch, shard_range
    assert per_rank_batch(512, 8) == 64
    with pytest.raises(RuntimeError):
        per_rank_batch(10, 4)  # reference raises when the global batch is not divisible by the world size
    for n, w in ((4096, 8), (50, 4), (7, 8), (0, 2)):
        cuts = [shard_range(n, r, w) for r in range(w)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
        assert max(hi - lo for lo, hi in cuts) - min(hi - lo for lo, hi in cuts) <= 1
