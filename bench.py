#!/usr/bin/env python
"""Headline benchmark: 256x256 tiles/s of the flow-matching hot path on 1/2/4/8 B200 (BASELINE.json `metric`).

    python bench.py --gpus N --steps K --warmup W                 # this repo's engine (train step, configs[1])
    python bench.py --mode sample ...                              # 50-evaluation Euler sampling (configs[2])
    python bench.py --impl reference ...                           # the reference's CPU path (fp32 oracle restatement)

A "step" (train) = FM sample -> UNet forward -> MSE -> backward -> gradient all-reduce (N>1) -> Adam, on one batch of
synthetic 3x256x256 tile pairs, batch 64 per GPU (config A of SURVEY.md).  A "step" (sample) = 50 Euler velocity
evaluations + state updates of one micro-batch of tiles.  One JSON line is printed by rank 0; see the task contract for
the keys.  `value` is timed with inputs resident in HBM; `e2e` includes the pinned-host -> device copy of each step's
inputs and the device -> host read of its result through the public LitModule API.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

if "reference" in sys.argv and "--impl" in sys.argv:
    # the CPU arm may use every host core: torch.distributed.run exports OMP_NUM_THREADS=1 to its ranks, and the OpenMP /
    # MKL runtimes size their pools from the environment when torch is imported -- so undo it BEFORE the import
    for _v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

FLOP_FWD_PER_TILE = 814.4e9        # 407.21 GMAC per 256^2 tile (config A)
FLOP_TRAIN_PER_TILE = 2443e9
L2_BYTES = 126 * 1024 * 1024


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tf=d["bf16_tflops_sustained"], tf_burst=d["bf16_tflops"], src="measured")
    return dict(hbm=6650.0, tf=1400.0, tf_burst=1590.0, src="fallback")


def _conv_traffic():
    """DRAM bytes per conv launch from the committed `ncu --set full` capture (profiles/), next to the algorithmic bytes
    of the same launches -- `roofline.traffic`.  None if the capture is not in the tree."""
    for name in ("r02_conv_traffic.json", "r01_conv_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            with open(p) as f:
                d = json.load(f)
            return d["traffic_per_launch"], {
                "algorithmic_bytes_per_launch": d["algorithmic_bytes_per_launch"], "ratio": d["ratio"],
                "source": "profiles/%s: %s -- a committed ncu --set full capture, NOT a measurement of this run"
                          % (name, d.get("what", "ncu --set full, conv launches"))}
    return None, None


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:  # noqa: BLE001
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def _dist_setup(n_gpus: int):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        torch.cuda.set_device(local)
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=180))
    return rank, world, local


def _max_over_ranks(ms: float, world: int, dev) -> float:
    if world == 1:
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _barrier(world):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def build_lit(device, dropout=0.1, class_cond=False):
    """config A (configs/model/conditional_flow_matching.yaml) or, class_cond, config B = the any2any model with a stain-class
    embedding in every block (configs/model/class_conditional_flow_matching.yaml: class_cond true, num_classes 3)."""
    import functools
    from stain2stain_b200.flow_matching import ConditionalFlowMatcher
    from stain2stain_b200.lit import ClassConditionalFlowMatchingLitModule, ConditionalFlowMatchingLitModule
    from stain2stain_b200.neural_ode import NeuralODE
    from stain2stain_b200.optim import FusedAdam
    from stain2stain_b200.unet import UNetModel
    torch.manual_seed(1984)
    kw = dict(class_cond=True, num_classes=3) if class_cond else {}
    net = UNetModel(dim=[3, 256, 256], num_channels=128, num_res_blocks=2, attention_resolutions="16,8",
                    dropout=dropout, use_scale_shift_norm=True, num_heads=4, num_head_channels=32,
                    channel_mult=[1, 2, 2, 4], **kw)
    _dezero(net)
    cls = ClassConditionalFlowMatchingLitModule if class_cond else ConditionalFlowMatchingLitModule
    lit = cls(net=net, flow_matcher=ConditionalFlowMatcher(sigma=0.0),
              solver=functools.partial(NeuralODE, solver="euler", sensitivity="adjoint", atol=1e-4, rtol=1e-4),
              optimizer=functools.partial(FusedAdam, lr=1e-4, weight_decay=0.0), scheduler=None)
    return lit.to(device)


def _dezero(net, seed=1984):
    """Zero-initialised convs re-drawn (SURVEY finding 7) so that every layer does real work with real values."""
    import math
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in net.parameters():
            if p.dim() >= 2 and float(p.abs().max()) == 0.0:
                p.copy_(torch.randn(p.shape, generator=g) / math.sqrt(p[0].numel()))


class _StepModule(torch.nn.Module):
    """What Lightning's DDP strategy does: DDP wraps a module whose forward is the LightningModule's training_step."""

    def __init__(self, lit, y=None):
        super().__init__()
        self.lit, self.y = lit, y

    def forward(self, x0, x1):
        return self.lit.training_step((x0, x1) if self.y is None else (x0, x1, self.y), 0)


def run_train(args, rank, world, local):
    from stain2stain_b200 import kernels as K
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    B = args.batch
    cc = args.mode == "classcond"
    lit = build_lit(dev, class_cond=cc)
    lit.train()
    opt = lit.configure_optimizers()["optimizer"]
    g = torch.Generator(device=dev).manual_seed(1984 + rank)
    # inputs larger than L2 (2 x 50 MB per step at B=64) and a fresh pair per step: no L2 reuse between steps
    n_sets = 4
    x0s = [torch.rand(B, 3, 256, 256, device=dev, generator=g) * 2 - 1 for _ in range(n_sets)]
    x1s = [torch.rand(B, 3, 256, 256, device=dev, generator=g) * 2 - 1 for _ in range(n_sets)]
    host0 = [x.cpu().pin_memory() for x in x0s[:2]]
    host1 = [x.cpu().pin_memory() for x in x1s[:2]]
    # class-conditional (configs[3]): target stain label per tile, y ~ randint(0, 3)
    ylab = torch.randint(0, 3, (B,), device=dev, generator=g) if cc else None
    yhost = ylab.cpu().pin_memory() if cc else None

    def batch_of(x0, x1):
        return (x0, x1, ylab) if cc else (x0, x1)

    def step_eager(x0, x1):  # the same step launched from Python (per-kernel CUDA events need eager launches)
        opt.zero_grad(set_to_none=True)
        loss = lit.training_step(batch_of(x0, x1), 0)
        loss.backward()
        opt.step()
        return loss

    # ---- roofline of the dominant kernel: per-launch CUDA events on the launching stream, one instrumented EAGER step
    # (events cannot be placed inside a graph replay; same kernels, same shapes).  Taken before the graphs are captured so
    # that the eager step's activations are freed again.  Every rank takes the steps, only rank 0 records events.
    prof = {}
    for _ in range(2):
        step_eager(x0s[1], x1s[1])
    torch.cuda.synchronize()
    if rank == 0:
        K.PROFILE = []
    step_eager(x0s[0], x1s[0])
    torch.cuda.synchronize()
    if rank == 0:
        prof = K.profile_summary(K.PROFILE)
        K.PROFILE = None
    opt.zero_grad(set_to_none=True)
    torch.cuda.empty_cache()
    _barrier(world)

    if args.no_graph:
        step_mod = _StepModule(lit, ylab)
        if world > 1:
            # ONE bucket: a single flat fp32 all-reduce (284 MB, < 1 ms over NVSwitch) right after backward.  The default
            # 25 MB buckets put ~12 ncclAllReduce kernels next to persistent 148-CTA tcgen05 kernels whose static tile
            # schedule turns every SM NCCL holds into a second wave (round 1: +9.4 ms per step at 8 GPUs).
            step_mod = torch.nn.parallel.DistributedDataParallel(step_mod, device_ids=[local], gradient_as_bucket_view=True,
                                                                 bucket_cap_mb=args.bucket_mb)

        def step(x0, x1):
            opt.zero_grad(set_to_none=True)
            loss = step_mod(x0, x1)
            loss.backward()
            opt.step()
            return loss
        how = "eager launches" + (f", torch DDP bucket_cap_mb={args.bucket_mb}" if world > 1 else "")
    else:
        # the step as CUDA graphs (stain2stain_b200/graphed.py): forward + backward + flat gradient gather [+ Adam] replayed;
        # at N > 1 one eager all-reduce of the flat 284 MB gradient buffer sits between the two graphs
        from stain2stain_b200.graphed import GraphedTrainStep
        gs = GraphedTrainStep(lit, opt, (B, 3, 256, 256), dev, process_group=dist.group.WORLD if world > 1 else None,
                              label_shape=(B,) if cc else None)

        def step(x0, x1):
            return gs(x0, x1, y=(yhost if not x0.is_cuda else ylab) if cc else None)
        how = "CUDA-graph replay" + (", one flat fp32 gradient all-reduce (NCCL) between the backward and the Adam graph"
                                     if world > 1 else "")

    for i in range(args.warmup):
        step(x0s[i % n_sets], x1s[i % n_sets])
    _barrier(world)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    K.LAUNCHES[0] = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = step(x0s[i % n_sets], x1s[i % n_sets])
    e1.record()
    _barrier(world)
    ms = _max_over_ranks(e0.elapsed_time(e1), world, dev)
    launches = K.LAUNCHES[0]
    clocks = sampler.stop() if rank == 0 else None
    last_loss = float(loss.detach())

    # ---- end to end through the public API: pinned host inputs -> device, loss back to the host, every step
    _barrier(world)
    e0.record()
    for i in range(args.steps):
        if args.no_graph:
            x0 = host0[i % 2].to(dev, non_blocking=True)
            x1 = host1[i % 2].to(dev, non_blocking=True)
        else:  # the graphed step copies pinned host memory straight into its static input buffers
            x0, x1 = host0[i % 2], host1[i % 2]
        loss_host = float(step(x0, x1).detach())  # device -> host read of the step's result
    e1.record()
    _barrier(world)
    ms_e2e = _max_over_ranks(e0.elapsed_time(e1), world, dev)

    tiles = B * world * args.steps
    out = {
        "metric": "256x256 tiles/s, conditional flow-matching train step (FM sample + UNet fwd + MSE + bwd + allreduce + Adam)",
        "value": tiles / (ms / 1e3), "unit": "tiles/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 forward / bf16 backward operands, fp32 accumulate (tcgen05 kind::f16)" if K.ACT == K.FMT_F16
                 else "bf16 operands, fp32 accumulate",
        "data": "synthetic U(-1,1) 3x256x256 tile pairs, random-init (de-zeroed) config-A UNet, seed 1984",
        "config": {"workload": ("configs[3]: stain-class-conditioned any2any flow matching (label embedding in every UNet block), "
                                "256x256 tiles, batch 64/GPU, DDP" if cc else
                                "configs[1]: simple flow-matching UNet training, 256x256 tiles, batch 64/GPU, DDP"),
                   "per_gpu_batch": B, "global_batch": B * world, "params": sum(p.numel() for p in lit.parameters()),
                   "parallelism": f"dp{world}", "dropout": 0.1, "optimizer": "Adam lr 1e-4 (fused, own kernel)",
                   "step_launch": how,
                   "l2": "inputs larger than L2 (2 x %.0f MB per step, 4 rotating sets)" % (B * 3 * 256 * 256 * 4 / 1e6)},
        "loss": last_loss,
        "e2e": {"value": tiles / (ms_e2e / 1e3), "unit": "tiles/s",
                "h2d_bytes_per_step": 2 * B * 3 * 256 * 256 * 4 + (8 * B if cc else 0), "d2h_bytes_per_step": 4, "loss": loss_host},
        "gpu_launches": launches,
    }
    roof = None
    if rank == 0:
        pk = _peaks()
        step_ms = ms / args.steps
        if "conv_igemm" in prof:
            c = prof["conv_igemm"]
            ach = c["flops"] / (c["ms"] / 1e3) / 1e12
            traffic, traffic_note = _conv_traffic()
            roof = {"bound": "tensor", "kernel": "conv_igemm_kernel (fwd + dgrad implicit GEMM)", "achieved": ach,
                    "peak": pk["tf"], "unit": "TFLOP/s", "frac": ach / pk["tf"], "traffic": traffic,
                    "traffic_note": traffic_note,
                    "peak_source": pk["src"] + " bf16_tflops_sustained", "launches_per_step": c["launches"],
                    "kernel_ms_per_step": c["ms"], "share_of_step": c["ms"] / step_ms,
                    "algorithmic_flops_per_step": c["flops"], "executed_flops_per_step": c["exec_flops"],
                    "executed_tflops": c["exec_flops"] / (c["ms"] / 1e3) / 1e12,
                    "note": "achieved = ALGORITHMIC flops (the reference's formulation: 9 taps at full resolution for the "
                            "Upsample convs) / CUDA-event time of the conv launches in one instrumented eager step; the "
                            "phase-decomposed Upsample convs execute 4/9 of their algorithmic MACs (executed_* fields)"}
        out["roofline"] = roof
        kern = {}
        for name, d in prof.items():
            e = {"launches": d["launches"], "ms": round(d["ms"], 3)}
            if d["flops"]:
                e["tflops"] = d["flops"] / (d["ms"] / 1e3) / 1e12
                e["frac_tensor_peak"] = e["tflops"] / pk["tf"]
                if d.get("exec_flops") and abs(d["exec_flops"] - d["flops"]) > 1e-6 * d["flops"]:
                    e["executed_tflops"] = d["exec_flops"] / (d["ms"] / 1e3) / 1e12
            if d["bytes"]:
                e["gbs"] = d["bytes"] / (d["ms"] / 1e3) / 1e9
                e["frac_hbm_peak"] = e["gbs"] / pk["hbm"]
            kern[name] = e
        out["kernels"] = kern
        out["profiled_kernel_ms_per_step"] = round(sum(d["ms"] for d in prof.values()), 3)
        out["peak_mem_gb"] = round(torch.cuda.max_memory_allocated() / 2**30, 2)
        out["model_flops_utilisation"] = (FLOP_TRAIN_PER_TILE * B / (step_ms / 1e3)) / 1e12 / pk["tf"]
        out["clocks"] = clocks
        out["cpu_baseline"] = cpu_baseline(sample_steps=1) if (world == 1 and not args.no_cpu) else None
    # ---- the metric's second half in the same line: 50-evaluation Euler sampling (configs[2]), tiles sharded by rank
    if not args.no_sample and not cc:
        del x0s, x1s, host0, host1
        opt.zero_grad(set_to_none=True)
        torch.cuda.empty_cache()
        rec = sample_record(args, lit, rank, world, local, micro_batch=args.sample_batch)
        if rank == 0:
            out["sample"] = rec
    return out


def sample_record(args, lit, rank, world, local, micro_batch=64):
    """configs[2]: 50-step Euler ODE sampling (`infer_simple_flowmatching`: lit.generate) of synthetic 256x256 tiles,
    the tile index range cut across ranks with `parallel.shard_range` (no collective on the data path).

    One "step" = one micro-batch of tiles taken through 50 velocity evaluations + state updates.  Two passes over this
    rank's shard: first half with the tiles resident in HBM (`value`), second half through the public API with each
    micro-batch copied from pinned host memory and the sampled tiles copied back to the host (`e2e`)."""
    from stain2stain_b200 import kernels as K
    from stain2stain_b200.parallel import shard_range
    dev = torch.device("cuda", local)
    evals = 50
    n_tiles = args.sample_tiles if args.sample_tiles else 512 * world   # 4096 tiles on 8 GPUs = configs[2]
    lo, hi = shard_range(n_tiles, rank, world)
    mine = hi - lo
    mb = min(micro_batch, mine)
    n_mb = max(2, mine // mb)            # micro-batches of this rank (ragged tails are not part of the synthetic workload)
    n_dev, n_e2e = (n_mb + 1) // 2, n_mb // 2
    lit.eval()
    g = torch.Generator(device=dev).manual_seed(1984 + 7919 * rank)
    n_sets = 3  # rotating input sets; one evaluation streams > 1 GB of activations, far beyond L2
    xs = [torch.rand(mb, 3, 256, 256, device=dev, generator=g) * 2 - 1 for _ in range(n_sets)]
    hosts = [x.cpu().pin_memory() for x in xs[:2]]
    out_host = torch.empty((mb, 3, 256, 256), dtype=torch.float32).pin_memory()

    def step(x):
        return lit.generate(x, num_steps=evals + 1)  # 51 grid points = 50 Euler evaluations, dt = 1/50

    step(xs[0])  # warm-up: packs the eval-mode operands, captures the CUDA graph
    _barrier(world)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    K.LAUNCHES[0] = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n_dev):
        out_img = step(xs[i % n_sets])
    e1.record()
    _barrier(world)
    ms = _max_over_ranks(e0.elapsed_time(e1), world, dev)
    launches = K.LAUNCHES[0]
    clocks = sampler.stop() if rank == 0 else None
    finite = bool(torch.isfinite(out_img).all())
    _barrier(world)
    e0.record()
    for i in range(n_e2e):
        out_host.copy_(step(hosts[i % 2].to(dev, non_blocking=True)), non_blocking=True)
    e1.record()
    _barrier(world)
    ms_e2e = _max_over_ranks(e0.elapsed_time(e1), world, dev)
    pk = _peaks()
    step_ms = ms / n_dev
    ach = FLOP_FWD_PER_TILE * evals * mb / (step_ms / 1e3) / 1e12
    return {
        "metric": "256x256 tiles/s, 50-evaluation Euler ODE sampling (CUDA-graph fused velocity eval + update)",
        "value": mb * n_dev * world / (ms / 1e3), "unit": "tiles/s", "n_gpus": world, "steps": n_dev, "warmup": 1,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 operands, fp32 accumulate, fp32 ODE state" if K.ACT == K.FMT_F16 else "bf16 operands, fp32 state",
        "data": "synthetic U(-1,1) 3x256x256 tiles, random-init (de-zeroed) config-A UNet, seed 1984",
        "config": {"workload": "configs[2]: 50-step Euler ODE sampling (generate(num_steps=51)) of %d synthetic 256x256 "
                               "tiles sharded across %d GPU(s) by parallel.shard_range; rank 0 holds tiles [%d, %d)"
                               % (n_tiles, world, lo, hi),
                   "tiles_total": n_tiles, "tiles_per_gpu": mine, "micro_batch": mb, "evaluations": evals,
                   "micro_batches_timed_resident": n_dev, "micro_batches_timed_e2e": n_e2e,
                   "parallelism": f"tile shards x{world} (no collective)",
                   "l2": "activations of one evaluation (>1 GB) exceed L2; 3 rotating input sets"},
        "e2e": {"value": mb * n_e2e * world / (ms_e2e / 1e3), "unit": "tiles/s", "steps": n_e2e,
                "h2d_bytes_per_step": mb * 3 * 256 * 256 * 4, "d2h_bytes_per_step": mb * 3 * 256 * 256 * 4},
        "gpu_launches": launches,
        "gpu_launches_note": "own kernels inside the CUDA-graph replays (%d replays per step)" % evals,
        "roofline": {"bound": "tensor", "kernel": "whole velocity evaluation (conv_igemm dominated)", "achieved": ach,
                     "peak": pk["tf"], "unit": "TFLOP/s", "frac": ach / pk["tf"], "traffic": None,
                     "peak_source": pk["src"] + " bf16_tflops_sustained",
                     "algorithmic_flops_per_step": FLOP_FWD_PER_TILE * evals * mb},
        "finite": finite,
        "clocks": clocks,
    }


def run_sample(args, rank, world, local):
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    lit = build_lit(dev)
    rec = sample_record(args, lit, rank, world, local, micro_batch=args.batch)
    rec["cpu_baseline"] = None
    return rec


def build_multitask_lit(dev):
    """configs/model/conditional_flow_matching_multitask_multiclass.yaml: features [64..1024], 5 classes, Adam 1e-4 / wd 1e-5."""
    import functools
    from stain2stain_b200 import multitask as mt
    from stain2stain_b200.flow_matching import ConditionalFlowMatcher
    from stain2stain_b200.neural_ode import NeuralODE
    from stain2stain_b200.optim import FusedAdam
    torch.manual_seed(1984)
    f = [64, 128, 256, 512, 1024]
    return mt.MultiTaskFlowMatchingLitModule(
        mt.SharedEncoder(3, f, True), mt.FlowMatchingDecoder(1024, f[:-1][::-1], 3, 256, True),
        mt.SegmentationDecoder(1024, f[:-1][::-1], 5, True), ConditionalFlowMatcher(0.0), num_classes=5,
        solver=functools.partial(NeuralODE, solver="euler"), optimizer=functools.partial(FusedAdam, lr=1e-4, weight_decay=1e-5),
        scheduler=None, log_images=False).to(dev)


def run_multitask(args, rank, world, local):
    """configs[4]: multitask multiclass model (shared encoder + flow / segmentation decoders), 512x512 tiles, train step
    = FM sample -> encoder+flow decoder -> MSE; second encoder pass -> segmentation decoder -> Dice+CE; backward; Adam."""
    import functools
    from stain2stain_b200 import kernels as K
    from stain2stain_b200 import multitask as mt
    from stain2stain_b200.flow_matching import ConditionalFlowMatcher
    from stain2stain_b200.neural_ode import NeuralODE
    from stain2stain_b200.optim import FusedAdam
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    B, S = args.batch, 512
    lit = build_multitask_lit(dev)
    if args.sync_bn and world > 1:  # what Lightning does for `sync_batchnorm: True` (configs/trainer/ddp.yaml:9)
        lit = torch.nn.SyncBatchNorm.convert_sync_batchnorm(lit)
    lit.train()

    class Step(torch.nn.Module):
        def __init__(self, lit):
            super().__init__()
            self.lit = lit

        def forward(self, x0, x1, m):
            return self.lit.training_step((x0, x1, m), 0)
    opt = lit.configure_optimizers()["optimizer"]
    g = torch.Generator(device=dev).manual_seed(1984 + rank)
    sets = [(torch.rand(B, 3, S, S, device=dev, generator=g) * 2 - 1, torch.rand(B, 3, S, S, device=dev, generator=g) * 2 - 1,
             torch.randint(0, 5, (B, 1, S, S), device=dev, generator=g).float()) for _ in range(2)]
    hosts = [tuple(t.cpu().pin_memory() for t in s) for s in sets]

    def step_eager(x0, x1, m):
        opt.zero_grad(set_to_none=True)
        loss = lit.training_step((x0, x1, m), 0)
        loss.backward()
        opt.step()
        return loss
    # instrumented eager step first (per-kernel CUDA events), then the graphs are captured
    prof = {}
    for _ in range(2):
        step_eager(*sets[1])
    torch.cuda.synchronize()
    if rank == 0:
        K.PROFILE = []
    step_eager(*sets[0])
    torch.cuda.synchronize()
    if rank == 0:
        prof = K.profile_summary(K.PROFILE)
        K.PROFILE = None
    opt.zero_grad(set_to_none=True)
    torch.cuda.empty_cache()
    _barrier(world)
    if args.no_graph:
        step_mod = Step(lit)
        if world > 1:
            step_mod = torch.nn.parallel.DistributedDataParallel(step_mod, device_ids=[local], gradient_as_bucket_view=True,
                                                                 bucket_cap_mb=args.bucket_mb)

        def step(x0, x1, m):
            opt.zero_grad(set_to_none=True)
            loss = step_mod(x0, x1, m)
            loss.backward()
            opt.step()
            return loss
    else:
        from stain2stain_b200.graphed import GraphedTrainStep
        gs = GraphedTrainStep(lit, opt, (B, 3, S, S), dev, process_group=dist.group.WORLD if world > 1 else None,
                              extra=[((B, 1, S, S), torch.float32)])

        def step(x0, x1, m):
            return gs(x0, x1, extra=[m])
    for i in range(args.warmup):
        step(*sets[i % 2])
    _barrier(world)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    K.LAUNCHES[0] = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = step(*sets[i % 2])
    e1.record()
    _barrier(world)
    ms = _max_over_ranks(e0.elapsed_time(e1), world, dev)
    launches = K.LAUNCHES[0]
    clocks = sampler.stop() if rank == 0 else None
    _barrier(world)
    e0.record()
    for i in range(args.steps):
        if args.no_graph:
            loss_host = float(step(*(t.to(dev, non_blocking=True) for t in hosts[i % 2])).detach())
        else:  # the graphed step copies pinned host memory straight into its static buffers
            loss_host = float(step(*hosts[i % 2]).detach())
    e1.record()
    _barrier(world)
    ms_e2e = _max_over_ranks(e0.elapsed_time(e1), world, dev)
    tiles = B * world * args.steps
    pk = _peaks()
    step_ms = ms / args.steps
    flop_tile = 2674e9  # config M at 512^2, fwd_flow + fwd_seg + backward
    out = {"metric": "512x512 tiles/s, multitask (flow + segmentation) train step", "value": tiles / (ms / 1e3),
           "unit": "tiles/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f16 forward / bf16 backward operands, fp32 accumulate",
           "data": "synthetic U(-1,1) 3x512x512 tile pairs + 5-class masks, default-init config-M model, seed 1984",
           "config": {"workload": "configs[4]: multi-task multi-class-loss model at 512x512 tiles", "per_gpu_batch": B,
                      "global_batch": B * world, "params": sum(p.numel() for p in lit.parameters()), "parallelism": f"dp{world}",
                      "batchnorm": "SyncBatchNorm (statistics over all ranks: 2 x 26 all-reduces of [C, 2] floats per step)"
                                   if (args.sync_bn and world > 1) else "per-rank batch statistics (plain strategy: ddp)",
                      "step_launch": "eager launches, torch DDP" if args.no_graph else
                                     "CUDA-graph replay (+ one flat gradient all-reduce between two graphs at N > 1)"},
           "loss": float(loss.detach()), "gpu_launches": launches,
           "e2e": {"value": tiles / (ms_e2e / 1e3), "unit": "tiles/s", "loss": loss_host,
                   "h2d_bytes_per_step": B * S * S * 4 * 7, "d2h_bytes_per_step": 4},
           "model_flops_utilisation": (flop_tile * B / (step_ms / 1e3)) / 1e12 / pk["tf"], "clocks": clocks,
           "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 2)}
    kern = {}
    for name, d in prof.items():
        e = {"launches": d["launches"], "ms": round(d["ms"], 3)}
        if d["flops"]:
            e["tflops"] = d["flops"] / (d["ms"] / 1e3) / 1e12
            e["frac_tensor_peak"] = e["tflops"] / pk["tf"]
        if d["bytes"]:
            e["gbs"] = d["bytes"] / (d["ms"] / 1e3) / 1e9
            e["frac_hbm_peak"] = e["gbs"] / pk["hbm"]
        kern[name] = e
    out["kernels"] = kern
    c = prof.get("conv_igemm")
    out["roofline"] = None if not c else {
        "bound": "tensor", "kernel": "conv_igemm (fwd + dgrad)", "achieved": c["flops"] / (c["ms"] / 1e3) / 1e12,
        "peak": pk["tf"], "unit": "TFLOP/s", "frac": c["flops"] / (c["ms"] / 1e3) / 1e12 / pk["tf"], "traffic": None}
    return out


# --------------------------------------------------------------------------------------------------- CPU reference arm
def _oracle_step_fn(batch=4):
    """configs[0]: `src/train.py trainer=cpu`-equivalent step on the fp32 oracle (the reference's packages are absent)."""
    from oracle import flow as oflow
    from oracle import unet as ounet
    torch.manual_seed(1984)
    net = ounet.UNetModel(**ounet.CONFIG_A)
    ounet.dezero_(net)
    net.train()
    fm = oflow.ConditionalFlowMatcher(0.0)
    opt = torch.optim.Adam(net.parameters(), lr=1e-4)
    g = torch.Generator().manual_seed(1984)
    x0 = torch.rand(batch, 3, 256, 256, generator=g) * 2 - 1
    x1 = torch.rand(batch, 3, 256, 256, generator=g) * 2 - 1

    def step():
        opt.zero_grad(set_to_none=True)
        loss = oflow.model_step(net, fm, (x0, x1))
        loss.backward()
        opt.step()
        return float(loss.detach())
    return step


def cpu_baseline(sample_steps=1, batch=4):
    _use_all_host_threads()
    step = _oracle_step_fn(batch)
    t0 = time.perf_counter()
    for _ in range(sample_steps):
        step()
    dt = time.perf_counter() - t0
    return {"value": batch * sample_steps / dt, "unit": "tiles/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{sample_steps} un-warmed train step(s) of configs[0] (config-A UNet, batch {batch}, fp32, "
                      f"fwd+bwd+Adam) on the host CPU, {os.cpu_count()} logical cores"}


def _use_all_host_threads():
    """torch.distributed.run exports OMP_NUM_THREADS=1 to every rank; the CPU arm is allowed all the host's cores."""
    n = os.cpu_count() or 1
    torch.set_num_threads(n)
    return torch.get_num_threads()


def run_reference(args, rank, world):
    """CPU arm: EXACTLY `--warmup` untimed and `--steps` timed steps of the reference's CPU path (fp32 oracle port), each step a
    bounded sample of the workload: a batch of 4 tiles (configs[0]) when the run then stays within ~2.5 minutes, else 2 or 1
    (the per-tile cost of the CPU path does not depend on the batch: 0.097 / 0.095 / 0.100 tiles/s at 1 / 2 / 4 on 8 threads)."""
    if rank != 0:
        return None
    _use_all_host_threads()
    per_tile_s, budget_s = 2.3, 150.0  # measured on the GPU boxes' 16 host threads: 8.9 s per batch-4 step
    batch = max(1, min(4, int(budget_s / (max(1, args.steps + args.warmup) * per_tile_s))))
    step = _oracle_step_fn(batch)
    for _ in range(args.warmup):
        step()
    steps = max(1, args.steps)
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    v = batch * steps / dt
    sample = (f"{steps} train steps of batch {batch} (configs[0]'s model and step; its batch is 4) after {args.warmup} warm-up, "
              f"fp32 oracle restatement of the torchcfm UNet (torchcfm/lightning are not installable here), "
              f"{torch.get_num_threads()} threads of {os.cpu_count()} logical cores")
    return {"impl": "reference",
            "metric": "256x256 tiles/s, conditional flow-matching train step (FM sample + UNet fwd + MSE + bwd + allreduce + Adam)",
            "value": v, "unit": "tiles/s", "n_gpus": world, "steps": steps, "warmup": args.warmup,
            "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic U(-1,1) 3x256x256 tile pairs, seed 1984",
            "config": {"workload": f"configs[1] model on the reference's CPU path (fp32 oracle port), bounded sample: batch {batch} "
                                   "per step on rank 0's host cores whatever --gpus is (the CPU path does not use the GPUs; "
                                   "tiles/s of this arm does not depend on N)",
                       "per_step_batch": batch, "threads": torch.get_num_threads()},
            "cpu_baseline": {"value": v, "unit": "tiles/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": sample},
            "e2e": {"value": v, "unit": "tiles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def _claim_stdout():
    """Only the JSON line may reach stdout: libraries (NCCL prints its version banner there) are sent to stderr."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    out_stream = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="train", choices=["train", "sample", "multitask", "classcond"])
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (train: 64) / micro-batch (sample: 64; 4096 tiles on 8 GPUs = 8 micro-batches per GPU)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-sample", action="store_true", help="train mode: skip the 50-step sampling record")
    ap.add_argument("--sample-tiles", type=int, default=0,
                    help="tiles of the sampling record over ALL ranks (default 512 per GPU: 4096 on 8 GPUs = configs[2])")
    ap.add_argument("--sample-batch", type=int, default=64, help="sampling micro-batch inside the train-mode run")
    ap.add_argument("--sync-bn", action="store_true", help="multitask mode, N > 1: SyncBatchNorm as under the reference's "
                    "trainer=ddp (configs/trainer/ddp.yaml:9); default = per-rank statistics as under its experiment configs")
    ap.add_argument("--no-graph", action="store_true", help="train mode: launch the step from Python (torch DDP at N > 1) "
                                                           "instead of replaying CUDA graphs")
    ap.add_argument("--bucket-mb", type=int, default=512,
                    help="DDP gradient bucket size; 512 = ONE flat all-reduce of the 284 MB of fp32 gradients")
    args = ap.parse_args()
    if args.batch is None:
        args.batch = {"train": 64, "sample": 64, "multitask": 16, "classcond": 64}[args.mode]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        out = run_reference(args, rank, world)
        if out is not None:
            print(json.dumps(out), file=out_stream, flush=True)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the engine has no CPU fallback); use --impl reference for the CPU arm")
    rank, world, local = _dist_setup(args.gpus)
    out = {"train": run_train, "sample": run_sample, "multitask": run_multitask, "classcond": run_train}[args.mode](
        args, rank, world, local)
    if rank == 0:
        print(json.dumps(out), file=out_stream, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
