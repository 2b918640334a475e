"""Two real GPUs, one process each (NCCL): the batch-sharded schedule reproduces the single-device
DetectionLoss bit for bit (scalars, heat-map gradients) and the sharded UDA losses sum to the
single-device value.  Skipped unless >= 2 CUDA devices are visible (gpurun --gpus 2)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, cfg_name, exchange, out):
    for p in (PKG, ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    from cnhead import sharded, synthetic, _lib as L
    from losses.centernet import DetectionLoss
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        cfg = synthetic.CONFIGS[cfg_name]
        B = 8                                # cfg5: 4 x 80 x 128^2 per rank = 1280 chunks -> the pre-count schedule
        kw = synthetic.loss_kwargs(cfg)
        whole = synthetic.make_inputs(cfg, batch=B)
        sl = sharded.shard_slice(B, rank, world)
        # sharded: this rank's slice only
        o = {k: v[sl].cuda().requires_grad_(True) for k, v in whole["output"].items()}
        b = {k: v[sl].cuda() for k, v in whole["batch"].items()}
        crit = sharded.make_sharded_loss(DetectionLoss)(exchange=exchange, **kw)
        work = dict(o)
        loss, stats = crit(work, b)
        loss.backward()
        if exchange.startswith("peers"):              # the kernel can be re-launched on the same mailboxes
            for _ in range(3):
                o2 = {k: v[sl].cuda().requires_grad_(True) for k, v in whole["output"].items()}
                l2, _s = crit(dict(o2), b)
                l2.backward()
                assert torch.equal(l2.detach(), loss.detach()) and torch.equal(o2["hm"].grad, o["hm"].grad)
        # single device over the whole batch (every rank does it for itself)
        o1 = {k: v.cuda().requires_grad_(True) for k, v in whole["output"].items()}
        b1 = {k: v.cuda() for k, v in whole["batch"].items()}
        work1 = dict(o1)
        loss1, stats1 = DetectionLoss(**kw)(work1, b1)
        loss1.backward()
        torch.cuda.synchronize()
        assert torch.equal(loss.detach(), loss1.detach()), (float(loss), float(loss1))
        for k in stats1:
            assert torch.equal(stats[k].detach(), stats1[k].detach()), k
        assert torch.equal(work["hm"], work1["hm"][sl])
        assert torch.equal(o["hm"].grad, o1["hm"].grad[sl])              # bit-identical heat-map gradient
        for k in ("wh", "reg"):
            ref = o1[k].grad[sl]
            assert (o[k].grad - ref).abs().max() <= 1e-6 * ref.abs().max().clamp_min(1e-30)
        # UDA loss: value = global loss on every rank, gradient = local part
        x = (torch.randn(B, 6, 32, 32, generator=torch.Generator().manual_seed(3)) * 1.5)
        xs = x[sl].cuda().requires_grad_(True)
        le = sharded.softmax_loss_sharded(xs, L.SOFTMAX_ENTROPY)
        le.backward()
        from cnhead import functional as F
        x1 = x.cuda().requires_grad_(True)
        l1 = F.softmax_loss(x1, L.SOFTMAX_ENTROPY)
        l1.backward()
        assert abs(float(le) - float(l1)) <= 1e-6 * abs(float(l1))
        assert (xs.grad - x1.grad[sl]).abs().max() <= 1e-6 * x1.grad.abs().max()
        out[rank] = float(loss)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("exchange", ["nccl", "peers", "peers_deferred"])
@pytest.mark.parametrize("cfg_name", ["cfg2", "cfg3", "cfg5"])
def test_sharded_loss_two_gpus(cfg_name, exchange):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), cfg_name, exchange, out), nprocs=world, join=True)
    assert len(out) == world and out[0] == out[1]


def _worker_graphed(rank, world, port, exchange, out):
    """cnhead.graphed.HostStep around the SHARDED loss: each rank stages its slice from pinned host memory and replays
    its captured step; loss, probabilities, heat-map gradient and detections equal the single-device eager calls."""
    for p in (PKG, ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    from cnhead import sharded, synthetic
    from cnhead.feeder import HostFeeder
    from cnhead.graphed import HostStep
    from losses.centernet import DetectionLoss
    from backends.decode import decode_detection
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        cfg = synthetic.CONFIGS["cfg2"]
        B = 8
        kw = synthetic.loss_kwargs(cfg)
        wholes = [synthetic.make_inputs(cfg, batch=B, hm_sigma=2.0, seed_offset=60 + i) for i in range(3)]
        sl = sharded.shard_slice(B, rank, world)
        host = HostFeeder.pinned_sets([({k: v[sl].contiguous() for k, v in w["output"].items()},
                                        {k: v[sl].contiguous() for k, v in w["batch"].items()}) for w in wholes])
        crit = sharded.make_sharded_loss(DetectionLoss)(exchange=exchange, **kw)

        def fn(o, b):
            o = {k: v.detach().requires_grad_(True) for k, v in o.items()}
            work = dict(o)
            loss, _ = crit(work, b)
            loss.backward()
            dets = decode_detection(work["hm"], work["wh"].detach(), work["reg"].detach(), K=cfg.K)
            return {"loss": loss.detach().reshape(1), "dets": dets, "grad_hm": o["hm"].grad, "prob": work["hm"]}

        step = HostStep(fn, torch.device("cuda", rank), fetch=("loss", "dets"), depth=2)
        step.stage(*host[0])
        for i in range(6):
            j = i % 3
            res = step.run()
            step.stage(*host[(i + 1) % 3])
            res.wait()
            o1 = {k: v.cuda().requires_grad_(True) for k, v in wholes[j]["output"].items()}
            b1 = {k: v.cuda() for k, v in wholes[j]["batch"].items()}
            work1 = dict(o1)
            loss1, _ = DetectionLoss(**kw)(work1, b1)
            loss1.backward()
            dets1 = decode_detection(work1["hm"], work1["wh"].detach(), work1["reg"].detach(), K=cfg.K)
            torch.cuda.synchronize()
            assert torch.equal(res.host["loss"], loss1.detach().reshape(1).cpu()), (i, float(res.host["loss"]), float(loss1))
            assert torch.equal(res.device["prob"], work1["hm"][sl]), i
            assert torch.equal(res.device["grad_hm"], o1["hm"].grad[sl]), i
            assert torch.equal(res.host["dets"], dets1[sl].cpu()), i
        assert step.graphed.n_graphs == 2
        out[rank] = float(res.host["loss"])
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("exchange", ["peers"])
def test_graphed_host_step_with_sharded_loss_two_gpus(exchange):
    """(exchange='nccl' is NOT supported under HostStep: with the two all-reduces of the NCCL schedule inside the
    captured step both ranks hung in this test -- cnhead/graphed.py says so, bench.py captures the e2e step only when
    the in-kernel exchange is available)"""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_graphed, args=(world, _free_port(), exchange, out), nprocs=world, join=True)
    assert len(out) == world and out[0] == out[1]
