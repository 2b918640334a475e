"""world_size-2 test of the sharded schedule's host logic on CPU (gloo): shard slicing, the
normaliser all-reduce, the integer all-reduce of the exact per-shard totals, and that combining
them reproduces the single-process loss.  The per-shard totals are produced here by the oracle (the
CUDA kernels that produce them on a GPU are covered by tests/test_gpu_parity.py and, across real
GPUs, by tests/test_gpu_multi.py)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def oracle_totals(output, batch, periodic):
    """this shard's exact totals in the layout of include/cnhead.h: int64[24], word q = hi and
    word 12+q = lo of the 2^-40 fixed-point sum (lo alone for the integer counts)."""
    import math
    import oracle
    q = [0.0] * 12
    ints = {}
    prob = oracle.sigmoid_clamp(output["hm"].double())
    pos, neg, npos = oracle.focal_terms(prob, batch["hm"].double())
    q[0], ints[1] = float(pos + neg), int(npos)
    for h, key in enumerate(("wh", "reg")):
        fmap = output[key].double()
        D = fmap.shape[1]
        pred = oracle.gather_rows(fmap, batch["ind"])
        m = batch["reg_mask"].unsqueeze(2).expand_as(pred).double()
        pred, tgt = pred * m, batch[key].double() * m
        ints[4 + 3 * h] = int(m.sum())
        if D == 3:
            q[2 + 3 * h] = float((pred[..., :2] - tgt[..., :2]).abs().sum())
            if periodic:
                pa = oracle.sigmoid_clamp(pred[..., 2:3]) * 2 * math.pi - math.pi
                ta = torch.deg2rad(tgt[..., 2:3])
                q[3 + 3 * h] = float((torch.remainder(pa - ta - math.pi / 2, math.pi) - math.pi / 2).abs().sum())
            else:
                q[3 + 3 * h] = float((oracle.sigmoid_clamp(pred[..., 2:3]) - oracle.sigmoid_clamp(tgt[..., 2:3])).abs().sum())
        else:
            q[2 + 3 * h] = float((pred - tgt).abs().sum())
    tot = torch.zeros(24, dtype=torch.int64)
    for i, v in enumerate(q):
        f = int(round(v * 2.0 ** 40))
        tot[i], tot[12 + i] = f >> 32, f & 0xffffffff
    for i, v in ints.items():
        tot[12 + i] = v
    return tot


def scalars_from_totals(tot, kw, D_wh):
    """what cnh_detloss_finalize computes (csrc/detloss.cu scalars_from_totals), in float64."""
    val = lambda i: (int(tot[i]) * 2 ** 32 + int(tot[12 + i])) / 2.0 ** 40
    cnt = lambda i: int(tot[12 + i])
    hm = (-val(0) if cnt(1) == 0 else -val(0) / cnt(1)) * kw["hm_weight"]
    wh = val(2) / (cnt(4) + 1e-4) * kw["wh_weight"]
    if D_wh == 3:
        wh = wh + val(3) / (cnt(4) + 1e-4) * kw.get("angle_weight", 1.0)
    off = val(5) / (cnt(7) + 1e-4) * kw["off_weight"]
    return hm + wh + off, hm, wh, off


def _worker(rank, world, port, cfg_name, results):
    for p in (PKG, ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import oracle
    from cnhead import sharded, synthetic
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        cfg = synthetic.CONFIGS[cfg_name]
        B = 4
        sl = sharded.shard_slice(B, rank, world)
        per = B // world
        mine = synthetic.make_inputs(cfg, batch=per, sample_offset=sl.start)      # only this rank's samples
        kw = synthetic.loss_kwargs(cfg)
        tot = oracle_totals(mine["output"], mine["batch"], cfg.periodic)
        norm = torch.tensor([float(tot[13]), float(tot[16]), float(tot[19]), 0.0], dtype=torch.float64)
        sharded.exchange_normalisers(norm)
        sharded.reduce_totals(tot)
        total = scalars_from_totals(tot, kw, cfg.wh_channels)
        whole = synthetic.make_inputs(cfg, batch=B)
        ref_loss, ref_stats, _ = oracle.detection_loss({k: v.double() for k, v in whole["output"].items()},
                                                       whole["batch"], **kw)
        assert abs(float(total[0]) - float(ref_loss)) <= 1e-9 * abs(float(ref_loss))
        assert float(norm[0]) == float((whole["batch"]["hm"] == 1).sum()) == float(tot[13])
        assert float(norm[1]) == float(whole["batch"]["reg_mask"].sum()) * cfg.wh_channels
        # every rank holds the same exact totals, hence the same scalars
        both = [torch.zeros_like(tot) for _ in range(world)]
        dist.all_gather(both, tot)
        assert all(torch.equal(both[0], b) for b in both)
        results[rank] = float(total[0])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("cfg_name", ["cfg2", "cfg3"])
def test_sharded_host_logic_world2(cfg_name):
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, cfg_name, results), nprocs=world, join=True)
    assert len(results) == world and results[0] == results[1]


def test_shard_slice():
    from cnhead import sharded
    assert sharded.shard_slice(128, 3, 8) == slice(48, 64)
    with pytest.raises(ValueError):
        sharded.shard_slice(10, 0, 4)


def test_single_process_is_a_no_op():
    from cnhead import sharded
    assert sharded.reduce_totals(torch.arange(24, dtype=torch.int64)) is None
    assert sharded.exchange_normalisers(torch.ones(4, dtype=torch.float64)) is None
