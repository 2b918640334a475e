"""world_size-2 test of the sharded schedule's host logic on CPU (gloo): shard slicing, the
normaliser all-reduce, the rank-ordered all-gather of per-sample partial rows, and that combining
them reproduces the single-process loss.  The per-sample rows are produced here by the oracle (the
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


def oracle_rows(output, batch, periodic):
    """per-sample partial rows with the layout of include/cnhead.h (float64)."""
    import oracle
    B = output["hm"].shape[0]
    rows = torch.zeros(B, 12, dtype=torch.float64)
    for b in range(B):
        prob = oracle.sigmoid_clamp(output["hm"][b:b + 1].double())
        pos, neg, npos = oracle.focal_terms(prob, batch["hm"][b:b + 1].double())
        rows[b, 0], rows[b, 1] = pos + neg, npos
        for h, (key, tkey) in enumerate((("wh", "wh"), ("reg", "reg"))):
            fmap = output[key][b:b + 1].double()
            D = fmap.shape[1]
            pred = oracle.gather_rows(fmap, batch["ind"][b:b + 1])
            m = batch["reg_mask"][b:b + 1].unsqueeze(2).expand_as(pred).double()
            pred, tgt = pred * m, batch[tkey][b:b + 1].double() * m
            rows[b, 4 + 3 * h] = m.sum()
            if D == 3:
                rows[b, 2 + 3 * h] = (pred[..., :2] - tgt[..., :2]).abs().sum()
                if periodic:
                    import math
                    pa = oracle.sigmoid_clamp(pred[..., 2:3]) * 2 * math.pi - math.pi
                    ta = torch.deg2rad(tgt[..., 2:3])
                    rows[b, 3 + 3 * h] = (torch.remainder(pa - ta - math.pi / 2, math.pi) - math.pi / 2).abs().sum()
                else:
                    rows[b, 3 + 3 * h] = (oracle.sigmoid_clamp(pred[..., 2:3]) - oracle.sigmoid_clamp(tgt[..., 2:3])).abs().sum()
            else:
                rows[b, 2 + 3 * h] = (pred - tgt).abs().sum()
    return rows


def combine_rows(rows, kw, D_wh):
    """what cnh_detloss_finalize computes (csrc/detloss.cu combine_partials), in float64."""
    s = rows.sum(0)
    hm = (-s[0] if s[1] == 0 else -s[0] / s[1]) * kw["hm_weight"]
    wh = s[2] / (s[4] + 1e-4) * kw["wh_weight"]
    if D_wh == 3:
        wh = wh + s[3] / (s[4] + 1e-4) * kw.get("angle_weight", 1.0)
    off = s[5] / (s[7] + 1e-4) * kw["off_weight"]
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
        rows = oracle_rows(mine["output"], mine["batch"], cfg.periodic)
        norm = torch.stack([rows[:, 1].sum(), rows[:, 4].sum(), rows[:, 7].sum(), torch.tensor(0.0, dtype=torch.float64)])
        sharded.exchange_normalisers(norm)
        all_rows = sharded.gather_partials(rows)
        assert all_rows.shape == (B, 12)
        assert torch.equal(all_rows[sl], rows)                                     # rank order == sample order
        total = combine_rows(all_rows, kw, cfg.wh_channels)
        whole = synthetic.make_inputs(cfg, batch=B)
        ref_loss, ref_stats, _ = oracle.detection_loss({k: v.double() for k, v in whole["output"].items()},
                                                       whole["batch"], **kw)
        assert abs(float(total[0]) - float(ref_loss)) <= 1e-9 * abs(float(ref_loss))
        assert float(norm[0]) == float((whole["batch"]["hm"] == 1).sum())
        assert float(norm[1]) == float(whole["batch"]["reg_mask"].sum()) * cfg.wh_channels
        # every rank holds the same scalars
        mine_t = torch.tensor([float(t) for t in total], dtype=torch.float64)
        both = [torch.zeros_like(mine_t) for _ in range(world)]
        dist.all_gather(both, mine_t)
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
    rows = torch.arange(24, dtype=torch.float64).reshape(2, 12)
    assert sharded.gather_partials(rows) is rows
    assert sharded.exchange_normalisers(torch.ones(4, dtype=torch.float64)) is None
