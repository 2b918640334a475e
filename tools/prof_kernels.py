"""Per-kernel device timings (CUDA events, graph replay over rotating buffer sets) for the hot-path
launches, per config and schedule.  Usage: python tools/prof_kernels.py [cfg2 cfg5 ...]"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "centernet-uda_b200"), ROOT]
import torch
import bench
from cnhead import _lib as L, synthetic

PEAK = 6547.8


def timed(fn, n_sets, iters=200, graph=True):
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for i in range(n_sets):
            fn(i)
        torch.cuda.synchronize()
        r = bench.GraphRunner(fn, n_sets, side, graph, 0)
        ms = r.timed(iters, 20, 1, torch.device("cuda"))
    return ms / iters * 1e3   # us


def main():
    names = sys.argv[1:] or ["cfg2", "cfg5"]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    out = {}
    for name in names:
        cfg = synthetic.CONFIGS[name]
        batch = cfg.batch if name != "cfg5" else 16
        probe = bench.BufferSet(synthetic.make_inputs(cfg, batch=batch, hm_sigma=2.0), cfg, dev)
        n_sets = 2 if probe.nbytes() > bench.L2_BYTES else max(2, min(16, -(-2 * bench.L2_BYTES // probe.nbytes())))
        sets = [probe] + [bench.BufferSet(synthetic.make_inputs(cfg, batch=batch, hm_sigma=2.0, seed_offset=1 + i), cfg, dev)
                          for i in range(n_sets - 1)]
        d = bench.DeviceStep(sets, cfg, 1, None)
        lib = d.lib
        hw = cfg.height * cfg.width
        loss_bytes = batch * (16 * cfg.classes * hw + 4 * (cfg.wh_channels + 2) * hw)
        dec_bytes = batch * 4 * cfg.classes * hw
        res = {}

        def fused(flags):
            def f(i):
                a = d.plain_args[i]
                a.flags = flags
                a.scalars = sets[i].scalars.data_ptr()
                L.check(lib.cnh_detloss_fused(C.byref(a), d.ws_loss.data_ptr(), d.ws_loss.numel(), L.stream_ptr()), "f")
            return f

        def fwd_only(i):
            a = d.plain_args[i]
            keep = (a.grad_hm, a.heads[0].grad, a.heads[1].grad)
            a.grad_hm, a.heads[0].grad, a.heads[1].grad = None, None, None
            L.check(lib.cnh_detloss_fused(C.byref(a), d.ws_loss.data_ptr(), d.ws_loss.numel(), L.stream_ptr()), "f")
            a.grad_hm, a.heads[0].grad, a.heads[1].grad = keep

        def count(i):
            a = d.plain_args[i]
            L.check(lib.cnh_detloss_count(C.byref(a), d.ws_loss.data_ptr(), d.ws_loss.numel(), L.stream_ptr()), "c")

        def main_(i):
            a = d.plain_args[i]
            a.scalars = None
            L.check(lib.cnh_detloss_main(C.byref(a), d.ws_loss.data_ptr(), d.ws_loss.numel(), L.stream_ptr()), "m")
            a.scalars = sets[i].scalars.data_ptr()

        def scale(i):
            L.check(lib.cnh_scale_inplace(C.byref(d.scale_args[i]), L.stream_ptr()), "s")

        def decode(i):
            L.check(lib.cnh_decode(C.byref(d.dec_args[i]), d.ws_dec.data_ptr(), d.ws_dec.numel(), L.stream_ptr()), "d")

        def copy(i):
            sets[i].grads[0].copy_(sets[i].hm)

        for i in range(n_sets):      # norm must hold valid values for main
            count(i)
        torch.cuda.synchronize()
        uda = []
        if cfg.target_domain:
            N, Cc, H, W, n_total = d.uda_dims
            def soft(j, mode):
                def f(i):
                    s_ = sets[i]
                    L.check(lib.cnh_softmax_loss(s_.tdom.data_ptr(), s_.tgrads[j].data_ptr(), s_.tloss[j:].data_ptr(), N, Cc, H, W,
                                                 n_total, mode, 0.0, d.ws_soft[j].data_ptr(), d.ws_soft[j].numel(), L.stream_ptr()), "s")
                return f
            uda = [("entropy_fwd_bwd", soft(0, L.SOFTMAX_ENTROPY), batch * 8 * cfg.classes * hw),
                   ("max_square_fwd_bwd", soft(1, L.SOFTMAX_MAX_SQUARE), batch * 8 * cfg.classes * hw)]
        for tag, fn, nbytes in (("fused_stash", fused(0), loss_bytes), ("fused_precount", fused(2), loss_bytes),
                                ("fused_stash_accurate", fused(1), loss_bytes), ("fwd_only", fwd_only, loss_bytes * 12 // 16),
                                ("count", count, batch * 4 * cfg.classes * hw), ("main", main_, loss_bytes),
                                ("scale_noop", scale, 0), ("decode", decode, dec_bytes),
                                ("torch_copy_hm", copy, batch * 8 * cfg.classes * hw), *uda,
                                ("loss_emitting_plus_decode_from_candidates", d.loss_decode_pair, loss_bytes + dec_bytes),
                                ("full_step", d.step, batch * cfg.bytes_per_sample())):
            us = timed(fn, n_sets)
            us_eager = timed(fn, n_sets, graph=False) if tag in ("fused_stash", "decode", "full_step") else None
            res[tag] = {"us": round(us, 2), "GBps": round(nbytes / us / 1e3, 1), "frac": round(nbytes / us / 1e3 / PEAK, 3),
                        "us_eager": None if us_eager is None else round(us_eager, 2)}
            print(name, tag, res[tag], flush=True)
        out[name] = res
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "prof_kernels.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
