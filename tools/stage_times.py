"""In-kernel stage timestamps (globaltimer, ns) for the fused loss and the decode launch on cfg2/cfg5."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "centernet-uda_b200"), ROOT]
import torch, bench
bench.DeviceStep.FUSE = "--emit" in sys.argv
from cnhead import _lib as L, synthetic
name = next((x for x in sys.argv[1:] if not x.startswith("-")), "cfg2")
cfg = synthetic.CONFIGS[name]
batch = cfg.batch if name != "cfg5" else 16
dev = torch.device("cuda", 0)
sets = [bench.BufferSet(synthetic.make_inputs(cfg, batch=batch, hm_sigma=2.0, seed_offset=i), cfg, dev) for i in range(4)]
d = bench.DeviceStep(sets, cfg, 1, None)
lib = d.lib
lib.cnh_debug_set_buffer.argtypes = [C.c_void_p]
dbg = torch.zeros(65536, 16, dtype=torch.int64, device=dev)
def show(tag, nslots, extra=()):
    torch.cuda.synchronize()
    t = dbg.cpu()
    used = ((t[:, 0] != 0) | (t[:, 5] != 0)).nonzero().flatten()
    t = t[used]
    t0 = t[:, 0][t[:, 0] != 0].min() if (t[:, 0] != 0).any() else t[:, 5][t[:, 5] != 0].min()
    print(f"--- {tag}: {len(used)} CTAs; kernel span {(t[:, :nslots].max() - t0).item() / 1e3:.2f} us")
    for sl in range(nslots):
        col = t[:, sl]
        ok = col != 0
        if ok.any():
            rel = (col[ok] - t0).float() / 1e3
            print(f"  stamp {sl:2d}: n={int(ok.sum()):4d}  min {rel.min():7.2f}  median {rel.median():7.2f}  max {rel.max():7.2f} us")
    for e in extra:
        col = t[:, e]; ok = t[:, 9] != 0
        if ok.any(): print(f"  value {e}: {col[ok].tolist()[:16]}")
for it in range(3):
    for i in range(4):
        d.step(i)
torch.cuda.synchronize()
lib.cnh_debug_set_buffer(dbg.data_ptr())
for rep in range(2):
    dbg.zero_(); torch.cuda.synchronize()
    d.loss_only(rep)
    show(f"detloss {name} rep{rep}", 8)
    t = dbg.cpu(); used = (t[:, 0] != 0) & (t[:, 1] != 0)
    if name == "cfg5":
        print("  plain consumers (tid 0): process_chunk us", round((t[t[:, 9] > 0, 5].float() / 1965.0).median().item(), 2))
    if name == "cfg5" and "--emit" in sys.argv:
        dbg.zero_(); torch.cuda.synchronize()
        L.check(lib.cnh_detloss_fused(C.byref(d.loss_args[rep]), d.ws_loss.data_ptr(), d.ws_loss.numel(), L.stream_ptr()), "f")
        show(f"detloss+emit {name} rep{rep}", 8)
        torch.cuda.synchronize()
        t = dbg.cpu(); u2 = t[:, 9] > 0
        f = lambda c: (t[u2, c].float() / 1965.0)        # SM cycles -> us at 1965 MHz
        med = lambda x: round(x.float().median().item(), 2)
        print("  EMIT consumers (tid 0): process_chunk", med(f(5)), "boot + pend_free", med(f(6)))
        print("  EMIT (clock64, us at 1965 MHz) medians/max: consumers blocked on full", med(f(8)), "on pend_free", med(f(6)), round(f(6).max().item(), 2),
              "| emitter: wait scanned", med(f(12)), "pending passes", med(f(11)), round(f(11).max().item(), 2), "full scans", med(f(13)), round(f(13).max().item(), 2),
              "| passes", med(t[u2, 15] & 0xffffffff), "full scans", med(t[u2, 15] >> 32), (t[u2, 15] >> 32).max().item(),
              "| keys kept per CTA median/max", med(t[u2, 14] & 0xffffffff), (t[u2, 14] & 0xffffffff).max().item())
        dbg.zero_(); torch.cuda.synchronize()
        d.decode_step(rep); torch.cuda.synchronize()
        show(f"finish-from-candidates {name} rep{rep}", 12)
        tt = dbg.cpu(); uu = tt[:, 13] > 0; print("  survivors m / selected:", tt[uu, 12].tolist()[:4], tt[uu, 13].tolist()[:4], "rank loop cycles", tt[uu, 14].tolist()[:6], "sync", (tt[uu, 15] >> 32).tolist()[:6], "pv wait + store", (tt[uu, 15] & 0xffffffff).tolist()[:6])
        t = dbg.cpu()
    if name == "cfg5":
        u2 = t[:, 9] > 0
        print("  consumer blocked on full (us) median/max:", (t[u2, 8].float() / 1e3).median().item(), (t[u2, 8].float() / 1e3).max().item(),
              "| chunks per CTA min/median/max:", (t[u2, 9] - 1).min().item(), (t[u2, 9] - 1).median().item(), (t[u2, 9] - 1).max().item(),
              "| producer blocked on empty (us) median:", (t[u2, 10].float() / 1e3).median().item())
    if used.any() and name != "cfg5":
        t0 = t[used, 0].min(); sm = t[used, 15]; done = (t[used, 1] - t0).float() / 1e3; start = (t[used, 0] - t0).float() / 1e3
        per = {}
        for s_, d_, st_ in zip(sm.tolist(), done.tolist(), start.tolist()): per.setdefault(s_, []).append((round(st_, 2), round(d_, 2)))
        by = {}
        for s_, v in per.items(): by.setdefault(len(v), []).append(max(d for _, d in v))
        for k_, v in sorted(by.items()): print(f"  SMs with {k_} CTAs: {len(v)}; chunk-done max per SM: min {min(v):.2f} median {sorted(v)[len(v)//2]:.2f} max {max(v):.2f}")
        worst = sorted(per.items(), key=lambda kv: -max(d for _, d in kv[1]))[:4]
        print("  slowest SMs (start, done):", worst)
    dbg.zero_(); torch.cuda.synchronize()
    L.check(lib.cnh_decode(C.byref(d.dec_args[rep]), d.ws_dec.data_ptr(), d.ws_dec.numel(), L.stream_ptr()), "d")
    show(f"decode {name} rep{rep}", 16)
    t = dbg.cpu(); used = t[:, 0] != 0
    if used.any():
        u = t[used].float() / 1e3
        print("  cluster loop, thread 0 per CTA (us) median [min..max]: " + " | ".join(
            f"{nm} {u[:, c].median():.2f} [{u[:, c].min():.2f}..{u[:, c].max():.2f}]"
            for nm, c in (("tile wait", 12), ("scan", 13), ("round barrier", 14), ("refill", 15), ("cluster cuts", 8))))
        lc = t[used][:, 10]
        print("  local cuts per CTA: median", int((lc >> 40).median()), "time (us) median", float(((lc & ((1 << 40) - 1)).float() / 1e3).median()))
    continue
    mg = t[:batch]; print("  merge m/got:", mg[:, 12].tolist()[:8], mg[:, 13].tolist()[:8])
    used = used & (torch.arange(t.shape[0]) >= batch)
    print("  per-CTA: tiles", t[used, 15].float().mean().item(), "heavy tiles", t[used, 12].float().mean().item(), "max", t[used, 12].max().item(),
          "| candidates/tile", (t[used, 13].float() / t[used, 15].float().clamp(min=1)).mean().item(),
          "| final thr (as prob) min/median", torch.tensor(t[used, 14].int().tolist(), dtype=torch.int32).view(torch.float32).min().item(),
          torch.tensor(t[used, 14].int().tolist(), dtype=torch.int32).view(torch.float32).median().item())
lib.cnh_debug_set_buffer(None)
lib.cnh_debug_decode_cluster.argtypes = [C.c_void_p]
print("decode cluster size*1000 + co-resident 8-CTA clusters:", lib.cnh_debug_decode_cluster(C.byref(d.dec_args[0])))
