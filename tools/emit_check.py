"""Detections decoded from the loss launch's candidates vs the regular decode of the same map (bench workload, eager)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "centernet-uda_b200"), ROOT]
import torch
import bench
from cnhead import synthetic
name = next((a for a in sys.argv[1:] if not a.startswith("-")), "cfg5")
B = int(next((a.split("=")[1] for a in sys.argv[1:] if a.startswith("--batch=")), 16))
bench.DeviceStep.FUSE = True
cfg = synthetic.CONFIGS[name]
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
w = bench.Workload(cfg, B, 0, 1, dev, False)
d = w.dstep
for rep in range(2):
    for i in range(w.n_sets):
        s = w.sets[i]
        with torch.cuda.stream(w.stream):
            d.step(i); w.stream.synchronize()
            a = s.dets.clone(); s.dets.zero_()
            d.decode_only(i); w.stream.synchronize()
            b = s.dets.clone()
        import torch.nn.functional as TF
        pm = TF.max_pool2d(s.prob, 3, 1, 1)
        keep = (pm == s.prob).float() * s.prob
        ref_sc, ref_ix = keep.view(keep.shape[0], -1).topk(cfg.K)
        ca = (a[:, :, 4] == ref_sc).all().item(); cb = (b[:, :, 4] == ref_sc).all().item()
        print(f"   scores vs torch top-K: from candidates {'OK' if ca else 'WRONG'}; regular decode {'OK' if cb else 'WRONG'}")
        if not ca:
            HW = cfg.height * cfg.width
            for bb in range(a.shape[0]):
                have = set(a[bb, :, 4].tolist())
                miss = [(float(sc), int(ix)) for sc, ix in zip(ref_sc[bb].tolist(), ref_ix[bb].tolist()) if sc not in have]
                if miss:
                    print(f"   sample {bb}: missing", [(round(sc, 5), "cls", ix // HW, "y", (ix % HW) // cfg.width, "x", ix % cfg.width,
                                                        "tile", (ix // 4096) % 4, "ticket", 319 - ix // 4096) for sc, ix in miss[:6]])
        if not torch.equal(a, b):
            bad = (a != b).any(dim=2)
            rows = bad.nonzero()
            print(f"rep {rep} set {i}: {int(bad.sum())} rows differ (G={d.cand.G}); first:", rows[:6].tolist())
            r = rows[0].tolist()
            print("   cand :", a[r[0], r[1]].tolist()); print("   plain:", b[r[0], r[1]].tolist())
            sa = set(map(tuple, a[r[0]][:, 4:6].tolist())); sb = set(map(tuple, b[r[0]][:, 4:6].tolist()))
            print("   in plain not in cand (score, cls):", sorted(sb - sa, reverse=True)[:5], "| in cand not in plain:", sorted(sa - sb, reverse=True)[:5])
        else:
            print(f"rep {rep} set {i}: identical (G={d.cand.G})")
