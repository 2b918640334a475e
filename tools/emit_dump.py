"""Diagnostic: the candidate workspace an emitting loss launch leaves, against torch's peaks of the same probability map."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "centernet-uda_b200"), ROOT]
import ctypes as C
import numpy as np
import torch
import torch.nn.functional as TF
import bench
from cnhead import synthetic, _lib as L
B = int(next((a.split("=")[1] for a in sys.argv[1:] if a.startswith("--batch=")), 16))
bench.DeviceStep.FUSE = True
cfg = synthetic.CONFIGS["cfg5"]
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
w = bench.Workload(cfg, B, 0, 1, dev, False)
d = w.dstep
up = lambda v: (v + 127) // 128 * 128
for rep in range(2):
    s = w.sets[0]
    with torch.cuda.stream(w.stream):
        L.check(d.lib.cnh_detloss_fused(C.byref(d.loss_args[0]), d.ws_loss.data_ptr(), d.ws_loss.numel(), L.stream_ptr()), "fused")
        w.stream.synchronize()
    G = d.cand.G
    ws = d.ws_cand.cpu().numpy()
    o = 0
    state = ws[o:o + B * 16].view(np.uint32).reshape(B, 4); o += up(B * 16)
    shist = ws[o:o + B * 64 * 4].view(np.uint32).reshape(B, 64); o += up(B * 64 * 4)
    fhist = ws[o:o + B * 4096 * 4].view(np.uint32).reshape(B, 4096); o += up(B * 4096 * 4)
    cnt = ws[o:o + B * G * 4].view(np.uint32).reshape(B, G); o += up(B * G * 4)
    slices = ws[o:o + B * G * 4096 * 8].view(np.uint64).reshape(B, G, 4096)
    print(f"rep {rep}: G={G} overflow={state[:,0].tolist()}")
    print("  cta_cnt sample 0:", cnt[0].tolist())
    pm = TF.max_pool2d(s.prob, 3, 1, 1)
    keep = ((pm == s.prob).float() * s.prob).view(B, -1)
    ref_sc, ref_ix = keep.topk(cfg.K)
    ref_sc, ref_ix = ref_sc.cpu().numpy(), ref_ix.cpu().numpy()
    HW = cfg.height * cfg.width
    for b in range(B):
        keys = np.concatenate([slices[b, j, :cnt[b, j]] for j in range(G)])
        flat = (0xffffffff - (keys & np.uint64(0xffffffff))).astype(np.int64)
        have = set(flat.tolist())
        dup = len(flat) - len(have)
        miss = [(float(sc), int(ix)) for sc, ix in zip(ref_sc[b], ref_ix[b]) if int(ix) not in have]
        thr_bin = None
        tot = int(fhist[b].sum()); tots = int(shist[b].sum())
        # K-th counted key's bin
        c = 0
        for fb in range(4095, -1, -1):
            c += int(fhist[b, fb])
            if c >= cfg.K:
                thr_bin = fb; break
        print(f"  b={b}: keys {len(flat)} dup {dup} counted fine {tot} super {tots} thr_bin {thr_bin} ({(thr_bin or 0)/4096:.5f}) "
              f"K-th score {ref_sc[b,-1]:.5f} missing {len(miss)}")
        for sc, ix in miss[:4]:
            jc = ix // 4096
            print(f"      missing score {sc:.5f} cls {ix//HW} y {(ix%HW)//128} x {ix%128} jc {jc} ticket {319-jc}")
    d.ws_cand[: d.lib.cnh_cand_state_bytes(C.byref(d.cand))].zero_()
    torch.cuda.synchronize()
