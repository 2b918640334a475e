#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference; the GPU box has no copy):

    python tests/golden/make_golden.py

Each .npz holds the exact inputs and the reference's outputs (loss scalars, clamped
probabilities, autograd gradients, decoded detections ...) for one small case.  The
oracle (oracle/head.py) is pinned against these in tests/test_oracle_golden.py and the
CUDA path is checked against them in tests/test_gpu_golden.py.
"""
import os
import sys
import warnings

import numpy as np
import torch

REF = os.environ.get("CNH_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

from losses.centernet import DetectionLoss            # noqa: E402
from losses.entropy import EntropyLoss                # noqa: E402
from losses.max_square import MaxSquareLoss           # noqa: E402
from backends.decode import decode_detection, _nms    # noqa: E402
from utils.image import entropy_map, gaussian_radius, draw_umich_gaussian  # noqa: E402

torch.set_num_threads(1)


def save(name, **arrays):
    out = {}
    for k, v in arrays.items():
        if torch.is_tensor(v):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: " + ", ".join(f"{k}{list(np.shape(v))}" for k, v in out.items()))


def head_case(seed, B, C, H, W, M, D, nobj, *, hm_mu=-2.19, hm_sigma=1.0, no_pos=False,
              dup=False, saturate=False, edge_angles=False, nk=0):
    g = torch.Generator().manual_seed(seed)
    out = {"hm": torch.randn(B, C, H, W, generator=g) * hm_sigma + hm_mu,
           "wh": torch.rand(B, D, H, W, generator=g) * 20.0,
           "reg": torch.rand(B, 2, H, W, generator=g)}
    if D == 3:
        out["wh"][:, 2] = torch.randn(B, H, W, generator=g) * 2.0
    if saturate:
        out["hm"].view(-1)[::7] = 12.0
        out["hm"].view(-1)[3::11] = -12.0
    gt = torch.rand(B, C, H, W, generator=g) ** 6          # skirts in (0,1), mostly small
    gt.view(-1)[::5] = 0.0
    bt = {"hm": gt, "reg_mask": torch.zeros(B, M, dtype=torch.uint8),
          "ind": torch.zeros(B, M, dtype=torch.int64),
          "wh": torch.zeros(B, M, D), "reg": torch.zeros(B, M, 2)}
    for b in range(B):
        n = nobj[b]
        idx = torch.randint(0, H * W, (n,), generator=g)
        if dup and n >= 3:
            idx[1] = idx[0]
            idx[2] = idx[0]
        cls = torch.randint(0, C, (n,), generator=g)
        bt["ind"][b, :n] = idx
        bt["reg_mask"][b, :n] = 1
        bt["wh"][b, :n, :2] = torch.rand(n, 2, generator=g) * 30 + 2
        if D == 3:
            ang = torch.rand(n, generator=g) * 180 - 90
            if edge_angles and n >= 2:
                ang[0], ang[1] = 90.0, -90.0
            bt["wh"][b, :n, 2] = ang
        bt["reg"][b, :n] = torch.rand(n, 2, generator=g)
        if not no_pos:
            gt[b].view(C, -1)[cls, idx] = 1.0
    if nk:
        out["kps"] = torch.randn(B, 2 * nk, H, W, generator=g) * 5
        bt["kps"] = torch.randn(B, M, 2 * nk, generator=g) * 5
        bt["kp_reg_mask"] = (torch.rand(B, M, 2 * nk, generator=g) > 0.4).to(torch.uint8)
        bt["kp_reg_mask"] *= bt["reg_mask"].unsqueeze(-1)
    return out, bt


def run_detloss(name, out, bt, grad_scale=1.0, **kw):
    crit = DetectionLoss(**kw)
    leaves = {k: v.clone().requires_grad_(True) for k, v in out.items()}
    work = {k: leaves[k] * 1.0 for k in leaves}        # non-leaf: the loss sigmoids hm in place
    tgt = {k: v.clone() for k, v in bt.items()}
    loss, stats = crit(work, tgt)
    (loss * grad_scale).backward()
    arrays = {f"in_{k}": v for k, v in out.items()}
    arrays.update({f"bt_{k}": v for k, v in bt.items()})
    arrays.update({f"grad_{k}": leaves[k].grad if leaves[k].grad is not None
                   else torch.zeros_like(leaves[k]) for k in leaves})
    arrays.update({f"stat_{k}": v for k, v in stats.items()})
    arrays["prob"] = work["hm"]
    arrays["grad_scale"] = grad_scale
    arrays["kw_names"] = np.array(list(kw.keys()))
    arrays["kw_vals"] = np.array([repr(v) for v in kw.values()])
    save(name, **arrays)


def main():
    base = dict(hm_weight=1.0, wh_weight=0.1, off_weight=1.0)
    o, b = head_case(1, 2, 3, 16, 20, 6, 2, [4, 2], dup=True)
    run_detloss("detloss_plain", o, b, **base)
    o, b = head_case(2, 2, 3, 16, 20, 6, 3, [5, 0], dup=True, edge_angles=True)
    run_detloss("detloss_angle_sigmoid", o, b, **base, angle_weight=0.7, periodic=False)
    o, b = head_case(3, 2, 3, 16, 20, 6, 3, [5, 3], dup=True, edge_angles=True)
    run_detloss("detloss_angle_periodic", o, b, **base, angle_weight=1.3, periodic=True)
    o, b = head_case(4, 2, 3, 16, 20, 6, 2, [3, 1], no_pos=True)
    run_detloss("detloss_no_positive", o, b, **base)
    o, b = head_case(5, 2, 3, 16, 20, 6, 2, [4, 4], saturate=True, hm_sigma=3.0)
    run_detloss("detloss_saturated", o, b, **base)
    o, b = head_case(6, 3, 5, 12, 12, 4, 2, [4, 1, 0])
    run_detloss("detloss_weights_gradscale", o, b, grad_scale=0.37,
                hm_weight=2.5, wh_weight=0.33, off_weight=0.5)
    o, b = head_case(7, 2, 2, 8, 8, 3, 2, [0, 0])
    run_detloss("detloss_all_masked", o, b, **base)
    o, b = head_case(8, 2, 3, 16, 16, 6, 2, [4, 3], nk=4)
    run_detloss("detloss_keypoints", o, b, **base, kp_weight=1.0,
                kp_indices=[[0, 1], [1, 2], [2, 3]], kp_distance_weight=0.1)
    o, b = head_case(9, 2, 3, 16, 16, 6, 2, [4, 3], nk=4)
    run_detloss("detloss_keypoints_l1dist", o, b, **base, kp_weight=0.8,
                kp_indices=[[0, 3], [1, 2]], kp_distance_weight=0.2, kp_distance_weight_l1=True)
    o, b = head_case(10, 2, 3, 16, 16, 6, 2, [4, 3], nk=3)
    run_detloss("detloss_keypoints_nodist", o, b, **base, kp_weight=1.0)

    # ---- decode --------------------------------------------------------------
    def dec_case(name, seed, B, C, H, W, D, K, rotated=False, use_reg=True, nk=0,
                 sigma=2.0, plateau=False, sparse=False):
        g = torch.Generator().manual_seed(seed)
        logits = torch.randn(B, C, H, W, generator=g) * sigma - 2.19
        if sparse:                                         # fewer peaks than K
            logits[:] = -30.0
            logits[:, :, 1, 1] = 1.0
            logits[:, 0, 2, 3] = 2.0
        if plateau:
            logits[:, :, 2:6, 3:9] = 30.0
        heat = torch.clamp(torch.sigmoid(logits), 1e-4, 1 - 1e-4)
        wh = torch.rand(B, D, H, W, generator=g) * 20
        if D == 3:
            wh[:, 2] = torch.randn(B, H, W, generator=g) * 2
        reg = torch.rand(B, 2, H, W, generator=g) if use_reg else None
        kps = torch.randn(B, 2 * nk, H, W, generator=g) * 4 if nk else None
        with torch.no_grad():
            res = decode_detection(heat.clone(), wh.clone(), None if reg is None else reg.clone(),
                                   kps=None if kps is None else kps.clone(), K=K, rotated=rotated)
            nms = _nms(heat.clone())
        arrays = dict(heat=heat, wh=wh, K=K, rotated=rotated, nms=nms)
        if reg is not None:
            arrays["reg"] = reg
        if kps is not None:
            arrays["kps"] = kps
            arrays["dets"], arrays["kps_out"] = res
        else:
            arrays["dets"] = res
        save(name, **arrays)

    dec_case("decode_plain", 20, 2, 3, 16, 20, 2, 10)
    dec_case("decode_rotated", 21, 2, 3, 16, 20, 3, 10, rotated=True)
    dec_case("decode_noreg", 22, 2, 2, 12, 12, 2, 7, use_reg=False)
    dec_case("decode_kps", 23, 2, 3, 16, 16, 2, 9, nk=3)
    dec_case("decode_fewer_peaks_than_K", 24, 2, 2, 8, 8, 2, 12, sparse=True)
    dec_case("decode_plateau", 25, 1, 2, 12, 12, 2, 20, plateau=True)
    dec_case("decode_K150", 26, 1, 6, 32, 32, 2, 150)

    # ---- UDA losses -----------------------------------------------------------
    def uda_case(name, seed, N, C, H, W, scale=1.5, eta=None, w=1.0):
        g = torch.Generator().manual_seed(seed)
        x = torch.randn(N, C, H, W, generator=g) * scale
        xe = x.clone().requires_grad_(True)
        le, _ = EntropyLoss(eta=eta)({"hm": xe}, None)
        le = le * 1.0
        le *= w
        le.backward()
        xm = x.clone().requires_grad_(True)
        lm, _ = MaxSquareLoss()({"hm": xm}, None)
        lm = lm * 1.0
        lm *= w
        lm.backward()
        xi = x.clone().requires_grad_(True)
        up = torch.randn(N, C, H, W, generator=g)
        im = entropy_map(xi)
        im.backward(up)
        save(name, x=x, eta=np.nan if eta is None else eta, w=w,
             entropy=le / w, entropy_grad=xe.grad, max_square=lm / w, max_square_grad=xm.grad,
             info_map=im, info_up=up, info_grad=xi.grad)

    uda_case("uda_c6", 30, 2, 6, 8, 12)
    uda_case("uda_c6_weighted", 31, 2, 6, 8, 12, w=0.001)
    uda_case("uda_c19_big_logits", 32, 1, 19, 6, 8, scale=30.0)
    uda_case("uda_c2_eta", 33, 2, 2, 8, 8, eta=2.0)
    uda_case("uda_c80", 34, 1, 80, 4, 8)

    g = torch.Generator().manual_seed(40)
    y = torch.randn(4, 1, 4, 4, generator=g) * 3
    for lab in (0, 1):
        yy = y.clone().requires_grad_(True)
        # losses/advent.py:10-18 cannot run on CPU tensors (get_device() == -1); this is
        # the computation its lines 8 and 16 perform.
        l = torch.nn.BCEWithLogitsLoss()(yy, torch.full_like(yy, float(lab)))
        l.backward()
        save(f"advent_label{lab}", y=y, label=lab, loss=l, grad=yy.grad)

    # ---- target rasteriser (row N2) -------------------------------------------
    H, W = 24, 32
    hm = np.zeros((H, W), dtype=np.float32)
    boxes = [(3, 2, 9.3, 14.9), (30, 22, 20.0, 8.0), (15, 12, 40.2, 33.3), (0, 0, 4.0, 4.0), (16, 12, 5.5, 5.1)]
    radii = []
    for cx, cy, bw, bh in boxes:
        r = max(0, int(gaussian_radius((np.ceil(bh), np.ceil(bw)))))
        radii.append(r)
        draw_umich_gaussian(hm, (cx, cy), r)
    save("raster_gaussians", hm=hm, boxes=np.array(boxes, dtype=np.float64), radii=np.array(radii))

    # ---- full target rasterisation of a batch (row N2) ----------------------------------------
    # datasets/coco.py cannot be imported here (imgaug / pycocotools are missing), so the per-object
    # statements of COCO.__getitem__ (datasets/coco.py:168-174, 191-215) are executed around the
    # reference's own gaussian_radius / draw_umich_gaussian, on boxes already in output-grid pixels.
    rng = np.random.RandomState(7)
    B, C, H, W, M = 3, 4, 48, 64, 12
    n_obj = np.array([12, 5, 0], dtype=np.int32)
    boxes = np.zeros((B, M, 4), dtype=np.float32)
    classes = rng.randint(0, C, size=(B, M)).astype(np.int32)
    for b in range(B):
        for k in range(M):
            x1, y1 = rng.uniform(-6, W - 2), rng.uniform(-6, H - 2)
            boxes[b, k] = (x1, y1, x1 + rng.uniform(0.3, 40), y1 + rng.uniform(0.3, 30))
    boxes[0, 3] = (10.0, 5.0, 10.0, 9.0)              # zero width: skipped
    boxes[0, 4] = (61.5, 40.0, 80.0, 60.0)            # clipped at the right/bottom border
    boxes[0, 5] = boxes[0, 6]                         # duplicate object
    classes[0, 5] = classes[0, 6]
    boxes[1, 0] = (0.0, 0.0, 63.0, 47.0)              # the whole map
    hm = np.zeros((B, C, H, W), dtype=np.float32)
    wh = np.zeros((B, M, 2), dtype=np.float32)
    reg = np.zeros((B, M, 2), dtype=np.float32)
    ind = np.zeros((B, M), dtype=np.int64)
    reg_mask = np.zeros((B, M), dtype=np.uint8)
    for b in range(B):
        for k in range(n_obj[b]):
            bbox = np.array([boxes[b, k, 0], boxes[b, k, 1], boxes[b, k, 2], boxes[b, k, 3]], dtype=np.float64)
            cls_id = int(classes[b, k])
            bbox[[0, 2]] = np.clip(bbox[[0, 2]], 0, W - 1)
            bbox[[1, 3]] = np.clip(bbox[[1, 3]], 0, H - 1)
            h, w = bbox[3] - bbox[1], bbox[2] - bbox[0]
            if h > 0 and w > 0:
                radius = gaussian_radius((np.ceil(h), np.ceil(w)))
                radius = max(0, int(radius))
                ct = np.array([(bbox[0] + bbox[2]) / 2, (bbox[1] + bbox[3]) / 2], dtype=np.float32)
                ct_int = ct.astype(np.int32)
                draw_umich_gaussian(hm[b, cls_id], ct_int, radius)
                wh[b, k] = 1. * w, 1. * h
                ind[b, k] = ct_int[1] * W + ct_int[0]
                reg[b, k] = ct - ct_int
                reg_mask[b, k] = 1
    save("raster_targets", boxes=boxes, classes=classes, n_obj=n_obj, hm=hm, wh=wh, reg=reg, ind=ind,
         reg_mask=reg_mask, C=C)


if __name__ == "__main__":
    main()
