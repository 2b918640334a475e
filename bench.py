#!/usr/bin/env python
"""bench.py -- heatmaps/s of the CenterNet head hot path (DetectionLoss fwd+bwd + decode).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

One step = one pass of the hot path over one synthetic batch of the workload (default cfg2: batch 16,
6 classes, 128x128 heat maps, max_detections 150, per GPU -- weak scaling): DetectionLoss forward,
its backward, and decode_detection of the same head tensors.
  value  : device-resident inputs, the three launches of a step (loss, then the backward scale on a side stream
           next to the decode) replayed from CUDA graphs over rotating buffer sets that together exceed the L2
           (HBM-cold), timed with CUDA events.
  e2e    : the same step through the reference-facing plugin API (cnhead.functional.raster_targets,
           losses.centernet.DetectionLoss, loss.backward(), backends.decode.decode_detection) with PINNED HOST inputs:
           H2D copy of the head maps and of the object lists the targets are rasterised from (cnhead.feeder.HostFeeder:
           step i+1's copy rides a copy stream under step i), D2H read of the loss and the detections inside the timed
           region, the host waiting for them every step.  The plugin calls run as ONE CUDA graph per feeder slot
           (cnhead.graphed.HostStep: captured from those very calls the first time a slot is used).
           e2e_eager: the same leg with the plugin calls made one by one from Python; e2e_dense_targets: the
           dataset's dense targets shipped instead of object lists; e2e_targets_only: head maps device-resident
           (as behind the reference's backbone), only the object lists shipped.
  roofline: the dominant kernel (fused detection-loss launch) timed alone with CUDA events.
  cpu_baseline / --impl reference: the oracle port of the reference's PyTorch path on the host cores.
  cfg5   : (every N) the same step on BASELINE config 5's per-GPU shard (16 x 80 x 128^2 of the 128-sample
           COCO-scale batch): ms_per_step, step_hbm_frac, per-kernel {us, frac}; at N > 1 the schedule
           north_star names (count -> all-reduce -> main -> all-reduce -> finalize over NCCL) with the
           all-reduce time.  --fuse auto: the step is timed with and without the loss launch emitting the decode's
           peak candidates; the faster is the block's number, the other is reported beside it.
  shapes : step_hbm_frac of the other named shapes (cfg1, cfg3, cfg4), short runs.
  sharded_parity (N > 1, outside the timed region): one sharded step per schedule is compared with a
           single-device launch over the all-gathered batch -- scalars, probabilities, heat-map gradients and
           detections must be bit-identical, else the run fails (rc != 0).
Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "centernet-uda_b200")
for _p in (PKG, ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

METRIC = "heatmaps/s (DetectionLoss fwd+bwd + decode)"
UNIT = "heatmaps/s"
L2_BYTES = 126 * 2 ** 20


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch override")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the cfg5 / shapes blocks and the sharded parity check")
    ap.add_argument("--fuse", default="auto", choices=["auto", "on", "off"],
                    help="on: the loss launch emits the decode's peak candidates (cnh_cand) where the library supports "
                         "it -- streaming launches of 128-wide maps, i.e. the cfg5 shard; the single-wave launch of cfg2 "
                         "never does -- and the decode then runs from them (DESIGN 4.4); off: loss + streaming decode; "
                         "auto: on, and the cfg5 block times both settings and reports the faster (and the other)")
    return ap.parse_args()


def workload_name(cfg, batch):
    extra = " + rotated/periodic angle head" if cfg.angle else ""
    if getattr(cfg, "advent", False):
        return (f"{cfg.name}: batch {batch} per GPU, {cfg.classes} classes, {cfg.height}x{cfg.width} heat maps, the head-path "
                f"kernels of one ADVENT step (uda/adversarial_entropy_minimization.py:77-152): source DetectionLoss fwd+bwd, "
                f"entropy_map of the target logits forward + backward (dense upstream gradient), entropy_map of the sigmoided "
                f"source map and of the target logits for the two discriminator passes, three AdventLoss fwd+bwd on "
                f"[{batch},1,4,4] discriminator logits (the discriminator network itself is torch/cuDNN: not on the path)")
    if cfg.target_domain:
        extra += " + EntropyLoss and MaxSquareLoss fwd+bwd on a target-domain batch"
    return (f"{cfg.name}: batch {batch} per GPU, {cfg.classes} classes, {cfg.height}x{cfg.width} heat maps, "
            f"DetectionLoss fwd+bwd + decode K={cfg.K}{extra}")


def config_dict(cfg, batch, world):
    """`config` of the JSON line -- the SAME dict in both arms (the reference arm runs `our arm's config`)."""
    return {"workload": workload_name(cfg, batch), "global_batch": batch * world,
            "l2": "GPU arm: rotating buffer sets that together exceed twice the 126 MiB L2, HBM-cold every step "
                  "(sets larger than the L2 rotate in pairs); CPU arm: one batch in host memory",
            "parallelism": "single device" if world == 1 else f"batch-sharded dp{world}, one process per GPU"}


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------
class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_step(oracle, data, kw, cfg):
    loss, stats, prob, grads = oracle.detection_loss_with_grads(data["output"], data["batch"], **kw)
    with torch.no_grad():
        dets = oracle.decode_two_stage(prob, data["output"]["wh"], data["output"]["reg"], K=cfg.K, rotated=cfg.rotated)
    return loss, dets


def time_cpu(cfg, batch, steps, warmup, budget_s=None):
    import oracle
    from cnhead import synthetic
    torch.set_num_threads(os.cpu_count() or 1)
    data = synthetic.make_inputs(cfg, batch=batch, hm_sigma=2.0)
    kw = synthetic.loss_kwargs(cfg)
    for _ in range(warmup):
        cpu_step(oracle, data, kw, cfg)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        cpu_step(oracle, data, kw, cfg)
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return batch * done / dt, dt / done * 1e3, done, torch.get_num_threads()


def run_reference(args, cfg, batch, rank, world):
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(1, args.warmup)
    value, ms, done, cores = time_cpu(cfg, batch, steps, warmup, budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(cfg, batch, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{done} steps of one {cfg.name} batch ({batch} samples) on the host: oracle port of "
                                   f"the reference's PyTorch-CPU DetectionLoss fwd+bwd + decode (the Python reference "
                                   f"cannot travel to the GPU box); one CPU process whatever --gpus says"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class BufferSet:
    """device-resident head maps, targets and outputs of one step."""

    def __init__(self, data, cfg, dev):
        from cnhead import _lib as L, functional as F
        o, b = data["output"], data["batch"]
        self.hm, self.wh, self.reg = (o[k].to(dev) for k in ("hm", "wh", "reg"))
        self.gt, self.ind, self.mask = b["hm"].to(dev), b["ind"].to(dev), b["reg_mask"].to(dev)
        self.wh_t, self.reg_t = b["wh"].to(dev), b["reg"].to(dev)
        self.prob = torch.empty_like(self.hm)
        self.grads = [torch.empty_like(self.hm), torch.empty_like(self.wh), torch.empty_like(self.reg)]
        self.scalars = torch.zeros(L.SCALARS, device=dev)
        self.totals = torch.zeros(L.TOTALS, dtype=torch.int64, device=dev)
        self.norm = torch.zeros(4, dtype=torch.float64, device=dev)
        self.ones = torch.ones(L.SCALARS, device=dev)          # upstream gradient of loss.backward()
        self.dets = torch.empty(self.hm.shape[0], cfg.K, 7 if cfg.rotated else 6, device=dev)
        mode = L.ANGLE_NONE if not cfg.angle else (L.ANGLE_PERIODIC if cfg.periodic else L.ANGLE_SIGMOID)
        self.heads = [F.HeadSpec(self.wh, self.wh_t, self.mask, 0.1, 1.0, mode),
                      F.HeadSpec(self.reg, self.reg_t, self.mask, 1.0)]
        self.tdom = None
        if cfg.target_domain:                      # cfg4: target-domain logits for EntropyLoss + MaxSquareLoss
            self.tdom = data["target"]["hm"].to(dev)
            self.tgrads = [torch.empty_like(self.tdom), torch.empty_like(self.tdom)]
            self.tloss = torch.zeros(2, device=dev)
        if getattr(cfg, "advent", False):          # ADVENT: maps, a dense upstream gradient, discriminator logits
            g = torch.Generator().manual_seed(99)
            B = self.hm.shape[0]
            self.maps = [torch.empty_like(self.tdom) for _ in range(3)]
            self.upstream = (torch.randn(self.tdom.shape, generator=g) * 1e-3).to(dev)
            self.disc_y = [torch.randn(B, 1, 4, 4, generator=g).to(dev) for _ in range(3)]
            self.disc_g = [torch.empty_like(y) for y in self.disc_y]
            self.adv_loss = torch.zeros(3, device=dev)

    def nbytes(self):
        ts = [self.hm, self.wh, self.reg, self.gt, self.prob] + self.grads
        if self.tdom is not None:
            ts += [self.tdom] + self.tgrads
        if hasattr(self, "maps"):
            ts += self.maps + [self.upstream]
        return sum(t.numel() * t.element_size() for t in ts)


class DeviceStep:
    """the launches of one step through the C ABI on device-resident buffers.

    schedule: 'single' (one GPU), 'peers' (sharded: normalisers exchanged INSIDE the fused launch through
    NVLink-mapped mailboxes; the totals by a one-warp launch forked next to decode) or 'nccl' (sharded:
    count -> all-reduce -> main -> all-reduce (forked next to decode) -> finalize)."""

    FUSE = True          # --fuse auto: the loss launch emits the decode's peak candidates when the shape allows it

    def __init__(self, sets, cfg, world, group, schedule="auto"):
        import ctypes as C
        from cnhead import _lib as L, functional as F, sharded
        self.C, self.L, self.sharded, self.world, self.group, self.cfg = C, L, sharded, world, group, cfg
        self.sets = sets
        self.lib = L.lib()
        dev = sets[0].hm.device
        self.loss_args, self.scale_args, self.dec_args = [], [], []
        # candidate emission (include/cnhead.h: cnh_cand): one workspace, used by the steps in stream order
        self.cand = L.Cand()
        self.ws_cand = torch.zeros(self.lib.cnh_cand_workspace_bytes(sets[0].hm.shape[0]) + 256, dtype=torch.uint8, device=dev)
        self.cand.workspace, self.cand.workspace_bytes = self.ws_cand.data_ptr(), self.ws_cand.numel()
        self.cand.K, self.cand.G = cfg.K, 0
        self.plain_args = []                      # the same launches without candidate emission (timed alone)
        for s in sets:
            a = F.fill_detloss_args(s.hm, s.gt, s.ind, s.heads, 1.0, s.prob, s.grads, s.scalars, s.totals,
                                    norm=s.norm, norm_out=s.norm, b_global=s.hm.shape[0] * world)
            self.plain_args.append(F.fill_detloss_args(s.hm, s.gt, s.ind, s.heads, 1.0, s.prob, s.grads, s.scalars, s.totals,
                                                       norm=s.norm, norm_out=s.norm, b_global=s.hm.shape[0] * world))
            if DeviceStep.FUSE and not getattr(cfg, "advent", False):   # (the ADVENT step has no decode to consume them)
                a.cand = C.pointer(self.cand)
            sc = L.ScaleArgs()
            sc.n_tensors = 3
            for i, t in enumerate(s.grads):
                sc.data[i], sc.count[i] = t.data_ptr(), t.numel()
                sc.fa[i], sc.fb[i] = s.ones.data_ptr(), None
            d = L.DecodeArgs()
            B, Cc, H, W = s.hm.shape
            d.B, d.C, d.H, d.W, d.K, d.D = B, Cc, H, W, cfg.K, s.wh.shape[1]
            d.rotated, d.nk = (1 if cfg.rotated else 0), 0
            d.heat, d.wh, d.reg, d.kps = s.prob.data_ptr(), s.wh.data_ptr(), s.reg.data_ptr(), None
            d.dets, d.inds_out, d.kps_out = s.dets.data_ptr(), None, None
            d.apply_sigmoid, d.box_scale = 0, 1.0
            self.loss_args.append(a)
            self.scale_args.append(sc)
            self.dec_args.append(d)
        self.ws_loss = torch.zeros(self.lib.cnh_detloss_workspace_bytes(C.byref(self.loss_args[0])) + 256,
                                   dtype=torch.uint8, device=dev)
        self.ws_dec = torch.zeros(self.lib.cnh_decode_workspace_bytes(C.byref(self.dec_args[0])) + 256,
                                  dtype=torch.uint8, device=dev)
        self.uda_scale, self.ws_soft = [], None
        self.advent = bool(getattr(cfg, "advent", False))
        if cfg.target_domain and not self.advent:
            s0 = sets[0]
            N, Cc, H, W = s0.tdom.shape
            self.uda_dims = (N, Cc, H, W, N * world)
            self.ws_soft = [torch.zeros(self.lib.cnh_softmax_workspace_bytes(N, Cc, H, W) + 256, dtype=torch.uint8, device=dev)
                            for _ in range(2)]
            for s in sets:
                sc = L.ScaleArgs()
                sc.n_tensors = 2
                for i, t in enumerate(s.tgrads):
                    sc.data[i], sc.count[i] = t.data_ptr(), t.numel()
                    sc.fa[i], sc.fb[i] = s.ones.data_ptr(), None
                self.uda_scale.append(sc)
        self.box = None
        self.schedule = "single"
        # the backward scale (and, sharded, what only the loss VALUE needs) runs on a side stream NEXT TO the decode:
        # neither reads what the other writes; joined before the step ends
        self.side = torch.cuda.Stream(device=dev)
        self.ev_fork, self.ev_join = torch.cuda.Event(), torch.cuda.Event()
        if world > 1:
            self.schedule = "nccl"
            self.single_wave = bool(self.lib.cnh_detloss_single_wave(C.byref(self.loss_args[0])))
            if schedule in ("auto", "peers"):
                try:
                    self.box = self.sharded.PeerMailbox.get(group)
                    self.schedule = "peers"
                except Exception as e:                      # noqa: BLE001
                    print(f"[bench] peer mailboxes unavailable ({e!r}); using the NCCL schedule", file=sys.stderr)
            # every rank must take the same branch
            flag = torch.tensor([1 if self.schedule == "peers" else 0], device=dev)
            torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
            if int(flag.item()) == 0:
                self.schedule, self.box = "nccl", None
        self.launches_per_step = {"single": 3, "peers": 4, "nccl": 5}[self.schedule] + (3 if self.ws_soft is not None else 0)
        if self.advent:
            self.launches_per_step += 7 - 1           # three maps forward, one backward, three BCE; no decode

    def fused_decode(self):
        return self.cand.G > 0

    def describe(self):
        tail = (" [the loss launch emits the decode's peak candidates; cnh_decode_candidates does not read the heat map]"
                if self.fused_decode() else "")
        return self._describe() + tail

    def _describe(self):
        if self.schedule == "single":
            return "single GPU: fused loss launch, then the backward scale (side stream) next to the decode"
        if self.schedule == "peers":
            where = ("single wave: the finaliser CTA posts, the chunk CTAs poll" if self.single_wave else
                     "pre-count schedule: CTA 0 posts after the count phase's grid barrier, every CTA polls")
            return (f"normalisers exchanged inside the fused loss launch through NVLink-mapped peer mailboxes "
                    f"({self.box.how}; {where}); totals traded by a 1-warp launch on a side stream next to decode")
        return ("cnh_detloss_count -> ncclAllReduce(4 doubles) -> cnh_detloss_main -> [side stream: "
                "ncclAllReduce(24 int64 exact totals) -> cnh_detloss_finalize] next to backward scale + decode")

    # ---- pieces (also timed alone) -------------------------------------------------------------------------
    def loss_only(self, i):
        """the fused loss launch alone, WITHOUT candidate emission (nothing would consume the candidates)"""
        C, L = self.C, self.L
        a = self.plain_args[i]
        L.check(self.lib.cnh_detloss_fused(C.byref(a), self.ws_loss.data_ptr(), self.ws_loss.numel(), L.stream_ptr()),
                "detloss_fused")

    def loss_decode_pair(self, i):
        """single GPU: the loss launch (emitting candidates when it can) + the decode that consumes them"""
        C, L = self.C, self.L
        a = self.loss_args[i]
        L.check(self.lib.cnh_detloss_fused(C.byref(a), self.ws_loss.data_ptr(), self.ws_loss.numel(), L.stream_ptr()),
                "detloss_fused")
        self.decode_step(i)

    def decode_step(self, i):
        """the decode of a step: from the loss launch's candidates when it emitted any, else the regular decode"""
        L = self.L
        if self.cand.G > 0:
            L.check(self.lib.cnh_decode_candidates(self.C.byref(self.dec_args[i]), self.C.byref(self.cand), L.stream_ptr()),
                    "decode_candidates")
        else:
            self.decode_only(i)

    def count_only(self, i):
        C, L = self.C, self.L
        a = self.loss_args[i]
        L.check(self.lib.cnh_detloss_count(C.byref(a), self.ws_loss.data_ptr(), self.ws_loss.numel(), L.stream_ptr()), "count")

    def main_only(self, i):
        C, L = self.C, self.L
        a = self.plain_args[i]
        a.scalars = None
        L.check(self.lib.cnh_detloss_main(C.byref(a), self.ws_loss.data_ptr(), self.ws_loss.numel(), L.stream_ptr()), "main")

    def allreduce_only(self, i):
        s = self.sets[i]
        self.sharded.exchange_normalisers(s.norm, self.group)
        self.sharded.reduce_totals(s.totals, self.group)

    def scale_only(self, i):
        self.L.check(self.lib.cnh_scale_inplace(self.C.byref(self.scale_args[i]), self.L.stream_ptr()), "scale")

    def decode_only(self, i):
        L = self.L
        L.check(self.lib.cnh_decode(self.C.byref(self.dec_args[i]), self.ws_dec.data_ptr(), self.ws_dec.numel(),
                                    L.stream_ptr()), "decode")

    def advent_tail(self, i, st):
        """uda/adversarial_entropy_minimization.py:91-133 after the source loss: generator pass (map forward, BCE against
        the source label, backward through the map with the discriminator's input gradient), then the two detached
        discriminator passes -- the source map taken over the ALREADY SIGMOIDED tensor (:116)."""
        L, lib, s = self.L, self.lib, self.sets[i]
        N, Cc, H, W = s.tdom.shape
        n = s.disc_y[0].numel()
        L.check(lib.cnh_entropy_map_fwd(s.tdom.data_ptr(), s.maps[0].data_ptr(), N, Cc, H, W, st), "entropy_map_fwd")
        L.check(lib.cnh_bce_const(s.disc_y[0].data_ptr(), s.disc_g[0].data_ptr(), s.adv_loss[0:].data_ptr(), n, 0.0, st), "bce")
        L.check(lib.cnh_entropy_map_bwd(s.tdom.data_ptr(), s.upstream.data_ptr(), s.tgrads[0].data_ptr(), N, Cc, H, W, st),
                "entropy_map_bwd")
        L.check(lib.cnh_entropy_map_fwd(s.prob.data_ptr(), s.maps[1].data_ptr(), N, Cc, H, W, st), "entropy_map_fwd")
        L.check(lib.cnh_bce_const(s.disc_y[1].data_ptr(), s.disc_g[1].data_ptr(), s.adv_loss[1:].data_ptr(), n, 0.0, st), "bce")
        L.check(lib.cnh_entropy_map_fwd(s.tdom.data_ptr(), s.maps[2].data_ptr(), N, Cc, H, W, st), "entropy_map_fwd")
        L.check(lib.cnh_bce_const(s.disc_y[2].data_ptr(), s.disc_g[2].data_ptr(), s.adv_loss[2:].data_ptr(), n, 1.0, st), "bce")

    # ---- the step -------------------------------------------------------------------------------------------
    def step(self, i):
        C, L = self.C, self.L
        st = L.stream_ptr()
        a, s = self.loss_args[i], self.sets[i]
        if self.schedule == "single":
            L.check(self.lib.cnh_detloss_fused(C.byref(a), self.ws_loss.data_ptr(), self.ws_loss.numel(), st), "fused")
        elif self.schedule == "peers":
            a.scalars = s.scalars.data_ptr()
            a.flags |= L.FLAG_DEFER_TOTALS            # the loss VALUE is completed next to scale + decode (below)
            L.check(self.lib.cnh_detloss_fused_peers(C.byref(a), C.byref(self.box.c), self.ws_loss.data_ptr(),
                                                     self.ws_loss.numel(), st), "fused_peers")
        else:
            a.scalars = None
            L.check(self.lib.cnh_detloss_count(C.byref(a), self.ws_loss.data_ptr(), self.ws_loss.numel(), st), "count")
            self.sharded.exchange_normalisers(s.norm, self.group)
            L.check(self.lib.cnh_detloss_main(C.byref(a), self.ws_loss.data_ptr(), self.ws_loss.numel(), st), "main")
            a.scalars = s.scalars.data_ptr()
        # fork: the backward scale and what only the loss VALUE needs (the totals' exchange and the scalars) run on a
        # side stream next to the decode; joined before the step ends
        main = torch.cuda.current_stream()
        self.ev_fork.record(main)
        self.side.wait_event(self.ev_fork)
        with torch.cuda.stream(self.side):
            L.check(self.lib.cnh_scale_inplace(C.byref(self.scale_args[i]), L.stream_ptr()), "scale")    # backward
            if self.schedule == "peers":
                L.check(self.lib.cnh_detloss_peers_finalize(C.byref(a), C.byref(self.box.c), self.ws_loss.data_ptr(),
                                                            self.ws_loss.numel(), L.stream_ptr()), "peers_finalize")
            elif self.schedule == "nccl":
                self.sharded.reduce_totals(s.totals, self.group)
                L.check(self.lib.cnh_detloss_finalize(C.byref(a), s.totals.data_ptr(), L.stream_ptr()), "finalize")
            self.ev_join.record(self.side)
        if self.advent:
            self.advent_tail(i, st)
        else:
            self.decode_step(i)
        main.wait_event(self.ev_join)
        if self.ws_soft is not None:               # cfg4: EntropyLoss and MaxSquareLoss fwd+bwd on the target batch
            N, Cc, H, W, n_total = self.uda_dims
            for j, mode in enumerate((L.SOFTMAX_ENTROPY, L.SOFTMAX_MAX_SQUARE)):
                L.check(self.lib.cnh_softmax_loss(s.tdom.data_ptr(), s.tgrads[j].data_ptr(), s.tloss[j:].data_ptr(), N, Cc,
                                                  H, W, n_total, mode, 0.0, self.ws_soft[j].data_ptr(),
                                                  self.ws_soft[j].numel(), st), "softmax_loss")
            L.check(self.lib.cnh_scale_inplace(C.byref(self.uda_scale[i]), st), "scale")         # their backward


def barrier(world):
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()


def aligned_start(world, dev):
    """barrier + synchronize.  (Tried: letting every rank leave at one agreed wall-clock instant 3 ms later -- the GPUs
    idle meanwhile and the first timed steps then run 3x slower: 103 instead of 28 us/step over a 20-step region at
    N = 8.  What aligns the ranks instead is an untimed pre-roll of the step graph right in front of the timed region,
    see GraphRunner.timed.)"""
    barrier(world)


def max_over_ranks(ms, world, dev):
    if world == 1:
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t.item())


class GraphRunner:
    """fn(i) for buffer set i, captured once per set plus one graph holding a whole round of all sets
    (so that replay cost on the host is amortised); falls back to eager launches if capture fails."""

    def __init__(self, fn, n_sets, stream, use_graph, rank):
        self.fn, self.n, self.graphs, self.round = fn, n_sets, None, None
        if not use_graph:
            return
        try:
            self.graphs = []
            for i in range(n_sets):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=stream):
                    fn(i)
                self.graphs.append(g)
            self.round = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.round, stream=stream):
                for i in range(n_sets):
                    fn(i)
        except Exception as e:                       # e.g. a collective that cannot be captured
            if rank == 0:
                print(f"[bench] graph capture failed ({type(e).__name__}: {e}); timing eager launches", file=sys.stderr)
            self.graphs, self.round = None, None
            torch.cuda.synchronize()

    def run(self, count):
        """exactly `count` steps, continuing the rotation over buffer sets"""
        if self.graphs is None:
            for i in range(count):
                self.fn(i % self.n)
            return
        for _ in range(count // self.n):
            self.round.replay()
        for i in range(count % self.n):
            self.graphs[i].replay()

    def timed(self, steps, warmup, world, dev):
        self.run(warmup)
        aligned_start(world, dev)
        if self.graphs is not None and steps >= self.n:
            # Pre-roll, untimed and NOT followed by a host synchronisation: one whole round of the step graph right in
            # front of the timed region (reported as "warmup_graph_steps").  The timed region replays this graph, whose
            # first launch uploads it to the device (hundreds of microseconds); and at N > 1 the steps' own rendezvous
            # (in-kernel mailboxes / NCCL) brings the ranks, which leave a host barrier tens of microseconds apart, into
            # lockstep before the start event is reached in stream order.
            self.round.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        self.run(steps)
        e1.record()
        barrier(world)
        return max_over_ranks(e0.elapsed_time(e1), world, dev)


def hbm_peak():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    if "hbm_gbs" in peaks:
        return float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy), of measured"
    return 6650.0, "6650 GB/s fallback (B200_PROFILING.md), of fallback"


class Workload:
    """buffer sets + launches + graphs of one named shape on this rank."""

    def __init__(self, cfg, batch, rank, world, dev, use_graph=True, schedule="auto"):
        from cnhead import synthetic
        self.cfg, self.batch, self.rank, self.world, self.dev = cfg, batch, rank, world, dev
        mk = lambda i: BufferSet(synthetic.make_inputs(cfg, batch=batch, hm_sigma=2.0, seed_offset=i,   # noqa: E731
                                                       sample_offset=rank * batch), cfg, dev)
        probe = mk(0)
        # rotating buffer sets: together > 2x L2 so every step reads its inputs from HBM
        n_sets = max(2, min(16, -(-2 * L2_BYTES // probe.nbytes())))
        if probe.nbytes() > L2_BYTES:
            n_sets = 2
        self.sets = [probe] + [mk(1 + i) for i in range(n_sets - 1)]
        self.n_sets, self.set_bytes = n_sets, probe.nbytes()
        self.dstep = DeviceStep(self.sets, cfg, world, None, schedule)
        self.stream = torch.cuda.Stream(device=dev)
        self.use_graph = use_graph
        self.runners = {}
        hw = cfg.height * cfg.width
        self.loss_bytes = batch * (16 * cfg.classes * hw + 4 * (cfg.wh_channels + 2) * hw)
        self.dec_bytes = batch * 4 * cfg.classes * hw
        self.step_bytes = batch * cfg.bytes_per_sample()
        with torch.cuda.stream(self.stream):
            for i in range(n_sets):                      # warm every kernel / tensor map before capture
                self.dstep.step(i)
            torch.cuda.synchronize()

    def runner(self, name, collective=False):
        if name not in self.runners:
            fn = getattr(self.dstep, name)
            with torch.cuda.stream(self.stream):
                for i in range(self.n_sets):
                    fn(i)
                torch.cuda.synchronize()
                self.runners[name] = GraphRunner(fn, self.n_sets, self.stream, self.use_graph, self.rank)
        return self.runners[name]

    def time(self, name, steps, warmup, collective):
        """ms per call of dstep.<name>; `collective`: every rank runs it in lockstep (max over ranks)."""
        r = self.runner(name)
        with torch.cuda.stream(self.stream):
            ms = r.timed(steps, warmup, self.world if collective else 1, self.dev)
        return ms / steps

    def launch_mode(self):
        r = self.runners.get("step")
        return "CUDA graph replay" if (r is not None and r.graphs) else "eager stream launches"

    def close(self):
        self.runners.clear()
        self.sets = None
        self.dstep = None
        torch.cuda.synchronize()
        torch.cuda.empty_cache()


def kernel_block(w, steps, warmup, peak):
    """per-kernel {us, frac} of one workload's launches, each timed alone (graph replay, HBM-cold rotation)."""
    out = {}

    def put(tag, name, nbytes, collective=False):
        us = w.time(name, steps, warmup, collective) * 1e3
        out[tag] = {"us": round(us, 2), "frac": round(nbytes / (us * 1e-6) / 1e9 / peak, 3) if nbytes else None}

    sched = w.dstep.schedule
    if sched == "single":
        put("loss_fused", "loss_only", w.loss_bytes)
        if w.dstep.fused_decode():
            put("loss_emitting_candidates_plus_decode_from_them", "loss_decode_pair", w.loss_bytes + w.dec_bytes)
    if sched == "nccl":
        hw = w.cfg.height * w.cfg.width
        put("loss_count", "count_only", w.batch * 4 * w.cfg.classes * hw)
        put("loss_main", "main_only", w.loss_bytes)
        put("allreduce_pair", "allreduce_only", 0, collective=True)
    put("scale_noop", "scale_only", 0)
    put("decode", "decode_only", w.dec_bytes)
    return out


def workload_block(cfg, batch, rank, world, dev, steps, warmup, use_graph, peak, kernels=True, schedule="auto"):
    """one named shape: step time (max over ranks), whole-job heatmaps/s, step_hbm_frac, per-kernel timings."""
    w = Workload(cfg, batch, rank, world, dev, use_graph, schedule)
    ms = w.time("step", steps, warmup, True)
    blk = {"workload": workload_name(cfg, batch), "ms_per_step": ms, "value": batch * world / (ms * 1e-3), "unit": UNIT,
           "steps": steps, "warmup": warmup, "step_algorithmic_bytes": w.step_bytes,
           "step_hbm_frac": w.step_bytes / (ms * 1e-3) / 1e9 / peak, "schedule": w.dstep.describe(),
           "launch": w.launch_mode(), "buffer_sets": f"{w.n_sets} x {w.set_bytes / 2**20:.0f} MiB",
           "gpu_launches_per_step": w.dstep.launches_per_step}
    if kernels:
        blk["kernels"] = kernel_block(w, max(20, steps // 2), warmup, peak)
    return blk, w


def emission_parity(w):
    """Outside any timed region: the detections a step decodes from the loss launch's candidates against the regular
    decode of the same probability map (set 0).  None: the launch emitted nothing; False: they differ (the caller then
    times the step without emission and says so)."""
    d, s = w.dstep, w.sets[0]
    with torch.cuda.stream(w.stream):
        d.step(0)
        w.stream.synchronize()
        if not d.fused_decode():
            return None
        from_cand = s.dets.clone()
        s.dets.zero_()
        d.decode_only(0)
        w.stream.synchronize()
    if not torch.equal(from_cand, s.dets):
        return False
    return "detections from the loss launch's candidates bit-identical to the regular decode of the same map"


def sharded_parity(w, rank, world, dev):
    """N > 1, outside any timed region: this rank's sharded step (set 0) against ONE single-device launch over the
    all-gathered batch.  Scalars, probabilities, heat-map gradients and detections must be bit-identical; the
    regression gradients (float atomics on duplicate centres) within 1e-6.  Raises on mismatch."""
    import ctypes as C
    from cnhead import _lib as L, functional as F
    dist = torch.distributed
    s, d, cfg = w.sets[0], w.dstep, w.cfg
    with torch.cuda.stream(w.stream):
        d.step(0)
    torch.cuda.synchronize()

    def gather(t):
        out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        dist.all_gather_into_tensor(out, t.contiguous())
        return out

    hm, wh, reg, gt, ind, mask, wh_t, reg_t = (gather(t) for t in (s.hm, s.wh, s.reg, s.gt, s.ind, s.mask, s.wh_t, s.reg_t))
    torch.cuda.synchronize()
    heads = [F.HeadSpec(wh, wh_t, mask, s.heads[0].weight, s.heads[0].angle_weight, s.heads[0].angle_mode),
             F.HeadSpec(reg, reg_t, mask, s.heads[1].weight)]
    prob = torch.empty_like(hm)
    grads = [torch.empty_like(hm), torch.empty_like(wh), torch.empty_like(reg)]
    scal = torch.zeros(L.SCALARS, device=dev)
    tot = torch.zeros(L.TOTALS, dtype=torch.int64, device=dev)
    a = F.fill_detloss_args(hm, gt, ind, heads, 1.0, prob, grads, scal, tot)
    ws = torch.zeros(L.lib().cnh_detloss_workspace_bytes(C.byref(a)) + 256, dtype=torch.uint8, device=dev)
    L.check(L.lib().cnh_detloss_fused(C.byref(a), ws.data_ptr(), ws.numel(), L.stream_ptr()), "fused (gathered batch)")
    dets = F.decode(prob, wh, reg, K=cfg.K, rotated=cfg.rotated)
    torch.cuda.synchronize()
    sl = slice(rank * w.batch, (rank + 1) * w.batch)
    res = {"schedule": d.schedule, "global_batch": int(hm.shape[0])}
    def values(t):                      # the 12 exact quantities as Python integers, hi * 2^32 + lo (an all-reduced
        t = t.tolist()                  # sum of carry-normalised pairs is compared by value)
        return [(t[q] << 32) + t[12 + q] for q in range(12)]

    checks = {"scalars": torch.equal(s.scalars[:6], scal[:6]), "totals": values(s.totals) == values(tot),
              "prob": torch.equal(s.prob, prob[sl]), "grad_hm": torch.equal(s.grads[0], grads[0][sl]),
              "dets": torch.equal(s.dets, dets[sl])}
    reg_rel = 0.0
    for mine, ref in ((s.grads[1], grads[1][sl]), (s.grads[2], grads[2][sl])):
        reg_rel = max(reg_rel, float((mine - ref).abs().max() / ref.abs().max().clamp_min(1e-30)))
    ok = all(checks.values()) and reg_rel <= 1e-6
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    res.update({k: ("bit-exact" if v else "MISMATCH") for k, v in checks.items()})
    res["reg_grads_max_rel"] = reg_rel
    res["loss"] = float(scal[0])
    if int(flag.item()) != 1:
        raise RuntimeError(f"sharded parity FAILED on some rank (rank {rank}: {res})")
    return res


def run_ours(args, cfg, batch, rank, local_rank, world):
    from cnhead import synthetic
    assert torch.cuda.is_available(), "bench.py needs a CUDA device for --impl ours"
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    # auto: the cfg5 block times its step under both settings and keeps the faster (reporting both); on / off: fixed
    DeviceStep.FUSE = args.fuse != "off"
    steps, warmup = args.steps, max(3, args.warmup)
    use_graph = not args.no_graph
    peak, peak_src = hbm_peak()

    # ---- value: device-resident, graph-replayed ------------------------------------------------------
    w = Workload(cfg, batch, rank, world, dev, use_graph)
    w.runner("step")
    if world == 1:
        w.runner("loss_only")
    clocks = Clocks(local_rank)
    clocks.start()
    ms_step = w.time("step", steps, warmup, True)
    # ---- roofline: the dominant kernel alone, same buffers, CUDA events on its stream -------------------
    ms_kernel = w.time("loss_only", steps, warmup, False) if world == 1 else None
    clk = clocks.stop()
    value = batch * world / (ms_step * 1e-3)
    launch_mode, schedule, n_sets, set_bytes = w.launch_mode(), w.dstep.describe(), w.n_sets, w.set_bytes
    launches_per_step, loss_bytes, step_bytes = w.dstep.launches_per_step, w.loss_bytes, w.step_bytes
    parity = {}
    if world > 1 and not args.no_extra and not getattr(cfg, "advent", False):
        parity[cfg.name] = sharded_parity(w, rank, world, dev)
    ep_main = None if getattr(cfg, "advent", False) else emission_parity(w)    # (--config cfg5: the main step emits)
    w.close()

    # ---- e2e: plugin API, pinned host inputs, H2D + D2H inside the timed region ------------------------
    # `e2e` (the headline leg): every input of the step from pinned host memory -- the head maps and, when the
    # targets are rasterisable on the device (no angle channel), the object lists they are made from; else the
    # dataset's dense targets.  e2e_dense_targets / e2e_targets_only: see run_e2e.
    # Every leg runs the step as ONE CUDA graph per feeder slot (cnhead.graphed.HostStep: the same plugin calls,
    # captured once); e2e_eager is the headline leg again with the plugin calls made one by one from Python.
    e2e = e2e_dense = e2e_tonly = e2e_eager = None
    if not args.no_e2e:
        graphed = not args.no_graph
        if graphed and world > 1:
            # HostStep captures the sharded step only with the in-kernel peer exchange (cnhead/graphed.py: the NCCL
            # schedule's all-reduces do not replay through it); every rank must take the same branch
            from cnhead import sharded
            ok = 1
            try:
                sharded.PeerMailbox.get(None)
            except Exception:                               # noqa: BLE001
                ok = 0
            flag = torch.tensor([ok], device=dev)
            torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
            graphed = bool(int(flag.item()))
            if not graphed and rank == 0:
                print("[bench] peer mailboxes unavailable: e2e legs run the plugin calls eagerly", file=sys.stderr)
        main_mode = "boxes" if not cfg.angle else "dense"
        e2e_dense = run_e2e(cfg, batch, rank, world, dev, steps, warmup, "dense", graphed)
        e2e = run_e2e(cfg, batch, rank, world, dev, steps, warmup, "boxes", graphed) if not cfg.angle else e2e_dense
        e2e_tonly = run_e2e(cfg, batch, rank, world, dev, steps, warmup, "targets_only", graphed)
        if graphed:
            e2e_eager = run_e2e(cfg, batch, rank, world, dev, steps, warmup, main_mode, False)

    # ---- the other named shapes: BASELINE config 5's per-GPU shard at every N; cfg1 / cfg3 / cfg4 ------
    extra = {}
    if not args.no_extra:
        x_steps, x_warm = 200, 20
        if cfg.name != "cfg5":
            c5 = synthetic.CONFIGS["cfg5"]
            blk, w5 = workload_block(c5, 16, rank, world, dev, x_steps, x_warm, use_graph, peak)
            ep = emission_parity(w5)
            bad = torch.tensor([1 if ep is False else 0], device=dev)
            if world > 1:
                torch.distributed.all_reduce(bad, op=torch.distributed.ReduceOp.MAX)
            if int(bad.item()):
                # never seen; a step whose detections differ is not a step: time it without emission and say so
                print("[bench] candidate emission: detections differ from the regular decode; cfg5 timed without it",
                      file=sys.stderr)
                w5.close()
                DeviceStep.FUSE = False
                blk, w5 = workload_block(c5, 16, rank, world, dev, x_steps, x_warm, use_graph, peak)
                blk["candidate_emission"] = "MISMATCH against the regular decode: switched off for this block"
            elif ep:
                blk["candidate_emission"] = ep
            if world > 1:
                parity["cfg5"] = sharded_parity(w5, rank, world, dev)
            sched5 = w5.dstep.schedule
            fuse_now = DeviceStep.FUSE and w5.dstep.fused_decode()
            w5.close()
            if args.fuse != "off" and not int(bad.item()):
                # the same step under the OTHER emission setting, timed on the same box: an A/B in every record
                DeviceStep.FUSE = not fuse_now
                alt, wa = workload_block(c5, 16, rank, world, dev, x_steps, x_warm, use_graph, peak, kernels=False)
                other = {"candidate_emission": "on" if wa.dstep.fused_decode() else "off",
                         "ms_per_step": alt["ms_per_step"], "step_hbm_frac": alt["step_hbm_frac"]}
                if wa.dstep.fused_decode():
                    other["parity"] = emission_parity(wa) or "MISMATCH against the regular decode"
                wa.close()
                mism = isinstance(other.get("parity"), str) and other["parity"].startswith("MISMATCH")
                if args.fuse == "auto" and other["ms_per_step"] < blk["ms_per_step"] and not mism:
                    # start-up calibration, as a deployment would do it: the faster setting is the block's number
                    mine = {"candidate_emission": "on" if fuse_now else "off", "ms_per_step": blk["ms_per_step"],
                            "step_hbm_frac": blk["step_hbm_frac"]}
                    blk["ms_per_step"], blk["step_hbm_frac"] = other["ms_per_step"], other["step_hbm_frac"]
                    blk["value"] = 16 * world / (blk["ms_per_step"] * 1e-3)
                    blk["schedule"], blk["gpu_launches_per_step"] = alt["schedule"], alt["gpu_launches_per_step"]
                    blk["candidate_emission"] = other.get("parity", "off")
                    other = mine
                blk["emission_setting"] = ("on" if (blk.get("candidate_emission") or "off") != "off" else "off") + \
                                          (" (the faster of the two on this box)" if args.fuse == "auto" else " (--fuse)")
                blk["other_emission_setting"] = other
                DeviceStep.FUSE = fuse_now
            if world > 1 and sched5 != "nccl":
                # the schedule north_star names, measured beside the in-kernel exchange: one NCCL all-reduce of
                # the normalisers between the count and the main launch, one of the exact totals next to decode
                nb, wn = workload_block(c5, 16, rank, world, dev, x_steps, x_warm, use_graph, peak, schedule="nccl")
                parity["cfg5_nccl"] = sharded_parity(wn, rank, world, dev)
                wn.close()
                blk["nccl_schedule"] = {k: nb[k] for k in ("ms_per_step", "value", "step_hbm_frac", "schedule", "launch",
                                                           "kernels", "gpu_launches_per_step")}
            extra["cfg5"] = blk
        shapes = {}
        for name in ("cfg1", "cfg3", "cfg4", "advent"):
            if name == cfg.name:
                continue
            c = synthetic.CONFIGS[name]
            blk, wx = workload_block(c, c.batch, rank, world, dev, x_steps, x_warm, use_graph, peak, kernels=False)
            wx.close()
            shapes[name] = {k: blk[k] for k in ("workload", "ms_per_step", "value", "step_hbm_frac", "schedule")}
        extra["shapes"] = shapes

    if rank != 0:
        return
    roof = None
    if ms_kernel:
        achieved = loss_bytes / (ms_kernel * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(cfg.name)
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": "detloss_kernel (fused sigmoid-clamp-focal fwd+bwd + gather-L1 heads)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "bytes_per_launch": loss_bytes, "us_per_launch": ms_kernel * 1e3, "peak_source": peak_src}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": config_dict(cfg, batch, world),
        "warmup_graph_steps": n_sets if (use_graph and steps >= n_sets) else 0,
        "run": {"buffer_sets": f"rotating {n_sets} buffer sets of {set_bytes / 2**20:.1f} MiB "
                               f"({n_sets * set_bytes / 2**20:.0f} MiB > 126 MiB L2): HBM-cold every step",
                "launch": launch_mode, "schedule": schedule},
        "step_algorithmic_bytes": step_bytes,
        "step_hbm_frac": step_bytes / (ms_step * 1e-3) / 1e9 / peak,
        "roofline": roof,
        "e2e": e2e,
        "e2e_dense_targets": e2e_dense,
        "e2e_targets_only": e2e_tonly,
        "e2e_eager": e2e_eager,
        "gpu_launches": launches_per_step * steps,
        "clocks": clk,
    }
    if ep_main is not None:
        line["candidate_emission"] = ep_main or "MISMATCH against the regular decode"
    line.update(extra)
    if parity:
        line["sharded_parity"] = parity
    if not args.no_cpu_baseline and world == 1:
        v, ms, done, cores = time_cpu(cfg, batch, 60, 2, budget_s=12.0)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{done} steps of one {cfg.name} batch ({batch} samples): oracle port of the "
                                          f"reference's PyTorch-CPU path, {ms:.1f} ms/step"}
    print(json.dumps(line), flush=True)


def run_e2e(cfg, batch, rank, world, dev, steps, warmup, mode="dense", graphed=True):
    """the step through the plugin API with HOST inputs.  mode: 'dense' -- head maps and the dataset's dense targets
    from pinned host memory; 'boxes' -- head maps and object lists (the targets are rasterised on the device,
    SURVEY 8f N2); 'targets_only' -- the head maps stay on the device, as they do behind the reference's backbone
    (train.py:148-150 moves only the batch), and only the targets (object lists when rasterisable) come from the host."""
    from cnhead import synthetic, sharded
    from cnhead.feeder import HostFeeder
    from cnhead import functional as F
    from losses.centernet import DetectionLoss
    from backends.decode import decode_detection
    kw = synthetic.loss_kwargs(cfg)
    crit = DetectionLoss(**kw) if world == 1 else sharded.make_sharded_loss(DetectionLoss)(**kw)
    from_boxes = mode == "boxes" or (mode == "targets_only" and not cfg.angle)
    heads_on_device = mode == "targets_only"
    n_host = 4
    host, dev_heads = [], []
    for i in range(n_host):
        d = synthetic.make_inputs(cfg, batch=batch, hm_sigma=2.0, seed_offset=20 + i, sample_offset=rank * batch)
        bt = d["batch"]
        if from_boxes:
            # the host ships object lists (boxes in heat-map pixels, classes, counts) and the
            # targets are rasterised on the device; the boxes are rebuilt from the synthetic targets.
            cx = (bt["ind"] % cfg.width).float() + bt["reg"][..., 0]
            cy = (bt["ind"] // cfg.width).float() + bt["reg"][..., 1]
            w, h = bt["wh"][..., 0], bt["wh"][..., 1]
            g = torch.Generator().manual_seed(77 + i)
            bt = {"boxes": torch.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], dim=-1).contiguous(),
                  "classes": torch.randint(0, cfg.classes, bt["ind"].shape, generator=g, dtype=torch.int32),
                  "n_obj": bt["reg_mask"].sum(1).to(torch.int32)}
        if heads_on_device:
            dev_heads.append({k: v.to(dev) for k, v in d["output"].items()})
            host.append((bt,))
        else:
            host.append((d["output"], bt))
    # pinned host staging carved from one large page-locked arena (steady 50 GB/s H2D; separate small
    # pin_memory() allocations copy at 20-40 GB/s depending on the box: tools/h2d_probe.py)
    host = HostFeeder.pinned_sets(host)
    h2d = sum(v.numel() * v.element_size() for grp in host[0] for v in grp.values())
    def head_step(o, b):
        """the calls a user of the plugin modules makes for one step (uda/base.py:31-82 around the head path)"""
        out = {k: v.detach().requires_grad_(True) for k, v in o.items()}
        work = dict(out)
        if from_boxes:
            b = F.raster_targets(b["boxes"], b["classes"], b["n_obj"], cfg.classes, cfg.height, cfg.width)
        loss, stats = crit(work, b)
        loss.backward()
        dets = decode_detection(work["hm"], work["wh"].detach(), work["reg"].detach(), K=cfg.K, rotated=cfg.rotated)
        return {"loss": loss.detach().reshape(1), "dets": dets, "grad_hm": out["hm"].grad}

    d2h = batch * cfg.K * (7 if cfg.rotated else 6) * 4 + 4
    if graphed:
        from cnhead.graphed import HostStep
        hstep = HostStep(head_step, dev, fetch=("loss", "dets"), depth=2)
        feeder = hstep.feeder

        def stage(i):
            hstep.stage(*host[i % n_host])

        def step(i):
            res = hstep.run(resident=(dev_heads[i % n_host],) if heads_on_device else None)
            stage(i + 1)                                    # H2D of the NEXT step rides the copy engine under this one
            res.wait()                                      # (queued while the graph runs) the caller reads loss and
            return res                                      # detections every step
    else:
        feeder = HostFeeder(dev, depth=2)
        dets_host = torch.empty(batch, cfg.K, 7 if cfg.rotated else 6).pin_memory()
        loss_host = torch.empty(1).pin_memory()

        def stage(i):
            feeder.put(*host[i % n_host])

        def step(i):
            stage(i + 1)                                    # H2D of the NEXT step rides the copy engine under this one
            got = feeder.get()
            if heads_on_device:
                o, b = dev_heads[i % n_host], got[0]
            else:
                o, b = got
            res = head_step(o, b)
            dets_host.copy_(res["dets"], non_blocking=True)
            loss_host.copy_(res["loss"], non_blocking=True)
            feeder.release()
            torch.cuda.current_stream().synchronize()      # the caller reads the results every step

    e2e_steps = min(steps, 1000)
    w = min(warmup, 10)
    stage(0)
    prime = 2 * n_host if graphed else 0                    # every (slot, resident set) pair meets once: graphs captured
    for i in range(prime + w):
        step(i)
    torch.cuda.current_stream().wait_stream(feeder.copy_stream)   # the primed copy of the next step is outside the region
    aligned_start(world, dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(e2e_steps):                              # exactly K puts (H2D) and K gets/compute/D2H inside
        step(prime + w + i)
    torch.cuda.current_stream().wait_stream(feeder.copy_stream)
    e1.record()
    barrier(world)
    ms = max_over_ranks(e0.elapsed_time(e1), world, dev)
    what = {"dense": "head maps + dense targets", "boxes": "head maps + object lists",
            "targets_only": "targets only (" + ("object lists" if from_boxes else "dense") + "); head maps device-resident"}[mode]
    return {"value": batch * world * e2e_steps / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": d2h, "ms_per_step": ms / e2e_steps, "steps": e2e_steps,
            "h2d_GBps": h2d / (ms / e2e_steps * 1e-3) / 1e9, "host_inputs": what,
            "launch": "one CUDA graph per feeder slot (cnhead.graphed.HostStep: the plugin calls captured once, replayed)"
                      if graphed else "eager plugin calls",
            "api": ("cnhead.graphed.HostStep = " if graphed else "") +
                   "cnhead.feeder.HostFeeder (double-buffered H2D from pinned memory on a copy stream) + "
                   + ("cnhead.functional.raster_targets (targets rasterised on the device from object lists) + "
                      if from_boxes else "") +
                   "losses.centernet.DetectionLoss + loss.backward() + backends.decode.decode_detection + D2H of "
                   "loss and detections, host-synchronised every step"}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    from cnhead import synthetic
    cfg = synthetic.CONFIGS[args.config]
    batch = args.batch or (cfg.batch if cfg.name != "cfg5" else 16)
    if getattr(cfg, "advent", False):              # kernels of the ADVENT step: no decode, no plugin-level e2e / CPU leg
        args.no_e2e, args.no_cpu_baseline = True, True
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "--config advent times the CUDA kernels of the ADVENT step only"}))
            return
    if args.impl == "reference":
        run_reference(args, cfg, batch, rank, world)
        return
    if world > 1:
        # one process per GPU, each on its own slice of the host's cores: eight host threads that all launch kernels and
        # stage pinned buffers otherwise migrate over the same few cores (e2e at N = 8)
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = max(1, len(cores) // world)
            os.sched_setaffinity(0, cores[local_rank * per:(local_rank + 1) * per] or cores)
        except (AttributeError, OSError):
            pass
        torch.cuda.set_device(local_rank)
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, cfg, batch, rank, local_rank, world)
    finally:
        if world > 1:
            torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
