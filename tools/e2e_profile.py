"""cProfile of the e2e step (host side) -- where does the per-step Python time go?"""
import cProfile, pstats, os, sys, time, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "centernet-uda_b200"), ROOT]
import torch, bench
from cnhead import synthetic
from cnhead.feeder import HostFeeder
from losses.centernet import DetectionLoss
from backends.decode import decode_detection
cfg = synthetic.CONFIGS["cfg2"]; batch = 16; dev = torch.device("cuda", 0)
kw = synthetic.loss_kwargs(cfg); crit = DetectionLoss(**kw)
host = []
for i in range(4):
    d = synthetic.make_inputs(cfg, batch=batch, hm_sigma=2.0, seed_offset=20 + i)
    host.append((d["output"], d["batch"]))
host = HostFeeder.pinned_sets(host)
dets_host = torch.empty(batch, cfg.K, 6).pin_memory(); loss_host = torch.empty(1).pin_memory()
feeder = HostFeeder(dev, depth=2)
T = {}
def lap(name, t0):
    t1 = time.perf_counter(); T[name] = T.get(name, 0.0) + (t1 - t0); return t1
def step(i, timing=False):
    t = time.perf_counter()
    feeder.put(*host[(i + 1) % 4]);                       t = lap("put", t) if timing else t
    o, b = feeder.get()
    out = {k: v.detach().requires_grad_(True) for k, v in o.items()}
    work = dict(out);                                     t = lap("get+leaf", t) if timing else t
    loss, stats = crit(work, b);                          t = lap("loss fwd", t) if timing else t
    loss.backward();                                      t = lap("backward", t) if timing else t
    dets = decode_detection(work["hm"], work["wh"].detach(), work["reg"].detach(), K=cfg.K); t = lap("decode", t) if timing else t
    dets_host.copy_(dets, non_blocking=True); loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
    feeder.release();                                     t = lap("d2h+release", t) if timing else t
    torch.cuda.current_stream().synchronize();            t = lap("sync", t) if timing else t
feeder.put(*host[0])
for i in range(20): step(i)
n = 300
t0 = time.perf_counter()
for i in range(n): step(i, True)
tot = time.perf_counter() - t0
print(f"per step {tot / n * 1e6:.0f} us:", {k: round(v / n * 1e6, 1) for k, v in T.items()})
class _Noop(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x): return x.sum()
    @staticmethod
    def backward(ctx, g): return None
xs = torch.zeros(4, device=dev, requires_grad=True)
def engine_floor(n=300):
    ys = [_Noop.apply(xs) for _ in range(n)]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for y in ys: y.backward()
    return (time.perf_counter() - t0) / n * 1e6
print(f"autograd engine floor (no-op Function, CUDA scalar): {engine_floor():.0f} us per backward()")
if hasattr(torch.autograd, "set_multithreading_enabled"):
    torch.autograd.set_multithreading_enabled(False)
    print(f"  with torch.autograd.set_multithreading_enabled(False): {engine_floor():.0f} us")
    T.clear(); t0 = time.perf_counter()
    for i in range(n): step(i, True)
    tot = time.perf_counter() - t0
    print(f"per step {tot / n * 1e6:.0f} us (single-threaded autograd):", {k: round(v / n * 1e6, 1) for k, v in T.items()})
    torch.autograd.set_multithreading_enabled(True)
pr = cProfile.Profile(); pr.enable()
for i in range(200): step(i)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22); print(s.getvalue()[:5000])
