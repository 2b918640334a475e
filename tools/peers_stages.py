"""In-kernel stage timestamps of the fused loss launch with the peer exchange (2+ GPUs, torchrun)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "centernet-uda_b200"), ROOT]
import torch, torch.distributed as dist, bench
from cnhead import _lib as L, synthetic, sharded
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
cfg = synthetic.CONFIGS["cfg2"]
dev = torch.device("cuda", rank)
sets = [bench.BufferSet(synthetic.make_inputs(cfg, batch=16, hm_sigma=2.0, seed_offset=i, sample_offset=rank * 16), cfg, dev) for i in range(4)]
d = bench.DeviceStep(sets, cfg, world, None)
lib = d.lib
lib.cnh_debug_set_buffer.argtypes = [C.c_void_p]
dbg = torch.zeros(4096, 16, dtype=torch.int64, device=dev)
def fused(i, peers):
    a = d.loss_args[i]
    a.scalars = sets[i].scalars.data_ptr()
    a.flags = L.FLAG_DEFER_TOTALS if peers else 0
    if peers:
        L.check(lib.cnh_detloss_fused_peers(C.byref(a), C.byref(d.box.c), d.ws_loss.data_ptr(), d.ws_loss.numel(), L.stream_ptr()), "p")
        L.check(lib.cnh_detloss_peers_finalize(C.byref(a), C.byref(d.box.c), d.ws_loss.data_ptr(), d.ws_loss.numel(), L.stream_ptr()), "f")
    else:
        L.check(lib.cnh_detloss_fused(C.byref(a), d.ws_loss.data_ptr(), d.ws_loss.numel(), L.stream_ptr()), "s")
for peers in (False, True):
    for it in range(6):
        fused(it % 4, peers)
    torch.cuda.synchronize(); dist.barrier()
    lib.cnh_debug_set_buffer(dbg.data_ptr())
    for rep in range(3):
        dbg.zero_(); torch.cuda.synchronize(); dist.barrier()
        fused(rep, peers)
        torch.cuda.synchronize()
        t = dbg.cpu()
        used = t[:, 0] != 0
        t = t[used]
        t0 = t[:, 0].min()
        if rank == 0 and rep == 2:
            print(f"--- peers={peers}: {int(used.sum())} CTAs, span {(t[:, :8].max() - t0).item() / 1e3:.2f} us")
            for sl in range(8):
                col = t[:, sl]; ok = col != 0
                if ok.any():
                    rel = (col[ok] - t0).float() / 1e3
                    print(f"  stamp {sl}: n={int(ok.sum()):4d} min {rel.min():6.2f} median {rel.median():6.2f} max {rel.max():6.2f}")
    lib.cnh_debug_set_buffer(None)
dist.destroy_process_group()
