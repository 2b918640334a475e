"""Pinned H2D bandwidth of the step's tensors: separate allocations vs slices of one large pinned arena,
touched or untouched (IOMMU / page-size effects differ between boxes)."""
import torch
torch.cuda.init()
dev = torch.device("cuda")
sizes = [6291456, 2097152, 2097152, 6291456, 19200, 2400, 19200, 19200]
def bw(host_list, dev_list, n=100):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(5):
            for h, d in zip(host_list, dev_list): d.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            for h, d in zip(host_list, dev_list): d.copy_(h, non_blocking=True)
        e1.record(); torch.cuda.synchronize()
    return round(sum(h.numel() for h in host_list) * n / (e0.elapsed_time(e1) * 1e-3) / 1e9, 1)
ds = [torch.empty(n, dtype=torch.uint8, device=dev) for n in sizes]
def carve(arena):
    out, off = [], 0
    for n in sizes:
        out.append(arena[off:off + n]); off += (n + 4095) // 4096 * 4096
    return out
for rep in range(2):
    sep = [torch.empty(n, dtype=torch.uint8).pin_memory() for n in sizes]
    print("separate, untouched      :", bw(sep, ds), "GB/s")
    for h in sep: h.fill_(1)
    print("separate, touched        :", bw(sep, ds), "GB/s")
    src = [torch.randint(0, 255, (n,), dtype=torch.uint8) for n in sizes]
    sep2 = [t.pin_memory() for t in src]
    print("separate, t.pin_memory() :", bw(sep2, ds), "GB/s")
    for mb in (32, 256):
        arena = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
        print(f"arena {mb:3d} MiB, untouched :", bw(carve(arena), ds), "GB/s")
        arena.fill_(1)
        print(f"arena {mb:3d} MiB, touched   :", bw(carve(arena), ds), "GB/s")
        del arena
