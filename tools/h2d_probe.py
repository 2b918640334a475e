import os, glob, torch
torch.cuda.init()
dev = torch.device("cuda")
p = torch.cuda.get_device_properties(0)
pci = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
try: gnode = int(open(f"/sys/bus/pci/devices/{pci}/numa_node").read())
except Exception as e: gnode = repr(e)
print("gpu pci", pci, "numa node", gnode, "allowed cpus", len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:4], "...")
def cpulist(s):
    out = set()
    for part in s.strip().split(","):
        if not part: continue
        a, _, b = part.partition("-")
        out |= set(range(int(a), int(b or a) + 1))
    return out
nodes = {}
for d in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
    nodes[int(d.rsplit("node", 1)[1])] = cpulist(open(d + "/cpulist").read())
print("nodes:", {k: len(v) for k, v in nodes.items()})
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
    return sum(h.numel() for h in host_list) * n / (e0.elapsed_time(e1) * 1e-3) / 1e9
ds = [torch.empty(n, dtype=torch.uint8, device=dev) for n in sizes]
allowed = os.sched_getaffinity(0)
for node, cpus in nodes.items():
    use = cpus & allowed
    if not use: print("node", node, "no allowed cpus"); continue
    os.sched_setaffinity(0, use)
    hs = [torch.empty(n, dtype=torch.uint8).pin_memory() for n in sizes]
    for h in hs: h.fill_(1)
    print("node", node, "cpus", len(use), "H2D 8 tensors:", round(bw(hs, ds), 1), "GB/s")
    del hs
os.sched_setaffinity(0, allowed)
