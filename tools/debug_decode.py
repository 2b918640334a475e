import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "centernet-uda_b200"), ROOT]
import torch, oracle
from cnhead import functional as F
B, C, H, W, K, rotated, sigma = 16, 6, 128, 128, 150, True, 2.0
g = torch.Generator().manual_seed(B * 1000 + C * 100 + H + W + K)
heat = oracle.sigmoid_clamp(torch.randn(B, C, H, W, generator=g) * sigma - 2.19)
wh = torch.rand(B, 3, H, W, generator=g) * 40
wh[:, 2] = torch.randn(B, H, W, generator=g)
reg = torch.rand(B, 2, H, W, generator=g)
for it in range(3):
    dets, inds = F.decode(heat.cuda(), wh.cuda(), reg.cuda(), K=K, rotated=rotated, return_inds=True)
    torch.cuda.synchronize()
    ref, rinds = oracle.decode_stable(heat, wh, reg, K=K, rotated=rotated)
    print(it, "inds equal:", torch.equal(inds.cpu(), rinds), "max diff", (dets.cpu() - ref).abs().max().item())
