"""Software profile of the chain kernel: where does each warp role of a CTA wait? (forward pass, config 2)"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sr_gan_fd_b200 as b200
from sr_gan_fd_b200 import lib
L = lib.load()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = b200.rrdbnet_x4(num_blocks=23).to(dev).eval()
lr = torch.rand(16, 3, 64, 64, device=dev)
with torch.no_grad():
    for _ in range(2): net(lr)
    L.b200sr_debug_set(64 | int(os.environ.get("DBG", 0)))
    net(lr); torch.cuda.synchronize()
    L.b200sr_debug_set(0)
buf = (C.c_ulonglong * (160 * 12))()
lib.check(L.b200sr_debug_read_profile(buf, 160 * 12))
full = np.array(buf, dtype=np.float64)
a = full[:148 * 12].reshape(148, 12)
extra = full[160 * 12 - 320: 160 * 12 - 320 + 296].reshape(148, 2)
names = ["prod: dependency wait", "prod: wait A slot free", "prod: wait W granules free", "prod: TOTAL",
         "mma: wait accumulator free", "mma: wait A tile landed", "mma: (weight-stage waits are inside the issue figure)", "mma: TOTAL",
         "epi: dependency wait + barriers", "epi: wait accumulator ready", "epi: TOTAL", "items per CTA"]
clk = 1.965e9
print("per-CTA average over 148 CTAs (ms at 1.965 GHz) | min | max")
for i, n in enumerate(names):
    col = a[:, i]
    if i == 11: print(f"{n:34s} {col.mean():9.1f} {col.min():9.0f} {col.max():9.0f}")
    else: print(f"{n:34s} {col.mean()/clk*1e3:9.3f} {col.min()/clk*1e3:9.3f} {col.max()/clk*1e3:9.3f}")
print(f"{'mma: entry parameter fetch':34s} {extra[:,0].mean()/clk*1e3:9.3f}")
print(f"{'mma: issue incl. weight-stage waits':34s} {extra[:,1].mean()/clk*1e3:9.3f}")
