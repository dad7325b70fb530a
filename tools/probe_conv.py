"""Timing probes: forward pass at config 2 with parts of the conv kernel disabled (debug flags)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sr_gan_fd_b200 as b200
from sr_gan_fd_b200 import lib
L = lib.load()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = b200.rrdbnet_x4(num_blocks=23).to(dev).eval()
lr = torch.rand(16, 3, 64, 64, device=dev)
def t_fwd():
    with torch.no_grad():
        for _ in range(2): net(lr)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): net(lr)
        e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5
names = {0: "normal", 1: "no epilogue traffic", 2: "no MMA", 3: "no MMA, no epilogue traffic", 16: "no dependency waits (wrong results)",
         19: "no MMA, no epilogue traffic, no dependency waits"}
for f in [0, 1, 2, 3, 16, 19]:
    L.b200sr_debug_set(f)
    print(f"flags {f:2d} {names[f]:50s} fwd {t_fwd():7.3f} ms", flush=True)
L.b200sr_debug_set(0)
