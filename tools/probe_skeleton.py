"""Forward chain with parts switched off (timing only): what does the bare per-entry skeleton cost?"""
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
def t(dbg, n=10):
    with torch.no_grad():
        net(lr); L.b200sr_debug_set(dbg); net(lr); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): net(lr)
        e1.record(); torch.cuda.synchronize()
        L.b200sr_debug_set(0)
    return e0.elapsed_time(e1) / n
for name, dbg in [("full", 0), ("no epilogue traffic (1)", 1), ("no MMA (2)", 2), ("no dependency waits (16)", 16), ("no loads (4)", 4),
                  ("no MMA, no epilogue traffic (3)", 3), ("no MMA/epilogue/loads (7)", 7), ("skeleton: no MMA/epilogue/loads/deps (23)", 23),
                  ("skeleton without the proxy fence (23+512)", 23 + 512), ("full without the proxy fence (512) [timing only]", 512), ("full", 0)]:
    print(f"{name:48s} {t(dbg):7.3f} ms")
