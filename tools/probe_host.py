"""Host enqueue cost vs GPU time for the forward pass; CUDA-graph replay removes the host side."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sr_gan_fd_b200 as b200
from sr_gan_fd_b200 import lib
L = lib.load()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = b200.rrdbnet_x4(num_blocks=23).to(dev).eval()
lr = torch.rand(16, 3, 64, 64, device=dev)
with torch.no_grad():
    for _ in range(3): net(lr)
    torch.cuda.synchronize()
    for flags in (0, 15):
        L.b200sr_debug_set(flags)
        net(lr); torch.cuda.synchronize()
        t0 = time.perf_counter(); net(lr); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        print(f"flags {flags}: host enqueue {1e3*(t1-t0):.3f} ms, until done {1e3*(t2-t0):.3f} ms")
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            net(lr)
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s):
                y = net(lr)
        torch.cuda.synchronize()
        for _ in range(2): g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): g.replay()
        e1.record(); torch.cuda.synchronize()
        print(f"flags {flags}: graph replay {e0.elapsed_time(e1)/5:.3f} ms per forward")
L.b200sr_debug_set(0)
