"""Host-side cost of one synchronous (loss read back every step) training step through the module API, config 2."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import sr_gan_fd_b200 as b200
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = b200.rrdbnet_x4(num_blocks=23).to(dev).train()
lr = torch.rand(16, 3, 64, 64, device=dev); gt = torch.rand(16, 3, 256, 256, device=dev)
for _ in range(3):
    net.zero_grad(set_to_none=True); F.l1_loss(net(lr), gt).backward()
torch.cuda.synchronize()
acc = {}
def lap(name, t0):
    t1 = time.perf_counter(); acc[name] = acc.get(name, 0.0) + (t1 - t0); return t1
N = 10
for _ in range(N):
    torch.cuda.synchronize()
    t = time.perf_counter()
    net.zero_grad(set_to_none=True); t = lap("zero_grad", t)
    y = net(lr); t = lap("forward (host enqueue)", t)
    loss = F.l1_loss(y, gt); t = lap("loss (host enqueue)", t)
    loss.backward(); t = lap("backward (host, incl. AccumulateGrad)", t)
    v = loss.item(); t = lap("loss.item() (wait for the GPU)", t)
tot = sum(acc.values())
for k, v in acc.items(): print(f"{k:42s} {v / N * 1e3:8.3f} ms")
print(f"{'TOTAL per step':42s} {tot / N * 1e3:8.3f} ms")
