"""Tiny fwd+bwd (+ dx, + inference at a second geometry) of the drop-in generator for compute-sanitizer runs:
    compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck python tools/sanitize_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import sr_gan_fd_b200 as b200

dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = b200.rrdbnet_x4(num_blocks=int(os.environ.get("NB", 1))).to(dev).train()
x = torch.rand(2, 3, 24, 16, device=dev, requires_grad=True)
gt = torch.rand(2, 3, 96, 64, device=dev)
for _ in range(2):
    net.zero_grad(set_to_none=True)
    F.l1_loss(net(x), gt).backward()
net.eval()
with torch.no_grad():
    y = net(torch.rand(1, 3, 40, 9, device=dev))
torch.cuda.synchronize()
print("ok", float(y.mean()), float(x.grad.abs().sum()))
