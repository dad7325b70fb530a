"""One discriminator update + one generator-update pass of the native U-Net discriminator (for ncu launch lists / captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sr_gan_fd_b200.discriminator import discriminator_unet
dev = torch.device("cuda", 0)
N, H = int(os.environ.get("DISC_BATCH", "16")), int(os.environ.get("DISC_SIZE", "256"))
torch.manual_seed(0)
d = discriminator_unet(in_channels=3, out_channels=1, channels=64).to(dev).train()
x = torch.rand(N, 3, H, H, device=dev)
dy = torch.randn(N, 1, H, H, device=dev) / (N * H * H)
for it in range(int(os.environ.get("DISC_ITERS", "2"))):
    for p in d.parameters(): p.requires_grad = True
    d.zero_grad(set_to_none=True)
    d(x).backward(dy)
    for p in d.parameters(): p.requires_grad = False
    xr = x.detach().requires_grad_(True)
    d(xr).backward(dy)
torch.cuda.synchronize()
print("ok")
