"""Minimal driver for ncu: 2 warm-up steps + 1 step of the BASELINE config-2 workload (fwd+bwd), nothing else."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sr_gan_fd_b200 as b200

dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = b200.rrdbnet_x4(num_blocks=int(os.environ.get("NB", 23))).to(dev).train()
bs = int(os.environ.get("BS", 16)); hw = int(os.environ.get("HW", 64))
lr = torch.rand(bs, 3, hw, hw, device=dev); gt = torch.rand(bs, 3, 4 * hw, 4 * hw, device=dev)
for i in range(int(os.environ.get("STEPS", 3))):
    net.zero_grad(set_to_none=True)
    torch.nn.functional.l1_loss(net(lr), gt).backward()
torch.cuda.synchronize()
print("done")
