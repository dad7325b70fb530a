"""torchrun --nproc-per-node 2 tools/dp_check.py : data-parallel equivalence on real GPUs.
2 ranks x 4 images (NCCL gradient all-reduce) must equal 1 rank x 8 images."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist, torch.nn.functional as F
import sr_gan_fd_b200 as b200
from sr_gan_fd_b200 import dist as b200dist

def rel_l2(a, b):  # (tools do not import the oracle: it is test infrastructure)
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
net = b200.rrdbnet_x4(num_blocks=3)
# (seeded reference init; the check is DP vs single replica, the weights themselves do not matter)
net = net.to(dev).train()
b200dist.broadcast_parameters(net)
g = torch.Generator().manual_seed(5)
lr = torch.rand(4 * world, 3, 32, 32, generator=g).to(dev)
gt = torch.rand(4 * world, 3, 128, 128, generator=g).to(dev)
# single-replica reference on every rank: whole global batch, no reducer
F.l1_loss(net(lr), gt).backward()
ref = torch.cat([p.grad.flatten() for p in net.parameters()]).clone()
net.zero_grad(set_to_none=True)
red = b200dist.make_data_parallel(net)
for it in range(3):
    net.zero_grad(set_to_none=True)
    sl = slice(4 * rank, 4 * rank + 4)
    F.l1_loss(net(lr[sl]), gt[sl]).backward()
got = torch.cat([p.grad.flatten() for p in net.parameters()])
torch.cuda.synchronize()
err = rel_l2(got, ref)
print(f"rank {rank}: DP vs single-replica flat-grad rel-L2 {err:.3e}; buckets {len(red.buckets_seen)//3} per step", flush=True)
assert err < 2e-3, err
dist.barrier()
dist.destroy_process_group()
