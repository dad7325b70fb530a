"""Quick CUDA-event timing of the generator fwd / fwd+bwd at BASELINE config 2 (not the contract bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sr_gan_fd_b200 as b200

dev = torch.device("cuda", 0)
torch.manual_seed(0)
nb = int(os.environ.get("NB", 23)); bs = int(os.environ.get("BS", 16)); hw = int(os.environ.get("HW", 64))
net = b200.rrdbnet_x4(num_blocks=nb).to(dev)
lr = torch.rand(bs, 3, hw, hw, device=dev); gt = torch.rand(bs, 3, 4 * hw, 4 * hw, device=dev)

def timeit(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

def fwd():
    with torch.no_grad(): net(lr)
def step():
    net.zero_grad(set_to_none=True)
    torch.nn.functional.l1_loss(net(lr), gt).backward()

net.eval(); t_f = timeit(fwd)
plan = net._runtime().last_plan
print(f"fwd  {t_f:8.3f} ms  {plan.flops_fwd / t_f / 1e9:8.1f} TFLOP/s  ({bs * 16 * hw * hw / t_f / 1e3:.1f} Mpix/s) launches {plan.launches_fwd}")
net.train(); t_s = timeit(step)
plan = net._runtime().last_plan
print(f"step {t_s:8.3f} ms  {(plan.flops_fwd + plan.flops_bwd) / t_s / 1e9:8.1f} TFLOP/s  ({bs / t_s * 1e3:.1f} img/s) launches {plan.launches_fwd + plan.launches_bwd}")
print(f"mem peak {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")
