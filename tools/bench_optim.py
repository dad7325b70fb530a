"""Timing of the step after the hot path: reference stack (torch Adam + AveragedModel avg_fn loop) vs FusedAdamEMA."""
import sys, os, time, copy, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.optim.swa_utils import AveragedModel
import sr_gan_fd_b200 as b200
from sr_gan_fd_b200.optim import FusedAdamEMA
dev = torch.device("cuda", 0)
torch.manual_seed(0)
a = b200.rrdbnet_x4(num_blocks=23).to(dev); b = copy.deepcopy(a)
decay = 0.99998
avg = lambda e, p, n: (1 - decay) * e + decay * p
ema_a = AveragedModel(a, avg_fn=avg); ema_b = AveragedModel(b, avg_fn=avg)
opt_a = torch.optim.Adam(a.parameters(), 2e-4, (0.9, 0.99), 1e-8, 0.0)
opt_b = FusedAdamEMA(b.parameters(), 2e-4, (0.9, 0.99), 1e-8, 0.0, ema_model=ema_b, ema_decay=decay)
for m in (a, b):
    for p in m.parameters(): p.grad = torch.randn_like(p) * 1e-3
def t(fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
def ref_step(): opt_a.step(); ema_a.update_parameters(a)
ms_ref = t(ref_step); ms_fused = t(opt_b.step)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); [opt_b.step() for _ in range(10)]; e1.record(); torch.cuda.synchronize()
gpu_ms = e0.elapsed_time(e1) / 10
n = sum(p.numel() for p in b.parameters())
print(json.dumps({"reference_adam_plus_ema_loop_ms": ms_ref, "fused_wall_ms": ms_fused, "fused_gpu_ms": gpu_ms,
                  "params": n, "fused_GBps": n * 36 / (gpu_ms * 1e-3) / 1e9}))
