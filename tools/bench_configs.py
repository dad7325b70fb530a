"""Secondary BASELINE.json configs on one GPU: C1 (16 x 32x32 inference), C4 (1 x 1024x1024 frame, whole and halo-tiled)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sr_gan_fd_b200 as b200
from sr_gan_fd_b200 import tile

def rel_l2(a, b):  # (tools do not import the oracle: it is test infrastructure)
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = b200.rrdbnet_x4(num_blocks=23).to(dev).eval()

def timeit(fn, iters, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

out = {}
with torch.no_grad():
    lr = torch.rand(16, 3, 32, 32, device=dev)
    t = timeit(lambda: net(lr), 20)
    p = net._runtime().last_plan
    out["C1_16x32x32"] = {"ms": t, "out_mpix_per_s": 16 * 128 * 128 / t / 1e3, "tflops": p.flops_fwd / t / 1e9}
    frame = torch.rand(1, 3, 1024, 1024, device=dev)
    t = timeit(lambda: net(frame), 3, warm=1)
    p = net._runtime().last_plan
    whole = net(frame)
    out["C4_1x1024x1024_whole"] = {"ms": t, "out_mpix_per_s": 4096 * 4096 / t / 1e3, "tflops": p.flops_fwd / t / 1e9,
                                   "workspace_gb": p.workspace_bytes / 2**30}
    for halo in (8, 16, 32):
        f = lambda: tile.tiled_forward(net, frame, 4, num_bands=8, halo=halo)[0]
        t = timeit(f, 2, warm=1)
        err = rel_l2(f(), whole)
        out[f"C4_tiled_8bands_halo{halo}"] = {"ms": t, "out_mpix_per_s": 4096 * 4096 / t / 1e3, "rel_l2_vs_whole": err,
                                              "redundant_rows": tile.redundant_fraction(1024, 8, halo)}
print(json.dumps(out, indent=1))
