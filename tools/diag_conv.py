"""GPU diagnostic: probes the tcgen05 conv / wgrad kernels with one-hot weights to localise layout bugs."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from sr_gan_fd_b200 import lib

L = lib.load()
dev = torch.device("cuda", 0)


def run_fwd(x, w, cout):
    n, cin, h, ww = x.shape
    xb = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    y = torch.zeros(n, h, ww, cout, dtype=torch.bfloat16, device=dev)
    scratch = torch.empty(L.b200sr_conv3x3_scratch_bytes(cin, cout), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    lib.check(L.b200sr_conv3x3_fwd(C.c_void_p(xb.data_ptr()), n, h, ww, cin, cin, C.c_void_p(w.data_ptr()), None, cout, 0,
                                   C.c_void_p(y.data_ptr()), cout, 0, C.c_void_p(scratch.data_ptr()), C.c_void_p(st)))
    torch.cuda.synchronize()
    return y.float().permute(0, 3, 1, 2)


def report(name, got, ref):
    err = (got - ref).abs().max().item()
    rel = ((got - ref).norm() / (ref.norm() + 1e-30)).item()
    print(f"{name:48s} max|err| {err:10.4e} rel {rel:10.4e} {'OK' if rel < 1e-2 else 'BAD'}", flush=True)
    return rel < 1e-2


torch.manual_seed(0)
n, cin, h, w, cout = 1, 64, 16, 8, 32
x = torch.randn(n, cin, h, w, device=dev).to(torch.bfloat16).float()
for (dy, dx) in [(1, 1), (0, 1), (2, 1), (1, 0), (1, 2), (0, 0), (2, 2)]:
    wt = torch.zeros(cout, cin, 3, 3, device=dev)
    for co in range(cout):
        wt[co, co, dy, dx] = 1.0  # out[co] = x[co] shifted
    ok = report(f"one-hot tap ({dy},{dx}) ch-identity", run_fwd(x, wt, cout), F.conv2d(x, wt, padding=1))
wt = torch.zeros(cout, cin, 3, 3, device=dev)
for co in range(cout):
    wt[co, 63 - co, 1, 1] = 1.0
report("centre tap, channel reversal (k 32..63)", run_fwd(x, wt, cout), F.conv2d(x, wt, padding=1))
wt = (torch.randn(cout, cin, 3, 3, device=dev) * 0.1).to(torch.bfloat16).float()
report("random weights 64->32 one tile", run_fwd(x, wt, cout), F.conv2d(x, wt, padding=1))
x2 = torch.randn(2, 192, 40, 24, device=dev).to(torch.bfloat16).float()
wt2 = (torch.randn(64, 192, 3, 3, device=dev) * 0.05).to(torch.bfloat16).float()
report("random weights 192->64 multi tile", run_fwd(x2, wt2, 64), F.conv2d(x2, wt2, padding=1))

# wgrad probes
def run_wgrad(x, dy):
    n, cin, h, ww = x.shape
    cout = dy.shape[1]
    xb = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    dyb = dy.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    dw = torch.full((cout, cin, 3, 3), 9.0, device=dev)
    scratch = torch.empty(L.b200sr_conv3x3_wgrad_scratch_bytes(cin, cout), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    lib.check(L.b200sr_conv3x3_wgrad(C.c_void_p(xb.data_ptr()), n, h, ww, cin, cin, C.c_void_p(dyb.data_ptr()), cout, cout,
                                     C.c_void_p(dw.data_ptr()), C.c_void_p(scratch.data_ptr()), C.c_void_p(st)))
    torch.cuda.synchronize()
    return dw

for (cin_, cout_, nn, hh, ww_) in [(64, 32, 1, 16, 8), (128, 64, 1, 16, 8), (128, 160, 2, 32, 24), (64, 16, 1, 16, 16)]:
    xx = torch.randn(nn, cin_, hh, ww_, device=dev).to(torch.bfloat16).float()
    dd = torch.randn(nn, cout_, hh, ww_, device=dev).to(torch.bfloat16).float()
    ref = torch.nn.grad.conv2d_weight(xx, (cout_, cin_, 3, 3), dd, padding=1)
    got = run_wgrad(xx, dd)
    report(f"wgrad cin {cin_} cout {cout_} {nn}x{hh}x{ww_}", got, ref)
    if cin_ == 64 and cout_ == 32:
        for t in range(9):
            e = ((got[:, :, t // 3, t % 3] - ref[:, :, t // 3, t % 3]).norm() / ref[:, :, t // 3, t % 3].norm()).item()
            print(f"    tap {t // 3},{t % 3}: rel {e:.3e}")
