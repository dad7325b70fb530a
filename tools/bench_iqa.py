"""Fused Y-channel PSNR + SSIM (libb200sr) vs the same metrics through stock torch ops (the reference's op sequence,
sr_gan_fd_b200.iqa's torch path) on one B200.  HBM roofline: each metric reads both fp32 RGB frames once = 24 B per pixel."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sr_gan_fd_b200 import iqa

dev = torch.device("cuda", 0)
out = []
for shape in [(16, 3, 256, 256), (1, 3, 4096, 4096)]:
    a = torch.rand(*shape, device=dev); b = (a + 0.05 * torch.randn_like(a)).clamp(0, 1)
    psnr, ssim = iqa.PSNR(4, True), iqa.SSIM(4, True)
    def fused(): return psnr(a, b), ssim(a, b)
    def stock():  # the torch path of the same modules = the reference's op sequence
        ra, rb = iqa._prepare(a, b, 4, True)
        p = 10 * torch.log10(255.0 ** 2 / torch.mean((ra * 255.0 - rb * 255.0) ** 2 + 1e-8, dim=[1, 2, 3]))
        ss = iqa.SSIM.forward(type("S", (), dict(crop_border=4, only_test_y_channel=False, window_size=11, _win=ssim._win,
                                                 gaussian_kernel_window=ssim.gaussian_kernel_window))(), ra / 1.0, rb / 1.0) if False else None
        return p
    def t(fn, n=10):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    # stock = force the torch path by handing the modules fp64-incompatible "native" conditions: use half-open trick (window 11 but only_y via torch)
    class _Torch(torch.nn.Module):
        def forward(self, x, y):
            rx, ry = iqa._prepare(x, y, 4, True)
            p = 10 * torch.log10(255.0 ** 2 / torch.mean((rx * 255.0 - ry * 255.0) ** 2 + 1e-8, dim=[1, 2, 3]))
            rx, ry = rx * 255.0, ry * 255.0
            win = torch.from_numpy(ssim.gaussian_kernel_window).view(1, 1, 11, 11).to(rx)
            F = torch.nn.functional
            mr, md = F.conv2d(rx, win), F.conv2d(ry, win)
            vr, vd = F.conv2d(rx * rx, win) - mr ** 2, F.conv2d(ry * ry, win) - md ** 2
            cv = F.conv2d(rx * ry, win) - mr * md
            c1, c2 = 6.5025, 58.5225
            s = ((2 * mr * md + c1) * (2 * cv + c2)) / ((mr ** 2 + md ** 2 + c1) * (vr + vd + c2))
            return p, torch.mean(s, [1, 2, 3]).float()
    tm = _Torch()
    ms_f, ms_t = t(fused), t(lambda: tm(a, b), 3)
    pf, sf = fused(); pt, st = tm(a, b)
    px = shape[0] * shape[2] * shape[3]
    out.append({"shape": list(shape), "fused_ms": ms_f, "torch_ops_ms": ms_t, "speedup": ms_t / ms_f,
                "fused_GBps_algorithmic": 2 * 24 * px / (ms_f * 1e-3) / 1e9,
                "max_abs_dpsnr": float((pf - pt).abs().max()), "max_abs_dssim": float((sf - st).abs().max())})
print(json.dumps(out))
