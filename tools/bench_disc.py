"""U-Net discriminator (BSRGAN/model.py:91-167) on one B200: the native path of sr_gan_fd_b200.discriminator vs the SAME module's
stock torch path (= the reference's op sequence on cuDNN: fp32/TF32, fp16 autocast as the reference scripts run it, and fp16
autocast + channels_last).  16 images of 256 x 256 (the HR side of BASELINE configs[4]); three passes as the GAN step uses them:
forward only, discriminator update (forward + all gradients but the input's), generator update (forward + input gradient, frozen D)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sr_gan_fd_b200.discriminator import discriminator_unet

dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
N, H = int(os.environ.get("DISC_BATCH", "16")), int(os.environ.get("DISC_SIZE", "256"))
torch.manual_seed(0)
d = discriminator_unet(in_channels=3, out_channels=1, channels=64).to(dev).train()
x = torch.rand(N, 3, H, H, device=dev)
dy = torch.randn(N, 1, H, H, device=dev) / (N * H * H)


def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def passes(autocast=None, channels_last=False):
    xin = x.contiguous(memory_format=torch.channels_last) if channels_last else x
    ctx = lambda: torch.autocast("cuda", dtype=autocast or torch.float16, enabled=autocast is not None)
    def fwd():
        with torch.no_grad(), ctx():
            return d(xin)
    def d_update():
        for p in d.parameters(): p.requires_grad = True
        d.zero_grad(set_to_none=True)
        with ctx():
            y = d(xin)
        y.float().backward(dy)
    def g_update():
        for p in d.parameters(): p.requires_grad = False
        xr = xin.detach().requires_grad_(True)
        with ctx():
            y = d(xr)
        y.float().backward(dy)
        return xr.grad
    return {"fwd_ms": t(fwd), "d_update_ms": t(d_update), "g_update_ms": t(g_update)}


out = {"config": {"batch": N, "size": H}}
d.use_native = True
out["b200"] = passes()
plan = d._runtime().last_plan
out["flops"] = {"fwd": plan.flops_fwd, "d_update": plan.flops_fwd + plan.flops_bwd_d, "g_update": plan.flops_fwd + plan.flops_bwd_g}
out["b200_tflops"] = {k: out["flops"][k.replace("_ms", "")] / (v * 1e-3) / 1e12 for k, v in out["b200"].items()}
d.use_native = False
out["torch_tf32"] = passes()
out["torch_fp16_autocast"] = passes(torch.float16)
d = d.to(memory_format=torch.channels_last)
out["torch_fp16_autocast_channels_last"] = passes(torch.float16, True)
best = {k: min(out[m][k] for m in ("torch_tf32", "torch_fp16_autocast", "torch_fp16_autocast_channels_last")) for k in out["b200"]}
out["speedup_vs_best_torch"] = {k: best[k] / out["b200"][k] for k in best}
print(json.dumps(out))
