"""VGG19 content loss (BSRGAN flavour: 5 nodes, no gradient; ESRGAN flavour: features.34 with d/dsr) on one B200: the tcgen05
chain-kernel path of sr_gan_fd_b200.vgg vs the same module's stock torch path (fp32/TF32, bf16 autocast).  Seeded random-init
VGG19 (no ImageNet weights offline).  16 sr + 16 gt images of 256 x 256."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torchvision.models as models
real = models.vgg19
def seeded(*a, **k):
    torch.manual_seed(1234); return real(weights=None)
models.vgg19 = seeded
from sr_gan_fd_b200 import vgg

dev = torch.device("cuda", 0)
NODES = ["features.2", "features.7", "features.16", "features.25", "features.34"]
MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
gt = torch.rand(16, 3, 256, 256, device=dev); sr = (gt + 0.1 * torch.randn_like(gt)).clamp(0, 1)
torch.backends.cudnn.benchmark = True

def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

out = {}
m = vgg.ContentLossMulti(NODES, MEAN, STD).to(dev)
def torch_multi(dtype=None):
    with torch.no_grad(), torch.autocast("cuda", dtype=dtype or torch.float16, enabled=dtype is not None):
        a, b = m.normalize(sr), m.normalize(gt)
        fa, fb = m.feature_extractor(a), m.feature_extractor(b)
        return [torch.nn.functional.l1_loss(fa[n], fb[n]) for n in NODES]
with torch.no_grad():
    out["multi_b200_ms"] = t(lambda: m(sr, gt))
out["multi_torch_tf32_ms"] = t(lambda: torch_multi())
out["multi_torch_fp16_autocast_ms"] = t(lambda: torch_multi(torch.float16))
out["multi_values_b200"] = m(sr, gt).flatten().tolist()
out["multi_values_torch"] = [float(v) for v in torch_multi()]
e = vgg.ContentLoss("features.34", MEAN, STD).to(dev)
def native_step():
    s = sr.clone().requires_grad_(True); e(s, gt).backward(); return s.grad
def torch_step(dtype=None):
    s = sr.clone().requires_grad_(True)
    with torch.autocast("cuda", dtype=dtype or torch.float16, enabled=dtype is not None):
        l = e._torch_forward(s, gt)
    l.float().backward(); return s.grad
out["single_fwd_bwd_b200_ms"] = t(native_step)
out["single_fwd_bwd_torch_tf32_ms"] = t(torch_step)
out["single_fwd_bwd_torch_fp16_autocast_ms"] = t(lambda: torch_step(torch.float16))
flops_fwd = 2 * 25.3e9 * 32
out["b200_fwd_tflops"] = flops_fwd / (out["multi_b200_ms"] * 1e-3) / 1e12
print(json.dumps(out))
