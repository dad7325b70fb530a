"""VGG19 content loss: drop-in ContentLoss (ESRGAN flavour, differentiable) and ContentLossMulti (BSRGAN flavour).
torchvision's pretrained weights cannot be downloaded here: models.vgg19 is patched to a SEEDED RANDOM-INIT network (same graph).
CPU: the torch path equals the reference classes.  GPU: the tcgen05 path vs the same module's fp32 torch path on the CPU.
Tolerances (bf16 operands through up to 16 conv layers, fp32 accumulation; none is stated by BASELINE.json for this loss):
loss values 1 % relative, d loss / d sr 5e-2 relative L2."""
import os

import pytest
import torch
import torchvision.models as models

from oracle import rrdbnet_oracle as orc

NODES = ["features.2", "features.7", "features.16", "features.25", "features.34"]
MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]


@pytest.fixture(autouse=True)
def seeded_vgg(monkeypatch):
    real = models.vgg19

    def seeded(*a, **k):
        state = torch.random.get_rng_state()
        torch.manual_seed(1234)
        m = real(weights=None)
        torch.random.set_rng_state(state)
        return m
    monkeypatch.setattr(models, "vgg19", seeded)


def _pair(n, h, w, seed=0):
    g = torch.Generator().manual_seed(seed)
    gt = torch.rand(n, 3, h, w, generator=g)
    sr = (gt + 0.1 * torch.randn(n, 3, h, w, generator=g)).clamp(0, 1)
    return sr, gt


@pytest.mark.skipif(not os.path.isfile("/root/reference/BSRGAN/model.py"), reason="reference tree not present")
def test_cpu_path_equals_reference_classes():
    from oracle import reference_loader as rl
    from sr_gan_fd_b200 import vgg
    sr, gt = _pair(2, 40, 48)
    ref_e = rl.load_module("esrgan").ContentLoss("features.34", MEAN, STD)
    mine_e = vgg.ContentLoss("features.34", MEAN, STD)
    assert list(mine_e.state_dict().keys()) == list(ref_e.state_dict().keys())
    s1 = sr.clone().requires_grad_(True); s2 = sr.clone().requires_grad_(True)
    a, b = mine_e(s1, gt), ref_e(s2, gt)
    assert torch.equal(a, b)
    a.backward(); b.backward()
    assert torch.equal(s1.grad, s2.grad)
    ref_b = rl.load_module("bsrgan").ContentLoss(NODES, MEAN, STD)
    mine_b = vgg.ContentLossMulti(NODES, MEAN, STD)
    ra, rb = mine_b(sr, gt), ref_b(sr, gt)
    assert ra.shape == rb.shape == (1, 5) and torch.equal(ra, rb) and not ra.requires_grad


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 64, 64), (1, 48, 80), (3, 33, 47)])
def test_gpu_multi_node_losses(shape):
    from sr_gan_fd_b200 import vgg
    dev = torch.device("cuda", 0)
    sr, gt = _pair(*shape, seed=3)
    m = vgg.ContentLossMulti(NODES, MEAN, STD)
    ref = m(sr, gt)                      # fp32 torch path on the CPU
    got = m.to(dev)(sr.to(dev), gt.to(dev)).cpu()
    rel = ((got - ref).abs() / ref.abs()).max().item()
    print(f"VGG multi-node losses {shape}: ref {ref.flatten().tolist()} max rel err {rel:.2e}")
    assert got.shape == (1, 5) and not got.requires_grad and rel <= 1e-2


@pytest.mark.gpu
@pytest.mark.parametrize("node,shape,indep", [("features.2", (2, 24, 24), False), ("features.7", (2, 32, 32), False), ("features.16", (1, 40, 56), False),
                                              ("features.34", (2, 64, 64), False), ("features.34", (2, 64, 64), True)])
def test_gpu_single_node_loss_and_input_gradient(node, shape, indep):
    from sr_gan_fd_b200 import vgg
    dev = torch.device("cuda", 0)
    sr, gt = _pair(*shape, seed=5)
    if indep:
        sr = torch.rand(sr.shape, generator=torch.Generator().manual_seed(77))
    m = vgg.ContentLoss(node, MEAN, STD)
    s_ref = sr.clone().requires_grad_(True)
    l_ref = m(s_ref, gt)
    (l_ref * 3.0).backward()
    md = m.to(dev)
    s_dev = sr.to(dev).requires_grad_(True)
    l = md(s_dev, gt.to(dev))
    (l * 3.0).backward()
    rel = abs(float(l) - float(l_ref)) / abs(float(l_ref))
    gerr = orc.rel_l2(s_dev.grad.cpu(), s_ref.grad)
    # yardstick: stock torch bf16 autocast of the SAME module on the same GPU.  The gradient of an L1 loss on ReLU features is
    # discontinuous in the forward values (ReLU masks / signs flip where a pre-activation or a feature difference is within the
    # forward rounding error), so any reduced-precision forward shows a depth-dependent gradient error against fp32
    s_ac = sr.to(dev).requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        l_ac = md._torch_forward(s_ac, gt.to(dev))
    (l_ac.float() * 3.0).backward()
    gerr_ac = orc.rel_l2(s_ac.grad.cpu(), s_ref.grad)
    rel_ac = abs(float(l_ac) - float(l_ref)) / abs(float(l_ref))
    print(f"VGG {node} {shape}: loss {float(l):.6f} vs {float(l_ref):.6f} (rel {rel:.2e}; torch bf16 autocast {rel_ac:.2e}), "
          f"d/dsr rel-L2 {gerr:.2e} (torch bf16 autocast {gerr_ac:.2e})")
    assert rel <= 1e-2
    assert gerr <= max(5e-2, 1.25 * gerr_ac), (gerr, gerr_ac)
    with torch.no_grad():
        assert abs(float(md(sr.to(dev), gt.to(dev))) - float(l)) < 1e-6
