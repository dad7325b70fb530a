"""BASELINE configs[4] with EVERY component native (1 GPU, no reference tree needed): one BSRGAN GAN iteration
(``BSRGAN/train_bsrgan.py:412-470`` as restated by tools/gan_step.py) run twice from the same seeds -- once with the U-Net discriminator
and the VGG19 content loss on libb200sr.so, once with the SAME drop-in modules sent through their stock torch paths (= the reference's
op sequences; fp16 autocast as the script runs them).  What the generator receives from the critics (the gradient w.r.t. ``sr``:
pixel + adversarial terms, x 65536) and what the discriminator's optimizer receives (its parameter gradients after the real + fake
backward passes) must agree within the 16-bit tolerance of tests/test_disc.py, the losses within 1e-3, and the fused Adam/EMA step must
move both models.  The generator itself is the native path in both runs (its parity is tests/test_model_gpu.py's subject)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu]


def rel_l2(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _one_step(native, optimizer="stock", ema=False, step_models=False):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gan_step
    dev = torch.device("cuda", 0)
    mode = "b200" if native else "torch"
    d_model, g_model, content = gan_step.build("b200", dev, seed=0, content=mode, disc=mode)
    d_model.train(); g_model.train()
    step = gan_step.GanStep(d_model, g_model, content, dev, optimizer=optimizer, ema=ema)
    g = torch.Generator().manual_seed(8)
    lr = torch.rand(4, 3, 32, 32, generator=g).to(dev)
    gt = torch.rand(4, 3, 128, 128, generator=g).to(dev)
    before = [p.detach().clone() for p in list(d_model.parameters()) + list(g_model.parameters())]
    sr, g_loss, d_loss = step(lr, gt, step_d=step_models, step_g=step_models, keep_sr_grad=True)
    d_grads = None
    if not step_models:  # the generator update froze the discriminator but left the gradients of its own update in place
        d_grads = {n: p.grad.detach().float().cpu() for n, p in d_model.named_parameters() if p.grad is not None}
    after = list(d_model.parameters()) + list(g_model.parameters())
    moved = sum(int(not torch.equal(a, b)) for a, b in zip(before, after))
    return dict(sr=sr.detach().float().cpu(), sr_grad=step.sr_grad.float().cpu(), g_loss=float(g_loss), d_loss=float(d_loss),
                d_grads=d_grads, moved=moved, total=len(before), step=step)


def test_native_critics_agree_with_their_torch_paths_inside_the_gan_step():
    a = _one_step(native=True)
    b = _one_step(native=False)
    assert torch.equal(a["sr"], b["sr"])  # same generator path, same seeds
    assert abs(a["g_loss"] - b["g_loss"]) <= 1e-3 * abs(b["g_loss"]) and abs(a["d_loss"] - b["d_loss"]) <= 1e-3 * abs(b["d_loss"])
    e_sr = rel_l2(a["sr_grad"], b["sr_grad"])
    worst = max(((rel_l2(a["d_grads"][n], b["d_grads"][n]), n) for n in b["d_grads"]))
    print(f"all-native GAN step vs torch critics: d loss/d sr rel-L2 {e_sr:.2e}, worst discriminator gradient {worst[1]} {worst[0]:.2e}, "
          f"g_loss {a['g_loss']:.4f}/{b['g_loss']:.4f}, d_loss {a['d_loss']:.4f}/{b['d_loss']:.4f}")
    assert float(a["sr_grad"].abs().max()) > 1.0  # carries the GradScaler's 65536
    assert e_sr <= 5e-2, e_sr  # the pixel term dominates; the adversarial term comes through two fp16 networks
    assert set(a["d_grads"]) == set(b["d_grads"]) and worst[0] <= 0.1, worst


def test_all_native_step_with_fused_optimizer_moves_both_models():
    r = _one_step(native=True, optimizer="fused", ema=True, step_models=True)
    assert r["moved"] == r["total"], (r["moved"], r["total"])
    assert all(torch.isfinite(torch.tensor([r["g_loss"], r["d_loss"]])))
    ema = r["step"].ema
    assert ema is not None and int(ema.n_averaged) == 1
