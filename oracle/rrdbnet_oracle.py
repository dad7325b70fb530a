"""ORACLE (test infrastructure, NOT product code) -- fp32 CPU restatement of the reference RRDBNet generator.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this file.  The product path (``sr_gan_fd_b200``) never routes through it.

What it restates (reference = MiNeves00/SR-GAN-FD, all four model folders share this network):

* dense block        ``ESRGAN/model.py:49-60``   (= ``BSRGAN/model.py:51-62``, ``Real_ESRGAN/model.py:131-142``)
* RRDB               ``ESRGAN/model.py:77-86``
* generator forward  ``ESRGAN/model.py:211-232`` (``BSRGAN/model.py:366-381``, ``Real_ESRGAN/model.py:246-263``)
* weight init        ``ESRGAN/model.py:237-243``
* PSNR / SSIM (Y)    ``ESRGAN/image_quality_assessment.py:361-541`` and ``ESRGAN/imgproc.py:409-434``

Parity pinning: the reference ships no golden vectors or tests for this path, so the oracle is
pinned by EXECUTING the reference classes in the build container (``tests/test_oracle_vs_reference.py``, bit-equal
on CPU fp32) and by the fixtures under ``tests/golden/`` which ``oracle/make_golden.py`` generated from the
imported reference ``model.py``.  On the GPU box ``/root/reference`` does not exist; the fixtures travel instead.

The arithmetic is a plain chain of ``torch.nn.functional.conv2d`` calls in fp32 on the CPU -- the same ATen kernels
the reference's ``nn.Conv2d`` modules dispatch to -- with explicit concatenation, LeakyReLU(0.2), the two 0.2
residual scalings, nearest x2 upsampling and the final clamp.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

NEG_SLOPE = 0.2
RES_SCALE = 0.2


# --------------------------------------------------------------------------------------------------------------
# parameter naming (state_dict layout of the reference)
# --------------------------------------------------------------------------------------------------------------
def num_upsamplings(upscale_factor: int, flavour: str = "esrgan") -> int:
    """How many ``upsampling{k}`` stages the reference builds.

    esrgan : ``ESRGAN/model.py:166-196`` -- x1: 0, x2: 1, x4: 2, x8: 3
    bsrgan : ``BSRGAN/model.py:337-346`` -- always upsampling1, upsampling2 only if x4
    real   : ``Real_ESRGAN/model.py:212-221`` -- always two (pixel-unshuffle front for x2 / x1)
    """
    if flavour == "esrgan":
        return {1: 0, 2: 1, 4: 2, 8: 3}[upscale_factor]
    if flavour == "bsrgan":
        return 2 if upscale_factor == 4 else 1
    if flavour == "real":
        return 2
    raise ValueError(flavour)


def conv_names(num_blocks: int, n_up: int) -> List[str]:
    """Conv layer prefixes in the reference's ``state_dict`` order."""
    names = ["conv1"]
    for b in range(num_blocks):
        for r in (1, 2, 3):
            for c in (1, 2, 3, 4, 5):
                names.append(f"trunk.{b}.rdb{r}.conv{c}")
    names.append("conv2")
    for u in range(1, n_up + 1):
        names.append(f"upsampling{u}.0")
    names.append("conv3.0")
    names.append("conv4")
    return names


def conv_shapes(in_channels: int, out_channels: int, channels: int, growth: int, num_blocks: int,
                n_up: int) -> "OrderedDict[str, Tuple[int, int]]":
    """name -> (Cout, Cin) for every conv, reference order."""
    shapes: "OrderedDict[str, Tuple[int, int]]" = OrderedDict()
    for name in conv_names(num_blocks, n_up):
        leaf = name.split(".")[-1]
        if name == "conv1":
            shapes[name] = (channels, in_channels)
        elif name.startswith("trunk."):
            k = int(leaf[-1])
            shapes[name] = (growth if k < 5 else channels, channels + growth * (k - 1))
        elif name == "conv4":
            shapes[name] = (out_channels, channels)
        else:
            shapes[name] = (channels, channels)
    return shapes


def init_params(seed: int = 0, in_channels: int = 3, out_channels: int = 3, channels: int = 64, growth: int = 32,
                num_blocks: int = 23, upscale_factor: int = 4, flavour: str = "esrgan",
                dtype: torch.dtype = torch.float32) -> "OrderedDict[str, torch.Tensor]":
    """Random init following ``ESRGAN/model.py:237-243``: kaiming_normal_ (fan_in, gain sqrt(2)) * 0.1, bias 0.

    NOTE: draws from a private generator, so values are NOT the ones ``torch.manual_seed(seed)`` + the reference
    constructor would produce (the drop-in module reproduces those by keeping real ``nn.Conv2d`` children).
    """
    g = torch.Generator().manual_seed(seed)
    if flavour == "real":
        if upscale_factor == 2:
            in_channels *= 4
        elif upscale_factor == 1:
            in_channels *= 16
    params: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, (co, ci) in conv_shapes(in_channels, out_channels, channels, growth, num_blocks,
                                      num_upsamplings(upscale_factor, flavour)).items():
        std = math.sqrt(2.0 / (ci * 9))
        params[name + ".weight"] = (torch.randn(co, ci, 3, 3, generator=g, dtype=dtype) * std) * 0.1
        params[name + ".bias"] = torch.zeros(co, dtype=dtype)
    return params


def in_range_fixture(params: Dict[str, torch.Tensor], gain: float = 3.7, out_bias: float = 0.5) -> Dict[str, torch.Tensor]:
    """SURVEY.md section 7.2 recipe: scale the head/tail conv weights so the SR output spans [0, 1] (random-init output is
    ~1e-4 and PSNR/SSIM/clamp masks are otherwise blind).  Returns a new dict."""
    out = OrderedDict((k, v.clone()) for k, v in params.items())
    for k in out:
        if k.endswith(".weight") and not k.startswith("trunk."):
            out[k] *= gain
    out["conv4.bias"] = torch.full_like(out["conv4.bias"], out_bias)
    return out


# --------------------------------------------------------------------------------------------------------------
# forward (functional)
# --------------------------------------------------------------------------------------------------------------
def _conv(x: torch.Tensor, p: Dict[str, torch.Tensor], name: str) -> torch.Tensor:
    return F.conv2d(x, p[name + ".weight"], p[name + ".bias"], stride=1, padding=1)


def _lrelu(x: torch.Tensor) -> torch.Tensor:
    return F.leaky_relu(x, NEG_SLOPE)


def rdb_forward(x: torch.Tensor, p: Dict[str, torch.Tensor], prefix: str) -> torch.Tensor:
    """``ESRGAN/model.py:49-60``."""
    o1 = _lrelu(_conv(x, p, prefix + ".conv1"))
    o2 = _lrelu(_conv(torch.cat([x, o1], 1), p, prefix + ".conv2"))
    o3 = _lrelu(_conv(torch.cat([x, o1, o2], 1), p, prefix + ".conv3"))
    o4 = _lrelu(_conv(torch.cat([x, o1, o2, o3], 1), p, prefix + ".conv4"))
    o5 = _conv(torch.cat([x, o1, o2, o3, o4], 1), p, prefix + ".conv5")
    return torch.add(torch.mul(o5, RES_SCALE), x)


def rrdb_forward(x: torch.Tensor, p: Dict[str, torch.Tensor], prefix: str) -> torch.Tensor:
    """``ESRGAN/model.py:77-86``."""
    out = rdb_forward(x, p, prefix + ".rdb1")
    out = rdb_forward(out, p, prefix + ".rdb2")
    out = rdb_forward(out, p, prefix + ".rdb3")
    return torch.add(torch.mul(out, RES_SCALE), x)


def count_blocks(p: Dict[str, torch.Tensor]) -> int:
    n = 0
    while f"trunk.{n}.rdb1.conv1.weight" in p:
        n += 1
    return n


def count_upsamplings(p: Dict[str, torch.Tensor]) -> int:
    n = 0
    while f"upsampling{n + 1}.0.weight" in p:
        n += 1
    return n


def rrdbnet_forward(p: Dict[str, torch.Tensor], x: torch.Tensor, pixel_unshuffle: int = 1,
                    return_pre_clamp: bool = False):
    """``ESRGAN/model.py:211-232``; ``pixel_unshuffle`` > 1 restates ``Real_ESRGAN/model.py:190-204,248``."""
    if pixel_unshuffle > 1:
        x = F.pixel_unshuffle(x, pixel_unshuffle)
    out1 = _conv(x, p, "conv1")
    out = out1
    for b in range(count_blocks(p)):
        out = rrdb_forward(out, p, f"trunk.{b}")
    out2 = _conv(out, p, "conv2")
    out = torch.add(out1, out2)
    for u in range(1, count_upsamplings(p) + 1):
        out = _lrelu(_conv(F.interpolate(out, scale_factor=2, mode="nearest"), p, f"upsampling{u}.0"))
    out = _lrelu(_conv(out, p, "conv3.0"))
    pre = _conv(out, p, "conv4")
    sr = torch.clamp(pre, 0.0, 1.0)
    if return_pre_clamp:
        return sr, pre
    return sr


def rrdbnet_l1_step(p: Dict[str, torch.Tensor], lr: torch.Tensor, gt: torch.Tensor, loss_scale: float = 1.0,
                    pixel_unshuffle: int = 1):
    """One L1 pre-training fwd+bwd (``ESRGAN/train_rrdbnet.py:256-261``): returns (sr, loss, grads dict)."""
    leaves = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in p.items())
    sr = rrdbnet_forward(leaves, lr, pixel_unshuffle)
    loss = F.l1_loss(sr, gt)
    (loss * loss_scale).backward()
    grads = OrderedDict((k, v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items())
    return sr.detach(), loss.detach(), grads


# --------------------------------------------------------------------------------------------------------------
# analytic work model
# --------------------------------------------------------------------------------------------------------------
def flops_per_lr_pixel(in_channels=3, out_channels=3, channels=64, growth=32, num_blocks=23, n_up=2,
                       backward: bool = False) -> int:
    """2 x MACs of the reference graph per LR pixel.  Backward = all wgrads + all dgrads except conv1's."""
    shapes = conv_shapes(in_channels, out_channels, channels, growth, num_blocks, n_up)
    total = 0
    for name, (co, ci) in shapes.items():
        res = 1
        if name.startswith("upsampling"):
            res = 4 ** int(name[len("upsampling")])
        elif name in ("conv3.0", "conv4"):
            res = 4 ** n_up
        macs = 9 * ci * co * res
        if backward:
            total += 2 * macs * (1 if name == "conv1" else 2)
        else:
            total += 2 * macs
    return total


# --------------------------------------------------------------------------------------------------------------
# IQA restatement: PSNR / SSIM on the Y channel
# --------------------------------------------------------------------------------------------------------------
def rgb_to_y(t: torch.Tensor) -> torch.Tensor:
    """``ESRGAN/imgproc.py:409-434`` with only_use_y_channel=True."""
    w = torch.tensor([[65.481], [128.553], [24.966]], dtype=t.dtype, device=t.device)
    y = torch.matmul(t.permute(0, 2, 3, 1), w).permute(0, 3, 1, 2) + 16.0
    return y / 255.0


def _gaussian_window(size: int = 11, sigma: float = 1.5) -> np.ndarray:
    # cv2.getGaussianKernel(size, sigma) for sigma > 0: exp(-(i-(n-1)/2)^2 / (2 sigma^2)), normalised
    ax = np.arange(size, dtype=np.float64) - (size - 1) / 2.0
    k = np.exp(-(ax ** 2) / (2.0 * sigma * sigma))
    k = k / k.sum()
    return np.outer(k, k)


def psnr_y(raw: torch.Tensor, dst: torch.Tensor, crop_border: int = 4) -> torch.Tensor:
    """``ESRGAN/image_quality_assessment.py:361-395`` (crop, Y, fp64, 10 log10(255^2 / mse))."""
    if crop_border > 0:
        raw = raw[:, :, crop_border:-crop_border, crop_border:-crop_border]
        dst = dst[:, :, crop_border:-crop_border, crop_border:-crop_border]
    raw = rgb_to_y(raw).to(torch.float64)
    dst = rgb_to_y(dst).to(torch.float64)
    mse = torch.mean((raw * 255.0 - dst * 255.0) ** 2 + 1e-8, dim=[1, 2, 3])
    return 10 * torch.log10(255.0 ** 2 / mse)


def ssim_y(raw: torch.Tensor, dst: torch.Tensor, crop_border: int = 4) -> torch.Tensor:
    """``ESRGAN/image_quality_assessment.py:421-505`` (11x11 gaussian sigma 1.5, valid conv, Y channel, fp64)."""
    if crop_border > 0:
        raw = raw[:, :, crop_border:-crop_border, crop_border:-crop_border]
        dst = dst[:, :, crop_border:-crop_border, crop_border:-crop_border]
    raw = rgb_to_y(raw).to(torch.float64) * 255.0
    dst = rgb_to_y(dst).to(torch.float64) * 255.0
    c1 = (0.01 * 255.0) ** 2
    c2 = (0.03 * 255.0) ** 2
    win = torch.from_numpy(_gaussian_window()).view(1, 1, 11, 11).to(raw)
    win = win.expand(raw.size(1), 1, 11, 11)
    g = raw.shape[1]
    mu_r = F.conv2d(raw, win, groups=g)
    mu_d = F.conv2d(dst, win, groups=g)
    var_r = F.conv2d(raw * raw, win, groups=g) - mu_r ** 2
    var_d = F.conv2d(dst * dst, win, groups=g) - mu_d ** 2
    cov = F.conv2d(raw * dst, win, groups=g) - mu_r * mu_d
    s = ((2 * mu_r * mu_d + c1) * (2 * cov + c2)) / ((mu_r ** 2 + mu_d ** 2 + c1) * (var_r + var_d + c2))
    return torch.mean(s, [1, 2, 3]).float()


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a - b|| / ||b|| in fp64 (b = reference)."""
    a = a.detach().double().flatten().cpu()
    b = b.detach().double().flatten().cpu()
    den = float(torch.linalg.norm(b))
    return float(torch.linalg.norm(a - b)) / (den if den > 0 else 1.0)
