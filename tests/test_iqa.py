"""Evaluation epilogue: drop-in PSNR / SSIM / tensor_to_image.  CPU: the torch path of the drop-ins against
the oracle restatement (itself pinned against the reference's image_quality_assessment.py in test_oracle.py) and, where the
reference tree is present, against the reference's own classes.  GPU: the fused kernels through the C ABI against the oracle.
Tolerances: PSNR 1e-6 dB, SSIM 1e-6 (fp64 arithmetic on both sides; BASELINE's bar is 0.01 dB / 1e-4)."""
import os

import numpy as np
import pytest
import torch

from oracle import rrdbnet_oracle as orc
from sr_gan_fd_b200 import iqa


def _pair(n, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    a = torch.rand(n, 3, h, w, generator=g)
    b = (a + 0.05 * torch.randn(n, 3, h, w, generator=g)).clamp(0, 1)
    return a, b


@pytest.mark.parametrize("shape,crop", [((2, 40, 52), 4), ((1, 23, 31), 0), ((3, 64, 64), 6)])
def test_cpu_path_matches_oracle(shape, crop):
    a, b = _pair(*shape, seed=1)
    assert torch.allclose(iqa.PSNR(crop, True)(a, b), orc.psnr_y(a, b, crop), rtol=0, atol=1e-9)
    assert torch.allclose(iqa.SSIM(crop, True)(a, b), orc.ssim_y(a, b, crop), rtol=0, atol=1e-6)
    assert iqa.PSNR(crop, True)(a, b).dtype == torch.float64 and iqa.SSIM(crop, True)(a, b).dtype == torch.float32


@pytest.mark.skipif(not os.path.isfile("/root/reference/ESRGAN/image_quality_assessment.py"), reason="reference tree not present")
def test_cpu_path_matches_reference_classes_incl_rgb_mode():
    from oracle import reference_loader as rl
    ref = rl.load_module("esrgan", "image_quality_assessment")
    a, b = _pair(2, 48, 40, seed=2)
    for only_y in (True, False):
        assert torch.allclose(iqa.PSNR(4, only_y)(a, b), ref.PSNR(4, only_y)(a, b), rtol=0, atol=1e-9)
        assert torch.allclose(iqa.SSIM(4, only_y)(a, b), ref.SSIM(4, only_y)(a, b), rtol=0, atol=1e-6)
    ref_img = rl.load_module("esrgan", "imgproc")
    t = torch.rand(1, 3, 9, 7) * 1.2 - 0.1
    assert np.array_equal(iqa.tensor_to_image(t, False, False), ref_img.tensor_to_image(t, False, False))
    assert np.array_equal(iqa.tensor_to_image(t, True, True), ref_img.tensor_to_image(t, True, True))


@pytest.mark.gpu
@pytest.mark.parametrize("shape,crop", [((2, 40, 52), 4), ((1, 23, 31), 0), ((3, 64, 64), 6), ((16, 256, 256), 4), ((1, 300, 517), 4)])
def test_gpu_fused_iqa_matches_oracle(shape, crop):
    dev = torch.device("cuda", 0)
    a, b = _pair(*shape, seed=3)
    p = iqa.PSNR(crop, True)(a.to(dev), b.to(dev))
    s = iqa.SSIM(crop, True)(a.to(dev), b.to(dev))
    p_ref, s_ref = orc.psnr_y(a, b, crop), orc.ssim_y(a, b, crop)
    assert p.dtype == torch.float64 and s.dtype == torch.float32 and p.is_cuda
    dp = (p.cpu() - p_ref).abs().max().item()
    ds = (s.cpu() - s_ref).abs().max().item()
    print(f"fused IQA {shape}: |dPSNR| {dp:.2e} dB, |dSSIM| {ds:.2e}")
    assert dp <= 1e-6 and ds <= 1e-6
    # identical images: PSNR saturates at the reference's 1e-8 floor, SSIM = 1
    assert torch.allclose(iqa.SSIM(crop, True)(a.to(dev), a.to(dev)).cpu(), torch.ones(shape[0]))
    assert torch.all(iqa.PSNR(crop, True)(a.to(dev), a.to(dev)) > 120)


@pytest.mark.gpu
def test_gpu_tensor_to_image_matches_reference_semantics():
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(5)
    t = torch.rand(1, 3, 67, 45, generator=g) * 1.2 - 0.1  # includes values outside [0, 1]
    for range_norm, half in ((False, False), (True, False), (False, True)):
        got = iqa.tensor_to_image(t.to(dev), range_norm, half)
        x = t.add(1.0).div(2.0) if range_norm else t
        x = x.to(dev).half() if half else x
        want = x.squeeze(0).permute(1, 2, 0).mul(255).clamp(0, 255).cpu().numpy().astype("uint8")
        assert got.shape == (67, 45, 3) and got.dtype == np.uint8
        assert np.array_equal(got, want), (range_norm, half, np.abs(got.astype(int) - want.astype(int)).max())
