"""Script-level proof: the reference's own ``inference.py`` / ``train_rrdbnet.py`` run UNCHANGED through the drop-in shim
(``python -m sr_gan_fd_b200.compat.run``).  Needs the reference tree (``/root/reference`` in the build container, or
``$SRGANFD_REFERENCE``, or the unmodified copy staged under ``baseline/_ref`` that travels to the GPU box); skipped only where none exists.  A log of the GPU variants on a B200 is
kept under ``profiles/r2_reference_scripts_gpu.log``."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
from sr_gan_fd_b200.compat._passthrough import reference_root
# $SRGANFD_REFERENCE, /root/reference (build container), or -- on the GPU box -- the unmodified copy staged by
# __graft_entry__.stage_reference() under baseline/_ref (git-ignored, travels with the snapshot)
REF = reference_root()
os.environ["SRGANFD_REFERENCE"] = REF
needs_ref = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "ESRGAN", "inference.py")), reason="reference tree not present")


def _run(args, cwd, timeout=1500):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), PYTHONDONTWRITEBYTECODE="1",
               SRGANFD_REFERENCE=REF)
    r = subprocess.run([sys.executable, "-m", "sr_gan_fd_b200.compat.run", "--reference", REF] + args, cwd=cwd, env=env,
                       capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-3000:]
    return r.stdout


def _checkpoint(path, num_blocks=23, seed=0):
    """A random-init checkpoint in the reference's format (ESRGAN/utils.py:70-75 reads checkpoint["state_dict"])."""
    import sr_gan_fd_b200 as b200
    torch.manual_seed(seed)
    net = b200.rrdbnet_x4(num_blocks=num_blocks)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    sd["conv4.bias"] += 0.5  # visible output instead of the ~1e-4 of a raw random init
    for k in ("conv1.weight", "conv2.weight", "upsampling1.0.weight", "upsampling2.0.weight", "conv3.0.weight", "conv4.weight"):
        sd[k] *= 3.7
    torch.save({"state_dict": sd}, path)


def _png_dir(path, n, size, seed):
    import cv2
    os.makedirs(path, exist_ok=True)
    rng = np.random.default_rng(seed)
    for i in range(n):
        img = (rng.random((size, size, 3)) * 255).astype(np.uint8)
        img = cv2.GaussianBlur(img, (0, 0), 2.0)
        cv2.imwrite(os.path.join(path, f"img_{i:03d}.png"), img)


@needs_ref
def test_inference_script_cpu_default_matches_stock_reference(tmp_path):
    """ESRGAN/inference.py with its DEFAULT --device_type cpu: CPU tensors take the module's own convs through torch (eager
    path), and the written PNG equals the one the stock reference model.py writes for the same checkpoint, byte for byte."""
    import cv2
    ckpt = str(tmp_path / "rand.pth.tar")
    _checkpoint(ckpt, num_blocks=23)
    lr_png = str(tmp_path / "lr.png")
    img = cv2.imread(os.path.join(REF, "ESRGAN", "figure", "baboon_lr.png"))
    cv2.imwrite(lr_png, img[:40, :36])  # a crop keeps the CPU run short; the full 120x123 frame works the same way
    outs = []
    for tag, extra in (("shim", []), ("stock", ["--stock"])):
        out = str(tmp_path / f"sr_{tag}.png")
        log = _run(extra + ["ESRGAN", "inference.py", "--inputs_path", lr_png, "--output_path", out, "--model_weights_path", ckpt],
                   cwd=str(tmp_path))
        assert "Build `rrdbnet_x4` model successfully." in log
        outs.append(cv2.imread(out))
    assert outs[0].shape == (160, 144, 3)
    assert np.array_equal(outs[0], outs[1])
    assert outs[0].std() > 5  # not a blank frame


@needs_ref
@pytest.mark.gpu
def test_inference_script_cuda(tmp_path):
    import cv2
    ckpt = str(tmp_path / "rand.pth.tar")
    _checkpoint(ckpt, num_blocks=23)
    lr_png = os.path.join(REF, "ESRGAN", "figure", "baboon_lr.png")
    outs = []
    for tag, extra in (("shim", []), ("stock", ["--stock"])):
        out = str(tmp_path / f"sr_{tag}.png")
        _run(extra + ["ESRGAN", "inference.py", "--inputs_path", lr_png, "--output_path", out, "--model_weights_path", ckpt,
                      "--device_type", "cuda"], cwd=str(tmp_path))
        outs.append(cv2.imread(out).astype(np.int32))
    h, w = cv2.imread(lr_png).shape[:2]
    assert outs[0].shape == outs[1].shape == (4 * h, 4 * w, 3)
    diff = np.abs(outs[0] - outs[1])
    print(f"inference.py cuda: B200 path vs stock torch/cuDNN path: max |d| {diff.max()} of 255, mean {diff.mean():.4f}")
    assert diff.max() <= 2 and diff.mean() < 0.1


@needs_ref
@pytest.mark.gpu
def test_train_rrdbnet_script_cuda(tmp_path):
    """ESRGAN/train_rrdbnet.py (autocast + GradScaler + Adam + AveragedModel + validate + save_checkpoint) for one epoch of
    three iterations on synthetic PNG folders, then resumes nothing -- the checkpoint it wrote must load back."""
    _png_dir(str(tmp_path / "train"), 12, 128, 1)
    _png_dir(str(tmp_path / "gt"), 2, 96, 2)
    import cv2
    os.makedirs(tmp_path / "lr", exist_ok=True)
    for f in os.listdir(tmp_path / "gt"):
        g = cv2.imread(str(tmp_path / "gt" / f))
        cv2.imwrite(str(tmp_path / "lr" / f), cv2.resize(g, (24, 24), interpolation=cv2.INTER_CUBIC))
    sets = {"train_gt_images_dir": str(tmp_path / "train"), "test_gt_images_dir": str(tmp_path / "gt"),
            "test_lr_images_dir": str(tmp_path / "lr"), "gt_image_size": 96, "batch_size": 4, "num_workers": 1, "epochs": 1,
            "train_print_frequency": 1, "lr_scheduler_step_size": 1}
    args = []
    for k, v in sets.items():
        args += ["--set", f"rrdbnet_config.{k}={v!r}"]
    log = _run(["--iqa"] + args + ["ESRGAN", "train_rrdbnet.py"], cwd=str(tmp_path))  # validate() also uses the fused PSNR / SSIM
    assert "Build `rrdbnet_x4` model successfully." in log
    assert "Epoch: [1][3/3]" in log or "[3/3]" in log, log[-1500:]
    ck = torch.load(str(tmp_path / "results" / "RRDBNet_x4" / "g_last.pth.tar"), map_location="cpu")
    assert len(ck["state_dict"]) == 702 and "optimizer" in ck and "ema_state_dict" in ck
    print(log[-1200:])
