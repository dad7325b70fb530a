"""Drop-in RRDBNet generators: same constructors, ``nn.Module`` tree and ``state_dict`` layout as the reference.

Reference interfaces mirrored here (MiNeves00/SR-GAN-FD):

* ``ESRGAN/model.py:29-86,144-243,301-322``      -> :class:`RRDBNet`, ``rrdbnet_x1/x2/x4/x8``
* ``BSRGAN/model.py:31-88,311-384,576-587``      -> :class:`BSRGAN`, ``bsrgan_x2/x4``   (A-ESRGAN/model.py:421-558 same)
* ``Real_ESRGAN/model.py:108-263,331-334``       -> :class:`RealRRDBNet` (exported as ``RRDBNet`` by the compat shim)

The submodules are real ``nn.Conv2d`` so ``torch.manual_seed(s)`` + constructor reproduces the reference's weights
bit for bit, and ``load_state_dict`` / ``utils.load_state_dict`` key+shape filtering behave identically.  Only
``forward`` differs: the whole network (forward and backward) is ONE ``torch.autograd.Function`` that calls the
C-ABI CUDA library (``include/b200sr.h``).  There is no CPU or eager-PyTorch fallback: a CPU tensor or a missing
``libb200sr.so`` raises.
"""
from __future__ import annotations

from typing import Any, List

import torch
from torch import nn

from . import function as _function

__all__ = [
    "RRDBNet", "BSRGAN", "RealRRDBNet",
    "rrdbnet_x1", "rrdbnet_x2", "rrdbnet_x4", "rrdbnet_x8", "bsrgan_x2", "bsrgan_x4", "real_rrdbnet_x4",
]


class _ResidualDenseBlock(nn.Module):
    """Parameter container for one dense block (``ESRGAN/model.py:38-47``).  Never called layer by layer."""

    def __init__(self, channels: int, growth_channels: int, reinit: bool = False) -> None:
        super().__init__()
        self.conv1 = nn.Conv2d(channels + growth_channels * 0, growth_channels, (3, 3), (1, 1), (1, 1))
        self.conv2 = nn.Conv2d(channels + growth_channels * 1, growth_channels, (3, 3), (1, 1), (1, 1))
        self.conv3 = nn.Conv2d(channels + growth_channels * 2, growth_channels, (3, 3), (1, 1), (1, 1))
        self.conv4 = nn.Conv2d(channels + growth_channels * 3, growth_channels, (3, 3), (1, 1), (1, 1))
        self.conv5 = nn.Conv2d(channels + growth_channels * 4, channels, (3, 3), (1, 1), (1, 1))
        self.leaky_relu = nn.LeakyReLU(0.2, True)
        self.identity = nn.Identity()
        if reinit:
            # Real_ESRGAN/model.py:128-129,144-150 re-initialises inside every block constructor, which changes
            # the RNG draw order relative to ESRGAN/BSRGAN for the same seed.
            for module in self.modules():
                if isinstance(module, nn.Conv2d):
                    nn.init.kaiming_normal_(module.weight)
                    module.weight.data *= 0.1
                    if module.bias is not None:
                        nn.init.constant_(module.bias, 0)

    def forward(self, x):  # pragma: no cover - guarded on purpose
        raise RuntimeError("dense blocks are executed by the fused B200 generator kernel, not layer by layer")


class _ResidualResidualDenseBlock(nn.Module):
    """``ESRGAN/model.py:72-75``."""

    def __init__(self, channels: int, growth_channels: int, reinit: bool = False) -> None:
        super().__init__()
        self.rdb1 = _ResidualDenseBlock(channels, growth_channels, reinit)
        self.rdb2 = _ResidualDenseBlock(channels, growth_channels, reinit)
        self.rdb3 = _ResidualDenseBlock(channels, growth_channels, reinit)

    def forward(self, x):  # pragma: no cover
        raise RuntimeError("RRDBs are executed by the fused B200 generator kernel, not layer by layer")


def _up_stage(channels: int) -> nn.Sequential:
    return nn.Sequential(nn.Conv2d(channels, channels, (3, 3), (1, 1), (1, 1)), nn.LeakyReLU(0.2, True))


class _GeneratorBase(nn.Module):
    """Shared forward: collects the convs in ``state_dict`` order and hands them to the native path."""

    upscale_factor: int
    _n_up: int
    _pixel_unshuffle: int = 1

    def _finish_init(self) -> None:
        for module in self.modules():
            if isinstance(module, nn.Conv2d):
                nn.init.kaiming_normal_(module.weight)
                module.weight.data *= 0.1
                if module.bias is not None:
                    nn.init.constant_(module.bias, 0)

    # --- native runtime state is never part of the module's persistent state --------------------------------
    def _runtime(self) -> "_function.GeneratorRuntime":
        rt = self.__dict__.get("_b200_runtime")
        if rt is None:
            rt = _function.GeneratorRuntime()
            self.__dict__["_b200_runtime"] = rt
        return rt

    def __getstate__(self):
        state = self.__dict__.copy()
        state.pop("_b200_runtime", None)
        return state

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k == "_b200_runtime":
                continue
            new.__dict__[k] = copy.deepcopy(v, memo)
        return new

    def zero_grad(self, set_to_none: bool = True) -> None:
        """Same contract as ``nn.Module.zero_grad``; the ``set_to_none`` path skips the generic per-parameter work (the
        training loop calls this every step, ESRGAN/train_rrdbnet.py:255, and the gradients are views of one flat buffer)."""
        if not set_to_none:
            return super().zero_grad(set_to_none=False)
        rt = self._runtime()
        if rt.convs is None:
            rt.convs = self._conv_list()  # every parameter of the generator belongs to one of these convs
        for conv in rt.convs:
            pd = conv._parameters
            pd["weight"].grad = None
            if pd["bias"] is not None:
                pd["bias"].grad = None

    def _conv_list(self) -> List[nn.Conv2d]:
        convs = [self.conv1]
        for rrdb in self.trunk:
            for rdb in (rrdb.rdb1, rrdb.rdb2, rrdb.rdb3):
                convs += [rdb.conv1, rdb.conv2, rdb.conv3, rdb.conv4, rdb.conv5]
        convs.append(self.conv2)
        for u in range(1, self._n_up + 1):
            convs.append(getattr(self, f"upsampling{u}")[0])
        convs += [self.conv3[0], self.conv4]
        return convs

    def net_desc(self) -> dict:
        first_rdb = self.trunk[0].rdb1
        return dict(
            in_channels=self.conv1.in_channels, out_channels=self.conv4.out_channels,
            channels=self.conv1.out_channels, growth=first_rdb.conv1.out_channels,
            num_blocks=len(self.trunk), n_up=self._n_up, pixel_unshuffle=self._pixel_unshuffle,
        )

    def _forward_impl(self, x: torch.Tensor) -> torch.Tensor:
        return _function.generator_forward(self, x)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self._forward_impl(x)


class RRDBNet(_GeneratorBase):
    """``ESRGAN/model.py:144-243`` -- same signature, same children, same init."""

    def __init__(self, in_channels: int = 3, out_channels: int = 3, channels: int = 64, growth_channels: int = 32,
                 num_blocks: int = 23, upscale_factor: int = 4) -> None:
        super().__init__()
        self.upscale_factor = upscale_factor
        self.conv1 = nn.Conv2d(in_channels, channels, (3, 3), (1, 1), (1, 1))
        self.trunk = nn.Sequential(*[_ResidualResidualDenseBlock(channels, growth_channels)
                                     for _ in range(num_blocks)])
        self.conv2 = nn.Conv2d(channels, channels, (3, 3), (1, 1), (1, 1))
        self._n_up = {1: 0, 2: 1, 4: 2, 8: 3}.get(upscale_factor, 0)
        for u in range(1, self._n_up + 1):
            setattr(self, f"upsampling{u}", _up_stage(channels))
        self.conv3 = nn.Sequential(nn.Conv2d(channels, channels, (3, 3), (1, 1), (1, 1)), nn.LeakyReLU(0.2, True))
        self.conv4 = nn.Conv2d(channels, out_channels, (3, 3), (1, 1), (1, 1))
        self._initialize_weights()

    def _initialize_weights(self) -> None:
        self._finish_init()


class BSRGAN(_GeneratorBase):
    """``BSRGAN/model.py:311-384`` (and its copy ``A-ESRGAN/model.py:489-552``): ctor kw ``num_rrdb``; always
    ``upsampling1``, ``upsampling2`` only for x4."""

    def __init__(self, in_channels: int = 3, out_channels: int = 3, channels: int = 64, growth_channels: int = 32,
                 num_rrdb: int = 23, upscale_factor: int = 4) -> None:
        super().__init__()
        self.upscale_factor = upscale_factor
        self.conv1 = nn.Conv2d(in_channels, channels, (3, 3), (1, 1), (1, 1))
        self.trunk = nn.Sequential(*[_ResidualResidualDenseBlock(channels, growth_channels)
                                     for _ in range(num_rrdb)])
        self.conv2 = nn.Conv2d(channels, channels, (3, 3), (1, 1), (1, 1))
        self.upsampling1 = _up_stage(channels)
        self._n_up = 1
        if upscale_factor == 4:
            self.upsampling2 = _up_stage(channels)
            self._n_up = 2
        self.conv3 = nn.Sequential(nn.Conv2d(channels, channels, (3, 3), (1, 1), (1, 1)), nn.LeakyReLU(0.2, True))
        self.conv4 = nn.Conv2d(channels, out_channels, (3, 3), (1, 1), (1, 1))
        self._finish_init()


class RealRRDBNet(_GeneratorBase):
    """``Real_ESRGAN/model.py:179-263``: positional ctor, always two upsamplings, pixel-unshuffle front for x2/x1."""

    def __init__(self, in_channels: int, out_channels: int, channels: int, growth_channels: int, num_rrdb: int,
                 upscale_factor: int) -> None:
        super().__init__()
        if upscale_factor == 2:
            in_channels *= 4
            downscale_factor = 2
        elif upscale_factor == 1:
            in_channels *= 16
            downscale_factor = 4
        else:
            downscale_factor = 1
        self.downsampling = nn.PixelUnshuffle(downscale_factor)
        self._pixel_unshuffle = downscale_factor
        self.conv1 = nn.Conv2d(in_channels, channels, (3, 3), (1, 1), (1, 1))
        self.trunk = nn.Sequential(*[_ResidualResidualDenseBlock(channels, growth_channels, reinit=True)
                                     for _ in range(num_rrdb)])
        self.conv2 = nn.Conv2d(channels, channels, (3, 3), (1, 1), (1, 1))
        self.upsampling1 = _up_stage(channels)
        self.upsampling2 = _up_stage(channels)
        self._n_up = 2
        self.conv3 = nn.Sequential(nn.Conv2d(channels, channels, (3, 3), (1, 1), (1, 1)), nn.LeakyReLU(0.2, True))
        self.conv4 = nn.Conv2d(channels, out_channels, (3, 3), (1, 1), (1, 1))
        self._finish_init()

    def _forward_impl(self, x: torch.Tensor) -> torch.Tensor:
        # nn.PixelUnshuffle is pure data movement on the 3-channel input (identity for x4); kept as the torch op.
        return _function.generator_forward(self, self.downsampling(x))


# --- factories (names the reference scripts look up with model.__dict__[arch]) ------------------------------------
def rrdbnet_x1(**kwargs: Any) -> RRDBNet:
    return RRDBNet(upscale_factor=1, **kwargs)


def rrdbnet_x2(**kwargs: Any) -> RRDBNet:
    return RRDBNet(upscale_factor=2, **kwargs)


def rrdbnet_x4(**kwargs: Any) -> RRDBNet:
    return RRDBNet(upscale_factor=4, **kwargs)


def rrdbnet_x8(**kwargs: Any) -> RRDBNet:
    return RRDBNet(upscale_factor=8, **kwargs)


def bsrgan_x2(**kwargs: Any) -> BSRGAN:
    print("* BSRGAN 2x")  # BSRGAN/model.py:577
    return BSRGAN(upscale_factor=2, **kwargs)


def bsrgan_x4(**kwargs: Any) -> BSRGAN:
    print("* BSRGAN 4x")  # BSRGAN/model.py:584
    return BSRGAN(upscale_factor=4, **kwargs)


def real_rrdbnet_x4(**kwargs: Any) -> RealRRDBNet:
    return RealRRDBNet(upscale_factor=4, **kwargs)
