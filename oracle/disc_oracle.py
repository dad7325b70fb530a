"""ORACLE (test infrastructure, NOT product code) -- fp32 CPU restatement of the reference U-Net discriminator.

Only ``tests/``, ``__graft_entry__.smoke()`` and bench / tool CHECK legs may import this file.  The product path
(``sr_gan_fd_b200``) never routes through it.

What it restates (reference = MiNeves00/SR-GAN-FD):

* ``DiscriminatorUNet.__init__``       ``BSRGAN/model.py:92-135``  (= ``Real_ESRGAN/model.py:30-73``): the parameter set
* ``DiscriminatorUNet._forward_impl``  ``BSRGAN/model.py:143-167`` (= ``Real_ESRGAN/model.py:81-105``): the graph
* spectral normalisation: a THIRD-PARTY dependency of the reference, ``torch.nn.utils.spectral_norm`` (PyTorch; the
  reference pins ``torch>=1.12.1``, this image has 2.11.0).  Its published algorithm (``torch/nn/utils/spectral_norm.py``,
  ``SpectralNorm.compute_weight``): with the weight flattened to ``[out, in*kh*kw]``, in training mode one power iteration
  ``v <- normalize(W^T u)``, ``u <- normalize(W v)`` (eps 1e-12, without gradient, the buffers ``weight_u`` / ``weight_v``
  are updated in place), then ``sigma = u . (W v)`` and ``weight = weight_orig / sigma`` (gradient flows through W in sigma,
  u and v are constants).  In eval mode the stored u, v are used as they are.

Parity pinning: the reference holds no golden vectors or tests for this path, so the oracle is pinned by EXECUTING the
reference class in the build container (``tests/test_oracle.py::test_disc_oracle_bit_equal_to_reference``: outputs, updated
power-iteration buffers and every parameter gradient bit-equal on CPU fp32) and by ``tests/golden/disc_unet*.pt`` which
``oracle/make_golden.py`` generated from the imported reference ``model.py``.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

NEG_SLOPE = 0.2
SN_EPS = 1e-12

# (name, kernel, stride, spectral-norm?, LeakyReLU?) in state_dict order
LAYERS = [
    ("conv1", 3, 1, False, False),
    ("down_block1.0", 4, 2, True, True),
    ("down_block2.0", 4, 2, True, True),
    ("down_block3.0", 4, 2, True, True),
    ("up_block1.0", 3, 1, True, True),
    ("up_block2.0", 3, 1, True, True),
    ("up_block3.0", 3, 1, True, True),
    ("conv2.0", 3, 1, True, True),
    ("conv3.0", 3, 1, True, True),
    ("conv4", 3, 1, False, False),
]


def layer_shapes(in_channels: int = 3, out_channels: int = 1, channels: int = 64) -> List[Tuple[int, int, int]]:
    """(cout, cin, k) per layer -- BSRGAN/model.py:102-135."""
    c = channels
    return [(64, in_channels, 3), (2 * c, c, 4), (4 * c, 2 * c, 4), (8 * c, 4 * c, 4), (4 * c, 8 * c, 3), (2 * c, 4 * c, 3),
            (c, 2 * c, 3), (c, c, 3), (c, c, 3), (out_channels, c, 3)]


def param_names() -> List[str]:
    """Learnable parameters in ``named_parameters()`` order of the reference module."""
    names = []
    for name, _, _, sn, _ in LAYERS:
        if sn:
            names.append(name + ".weight_orig")
        else:
            names += [name + ".weight", name + ".bias"]
    # torch registers a conv's bias BEFORE spectral_norm re-registers weight_orig, hence for plain convs (weight, bias)
    return names


def spectral_weight(weight_orig: torch.Tensor, u: torch.Tensor, v: torch.Tensor, training: bool):
    """``SpectralNorm.compute_weight`` (n_power_iterations = 1, dim = 0).  Returns (weight, new_u, new_v)."""
    wm = weight_orig.flatten(1)
    if training:
        with torch.no_grad():
            v = F.normalize(torch.mv(wm.t(), u), dim=0, eps=SN_EPS)
            u = F.normalize(torch.mv(wm, v), dim=0, eps=SN_EPS)
    sigma = torch.dot(u, torch.mv(wm, v))
    return weight_orig / sigma, u, v


def effective_weights(state: Dict[str, torch.Tensor], training: bool):
    """All conv weights as the forward uses them + the power-iteration buffers after this forward."""
    weights, new_buffers = OrderedDict(), OrderedDict()
    for name, _, _, sn, _ in LAYERS:
        if sn:
            w, u, v = spectral_weight(state[name + ".weight_orig"], state[name + ".weight_u"], state[name + ".weight_v"], training)
            weights[name] = w
            new_buffers[name + ".weight_u"], new_buffers[name + ".weight_v"] = u, v
        else:
            weights[name] = state[name + ".weight"]
    return weights, new_buffers


def _up(t: torch.Tensor) -> torch.Tensor:
    return F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=False)


def forward_from_weights(weights: Dict[str, torch.Tensor], state: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """The graph of ``BSRGAN/model.py:143-167`` given the effective weights."""
    def conv(i: int, t: torch.Tensor) -> torch.Tensor:
        name, _, stride, sn, act = LAYERS[i]
        bias = None if sn else state[name + ".bias"]
        t = F.conv2d(t, weights[name], bias, stride=stride, padding=1)
        return F.leaky_relu(t, NEG_SLOPE) if act else t

    out1 = conv(0, x)
    down1 = conv(1, out1)
    down2 = conv(2, down1)
    down3 = conv(3, down2)
    up1 = conv(4, _up(down3)) + down2
    up2 = conv(5, _up(up1)) + down1
    up3 = conv(6, _up(up2)) + out1
    return conv(9, conv(8, conv(7, up3)))


def forward(state: Dict[str, torch.Tensor], x: torch.Tensor, training: bool = True):
    """state: the reference module's ``state_dict`` (weight_orig / weight_u / weight_v for the spectral-norm convs).
    Returns (logits, buffers after the forward)."""
    weights, new_buffers = effective_weights(state, training)
    return forward_from_weights(weights, state, x), new_buffers


def forward_backward(state: Dict[str, torch.Tensor], x: torch.Tensor, dy: torch.Tensor, training: bool = True, input_grad: bool = False):
    """One forward + backward with the upstream gradient ``dy`` on the logits.  Returns (logits, gradients by parameter
    name [w.r.t. weight_orig for the spectral-norm convs], gradients w.r.t. the EFFECTIVE weights by layer name, dx or None,
    buffers after the forward)."""
    st = {k: v.detach().clone().float() for k, v in state.items()}
    leaves = {k: st[k].requires_grad_(True) for k in param_names()}
    st.update(leaves)
    xin = x.detach().clone().float().requires_grad_(input_grad)
    weights, new_buffers = effective_weights(st, training)
    for name, w in weights.items():
        if not w.is_leaf:
            w.retain_grad()
    y = forward_from_weights(weights, st, xin)
    y.backward(dy.float())
    grads = OrderedDict((k, leaves[k].grad.detach().clone()) for k in param_names())
    eff = OrderedDict((name, w.grad.detach().clone()) for name, w in weights.items())
    return y.detach(), grads, eff, (xin.grad.detach().clone() if input_grad else None), new_buffers
