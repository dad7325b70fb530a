"""GPU parity of the fused GradScaler-unscale + Adam + EMA step against the reference's own
optimizer stack: torch.optim.Adam + AveragedModel(avg_fn) exactly as ESRGAN/train_rrdbnet.py:182-202,263-267 builds them."""
import copy

import pytest
import torch
import torch.nn.functional as F
from torch.optim.swa_utils import AveragedModel

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


def _models():
    import sr_gan_fd_b200 as b200
    torch.manual_seed(0)
    a = b200.rrdbnet_x4(num_blocks=1).to(DEV)
    b = copy.deepcopy(a)
    return a, b


@pytest.mark.parametrize("weight_decay,use_scaler", [(0.0, False), (1e-2, False), (0.0, True)])
def test_fused_adam_ema_matches_torch(weight_decay, use_scaler):
    from sr_gan_fd_b200.optim import FusedAdamEMA
    decay = 0.99998
    ref, mine = _models()
    ema_avg = lambda avg, p, n: (1 - decay) * avg + decay * p
    ref_ema = AveragedModel(ref, avg_fn=ema_avg)
    mine_ema = AveragedModel(mine, avg_fn=ema_avg)
    opt_ref = torch.optim.Adam(ref.parameters(), 2e-4, (0.9, 0.99), 1e-8, weight_decay)
    opt_mine = FusedAdamEMA(mine.parameters(), 2e-4, (0.9, 0.99), 1e-8, weight_decay, ema_model=mine_ema, ema_decay=decay)
    scaler_ref = torch.amp.GradScaler("cuda", enabled=use_scaler)
    scaler_mine = torch.amp.GradScaler("cuda", enabled=use_scaler)
    g = torch.Generator().manual_seed(3)
    for it in range(4):
        # identical synthetic gradients for both replicas (the generator path itself is tested elsewhere)
        grads = [torch.randn(p.shape, generator=g).to(DEV) * 1e-3 for p in ref.parameters()]
        if use_scaler and it == 2:
            grads[5][0] = float("inf")  # GradScaler must skip this step on both sides
        scale = 65536.0 if use_scaler else 1.0
        for model in (ref, mine):
            for p, gr in zip(model.parameters(), grads):
                p.grad = (gr * scale).clone()
        if use_scaler:
            # emulate scaler.scale(loss).backward(): grads are already scaled; step/unscale through the scaler
            if it == 0:
                scaler_ref.scale(torch.ones(1, device=DEV)); scaler_mine.scale(torch.ones(1, device=DEV))  # lazy init
            scaler_ref.step(opt_ref); scaler_ref.update()
            scaler_mine.step(opt_mine); scaler_mine.update()
            ref_ema.update_parameters(ref)  # the reference script updates the EMA unconditionally (train_rrdbnet.py:267)
        else:
            opt_ref.step()
            opt_mine.step()
            ref_ema.update_parameters(ref)
    for (n, a), b in zip(ref.named_parameters(), mine.parameters()):
        assert torch.allclose(a, b, rtol=2e-5, atol=1e-8), n
    for a, b in zip(ref_ema.module.parameters(), mine_ema.module.parameters()):
        assert torch.allclose(a, b, rtol=2e-5, atol=1e-8)
    assert int(mine_ema.n_averaged) == int(ref_ema.n_averaged)


def test_fused_step_triggers_weight_repack_and_trains():
    """The in-place kernel update must invalidate the generator's packed weights; loss goes down over a few steps."""
    import sr_gan_fd_b200 as b200
    from sr_gan_fd_b200.optim import FusedAdamEMA
    from oracle import rrdbnet_oracle as orc
    torch.manual_seed(0)
    net = b200.rrdbnet_x4(num_blocks=1)
    net.load_state_dict(orc.in_range_fixture({k: v.clone() for k, v in net.state_dict().items()}))
    net = net.to(DEV).train()
    opt = FusedAdamEMA(net.parameters(), lr=1e-3)
    lr = torch.rand(2, 3, 16, 16, device=DEV)
    gt = torch.rand(2, 3, 64, 64, device=DEV)
    losses = []
    for _ in range(6):
        net.zero_grad(set_to_none=True)
        loss = F.l1_loss(net(lr), gt)
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert losses[-1] < losses[0], losses
    assert len(set(round(x, 9) for x in losses)) > 3  # outputs changed step to step => weights were re-packed
