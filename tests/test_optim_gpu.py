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


def test_state_dict_round_trip_matches_torch():
    """The reference's resume path (ESRGAN/utils.py:53: optimizer.load_state_dict, ema_model.load_state_dict): save after two
    steps, rebuild everything, load, take two more steps -- must equal torch.optim.Adam + AveragedModel doing the same."""
    import io
    from sr_gan_fd_b200.optim import FusedAdamEMA
    decay = 0.999
    ema_avg = lambda avg, p, n: (1 - decay) * avg + decay * p

    def build(fused):
        ref, mine = _models()
        model = mine if fused else ref
        ema = AveragedModel(model, avg_fn=ema_avg)
        if fused:
            opt = FusedAdamEMA(model.parameters(), 2e-4, (0.9, 0.99), 1e-8, 0.0, ema_model=ema, ema_decay=decay)
        else:
            opt = torch.optim.Adam(model.parameters(), 2e-4, (0.9, 0.99), 1e-8, 0.0)
        return model, ema, opt

    g = torch.Generator().manual_seed(9)
    shapes = [p.shape for p in _models()[0].parameters()]
    all_grads = [[torch.randn(s, generator=g).to(DEV) * 1e-3 for s in shapes] for _ in range(4)]

    def run(fused):
        model, ema, opt = build(fused)

        def steps(its):
            for it in its:
                for p, gr in zip(model.parameters(), all_grads[it]):
                    p.grad = gr.clone()
                opt.step()
                if not fused:
                    ema.update_parameters(model)
        steps([0, 1])
        buf = io.BytesIO()
        torch.save({"model": model.state_dict(), "ema": ema.state_dict(), "opt": opt.state_dict()}, buf)
        buf.seek(0)
        ckpt = torch.load(buf, map_location="cpu")  # checkpoints come back as CPU tensors on a fresh process
        del model, ema, opt
        model, ema, opt = build(fused)  # fresh objects, as after a restart
        model.load_state_dict(ckpt["model"]); ema.load_state_dict(ckpt["ema"]); opt.load_state_dict(ckpt["opt"])
        steps([2, 3])
        return model, ema

    ref, ref_ema = run(False)
    mine, mine_ema = run(True)
    for (n, a), b in zip(ref.named_parameters(), mine.parameters()):
        assert torch.allclose(a, b, rtol=2e-5, atol=1e-8), n
    for a, b in zip(ref_ema.module.parameters(), mine_ema.module.parameters()):
        assert torch.allclose(a, b, rtol=2e-5, atol=1e-8)
    assert int(mine_ema.n_averaged) == int(ref_ema.n_averaged) == 4


def test_ema_module_forward_sees_fused_updates():
    """The fused kernel also writes the EMA copy: its generator runtime must re-pack (version counters bumped)."""
    import sr_gan_fd_b200 as b200
    from sr_gan_fd_b200.optim import FusedAdamEMA
    torch.manual_seed(0)
    net = b200.rrdbnet_x4(num_blocks=1).to(DEV).train()
    ema = AveragedModel(net, avg_fn=lambda avg, p, n: 0.5 * avg + 0.5 * p)
    opt = FusedAdamEMA(net.parameters(), lr=1e-2, ema_model=ema, ema_decay=0.5)
    x = torch.rand(1, 3, 16, 16, device=DEV)
    with torch.no_grad():
        before = ema.module(x).clone()  # packs the EMA weights once
    for _ in range(3):
        net.zero_grad(set_to_none=True)
        net(x).mean().backward()
        opt.step()
    with torch.no_grad():
        after = ema.module(x)
        fresh = copy.deepcopy(ema.module)(x)  # fresh runtime: packs the current EMA weights
    assert torch.equal(after, fresh)
    assert not torch.equal(after, before)
