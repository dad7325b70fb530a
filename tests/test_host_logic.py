"""CPU: host-side logic of the drop-in (module protocol, C-ABI symbols, plan bookkeeping).  No GPU compute calls."""
import copy
import ctypes as C
import io
import os
import pickle
import re

import pytest
import torch

import sr_gan_fd_b200 as b200
from oracle import rrdbnet_oracle as orc
from sr_gan_fd_b200 import lib as b200lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    from sr_gan_fd_b200.build import build_native
    build_native()
    return b200lib.load()


def test_header_symbols_are_exported(L):
    hdr = open(os.path.join(ROOT, "include", "b200sr.h")).read()
    declared = set(re.findall(r"\b(b200sr_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"b200sr_bucket_cb"}
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/b200sr.h but not exported by libb200sr.so"
    assert set(b200lib.SIGNATURES) == declared
    assert L.b200sr_version() >= 100


def _plan(L, training, batch=16, h=64, w=64, blocks=23, n_up=2, **kw):
    nd = b200lib.NetDesc(kw.get("cin", 3), kw.get("cout", 3), kw.get("channels", 64), kw.get("growth", 32), blocks, n_up, batch, h, w,
                         1 if training else 0)
    handle = C.c_void_p()
    rc = L.b200sr_plan_create(C.byref(nd), C.byref(handle))
    return rc, handle


def test_plan_bookkeeping_matches_work_model(L):
    rc, h = _plan(L, True)
    assert rc == 0
    try:
        assert L.b200sr_num_params(h) == 702
        assert L.b200sr_param_numel(h) == 16_697_987
        px = 16 * 64 * 64
        assert L.b200sr_flops(h, 0) == orc.flops_per_lr_pixel() * px
        assert L.b200sr_flops(h, 1) == orc.flops_per_lr_pixel(backward=True) * px
        assert L.b200sr_num_launches(h, 0) == 2  # ingest + ONE chain launch for all 351 convs
        assert L.b200sr_workspace_bytes(h) > 69 * px * 192 * 2
        assert L.b200sr_packed_bytes(h) > 2 * 16_697_987
    finally:
        L.b200sr_plan_destroy(h)
    rc, h = _plan(L, False)
    assert rc == 0
    assert L.b200sr_workspace_bytes(h) < 2 * 2 ** 30
    L.b200sr_plan_destroy(h)


def test_plan_rejects_unsupported_configs(L):
    for kw in (dict(channels=32), dict(growth=16), dict(cout=17), dict(cin=0)):
        rc, h = _plan(L, False, **kw)
        assert rc < 0 and L.b200sr_last_error()
    rc, h = _plan(L, False, blocks=0)
    assert rc < 0


def test_state_dict_layout():
    net = b200.rrdbnet_x4()
    sd = net.state_dict()
    assert len(sd) == 702 and sum(v.numel() for v in sd.values()) == 16_697_987
    names = orc.conv_names(23, 2)
    assert list(sd.keys()) == [n + s for n in names for s in (".weight", ".bias")]
    shapes = orc.conv_shapes(3, 3, 64, 32, 23, 2)
    for n, (co, ci) in shapes.items():
        assert tuple(sd[n + ".weight"].shape) == (co, ci, 3, 3) and sd[n + ".weight"].dtype == torch.float32
    assert sum(v.numel() for v in b200.bsrgan_x2().state_dict().values()) == 16_661_059


def test_module_protocol_cpu():
    torch.manual_seed(0)
    net = b200.rrdbnet_x4(num_blocks=1)
    net._runtime()  # native runtime state must not leak into copies / pickles / state_dict
    ema = torch.optim.swa_utils.AveragedModel(net, avg_fn=lambda a, m, n: 0.999 * a + 0.001 * m)
    ema.update_parameters(net)
    assert "_b200_runtime" not in ema.module.__dict__
    blob = pickle.dumps(net)
    net2 = pickle.loads(blob)
    assert all(torch.equal(a, b) for a, b in zip(net.state_dict().values(), net2.state_dict().values()))
    buf = io.BytesIO()
    torch.save({"state_dict": net.state_dict()}, buf)
    buf.seek(0)
    net3 = b200.rrdbnet_x4(num_blocks=1)
    net3.load_state_dict(torch.load(buf)["state_dict"])
    net.zero_grad(set_to_none=True)
    net.train(); net.eval()
    assert len(list(net.buffers())) == 0
    assert copy.deepcopy(net).conv1.weight.data_ptr() != net.conv1.weight.data_ptr()


@pytest.mark.parametrize("factory,kw,unshuffle", [
    ("rrdbnet_x4", dict(num_blocks=2), 1), ("rrdbnet_x2", dict(num_blocks=1), 1), ("bsrgan_x2", dict(num_rrdb=1), 1),
    ("RealRRDBNet", dict(in_channels=3, out_channels=3, channels=64, growth_channels=32, num_rrdb=1, upscale_factor=2), 2),
])
def test_cpu_tensors_take_the_eager_torch_path(factory, kw, unshuffle):
    """CPU input -> the module's own nn.Conv2d children through torch (ESRGAN/inference.py defaults to
    --device_type cpu).  Bit-equal to the fp32 restatement, differentiable w.r.t. parameters and input."""
    from oracle import rrdbnet_oracle as orc
    torch.manual_seed(3)
    net = getattr(b200, factory)(**kw)
    params = orc.in_range_fixture({k: v.detach().clone() for k, v in net.state_dict().items()})
    net.load_state_dict(params)
    x = torch.rand(2, 3, 12, 8, requires_grad=True)
    y = net(x)
    ref = orc.rrdbnet_forward(params, x.detach(), unshuffle)
    assert torch.equal(y.detach(), ref)
    y.mean().backward()
    assert x.grad is not None and float(x.grad.abs().sum()) > 0
    assert all(p.grad is not None for p in net.parameters())


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """The CUDA path never falls back: without libb200sr.so loading raises (checked without touching a GPU)."""
    from sr_gan_fd_b200 import lib as _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setenv("B200SR_LIB", str(tmp_path / "missing.so"))
    with pytest.raises(RuntimeError, match="missing"):
        _lib.load()


def test_factories_and_kwargs():
    assert b200.rrdbnet_x2(num_blocks=1).upscale_factor == 2
    assert hasattr(b200.rrdbnet_x8(num_blocks=1), "upsampling3")
    assert not hasattr(b200.rrdbnet_x1(num_blocks=1), "upsampling1")
    m = b200.bsrgan_x2(num_rrdb=1)
    assert hasattr(m, "upsampling1") and not hasattr(m, "upsampling2")
    r = b200.RealRRDBNet(3, 3, 64, 32, 1, 2)
    assert r.conv1.in_channels == 12 and hasattr(r, "upsampling2")


def test_compat_shims_export_reference_names():
    import importlib
    e = importlib.import_module("sr_gan_fd_b200.compat.esrgan_model")
    for n in ("RRDBNet", "rrdbnet_x1", "rrdbnet_x2", "rrdbnet_x4", "rrdbnet_x8"):
        assert n in e.__dict__
    bs = importlib.import_module("sr_gan_fd_b200.compat.bsrgan_model")
    for n in ("BSRGAN", "bsrgan_x2", "bsrgan_x4"):
        assert n in bs.__dict__
    re_ = importlib.import_module("sr_gan_fd_b200.compat.real_esrgan_model")
    assert "RRDBNet" in re_.__dict__ and "rrdbnet_x4" in re_.__dict__
    ae = importlib.import_module("sr_gan_fd_b200.compat.a_esrgan_model")
    assert ae.BSRGAN is b200.BSRGAN and ae.bsrgan_x2 is b200.bsrgan_x2
    import os
    if os.path.isfile("/root/reference/A-ESRGAN/model.py"):  # the folder's other entry points are the reference's own objects
        for n in ("uNetDiscriminatorAesrgan", "bsrgantrans_x2", "gen_rrdb2x", "ContentLoss"):
            assert n in ae.__dict__
        assert ae.__dict__["uNetDiscriminatorAesrgan"].__module__.startswith("_srganfd_reference")


def test_schedule_dependent_bookkeeping(L):
    """The dense-block schedule depends on the geometry: small batches use the windowed re-association, a frame with more
    8x32 items than SMs the per-conv schedule -- different packed layouts; large batches cut the chains; the backward pass
    is two data-gradient chains (HR tail, then the trunk next to the tail's weight gradients) plus the weight-gradient launches."""
    L.b200sr_pack_layout_id.restype = C.c_uint64
    rc, small = _plan(L, False, batch=16, h=64, w=64)
    rc2, frame = _plan(L, False, batch=1, h=512, w=512)
    rc3, small2 = _plan(L, False, batch=8, h=64, w=64)
    assert rc == 0 and rc2 == 0 and rc3 == 0
    try:
        assert L.b200sr_pack_layout_id(small) != 0
        assert L.b200sr_pack_layout_id(small) == L.b200sr_pack_layout_id(small2)   # same schedule, other batch
        assert L.b200sr_pack_layout_id(small) != L.b200sr_pack_layout_id(frame)    # windowed vs per-conv
        assert L.b200sr_packed_bytes(small) != L.b200sr_packed_bytes(frame)
    finally:
        for h in (small, frame, small2):
            L.b200sr_plan_destroy(h)
    rc, big = _plan(L, True, batch=64)
    rc2, base = _plan(L, True, batch=16)
    assert rc == 0 and rc2 == 0
    try:
        assert L.b200sr_num_launches(base, 0) == 2
        assert L.b200sr_num_launches(big, 0) > 2                                    # 351 layers x 8 groups: the chain is cut
        # backward: ingest + TWO chains (cut once behind the HR tail) + 77 weight-gradient launches + bias grads + add + one merged unpack
        n1, n2 = L.b200sr_num_launches(base, 1), L.b200sr_num_launches(base, 2)
        assert n1 == 88, n1
        assert n2 - n1 == 24                                                         # 25 gradient buckets instead of one unpack
    finally:
        L.b200sr_plan_destroy(big)
        L.b200sr_plan_destroy(base)


def test_reference_root_and_staged_reference_generator():
    """The reference tree resolves to $SRGANFD_REFERENCE, /root/reference or the staged copy under baseline/_ref, and bench.py's
    reference arm builds the reference's OWN rrdbnet_x4 from the staged copy (cpu_baseline.kind = "reference") when it exists."""
    import importlib.util
    import sys
    from sr_gan_fd_b200.compat._passthrough import reference_root
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    old = os.environ.pop("SRGANFD_REFERENCE", None)
    try:
        r = reference_root()
        assert r == "/root/reference" if os.path.isdir("/root/reference") else r == os.path.join(root, "baseline", "_ref")
        os.environ["SRGANFD_REFERENCE"] = "/somewhere/else"
        assert reference_root() == "/somewhere/else"
    finally:
        os.environ.pop("SRGANFD_REFERENCE", None)
        if old is not None:
            os.environ["SRGANFD_REFERENCE"] = old
    if not os.path.isfile(os.path.join(root, "baseline", "_ref", "ESRGAN", "model.py")):
        pytest.skip("no staged reference copy (run __graft_entry__.build() where /root/reference exists)")
    spec = importlib.util.spec_from_file_location("_bench_for_test", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    sys.modules["_bench_for_test"] = bench
    spec.loader.exec_module(bench)
    net = bench._reference_generator()
    assert net is not None and type(net).__module__ == "_reference_esrgan_model"   # the reference's class, not the drop-in
    assert sum(p.numel() for p in net.parameters()) == 16_697_987
    mine = b200.rrdbnet_x4()
    assert list(net.state_dict().keys()) == list(mine.state_dict().keys())
