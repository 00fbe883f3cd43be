"""Packed rows (arb_set_pack_rows, include/allrank_b200.h): the encoder runs over the items below every slate's extent
only.  Against the dense layout -- the computation the reference performs (model.py:62-92 over all B * S rows) -- the
scores of every item below the extent must agree BIT FOR BIT (every row-wise kernel and GEMM row is independent of the
rows around it, and the attention kernels see the same query / key tiles of the slate), the items beyond it score 0,
and every parameter gradient agrees up to the fp32 summation order of the split-K weight-gradient reductions."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _model(N=2, d=128, h=4, dff=256, F=136, sizes=None, act=None, input_norm=False, out_act=None, dtype="tf32"):
    from allrank_b200.model import make_model
    torch.manual_seed(3)
    m = make_model(fc_model={"sizes": sizes or [d], "input_norm": input_norm, "activation": act, "dropout": 0.0},
                   transformer={"N": N, "d_ff": dff, "h": h, "positional_encoding": None, "dropout": 0.0},
                   post_model={"d_output": 1, "output_activation": out_act}, n_features=F, compute_dtype=dtype)
    gen = torch.Generator().manual_seed(4)
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn(p.shape, generator=gen))
    return m.cuda()


def _slates(B, S, F=136, seed=19, mean_len=None, std_len=None):
    from allrank_b200.synth import make_slates
    x, y, _ = make_slates(B, S, F, seed=seed, mean_len=mean_len or S / 2, std_len=std_len or S / 4)
    g = torch.Generator().manual_seed(seed + 1)
    y[0] = torch.ones(S)                                   # one full slate
    x[0] = torch.randn(S, F, generator=g)
    if B > 2:                                              # a slate of a single item
        y[2] = -1.0
        y[2, 0] = 1.0
        x[2, 1:] = 0.0
    if B > 3 and S > 8:                                    # a padded item INSIDE a slate (mask true below the extent)
        y[3, :8] = 1.0
        x[3, :8] = torch.randn(8, F, generator=g)
        y[3, 5] = -1.0
    return x, y


def _extents(y):
    S = y.shape[1]
    pos = torch.arange(S, device=y.device)[None, :]
    return torch.where(y != -1, pos + 1, torch.zeros_like(pos)).max(1).values


def _run(model, x, y, w, pack, train=True):
    from allrank_b200 import _lib
    lib = _lib.lib()
    lib.arb_set_pack_rows(pack)
    try:
        model.train(train)
        model.zero_grad(set_to_none=True)
        s = model(x, y == -1, None)
        if not train:
            return s.detach().clone(), None
        (s * w).sum().backward()
        return s.detach().clone(), model.flat_gradients.clone()
    finally:
        lib.arb_set_pack_rows(_lib.default_pack_rows())


@pytest.mark.parametrize("B,S", [(24, 240), (7, 120), (5, 16), (33, 48), (300, 240)])
@pytest.mark.parametrize("train", [True, False])
def test_packed_rows_equal_dense_rows(B, S, train):
    model = _model()
    x, y = _slates(B, S)
    x, y = x.cuda(), y.cuda()
    w = torch.randn(B, S, generator=torch.Generator().manual_seed(2)).cuda()
    w = torch.where(y == -1, torch.zeros_like(w), w)
    dense_s, dense_g = _run(model, x, y, w, 0, train)
    pack_s, pack_g = _run(model, x, y, w, 1, train)
    ext = _extents(y)
    pos = torch.arange(S, device="cuda")[None, :]
    below = pos < ext[:, None]
    beyond = pos >= ((ext + 15) // 16 * 16)[:, None]
    assert torch.isfinite(pack_s).all()
    assert torch.equal(pack_s[below], dense_s[below])
    assert (pack_s[beyond] == 0).all()
    if train:
        gap = (pack_g - dense_g).abs().max().item()
        assert gap <= 4e-6 * dense_g.abs().max().item(), gap


def test_packed_rows_drop_the_gradient_of_items_beyond_the_extent_and_survive_empty_slates():
    """The items beyond a slate's extent have the constant score 0: a gradient placed on them reaches no parameter
    (the dense layout would propagate it through the padded row's own activations); a slate without any item -- which
    the dense layout, like the reference, turns into NaN rows (quirk Q2) -- contributes nothing."""
    model = _model()
    B, S = 12, 240
    x, y = _slates(B, S)
    y[5] = -1.0
    x[5] = 0.0
    x, y = x.cuda(), y.cuda()
    w = torch.randn(B, S, generator=torch.Generator().manual_seed(2)).cuda()
    w = torch.where(y == -1, torch.zeros_like(w), w)
    s0, g0 = _run(model, x, y, w, 1)
    assert torch.isfinite(s0).all() and torch.isfinite(g0).all() and (s0[5] == 0).all()
    ext = _extents(y)
    w2 = w.clone()
    b = int((ext < S - 32).nonzero()[0])
    w2[b, S - 1] = 0.7
    s1, g1 = _run(model, x, y, w2, 1)
    assert torch.equal(s0, s1)
    assert (g1 - g0).abs().max().item() <= 4e-6 * g0.abs().max().item()
    # without the empty slate the packed gradients equal the dense ones
    keep = torch.arange(B, device="cuda") != 5
    _, gd = _run(model, x[keep], y[keep], w[keep], 0)
    _, gp = _run(model, x[keep], y[keep], w[keep], 1)
    assert (gp - gd).abs().max().item() <= 4e-6 * gd.abs().max().item()
    assert (g0 - gd).abs().max().item() <= 4e-6 * gd.abs().max().item()


@pytest.mark.parametrize("kw", [dict(sizes=[96, 128], act="ReLU", input_norm=True), dict(act="Tanh", out_act="Sigmoid"),
                                dict(N=1, d=64, h=4, dff=64), dict(dtype="bf16"), dict(N=3, d=256, h=8, dff=512)])
def test_packed_rows_equal_dense_rows_for_other_models(kw):
    """FC block with activations / input norm, head width 16, bf16 mode, a deeper and wider encoder."""
    model = _model(**kw)
    B, S = 40, 240
    x, y = _slates(B, S)
    x, y = x.cuda(), y.cuda()
    w = torch.randn(B, S, generator=torch.Generator().manual_seed(2)).cuda()
    w = torch.where(y == -1, torch.zeros_like(w), w)
    dense_s, dense_g = _run(model, x, y, w, 0)
    pack_s, pack_g = _run(model, x, y, w, 1)
    below = torch.arange(S, device="cuda")[None, :] < _extents(y)[:, None]
    assert torch.equal(pack_s[below], dense_s[below])
    gap = (pack_g - dense_g).abs().max().item()
    assert gap <= (2e-3 if kw.get("dtype") == "bf16" else 4e-6) * dense_g.abs().max().item(), gap


def test_packed_rows_train_like_dense_rows():
    """Three SGD steps on approxNDCGLoss (the headline configuration's loss) from the same initialisation: same losses,
    same parameters.  (SGD, not Adam: Adam divides by |g|, which turns the 1e-6 summation-order differences of the
    near-zero gradient entries into O(lr) parameter differences.)"""
    from allrank_b200 import _lib
    from allrank_b200.losses import approxNDCGLoss
    B, S = 64, 240
    x, y = _slates(B, S)
    x, y = x.cuda(), y.cuda()
    finals = []
    for pack in (0, 1):
        model = _model()
        opt = torch.optim.SGD(model.parameters(), lr=0.05)
        _lib.lib().arb_set_pack_rows(pack)
        try:
            model.train()
            losses = []
            for _ in range(3):
                loss = approxNDCGLoss(model(x, y == -1, None), y)
                opt.zero_grad()
                loss.backward()
                opt.step()
                losses.append(loss.item())
        finally:
            _lib.lib().arb_set_pack_rows(_lib.default_pack_rows())
        finals.append((losses, model.flat_parameters.clone()))
    assert finals[0][0][0] != finals[0][0][2]                       # the steps moved the model
    for a, b in zip(finals[0][0], finals[1][0]):
        assert abs(a - b) <= 2e-5 * abs(a), (a, b)
    assert (finals[0][1] - finals[1][1]).abs().max().item() <= 1e-5 * finals[0][1].abs().max().item()


@pytest.mark.parametrize("pack", [1, 0])
def test_relu_bit_mask_backward_equals_the_mask_tile(pack):
    """arb_set_relu_bits: the FFN's ReLU backward from the 1-bit-per-unit mask the W1 epilogue writes must equal the
    backward that re-reads the fp32 hidden activation as a mask tile -- same scores bit for bit, gradients up to the
    summation order of the split-K reductions; packed and dense rows."""
    import ctypes
    from allrank_b200 import _lib
    lib = _lib.lib()
    lib.arb_set_relu_bits.argtypes = [ctypes.c_int32]
    model = _model(N=2, d=128, h=4, dff=512)
    B, S = 48, 240
    x, y = _slates(B, S)
    x, y = x.cuda(), y.cuda()
    w = torch.randn(B, S, generator=torch.Generator().manual_seed(2)).cuda()
    w = torch.where(y == -1, torch.zeros_like(w), w)
    outs = []
    for bits in (0, 1):
        lib.arb_set_relu_bits(bits)
        try:
            outs.append(_run(model, x, y, w, pack))
        finally:
            lib.arb_set_relu_bits(_lib.default_relu_bits())
    assert torch.equal(outs[0][0], outs[1][0])
    gap = (outs[0][1] - outs[1][1]).abs().max().item()
    assert gap <= 4e-6 * outs[0][1].abs().max().item(), gap
    assert outs[1][1].abs().max().item() > 0
