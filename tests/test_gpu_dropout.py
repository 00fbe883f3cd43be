"""Dropout fused into the scorer kernels (counter-based masks regenerated in backward).

The mask stream cannot equal torch's Philox stream, so parity with the reference under dropout is statistical
(SURVEY.md section 7).  What is checked: nn.Dropout semantics (train() only, repeatable under torch.manual_seed,
keep rate and 1/(1-p) scaling), and that backward uses exactly the forward's masks (directional finite differences
with the seed held fixed)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def make(F=136, d=128, N=2, h=4, dff=512, p=0.3, p_fc=0.0):
    from allrank_b200.model import make_model
    tcfg = {"N": N, "d_ff": dff, "h": h, "positional_encoding": None, "dropout": p} if N > 0 else None
    return make_model(fc_model={"sizes": [d], "input_norm": False, "activation": None, "dropout": p_fc},
                      transformer=tcfg, post_model={"d_output": 1, "output_activation": None}, n_features=F).cuda()


def data(B=6, S=64, F=136, seed=5):
    from allrank_b200.synth import make_slates
    x, y, _ = make_slates(B, S, n_features=F, seed=seed, mean_len=0.7 * S, std_len=0.2 * S)
    return x.cuda(), (y == -1).cuda()


def test_train_eval_semantics_and_repeatability():
    torch.manual_seed(1)
    model = make(p=0.3, p_fc=0.1)
    x, mask = data()
    with torch.no_grad():
        e1 = model.eval()(x, mask, None)
        e2 = model.eval()(x, mask, None)
        assert torch.equal(e1, e2)                       # eval: dropout off, deterministic
        model.train()
        torch.manual_seed(7); a = model(x, mask, None)
        torch.manual_seed(7); b = model(x, mask, None)
        torch.manual_seed(8); c = model(x, mask, None)
    assert torch.equal(a, b)                              # same seed -> same masks
    assert not torch.equal(a, c)                          # different seed -> different masks
    assert not torch.equal(a, e1)
    assert torch.isfinite(a).all()


def test_fc_dropout_keep_rate_and_scale():
    """FC-only model with a one-hot head: scores ARE one column of dropout(FC(x)), so the mask is observable."""
    torch.manual_seed(2)
    p = 0.25
    model = make(F=20, d=64, N=0, p=0.0, p_fc=p)
    x, mask = data(B=64, S=120, F=20)
    with torch.no_grad():
        model.output_layer.w_1.weight.zero_()
        model.output_layer.w_1.weight[0, 5] = 1.0
        model.output_layer.w_1.bias.zero_()
        ref = model.eval()(x, mask, None)                 # = FC(x)[:, :, 5]
        torch.manual_seed(3)
        out = model.train()(x, mask, None)
    kept = out != 0
    frac = kept.float().mean().item()
    assert abs(frac - (1 - p)) < 0.02, frac
    assert torch.allclose(out[kept], ref[kept] / (1 - p), rtol=1e-5, atol=1e-6)


from tests.dropout_masks import mask_tensor as _mask  # noqa: E402


@pytest.mark.parametrize("p,p_fc", [(0.3, 0.0), (0.15, 0.2)])
def test_forward_and_backward_match_reference_maths_under_the_same_masks(p, p_fc):
    """Regenerate the kernels' dropout masks on the host, feed them to the eager functional scorer (same maths as the
    reference, oracle/tf32_emulation.py with TF32-rounded matmuls) and compare scores and every parameter gradient:
    proves the fused forward applies the masks where the reference applies dropout and that backward reuses them."""
    from oracle.tf32_emulation import scorer_forward
    F, d, N, h, dff, B, S = 136, 64, 2, 2, 128, 5, 48
    torch.manual_seed(4)
    model = make(F=F, d=d, N=N, h=h, dff=dff, p=p, p_fc=p_fc).train()
    x, mask = data(B=B, S=S)
    w = torch.randn(B, S)
    torch.manual_seed(21)
    call_seed = int(torch.randint(0, 2 ** 62, (1,)).item())      # what LTRModel._draw_seed will draw
    torch.manual_seed(21)
    scores = model(x, mask, None)
    (scores * w.cuda()).sum().backward()
    R = B * S
    drop = {(0, "fc"): _mask((R, d), call_seed, 0, 0, p_fc)}
    for l in range(N):
        drop[(l, "attn_p")] = _mask((B, h, S, S), call_seed, l, 1, p)
        drop[(l, "attn_out")] = _mask((R, d), call_seed, l, 2, p)
        drop[(l, "ffn_hid")] = _mask((R, dff), call_seed, l, 3, p)
        drop[(l, "ffn_out")] = _mask((R, d), call_seed, l, 4, p)
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    ref = scorer_forward(sd, x.cpu(), mask.cpu(), N, h, None, "rna", drop=drop)
    (ref * w).sum().backward()
    err = (ref.detach() - scores.detach().cpu()).abs().max().item()
    assert err <= 3e-3 * max(1.0, ref.abs().max().item()), err
    floor = 1e-2 * max(v.grad.abs().max().item() for v in sd.values())
    worst = 0.0
    for k, q in model.named_parameters():
        a, r = q.grad.cpu().double().numpy(), sd[k].grad.double().numpy()
        fro = np.linalg.norm(a - r) / max(np.linalg.norm(r), floor * np.sqrt(r.size))
        worst = max(worst, fro)
        # gradients below a ReLU inherit the forward's TF32 rounding through units whose pre-activation is within that
        # noise of zero (derivative 1 on one side, 0 on the other): a few per cent here, an order more would be a mask
        # mismatch (tests/test_shipped_configs.py explains the measurement)
        last = f"encoder.layers.{N - 1}.feed_forward.w_2"
        tight = k.startswith(last) or k.startswith("encoder.norm") or k.startswith("output_layer")
        assert fro <= (3e-2 if tight else 1e-1), (k, fro)
    print("dropout p =", p, "fc", p_fc, ": score err", err, "worst grad rel err", worst)


def test_training_with_dropout_reduces_loss():
    from allrank_b200 import losses
    from allrank_b200.optim import FlatAdam
    from allrank_b200.synth import make_slates
    torch.manual_seed(6)
    model = make(F=136, d=64, N=1, h=2, dff=128, p=0.2).train()
    x, y, _ = make_slates(64, 60, seed=9, mean_len=40, std_len=10)
    x, y = x.cuda(), y.cuda()
    opt = FlatAdam(model, lr=2e-3)
    first = last = None
    for i in range(30):
        loss = losses.listNet(model(x, y == -1, None), y)
        loss.backward(); opt.step(); opt.zero_grad()
        first = loss.item() if first is None else first
        last = loss.item()
    assert np.isfinite(last) and last < first
