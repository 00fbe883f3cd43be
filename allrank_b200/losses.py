"""Drop-in replacements for the listwise members of allrank.models.losses -- same names, positional
order, keyword arguments and defaults as the reference:

    listNet         /root/reference/allrank/models/losses/listNet.py:8
    listMLE         .../listMLE.py:7
    approxNDCGLoss  .../approxNDCG.py:7
    lambdaLoss      .../lambdaLoss.py:7   (7 weighing schemes :84-114)
    neuralNDCG      .../neuralNDCG.py:10  (deterministic NeuralSort + Sinkhorn, loss_utils.py:8-67)

Each call is ONE fused forward+backward kernel launch (+ a tiny batch-finalise launch) through the C ABI;
the result is a 0-dim tensor on y_pred's device attached to autograd (gradient w.r.t. y_pred only), which is
what allrank/training/train_utils.py:20-29 needs (`loss.backward()`, `loss.item()`).
No CPU fallback: CPU tensors raise.
"""
import torch

from . import _lib
from .metrics import discount_table

PADDED_Y_VALUE = -1   # allrank/data/dataset_loading.py:15
DEFAULT_EPS = 1e-10   # allrank/models/losses/__init__.py:1

_SCHEMES = {
    None: 0,
    "ndcgLoss1_scheme": 1,
    "ndcgLoss2_scheme": 2,
    "lambdaRank_scheme": 3,
    "ndcgLoss2PP_scheme": 4,
    "rankNet_scheme": 5,
    "rankNetWeightedByGTDiff_scheme": 6,
    "rankNetWeightedByGTDiffPowed_scheme": 7,
}


class _FusedLoss(torch.autograd.Function):
    """forward runs the fused kernel (loss + d loss/d y_pred); backward scales the saved gradient."""

    @staticmethod
    def forward(ctx, y_pred, launcher):
        need_grad = ctx.needs_input_grad[0]
        scores = y_pred.detach().float().contiguous()
        loss = torch.empty((), dtype=torch.float32, device=y_pred.device)
        grad = torch.empty_like(scores) if need_grad else None
        with torch.cuda.device(y_pred.device):
            launcher(scores, loss, grad)
        ctx.in_dtype = y_pred.dtype
        if need_grad:
            ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, upstream):
        (grad,) = ctx.saved_tensors
        return (upstream * grad).to(ctx.in_dtype), None


def _run(y_pred, y_true, launcher):
    _lib.require_cuda(y_pred, y_true)
    if y_pred.dim() != 2 or y_pred.shape != y_true.shape:
        raise ValueError("y_pred and y_true must both be [batch_size, slate_length]")
    if y_pred.shape[0] == 0:
        raise ValueError("empty batch")
    labels = y_true.detach().float().contiguous()
    B, S = labels.shape
    scratch = torch.empty(2 * B, dtype=torch.float32, device=labels.device)

    def bound(scores, loss, grad):
        launcher(scores, labels, B, S, loss, grad, scratch)

    # needs_input_grad is decided by autograd; grad is only computed when y_pred requires it
    return _FusedLoss.apply(y_pred, bound)


def listNet(y_pred, y_true, eps=DEFAULT_EPS, padded_value_indicator=PADDED_Y_VALUE):
    """ListNet loss -- listNet.py:8-30."""

    def launch(s, t, B, S, loss, grad, scratch):
        rc = _lib.lib().arb_listnet(_lib.ptr(s), _lib.ptr(t), B, S, float(eps), float(padded_value_indicator),
                                    _lib.ptr(loss), _lib.ptr(grad), _lib.ptr(scratch), _lib.stream_ptr(s.device))
        _lib.check(rc, "arb_listnet")

    return _run(y_pred, y_true, launch)


def listMLE(y_pred, y_true, eps=DEFAULT_EPS, padded_value_indicator=PADDED_Y_VALUE, perm=None, order=None):
    """ListMLE loss -- listMLE.py:7-38.

    The reference shuffles columns with `torch.randperm(S)` drawn from the global CPU RNG (:17) before its
    (unstable) label sort; we draw the same permutation from the same RNG.  Extensions (not in the reference
    signature): `perm` fixes the shuffle, `order` ([B,S] int) feeds a realised sort order of the shuffled
    labels (parity hook).
    """
    S = y_pred.shape[-1]
    if perm is None:
        perm = torch.randperm(S)
    perm_dev = perm.to(device=y_pred.device, dtype=torch.int64).contiguous()
    order_dev = None if order is None else order.to(device=y_pred.device, dtype=torch.int32).contiguous()

    def launch(s, t, B, S_, loss, grad, scratch):
        rc = _lib.lib().arb_listmle(_lib.ptr(s), _lib.ptr(t), B, S_, float(eps), float(padded_value_indicator),
                                    _lib.ptr(perm_dev), _lib.ptr(order_dev), _lib.ptr(loss), _lib.ptr(grad),
                                    _lib.ptr(scratch), _lib.stream_ptr(s.device))
        _lib.check(rc, "arb_listmle")

    return _run(y_pred, y_true, launch)


def approxNDCGLoss(y_pred, y_true, eps=DEFAULT_EPS, padded_value_indicator=PADDED_Y_VALUE, alpha=1.):
    """ApproxNDCG loss (no @k truncation) -- approxNDCG.py:7-53."""

    def launch(s, t, B, S, loss, grad, scratch):
        rc = _lib.lib().arb_approx_ndcg(_lib.ptr(s), _lib.ptr(t), B, S, float(eps), float(padded_value_indicator),
                                        float(alpha), _lib.ptr(loss), _lib.ptr(grad), _lib.ptr(scratch),
                                        _lib.stream_ptr(s.device))
        _lib.check(rc, "arb_approx_ndcg")

    return _run(y_pred, y_true, launch)


def lambdaLoss(y_pred, y_true, eps=DEFAULT_EPS, padded_value_indicator=PADDED_Y_VALUE, weighing_scheme=None, k=None,
               sigma=1., mu=10., reduction="sum", reduction_log="binary"):
    """LambdaLoss framework -- lambdaLoss.py:7-81; same error conventions (:72, :79, KeyError at :61)."""
    if weighing_scheme not in _SCHEMES:
        raise KeyError(weighing_scheme)
    if reduction_log not in ("natural", "binary"):
        raise ValueError("Reduction logarithm base can be either natural or binary")
    if reduction not in ("sum", "mean"):
        raise ValueError("Reduction method can be either sum or mean")
    scheme = _SCHEMES[weighing_scheme]
    kk = 0 if k is None else int(k)
    if k is not None and kk <= 0:
        raise ValueError("k must be a positive rank or None")

    def launch(s, t, B, S, loss, grad, scratch):
        rc = _lib.lib().arb_lambda_loss(_lib.ptr(s), _lib.ptr(t), B, S, float(eps), float(padded_value_indicator),
                                        scheme, kk, float(sigma), float(mu), 1 if reduction == "mean" else 0,
                                        1 if reduction_log == "natural" else 0, _lib.ptr(loss), _lib.ptr(grad),
                                        _lib.ptr(scratch), _lib.stream_ptr(s.device))
        _lib.check(rc, "arb_lambda_loss")

    return _run(y_pred, y_true, launch)


def _neural_ndcg_launch(scores_2d, labels_2d, pad, temperature, powered, kk, max_iter, tol):
    """Deterministic NeuralSort + Sinkhorn + soft DCG on [N,S] rows (N = slates, or samples x slates)."""
    S = scores_2d.shape[-1]
    dev = scores_2d.device
    disc = discount_table(S, dev)
    ws_bytes = int(_lib.lib().arb_neural_ndcg_workspace_bytes(scores_2d.shape[0], S, int(max_iter)))
    ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=dev) if ws_bytes else None

    def launch(s, t, B, S_, loss, grad, scratch):
        rc = _lib.lib().arb_neural_ndcg(_lib.ptr(s), _lib.ptr(t), B, S_, _lib.ptr(disc), float(pad), float(temperature),
                                        int(powered), kk, int(max_iter), float(tol), _lib.ptr(loss),
                                        _lib.ptr(grad), _lib.ptr(scratch), _lib.ptr(ws), ws_bytes,
                                        _lib.stream_ptr(s.device))
        _lib.check(rc, "arb_neural_ndcg")

    return _run(scores_2d, labels_2d, launch)


def _gumbel_perturbed(y_pred, n_samples, beta, log_scores, eps=1e-10):
    """stochastic_neural_sort's score perturbation (loss_utils.py:96-104): O(S) host-side tensor algebra, kept in
    PyTorch so autograd carries the gradient back to y_pred; the O(S^2) part stays in the fused kernel."""
    s_positive = y_pred + torch.abs(y_pred.min())
    u = torch.rand([n_samples, y_pred.shape[0], y_pred.shape[1]], device=y_pred.device)
    samples = beta * (-torch.log(-torch.log(u + eps) + eps))          # sample_gumbel, loss_utils.py:70-81
    if log_scores:
        s_positive = torch.log(s_positive + eps)
    return (s_positive.unsqueeze(0) + samples).reshape(n_samples * y_pred.shape[0], y_pred.shape[1])


def neural_sort_matrices(y_pred, y_true, temperature=1., max_iter=50, tol=1e-6, padded_value_indicator=PADDED_Y_VALUE):
    """Parity hook (not part of the reference's surface): the matrices the fused neuralNDCG kernel works with, for
    slates of at most 128 items -- (deterministic_neural_sort(y_pred, tau, mask), sinkhorn_scaling(., mask, tol,
    max_iter)) of loss_utils.py:34-67 / :8-31, as [B,S,S] tensors whose padded rows / columns are zero."""
    import ctypes
    _lib.require_cuda(y_pred, y_true)
    s = y_pred.detach().float().contiguous()
    t = y_true.detach().float().contiguous()
    B, S = s.shape
    p0 = torch.zeros(B, S, S, device=s.device)
    p = torch.zeros(B, S, S, device=s.device)
    scratch = torch.empty(2 * B, device=s.device)
    disc = discount_table(S, s.device)
    fn = _lib.lib().arb_neural_sort_debug
    fn.restype = ctypes.c_int32
    c_p, c_i, c_f = ctypes.c_void_p, ctypes.c_int32, ctypes.c_float
    fn.argtypes = [c_p, c_p, c_i, c_i, c_p, c_f, c_f, c_i, c_f, c_p, c_p, c_p, c_p]
    with torch.cuda.device(s.device):
        rc = fn(_lib.ptr(s), _lib.ptr(t), B, S, _lib.ptr(disc), float(padded_value_indicator), float(temperature),
                int(max_iter), float(tol), _lib.ptr(p0), _lib.ptr(p), _lib.ptr(scratch), _lib.stream_ptr(s.device))
    _lib.check(rc, "arb_neural_sort_debug")
    return p0, p


def neuralNDCG(y_pred, y_true, padded_value_indicator=PADDED_Y_VALUE, temperature=1., powered_relevancies=True, k=None,
               stochastic=False, n_samples=32, beta=0.1, log_scores=True, max_iter=50, tol=1e-6, _gain_mode=None):
    """NeuralNDCG loss -- neuralNDCG.py:10-70.

    `max_iter` / `tol` are the Sinkhorn parameters the reference hard-codes to 50 / 1e-6 (:41-42); they are
    exposed here (defaults unchanged) because BASELINE.json also quotes a 30-iteration configuration.
    stochastic=True draws n_samples Gumbel perturbations per slate (loss_utils.py:84-112) and averages over
    samples x slates-with-idcg>0 exactly like neuralNDCG.py:69 (statistical parity: the RNG stream differs).
    """
    kk = 0 if k is None else int(k)
    if stochastic:
        scores = _gumbel_perturbed(y_pred, int(n_samples), float(beta), bool(log_scores))
        labels = y_true.repeat(int(n_samples), 1)
    else:
        scores, labels = y_pred, y_true
    mode = _gain_mode if _gain_mode is not None else (1 if powered_relevancies else 0)
    return _neural_ndcg_launch(scores, labels, padded_value_indicator, temperature, mode, kk, max_iter, tol)


def neuralNDCG_transposed(y_pred, y_true, padded_value_indicator=PADDED_Y_VALUE, temperature=1.,
                          powered_relevancies=True, k=None, stochastic=False, n_samples=32, beta=0.1, log_scores=True,
                          max_iter=50, tol=1e-6):
    """NeuralNDCG Transposed -- neuralNDCG.py:73-136.  It evaluates sum_i g_i (P^T disc)_i, which is the same
    bilinear form sum_{j,i} disc_j P[j,i] g_i as neuralNDCG (:48-55) on the same Sinkhorn-scaled matrix, with the
    same mean over slates with idcg != 0, so both names map onto one kernel; the transposed signature exposes
    max_iter / tol (:75).  Quirk kept: with powered_relevancies=False the reference still normalises by the
    2^x-1 ideal DCG (:126-128)."""
    return neuralNDCG(y_pred, y_true, padded_value_indicator, temperature, powered_relevancies, k, stochastic,
                      n_samples, beta, log_scores, max_iter, tol, _gain_mode=1 if powered_relevancies else 2)


def rankNet(y_pred, y_true, padded_value_indicator=PADDED_Y_VALUE, weight_by_diff=False, weight_by_diff_powed=False):
    """RankNet -- rankNet.py:31-79 (BCE-with-logits over every ordered pair with t_i > t_j, mean over pairs)."""
    mode = 1 if weight_by_diff else (2 if weight_by_diff_powed else 0)

    def launch(s, t, B, S, loss, grad, scratch):
        rc = _lib.lib().arb_ranknet(_lib.ptr(s), _lib.ptr(t), B, S, float(padded_value_indicator), mode, _lib.ptr(loss),
                                    _lib.ptr(grad), _lib.ptr(scratch), _lib.stream_ptr(s.device))
        _lib.check(rc, "arb_ranknet")

    return _run(y_pred, y_true, launch)


def rankNet_weightByGTDiff(y_pred, y_true, padded_value_indicator=PADDED_Y_VALUE):
    """rankNet.py:9-17"""
    return rankNet(y_pred, y_true, padded_value_indicator, weight_by_diff=True)


def rankNet_weightByGTDiff_pow(y_pred, y_true, padded_value_indicator=PADDED_Y_VALUE):
    """rankNet.py:20-28"""
    return rankNet(y_pred, y_true, padded_value_indicator, weight_by_diff=False, weight_by_diff_powed=True)


def _pointwise(y_pred, y_true, pad, mode, param, eps):
    def launch(s, t, B, S, loss, grad, scratch):
        rc = _lib.lib().arb_pointwise_loss(_lib.ptr(s), _lib.ptr(t), B, S, float(pad), mode, float(param), float(eps),
                                           _lib.ptr(loss), _lib.ptr(grad), _lib.ptr(scratch), _lib.stream_ptr(s.device))
        _lib.check(rc, "arb_pointwise_loss")

    return _run(y_pred, y_true, launch)


def binary_listNet(y_pred, y_true, eps=DEFAULT_EPS, padded_value_indicator=PADDED_Y_VALUE):
    """ListNet for binary labels (labels normalised by their sum) -- binary_listNet.py:8-33."""
    return _pointwise(y_pred, y_true, padded_value_indicator, 0, 0.0, eps)


def pointwise_rmse(y_pred, y_true, no_of_levels, padded_value_indicator=PADDED_Y_VALUE):
    """Pointwise RMSE between labels and no_of_levels * scores -- pointwise.py:6-32."""
    return _pointwise(y_pred, y_true, padded_value_indicator, 1, float(no_of_levels), 0.0)


def bce(y_pred, y_true, padded_value_indicator=PADDED_Y_VALUE):
    """Binary cross-entropy on probabilities, padded items ignored, summed per slate and divided by the number of
    non-empty slates -- bce.py:8-32.  (With padded targets the reference itself only runs on torch < 2:
    nn.BCELoss now rejects the -1 targets before they are masked; the intended semantics are implemented.)"""
    return _pointwise(y_pred, y_true, padded_value_indicator, 2, 0.0, 0.0)


def with_ordinals(y, n, padded_value_indicator=PADDED_Y_VALUE):
    """Labels [B,S] -> ordinal targets [B,S,n]: level j is 1 where y >= j+1, padded items keep the padding value --
    ordinal.py:8-22.  (Target preparation only; the loss kernel derives the targets on the fly.)"""
    _lib.require_cuda(y)
    levels = torch.arange(1, n + 1, dtype=torch.float32, device=y.device)
    spread = y.unsqueeze(2).repeat(1, 1, n)
    out = (spread >= levels).float()
    out[spread == padded_value_indicator] = padded_value_indicator
    return out


def ordinal(y_pred, y_true, n, padded_value_indicator=PADDED_Y_VALUE):
    """Ordinal loss -- ordinal.py:25-50: y_pred [B,S,n] are per-level probabilities (a d_output = n model with a
    Sigmoid head), BCE against with_ordinals(y_true, n), summed over levels and valid items and divided by the
    number of valid items.  (On torch >= 2 the reference itself rejects padded slates: nn.BCELoss checks the -1
    targets before they are masked; the intended semantics are implemented.)"""
    _lib.require_cuda(y_pred, y_true)
    n = int(n)
    if y_pred.dim() != 3 or y_pred.shape[:2] != y_true.shape or y_pred.shape[2] != n:
        raise ValueError("ordinal: y_pred must be [batch_size, slate_length, n] and y_true [batch_size, slate_length]")
    if y_pred.shape[0] == 0:
        raise ValueError("empty batch")
    labels = y_true.detach().float().contiguous()
    B, S = labels.shape
    scratch = torch.empty(2 * B, dtype=torch.float32, device=labels.device)

    def launch(probs, loss, grad):
        rc = _lib.lib().arb_ordinal(_lib.ptr(probs), _lib.ptr(labels), B, S, n, float(padded_value_indicator),
                                    _lib.ptr(loss), _lib.ptr(grad), _lib.ptr(scratch), _lib.stream_ptr(probs.device))
        _lib.check(rc, "arb_ordinal")

    return _FusedLoss.apply(y_pred, launch)


__all__ = ["listNet", "listMLE", "approxNDCGLoss", "lambdaLoss", "neuralNDCG", "neuralNDCG_transposed", "rankNet",
           "rankNet_weightByGTDiff", "rankNet_weightByGTDiff_pow", "binary_listNet", "pointwise_rmse", "bce", "ordinal",
           "with_ordinals", "DEFAULT_EPS",
           "PADDED_Y_VALUE"]
