"""Oracle restatement of allRank's listwise losses (TEST INFRASTRUCTURE, see oracle/__init__.py).

Reference files (all under /root/reference/allrank/models/losses/):
  listNet.py:8-30        listMLE.py:7-38        approxNDCG.py:7-53
  lambdaLoss.py:7-114    neuralNDCG.py:10-70    loss_utils.py:8-67

All functions are differentiable eager PyTorch (autograd gives the reference
gradient w.r.t. y_pred); they run on whatever device/dtype the inputs have
(the reference hard-wires cuda:0 inside NeuralSort, model_utils.py:13-18 -- the
restatement follows the input tensor instead, which is the only deliberate
deviation and has no numerical effect).
"""
import math

import torch

from .metrics_ref import dcg as _dcg

PAD = -1          # allrank/data/dataset_loading.py:15
EPS = 1e-10       # allrank/models/losses/__init__.py:1  DEFAULT_EPS


# --------------------------------------------------------------------------- listNet
def listNet(y_pred, y_true, eps=EPS, padded_value_indicator=PAD):
    # listNet.py:17-30 : cross entropy between softmax(labels) and softmax(scores), pads -> -inf
    is_pad = y_true == padded_value_indicator
    s = y_pred.masked_fill(is_pad, float("-inf"))
    t = y_true.masked_fill(is_pad, float("-inf"))
    p_s = torch.softmax(s, dim=1) + eps
    p_t = torch.softmax(t, dim=1)
    return torch.mean(-torch.sum(p_t * torch.log(p_s), dim=1))


# --------------------------------------------------------------------------- listMLE
def listMLE(y_pred, y_true, eps=EPS, padded_value_indicator=PAD, perm=None, order=None):
    """listMLE.py:16-38.

    `perm` : the column shuffle (reference draws torch.randperm from the global CPU RNG, :17).
    `order`: optional [B,S] realised sort order of the shuffled labels (debug hook so a
             kernel can be fed the oracle's tie resolution).
    """
    n = y_pred.shape[-1]
    if perm is None:
        perm = torch.randperm(n)
    perm = perm.to(y_pred.device)
    s = y_pred[:, perm]
    t = y_true[:, perm]
    if order is None:
        t_sorted, order = t.sort(descending=True, dim=-1)
    else:
        t_sorted = t.gather(1, order)
    is_pad = t_sorted == padded_value_indicator
    z = s.gather(1, order).masked_fill(is_pad, float("-inf"))
    z = z - z.max(dim=1, keepdim=True).values
    tail = torch.cumsum(z.exp().flip(dims=[1]), dim=1).flip(dims=[1])   # sum_{j>=i} exp(z_j)
    per_item = (torch.log(tail + eps) - z).masked_fill(is_pad, 0.0)
    return torch.mean(torch.sum(per_item, dim=1))


def listMLE_realised_order(y_true, perm):
    """The (perm, order) pair the reference would realise for these labels on this host."""
    t = y_true[:, perm]
    return t.sort(descending=True, dim=-1).indices


# --------------------------------------------------------------------------- shared pieces
def _sorted_views(y_pred, y_true, pad):
    """approxNDCG.py:19-33 == lambdaLoss.py:25-39: pads -> -inf, sort by score, labels in score order."""
    is_pad = y_true == pad
    s = y_pred.masked_fill(is_pad, float("-inf"))
    t = y_true.masked_fill(is_pad, float("-inf"))
    s_sorted, order = s.sort(descending=True, dim=-1)
    t_ideal = t.sort(descending=True, dim=-1).values
    t_by_s = t.gather(1, order)
    return s_sorted, t_by_s, t_ideal


def _log2_positions(n, device):
    return torch.log2(1.0 + torch.arange(1, n + 1, device=device).float())[None, :]


# --------------------------------------------------------------------------- approxNDCG
def approxNDCGLoss(y_pred, y_true, eps=EPS, padded_value_indicator=PAD, alpha=1.0):
    # approxNDCG.py:17-53 (no @k truncation; maxDCG clamped at eps)
    s_sorted, t_by_s, t_ideal = _sorted_views(y_pred, y_true, padded_value_indicator)
    n = y_pred.shape[1]
    label_gap = t_by_s[:, :, None] - t_by_s[:, None, :]
    pair_ok = torch.isfinite(label_gap)
    pair_ok = pair_ok & ~torch.eye(n, dtype=torch.bool, device=y_pred.device)[None]
    t_by_s = t_by_s.clamp(min=0.0)
    t_ideal = t_ideal.clamp(min=0.0)
    D = _log2_positions(n, y_pred.device)
    max_dcg = torch.sum((torch.pow(2, t_ideal) - 1) / D, dim=-1).clamp(min=eps)
    G = (torch.pow(2, t_by_s) - 1) / max_dcg[:, None]
    gap = s_sorted[:, :, None] - s_sorted[:, None, :]
    gap = torch.where(pair_ok, gap, torch.zeros_like(gap))
    soft_rank = 1.0 + torch.sum(pair_ok.float() * torch.sigmoid(-alpha * gap).clamp(min=eps), dim=-1)
    return -torch.mean(torch.sum(G / torch.log2(1.0 + soft_rank), dim=-1))


# --------------------------------------------------------------------------- lambdaLoss
def _w_ndcgLoss1(G, D, mu, t):        # lambdaLoss.py:84-85
    return (G / D)[:, :, None]


def _w_ndcgLoss2(G, D, mu, t):        # lambdaLoss.py:88-94 (Toeplitz |1/D(|i-j|) - 1/D(|i-j|+1)|, diag 0)
    n = G.shape[1]
    pos = torch.arange(1, n + 1, device=G.device)
    lag = torch.abs(pos[:, None] - pos[None, :])
    toe = torch.abs(torch.pow(torch.abs(D[0, lag - 1]), -1.0) - torch.pow(torch.abs(D[0, lag]), -1.0))
    toe = toe * (1 - torch.eye(n, device=G.device))  # diagonal (lag 0 wraps to D[-1], quirk Q8) zeroed
    return toe[None] * torch.abs(G[:, :, None] - G[:, None, :])


def _w_lambdaRank(G, D, mu, t):       # lambdaLoss.py:97-98
    inv = torch.pow(D, -1.0)
    return torch.abs(inv[:, :, None] - inv[:, None, :]) * torch.abs(G[:, :, None] - G[:, None, :])


def _w_ndcgLoss2PP(G, D, mu, t):      # lambdaLoss.py:101-102
    return mu * _w_ndcgLoss2(G, D, mu, t) + _w_lambdaRank(G, D, mu, t)


def _w_rankNet(G, D, mu, t):          # lambdaLoss.py:105-106
    return 1.0


def _w_rankNetGTDiff(G, D, mu, t):    # lambdaLoss.py:109-110
    return torch.abs(t[:, :, None] - t[:, None, :])


def _w_rankNetGTDiffPowed(G, D, mu, t):  # lambdaLoss.py:113-114
    return torch.abs(torch.pow(t[:, :, None], 2) - torch.pow(t[:, None, :], 2))


WEIGHING_SCHEMES = {
    "ndcgLoss1_scheme": _w_ndcgLoss1,
    "ndcgLoss2_scheme": _w_ndcgLoss2,
    "lambdaRank_scheme": _w_lambdaRank,
    "ndcgLoss2PP_scheme": _w_ndcgLoss2PP,
    "rankNet_scheme": _w_rankNet,
    "rankNetWeightedByGTDiff_scheme": _w_rankNetGTDiff,
    "rankNetWeightedByGTDiffPowed_scheme": _w_rankNetGTDiffPowed,
}


def lambdaLoss(y_pred, y_true, eps=EPS, padded_value_indicator=PAD, weighing_scheme=None, k=None,
               sigma=1.0, mu=10.0, reduction="sum", reduction_log="binary"):
    # lambdaLoss.py:22-81
    s_sorted, t_by_s, t_ideal = _sorted_views(y_pred, y_true, padded_value_indicator)
    n = y_pred.shape[1]
    label_gap = t_by_s[:, :, None] - t_by_s[:, None, :]
    pair_ok = torch.isfinite(label_gap)
    if weighing_scheme != "ndcgLoss1_scheme":
        pair_ok = pair_ok & (label_gap > 0)
    top_k = torch.zeros((n, n), dtype=torch.bool, device=y_pred.device)
    top_k[:k, :k] = True

    t_by_s = t_by_s.clamp(min=0.0)
    t_ideal = t_ideal.clamp(min=0.0)
    D = _log2_positions(n, y_pred.device)
    max_dcg = torch.sum(((torch.pow(2, t_ideal) - 1) / D)[:, :k], dim=-1).clamp(min=eps)
    G = (torch.pow(2, t_by_s) - 1) / max_dcg[:, None]

    if weighing_scheme is None:
        w = 1.0
    else:
        w = WEIGHING_SCHEMES[weighing_scheme](G, D, mu, t_by_s)  # KeyError like the reference's globals()[...]

    gap = (s_sorted[:, :, None] - s_sorted[:, None, :]).clamp(min=-1e8, max=1e8)
    prob = (torch.sigmoid(sigma * gap).clamp(min=eps) ** w).clamp(min=eps)
    if reduction_log == "natural":
        terms = torch.log(prob)
    elif reduction_log == "binary":
        terms = torch.log2(prob)
    else:
        raise ValueError("Reduction logarithm base can be either natural or binary")
    picked = terms[pair_ok & top_k]
    if reduction == "sum":
        return -torch.sum(picked)
    if reduction == "mean":
        return -torch.mean(picked)
    raise ValueError("Reduction method can be either sum or mean")


# --------------------------------------------------------------------------- NeuralSort / Sinkhorn / neuralNDCG
def deterministic_neural_sort(s, tau, mask):
    """loss_utils.py:34-67.  s [B,S,1], mask [B,S] (True = padded) -> P_hat [B,S,S] (rows = ranks)."""
    dev = s.device
    n = s.shape[1]
    either = mask[:, :, None] | mask[:, None, :]
    both = mask[:, :, None] & mask[:, None, :]
    s_far = s.masked_fill(mask[:, :, None], -1e8)
    absdiff = torch.abs(s_far - s_far.permute(0, 2, 1)).masked_fill(either, 0.0)
    ones = torch.ones((n, 1), dtype=s.dtype, device=dev)   # (the reference hard-wires fp32 here; fp64 inputs are an oracle extension)
    row_tot = torch.matmul(absdiff, torch.matmul(ones, ones.t()))            # [b,i,j] = sum_k |s_i - s_k|
    n_valid = n - mask.sum(dim=1)                                            # [B]
    ranks = torch.arange(n, device=dev)[None, :]
    coef = (n_valid[:, None] + 1 - 2 * (ranks + 1)).to(s.dtype)
    coef = torch.where(ranks < n_valid[:, None], coef, torch.zeros_like(coef))  # zero beyond the valid count
    s_zero = s.masked_fill(mask[:, :, None], 0.0)
    lin = torch.matmul(s_zero, coef.unsqueeze(-2))                           # [b,i,j] = s_i * coef_j
    logits = (lin - row_tot).permute(0, 2, 1)                                # [b, rank j, item i]
    logits = logits.masked_fill(either, float("-inf")).masked_fill(both, 1.0)
    return torch.softmax(logits / tau, dim=-1)


def sinkhorn_scaling(mat, mask=None, tol=1e-6, max_iter=50):
    """loss_utils.py:8-31: alternate column / row normalisation, global convergence test."""
    if mask is not None:
        either = mask[:, None, :] | mask[:, :, None]
        both = mask[:, None, :] & mask[:, :, None]
        mat = mat.masked_fill(either, 0.0).masked_fill(both, 1.0)
    for _ in range(max_iter):
        mat = mat / mat.sum(dim=1, keepdim=True).clamp(min=EPS)
        mat = mat / mat.sum(dim=2, keepdim=True).clamp(min=EPS)
        if torch.max(torch.abs(mat.sum(dim=2) - 1.0)) < tol and torch.max(torch.abs(mat.sum(dim=1) - 1.0)) < tol:
            break
    if mask is not None:
        mat = mat.masked_fill(either, 0.0)
    return mat


def neuralNDCG(y_pred, y_true, padded_value_indicator=PAD, temperature=1.0, powered_relevancies=True, k=None,
               stochastic=False, n_samples=32, beta=0.1, log_scores=True, max_iter=50, tol=1e-6):
    """neuralNDCG.py:27-70 (deterministic variant; max_iter/tol are hard-coded 50/1e-6 there, :41-42)."""
    if stochastic:
        raise NotImplementedError("oracle restates the deterministic variant only")
    if k is None:
        k = y_true.shape[1]
    mask = y_true == padded_value_indicator
    P = deterministic_neural_sort(y_pred.unsqueeze(-1), tau=temperature, mask=mask)
    P = sinkhorn_scaling(P, mask, tol=tol, max_iter=max_iter)
    P = P.masked_fill(mask[:, :, None] | mask[:, None, :], 0.0)
    rel = y_true.masked_fill(mask, 0.0)
    gains = torch.pow(2.0, rel) - 1.0 if powered_relevancies else rel
    soft_sorted = torch.matmul(P, gains.unsqueeze(-1)).squeeze(-1)
    disc = (torch.tensor(1.0) / torch.log2(torch.arange(y_true.shape[-1], dtype=torch.float) + 2.0)).to(y_pred.device, y_pred.dtype)
    disc_gains = (soft_sorted * disc)[:, :k]
    if powered_relevancies:
        idcg = _dcg(y_true, y_true, ats=[k]).squeeze(1)
    else:
        idcg = _dcg(y_true, y_true, ats=[k], gain_function=lambda x: x).squeeze(1)
    val = disc_gains.sum(dim=-1) / (idcg + EPS)
    dead = idcg == 0.0
    val = val.masked_fill(dead, 0.0)
    if bool(dead.all()):
        return torch.tensor(0.0)
    return -1.0 * (val.sum() / (~dead).sum())


def neuralNDCG_transposed(y_pred, y_true, padded_value_indicator=PAD, temperature=1.0, powered_relevancies=True,
                          k=None, stochastic=False, n_samples=32, beta=0.1, log_scores=True, max_iter=50, tol=1e-6):
    """neuralNDCG.py:73-136: expected discounts P^T disc (truncated at k), times gains, over (idcg + eps)."""
    if stochastic:
        raise NotImplementedError("oracle restates the deterministic variant only")
    if k is None:
        k = y_true.shape[1]
    mask = y_true == padded_value_indicator
    P = deterministic_neural_sort(y_pred.unsqueeze(-1), tau=temperature, mask=mask)
    P = sinkhorn_scaling(P, mask, tol=tol, max_iter=max_iter)
    disc = (torch.tensor(1) / torch.log2(torch.arange(y_true.shape[-1], dtype=torch.float) + 2.0)).to(y_pred.device, y_pred.dtype)
    disc = disc.clone()
    disc[k:] = 0.0
    expected = torch.matmul(P.permute(0, 2, 1), disc[None, :, None]).squeeze(-1)
    gains = torch.pow(2.0, y_true) - 1 if powered_relevancies else y_true
    idcg = _dcg(y_true, y_true, ats=[k]).squeeze(1)
    val = (gains * expected).sum(dim=1) / (idcg + EPS)
    dead = idcg == 0.0
    val = val.masked_fill(dead, 0.0)
    if bool(dead.all()):
        return torch.tensor(0.0)
    return -1.0 * (val.sum() / (~dead).sum())


# --------------------------------------------------------------------------- SURVEY 8(f) rank-1 "next" losses
def rankNet(y_pred, y_true, padded_value_indicator=PAD, weight_by_diff=False, weight_by_diff_powed=False):
    # rankNet.py:31-79: BCE-with-logits (target 1) over ordered pairs of real items with t_i > t_j, mean over pairs
    is_pad = y_true == padded_value_indicator
    t = y_true.masked_fill(is_pad, float("-inf"))
    gap_t = t[:, :, None] - t[:, None, :]
    gap_s = y_pred[:, :, None] - y_pred[:, None, :]
    chosen = (gap_t > 0) & ~torch.isinf(gap_t)
    weight = None
    if weight_by_diff:
        weight = torch.abs(gap_t)[chosen]
    elif weight_by_diff_powed:
        weight = torch.abs(t[:, :, None] ** 2 - t[:, None, :] ** 2)[chosen]
    x = gap_s[chosen]
    return torch.nn.functional.binary_cross_entropy_with_logits(x, torch.ones_like(x), weight=weight)


def rankNet_weightByGTDiff(y_pred, y_true, padded_value_indicator=PAD):       # rankNet.py:9-17
    return rankNet(y_pred, y_true, padded_value_indicator, weight_by_diff=True)


def rankNet_weightByGTDiff_pow(y_pred, y_true, padded_value_indicator=PAD):   # rankNet.py:20-28
    return rankNet(y_pred, y_true, padded_value_indicator, weight_by_diff_powed=True)


def binary_listNet(y_pred, y_true, eps=EPS, padded_value_indicator=PAD):
    # binary_listNet.py:17-33: labels normalised by their sum (1 if the sum is 0), softmax of the scores
    is_pad = y_true == padded_value_indicator
    s = y_pred.masked_fill(is_pad, float("-inf"))
    t = y_true.masked_fill(is_pad, 0.0)
    norm = t.sum(dim=-1, keepdim=True)
    norm = torch.where(norm == 0.0, torch.ones_like(norm), norm)
    return torch.mean(-torch.sum((t / norm) * torch.log(torch.softmax(s, dim=1) + eps), dim=1))


def pointwise_rmse(y_pred, y_true, no_of_levels, padded_value_indicator=PAD):
    # pointwise.py:16-32
    is_pad = y_true == padded_value_indicator
    err = (y_true.masked_fill(is_pad, 0.0) - no_of_levels * y_pred.masked_fill(is_pad, 0.0)) ** 2
    return torch.mean(torch.sqrt(err.sum(dim=1) / (~is_pad).float().sum(dim=1)))


def bce(y_pred, y_true, padded_value_indicator=PAD):
    # bce.py:17-32 (intended semantics: padded items contribute 0; torch >= 2 rejects -1 targets in nn.BCELoss, so the
    # per-element loss is restated with the same -100 log clamp)
    is_pad = y_true == padded_value_indicator
    t = y_true.masked_fill(is_pad, 0.0)
    p = y_pred
    per = -(t * torch.log(p).clamp(min=-100.0) + (1 - t) * torch.log(1 - p).clamp(min=-100.0))
    per = per.masked_fill(is_pad, 0.0)
    non_empty = ((~is_pad).sum(dim=-1) > 0).float().sum()
    return per.sum() / non_empty


def with_ordinals(y, n, padded_value_indicator=PAD):
    # ordinal.py:8-22: level j (1-based) is 1 where y >= j; padded items keep the padding value
    levels = torch.arange(1, n + 1, dtype=torch.float32)
    spread = y.unsqueeze(2).repeat(1, 1, n)
    out = (spread >= levels).float()
    return torch.where(spread == padded_value_indicator, torch.full_like(out, float(padded_value_indicator)), out)


def ordinal(y_pred, y_true, n, padded_value_indicator=PAD):
    # ordinal.py:25-50 (intended semantics, as for bce above: padded items contribute 0): BCE of the [B,S,n] level
    # probabilities against with_ordinals(y_true), summed, divided by the number of valid items
    t = with_ordinals(y_true, n, padded_value_indicator)
    is_pad = t == padded_value_indicator
    t = t.masked_fill(is_pad, 0.0)
    per = -(t * torch.log(y_pred).clamp(min=-100.0) + (1 - t) * torch.log(1 - y_pred).clamp(min=-100.0))
    per = per.masked_fill(is_pad, 0.0)
    n_valid_items = ((~is_pad).sum(dim=2) > 0).float().sum()
    return per.sum() / n_valid_items


LOSSES = {
    "listNet": listNet,
    "listMLE": listMLE,
    "approxNDCGLoss": approxNDCGLoss,
    "lambdaLoss": lambdaLoss,
    "neuralNDCG": neuralNDCG,
    "neuralNDCG_transposed": neuralNDCG_transposed,
    "rankNet": rankNet,
    "rankNet_weightByGTDiff": rankNet_weightByGTDiff,
    "rankNet_weightByGTDiff_pow": rankNet_weightByGTDiff_pow,
    "ordinal": ordinal,
    "binary_listNet": binary_listNet,
    "pointwise_rmse": pointwise_rmse,
    "bce": bce,
}

LN2 = math.log(2.0)
