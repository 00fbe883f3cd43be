"""Oracle restatement of the slate movers either side of the scorer (TEST INFRASTRUCTURE, see oracle/__init__.py).

Reference: /root/reference/allrank/data/dataset_loading.py     LibSVMDataset grouping :104-111, FixLength :32-93
           /root/reference/allrank/inference/inference_utils.py __rank_slates :37-60
           /root/reference/allrank/training/train_utils.py      metric_on_epoch / compute_metrics :37-56

`fix_length` draws from numpy's global stream with the same calls in the same order as the reference, so with the
same np.random.seed it reproduces the reference's slates exactly -- pinned by tests/golden/slates.npz
(oracle/make_golden.py).  The CUDA path uses a different random stream: sampled slates are compared through the
rules below (properties + distribution), padded slates bit for bit.
"""
import numpy as np
import torch

PAD_Y = -1
PAD_INDEX = -1


def group_offsets(query_ids):
    # dataset_loading.py:106-111: groups in order of first appearance, each `count` consecutive rows
    _, first, counts = np.unique(np.asarray(query_ids), return_index=True, return_counts=True)
    return np.concatenate([[0], np.cumsum(counts[np.argsort(first)])]).astype(np.int64)


def _pad(x, y, n, slate_length):
    # dataset_loading.py:76-93
    fx = np.pad(x, ((0, slate_length - n), (0, 0)), "constant")
    fy = np.pad(y, (0, slate_length - n), "constant", constant_values=PAD_Y)
    idx = np.pad(np.arange(0, n), (0, slate_length - n), "constant", constant_values=PAD_INDEX)
    return fx, fy, idx


def _sample(x, y, n, slate_length):
    # dataset_loading.py:55-74
    idx = np.random.choice(n, slate_length, replace=False)
    fy = y[idx]
    if fy.sum() == 0:
        if y.sum() == 1:
            idx = np.concatenate([np.random.choice(idx, slate_length - 1, replace=False), [np.argmax(y)]])
            fy = y[idx]
        elif y.sum() > 0:
            return _sample(x, y, n, slate_length)
    return x[idx], fy, idx


def fix_length(x, y, slate_length):
    # dataset_loading.py:41-53: shorter -> pad, otherwise (including equal length) -> sample
    n = len(y)
    return _pad(x, y, n, slate_length) if n < slate_length else _sample(x, y, n, slate_length)


def sample_is_admissible(y_query, idx, slate_length):
    """The rules a sampled slate obeys whatever the random stream (dataset_loading.py:55-74)."""
    idx = np.asarray(idx)
    n = len(y_query)
    if len(idx) != slate_length or idx.min() < 0 or idx.max() >= n or len(np.unique(idx)) != slate_length:
        return False
    if y_query.sum() > 0 and y_query[idx].sum() == 0:
        return False                      # a relevant item exists but the sample has none: must have been redrawn
    return True


def rank_batch(scores, X, y_true):
    # inference_utils.py:49-57
    scores = scores.clone()
    scores[y_true == PAD_Y] = float("-inf")
    _, order = scores.sort(descending=True, dim=-1)
    order_x = torch.unsqueeze(order, -1).repeat_interleave(X.shape[-1], -1)
    return torch.gather(X, dim=1, index=order_x), torch.gather(y_true, dim=1, index=order), order


def metric_on_epoch(metric, score_fn, batches):
    # train_utils.py:37-46: mean over all slates of the per-slate metric rows
    rows = [metric(score_fn(xb, yb == PAD_Y, idx), yb) for xb, yb, idx in batches]
    return torch.mean(torch.cat(rows), dim=0).numpy()
