"""Epoch-level evaluation -- the counterpart of allrank/training/train_utils.py:32-56
(`metric_on_batch`, `metric_on_epoch`, `compute_metrics`).

The reference runs one full scoring pass over the loader PER METRIC NAME and concatenates every batch's
[B, len(ats)] result before averaging (train_utils.py:37-43, :49-53).  `compute_metrics` here scores each batch once,
evaluates all requested metrics from those scores in one metrics-kernel launch per metric, and keeps only running sums
on the device; the returned dict has the reference's keys ("ndcg_5", ...) and numpy float32 values.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import metrics as metrics_module

PADDED_Y_VALUE = -1


def reduce_epoch_sums(totals, count, group=None):
    """Data-parallel evaluation (one process per GPU, each scoring its own shard of the loader): sum the per-metric
    running sums and the slate count over all ranks in ONE small all-reduce, so that every rank
    returns the mean over the whole loader.  No-op without an initialised process group / with a single rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return totals, count
    names = list(totals)
    ref = totals[names[0]]
    flat = torch.cat([totals[n].reshape(-1).double() for n in names] +
                     [torch.tensor([float(count)], dtype=torch.float64, device=ref.device)])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    out, at = {}, 0
    for n in names:
        k = totals[n].numel()
        out[n] = flat[at:at + k].reshape(totals[n].shape)
        at += k
    return out, int(round(flat[at].item()))


def metric_on_batch(metric, model, xb, yb, indices):
    """train_utils.py:32-34."""
    mask = (yb == PADDED_Y_VALUE)
    return metric(model.score(xb, mask, indices), yb)


def metric_on_epoch(metric, model, dl, dev):
    """train_utils.py:37-46: mean over all slates of the loader of metric(...)[B, n_ats]; running sums, no torch.cat."""
    total, count = None, 0
    for xb, yb, indices in dl:
        vals = metric_on_batch(metric, model, xb.to(device=dev), yb.to(device=dev), indices.to(device=dev))
        part = vals.double().sum(dim=0)
        total = part if total is None else total + part
        count += vals.shape[0]
    sums, count = reduce_epoch_sums({"m": total}, count)
    return (sums["m"] / count).float().cpu().numpy()


def compute_metrics(metrics, model, dl, dev):
    """train_utils.py:49-56 with one scoring pass per batch shared by every metric name."""
    names = list(metrics.items())
    totals = {name: None for name, _ in names}
    count = 0
    for xb, yb, indices in dl:
        xb, yb, indices = xb.to(device=dev), yb.to(device=dev), indices.to(device=dev)
        with torch.no_grad():        # metrics never back-propagate; dropout still follows model.training
            scores = model.score(xb, yb == PADDED_Y_VALUE, indices)
        for name, ats in names:
            vals = getattr(metrics_module, name)(scores, yb, ats=ats)
            part = vals.double().sum(dim=0)
            totals[name] = part if totals[name] is None else totals[name] + part
        count += yb.shape[0]
    totals, count = reduce_epoch_sums(totals, count)
    out = {}
    for name, ats in names:
        values = (totals[name] / count).float().cpu().numpy()
        out.update({"{metric_name}_{at}".format(metric_name=name, at=at): v for at, v in zip(ats, values)})
    return out


__all__ = ["metric_on_batch", "metric_on_epoch", "compute_metrics", "reduce_epoch_sums"]
