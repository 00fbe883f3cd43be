"""world_size-2 gloo test of the data-parallel host logic (allrank_b200/ddp.py) on CPU tensors:
rank replicas start identical after the parameter broadcast, and the single flat-bucket all-reduce reproduces
the single-process gradient of the concatenated batch for a mean-over-batch loss (average=True) and for a
sum-reduced loss (average=False).   The kernels themselves need a GPU; here the "model" is a flat parameter
vector with a hand-written listwise gradient so that only the collective logic is under test."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp



def _by_value(obj):
    """Tensors cross the result queue as numpy arrays: a torch tensor is passed by file descriptor, which the parent
    can no longer fetch once the worker has exited (a race the tests lost on a loaded machine)."""
    if isinstance(obj, torch.Tensor):
        return ("__tensor__", obj.detach().cpu().numpy())
    if isinstance(obj, dict):
        return {k: _by_value(v) for k, v in obj.items()}
    if isinstance(obj, (tuple, list)):
        return type(obj)(_by_value(v) for v in obj)
    return obj


def _from_value(obj):
    if isinstance(obj, tuple) and len(obj) == 2 and isinstance(obj[0], str) and obj[0] == "__tensor__":
        return torch.from_numpy(obj[1].copy())
    if isinstance(obj, dict):
        return {k: _from_value(v) for k, v in obj.items()}
    if isinstance(obj, (tuple, list)):
        return type(obj)(_from_value(v) for v in obj)
    return obj


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class FlatToy:
    """Minimal stand-in with the two attributes FlatDDP uses (flat_parameters / flat_gradients)."""

    def __init__(self, n, seed):
        g = torch.Generator().manual_seed(seed)
        self.flat_parameters = torch.randn(n, generator=g)
        self.flat_gradients = torch.zeros(n)

    def loss_and_grad(self, x, y, reduction):
        # listNet-like: softmax cross entropy of scores x @ w against softmax(labels), per slate
        w = self.flat_parameters.clone().requires_grad_(True)
        scores = x @ w
        per_slate = -(torch.softmax(y, 1) * torch.log_softmax(scores, 1)).sum(1)
        loss = per_slate.mean() if reduction == "mean" else per_slate.sum()
        loss.backward()
        self.flat_gradients.copy_(w.grad)
        return loss.item()


def _worker(rank, world, port, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from allrank_b200.ddp import FlatDDP
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(123)
    x_all = torch.randn(8, 6, 5, generator=g)          # 8 slates x 6 items x 5 features
    y_all = torch.randint(0, 5, (8, 6), generator=g).float()
    shard = slice(rank * 4, rank * 4 + 4)
    results = {}
    for reduction, average in (("mean", True), ("sum", False)):
        model = FlatToy(5, seed=100 + rank)             # ranks start DIFFERENT on purpose
        ddp = FlatDDP(model, average=average)
        ddp.sync_parameters()
        params_after_sync = model.flat_parameters.clone()
        model.loss_and_grad(x_all[shard], y_all[shard], reduction)
        scale = ddp.reduce_gradients()
        assert scale == 1.0
        results[reduction] = (params_after_sync, model.flat_gradients.clone())
        # folded averaging: pure sum reduce, scale handed to the optimiser
        model.loss_and_grad(x_all[shard], y_all[shard], reduction)
        scale = ddp.reduce_gradients(fold_average_into_optimizer=True)
        results[reduction + "_folded"] = (scale, model.flat_gradients.clone())
    if rank == 0:
        ref = FlatToy(5, seed=100)
        single = {}
        for reduction in ("mean", "sum"):
            ref.loss_and_grad(x_all, y_all, reduction)
            single[reduction] = ref.flat_gradients.clone()
        out_q.put(_by_value((results, single, ref.flat_parameters.clone())))
    else:
        out_q.put(_by_value(({k: v[0] if not isinstance(v[0], float) else None for k, v in results.items()}, None, None)))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_bucket_allreduce_matches_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [_from_value(q.get(timeout=120)) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rank0 = next(g for g in got if g[1] is not None)
    other = next(g for g in got if g[1] is None)
    results, single, ref_params = rank0
    # broadcast: both ranks hold rank 0's parameters
    assert torch.equal(results["mean"][0], ref_params)
    assert torch.equal(other[0]["mean"], ref_params)
    # mean-over-batch loss: averaged shard gradients == gradient of the full batch
    assert torch.allclose(results["mean"][1], single["mean"], rtol=1e-5, atol=1e-7)
    # sum-reduced loss (lambdaLoss default): summed shard gradients == gradient of the full batch
    assert torch.allclose(results["sum"][1], single["sum"], rtol=1e-5, atol=1e-7)
    # folded averaging returns the optimiser scale instead of touching the buffer
    scale, summed = results["mean_folded"]
    assert scale == 0.5
    assert torch.allclose(summed * scale, single["mean"], rtol=1e-5, atol=1e-7)
    scale, summed = results["sum_folded"]
    assert scale == 1.0


def _metric_worker(rank, world, port, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from allrank_b200.training import reduce_epoch_sums
    g = torch.Generator().manual_seed(5)
    rows = torch.rand(10, 3, generator=g).double()          # per-slate metric rows of the WHOLE loader
    mine = rows[:7] if rank == 0 else rows[7:]                # uneven shards: 7 + 3 slates
    totals, count = reduce_epoch_sums({"ndcg": mine.sum(0), "mrr": 2 * mine.sum(0)}, mine.shape[0])
    out_q.put(_by_value((rank, totals["ndcg"] / count, totals["mrr"] / count, count, rows.mean(0))))
    dist.destroy_process_group()


def test_epoch_metric_sums_reduce_to_the_whole_loader_mean():
    """Each rank scores its shard; one all-reduce of (sums, count) gives every rank the mean over all slates --
    what train_utils.metric_on_epoch's torch.mean(torch.cat(...)) returns in a single process (train_utils.py:37-43)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_metric_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [_from_value(q.get(timeout=120)) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ndcg, mrr, count, expect in got:
        assert count == 10
        assert torch.allclose(ndcg, expect) and torch.allclose(mrr, 2 * expect)


def test_epoch_metric_reduce_is_a_no_op_without_a_process_group():
    from allrank_b200.training import reduce_epoch_sums
    t = {"ndcg": torch.tensor([1.0, 2.0])}
    out, n = reduce_epoch_sums(t, 4)
    assert out is t and n == 4


def _weighted_worker(rank, world, port, out_q):
    """neuralNDCG-style mean over a DATA-DEPENDENT subset of slates (those with a relevant item), uneven shards."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from allrank_b200.ddp import FlatDDP, loss_weight
    g = torch.Generator().manual_seed(77)
    x_all = torch.randn(9, 6, 5, generator=g)
    y_all = torch.randint(0, 3, (9, 6), generator=g).float()
    y_all[1] = 0.0                                   # slates without any relevant item: excluded from the mean
    y_all[2, :3] = 0.0; y_all[2, 3:] = -1.0          # ... also when the rest is padding
    y_all[7] = 0.0
    shard = slice(0, 5) if rank == 0 else slice(5, 9)            # 5 + 4 slates, 3 + 3 of them counted

    def grad_of(model, x, y):
        w = model.flat_parameters.clone().requires_grad_(True)
        keep = ((y > 0) & (y != -1)).any(1)
        per_slate = -(torch.softmax(y, 1) * torch.log_softmax(x @ w, 1)).sum(1)
        loss = per_slate[keep].sum() / keep.sum()                # mean over the counted slates only
        loss.backward()
        model.flat_gradients.copy_(w.grad)

    model = FlatToy(5, seed=100 + rank)
    ddp = FlatDDP(model)
    ddp.sync_parameters()
    grad_of(model, x_all[shard], y_all[shard])
    w = loss_weight("neuralNDCG", y_all[shard])
    ddp.reduce_gradients(local_weight=w)
    reduced = model.flat_gradients.clone()
    ref = FlatToy(5, seed=100)
    grad_of(ref, x_all, y_all)
    out_q.put(_by_value((rank, float(w), reduced, ref.flat_gradients.clone(), float(loss_weight("listNet", y_all[shard])))))
    dist.barrier()
    dist.destroy_process_group()


def test_numerator_and_count_reduction_for_means_over_a_data_dependent_subset():
    """neuralNDCG needs the GLOBAL count of idcg != 0 slates (neuralNDCG.py:62-69): all-reduce the
    numerator and the count, not the per-rank means."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_weighted_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [_from_value(q.get(timeout=120)) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, w, reduced, single, wb in got:
        assert w == 3.0 and wb == (5.0 if rank == 0 else 4.0)
        assert torch.allclose(reduced, single, rtol=1e-5, atol=1e-7)
