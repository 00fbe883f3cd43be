"""CUDA parity of the slate movers (csrc/slates.cu through the C ABI): device slate assembly against the reference's
LibSVMDataset + FixLength outputs (bit-exact for padded slates; rules + distribution for sampled ones), the
DataLoader-shaped iterator, the rank_slates gather, and the epoch-level metric accumulation."""
import os

import numpy as np
import pytest
import torch

from oracle import slates_ref

pytestmark = pytest.mark.gpu

CORPUS = os.path.join(os.path.dirname(__file__), "golden", "slates_corpus.txt")


@pytest.fixture(scope="module")
def store():
    from allrank_b200.data import SlateStore
    return SlateStore.from_svm_file(CORPUS, device="cuda")


def host_groups(store):
    off = store.offsets_host
    X, y = store.docs_x.cpu().numpy(), store.docs_y.cpu().numpy()
    return [(X[off[i]:off[i + 1]], y[off[i]:off[i + 1]]) for i in range(len(off) - 1)]


def test_store_matches_reference_dataset_shape(store, golden):
    g = golden("slates")
    assert store.shape == [int(v) for v in g["shape"]]
    assert np.array_equal(np.diff(store.offsets_host), g["lengths"])
    assert len(store) == len(g["lengths"])


def test_padded_slates_are_bit_identical_to_the_reference(store, golden):
    g = golden("slates")
    S = int(g["slate_length"])
    Q = len(store)
    x, y, idx = store.assemble(torch.arange(Q), S, seed=1)
    assert x.dtype == torch.float32 and y.dtype == torch.float32 and idx.dtype == torch.int64
    n_padded = 0
    for qi in range(Q):
        if g["lengths"][qi] < S:
            n_padded += 1
            assert np.array_equal(x[qi].cpu().numpy(), g[f"q{qi}_x"]), qi
            assert np.array_equal(y[qi].cpu().numpy(), g[f"q{qi}_y"]), qi
            assert np.array_equal(idx[qi].cpu().numpy(), g[f"q{qi}_idx"]), qi
    assert n_padded >= 4
    # validation transform: everything padded to the longest query (dataset_loading.py:185-192)
    longest = store.longest_query_length
    x, y, idx = store.assemble(torch.arange(Q), longest, seed=2)
    for qi in range(Q):
        if f"v{qi}_x" in g.files:
            assert np.array_equal(x[qi].cpu().numpy(), g[f"v{qi}_x"])
            assert np.array_equal(y[qi].cpu().numpy(), g[f"v{qi}_y"])
            assert np.array_equal(idx[qi].cpu().numpy(), g[f"v{qi}_idx"])


def test_sampled_slates_obey_the_reference_rules(store):
    groups = host_groups(store)
    S = 20
    Q = len(store)
    seen_single, seen_orders = 0, set()
    for seed in range(40):
        x, y, idx = store.assemble(torch.arange(Q), S, seed=seed)
        x, y, idx = x.cpu().numpy(), y.cpu().numpy(), idx.cpu().numpy()
        for qi, (xs, ys) in enumerate(groups):
            if len(ys) < S:
                continue
            assert slates_ref.sample_is_admissible(ys, idx[qi], S), (seed, qi)
            assert np.array_equal(x[qi], xs[idx[qi]]) and np.array_equal(y[qi], ys[idx[qi]])
            if ys.sum() == 1:
                seen_single += 1
                assert ys[idx[qi]].sum() == 1                     # the single relevant item is always kept (:66-68)
            seen_orders.add((qi, tuple(idx[qi][:3])))
    assert seen_single >= 40
    assert len(seen_orders) > 100                                   # different seeds give different samples


def test_same_seed_same_sample_and_unknown_query_is_all_padding(store):
    q = torch.tensor([4, 9, 4])
    a = store.assemble(q, 20, seed=5)
    b = store.assemble(q, 20, seed=5)
    assert all(torch.equal(u, v) for u, v in zip(a, b))
    assert not torch.equal(a[2][0], a[2][2])                        # same query twice in a batch: independent samples
    x, y, idx = store.assemble(torch.tensor([10 ** 6], device="cuda"), 8, seed=0)   # device tensor: kernel-side check
    assert (x == 0).all() and (y == -1).all() and (idx == -1).all()
    with pytest.raises(IndexError):
        store.assemble(torch.tensor([len(store)]), 8)


def test_sampling_is_uniform_over_items():
    """A 50-item query with no relevant items sampled to 10: every item is picked with probability 1/5, every
    position of the slate is uniform over items (np.random.choice(n, S, replace=False) semantics)."""
    from allrank_b200.data import SlateStore
    n, S, trials = 50, 10, 4096
    X = np.arange(n, dtype=np.float32)[:, None] * np.ones((1, 4), dtype=np.float32)
    st = SlateStore(X, np.zeros(n, dtype=np.float32), np.zeros(n, dtype=np.int64), device="cuda")
    _, _, idx = st.assemble(torch.zeros(trials, dtype=torch.int64), S, seed=123)
    idx = idx.cpu().numpy()
    counts = np.bincount(idx.ravel(), minlength=n)
    expected = trials * S / n
    assert np.abs(counts - expected).max() < 6 * np.sqrt(expected)            # ~6 sigma (binomial)
    first = np.bincount(idx[:, 0], minlength=n)
    assert np.abs(first - trials / n).max() < 6 * np.sqrt(trials / n)


def test_single_relevant_item_distribution():
    """Query of 30 with one label-1 item, slate 10: kept in every slate; when the plain sample missed it, it sits in
    the last position (dataset_loading.py:67)."""
    from allrank_b200.data import SlateStore
    n, S, trials = 30, 10, 2048
    y = np.zeros(n, dtype=np.float32)
    y[17] = 1.0
    st = SlateStore(np.random.RandomState(0).randn(n, 8).astype(np.float32), y, np.zeros(n, dtype=np.int64), device="cuda")
    _, ys, idx = st.assemble(torch.zeros(trials, dtype=torch.int64), S, seed=9)
    idx = idx.cpu().numpy()
    assert (ys.sum(dim=1) == 1).all()
    last = (idx[:, -1] == 17).mean()
    # P(sample already holds it) = 1/3, and then it is at a uniform position; else it is appended last: 2/3 + 1/30
    assert abs(last - (2 / 3 + 1 / 30)) < 0.05


def test_label_sum_above_one_is_redrawn_not_patched():
    from allrank_b200.data import SlateStore
    n, S, trials = 40, 5, 1024
    y = np.zeros(n, dtype=np.float32)
    y[3] = 2.0                                                      # sum = 2 -> the redraw branch (:69-70)
    st = SlateStore(np.zeros((n, 4), dtype=np.float32), y, np.zeros(n, dtype=np.int64), device="cuda")
    _, ys, idx = st.assemble(torch.zeros(trials, dtype=torch.int64), S, seed=4)
    assert (ys.sum(dim=1) == 2).all()
    pos = (idx.cpu().numpy() == 3).argmax(axis=1)
    assert np.abs(np.bincount(pos, minlength=S) - trials / S).max() < 6 * np.sqrt(trials / S)   # uniform position


def test_loader_iterates_every_query_once_like_a_dataloader(store):
    from allrank_b200.data import DeviceSlateLoader, create_data_loaders
    store.slate_length = 20
    torch.manual_seed(3)
    dl = DeviceSlateLoader(store, batch_size=5, shuffle=True)
    assert len(dl) == 3 and dl.dataset is store
    batches = list(dl)
    assert [b[0].shape[0] for b in batches] == [5, 5, 2]
    assert all(b[0].shape[1:] == (20, store.n_features) and b[1].shape[1] == 20 and b[2].shape[1] == 20 for b in batches)
    lengths = sorted(int((b[1][i] != -1).sum()) for b in batches for i in range(b[1].shape[0]))
    assert lengths == sorted(min(int(v), 20) for v in np.diff(store.offsets_host))
    torch.manual_seed(3)
    again = list(DeviceSlateLoader(store, batch_size=5, shuffle=True))
    assert all(torch.equal(u[2], v[2]) for u, v in zip(batches, again))      # repeatable under torch.manual_seed
    train_dl, val_dl = create_data_loaders(store, store, num_workers=1, batch_size=4)
    assert train_dl.shuffle and not val_dl.shuffle and len(val_dl) == 3


def test_rank_batch_matches_reference_gather(store):
    from allrank_b200 import inference
    S = store.longest_query_length
    x, y, _ = store.assemble(torch.arange(len(store)), S, seed=0)
    scores = torch.randn(y.shape, generator=torch.Generator().manual_seed(8)).cuda()
    order = __import__("allrank_b200.metrics", fromlist=["ranking"]).ranking(scores, y)
    rx, ry = inference.reorder_slates(x, y, order)
    ex, ey, eorder = slates_ref.rank_batch(scores.cpu(), x.cpu(), y.cpu())
    assert torch.equal(ry.cpu(), ey)
    valid = (ey != -1)
    assert torch.equal(order.cpu().long()[valid], eorder[valid])              # padded tail: order unspecified
    assert torch.equal(rx.cpu()[valid], ex[valid])
    assert (rx.cpu()[~valid] == 0).all()


def test_rank_slates_and_epoch_metrics_with_a_model(store):
    from allrank_b200 import inference, metrics, training
    from allrank_b200.data import DeviceSlateLoader
    from allrank_b200.model import make_model
    torch.manual_seed(1)
    model = make_model(fc_model={"sizes": [32], "input_norm": False, "activation": None, "dropout": 0.0},
                       transformer={"N": 1, "d_ff": 64, "h": 2, "positional_encoding": None, "dropout": 0.0},
                       post_model={"d_output": 1, "output_activation": None}, n_features=store.n_features).cuda().eval()
    dl = DeviceSlateLoader(store, batch_size=5, slate_length=store.longest_query_length, shuffle=False)

    class Cfg:
        class data:
            batch_size = 5
    ranked = inference.rank_slates({"vali": dl}, model, Cfg)
    rx, ry = ranked["vali"]
    assert rx.device.type == "cpu" and tuple(rx.shape) == (len(store), store.longest_query_length, store.n_features)
    # ranked labels are exactly the sequence the metrics see: DCG of the ranked labels in place == dcg(scores, y)
    ats = [1, 5, 10]
    got = training.compute_metrics({"ndcg": ats, "mrr": ats}, model, dl, torch.device("cuda"))
    assert sorted(got) == sorted([f"ndcg_{a}" for a in ats] + [f"mrr_{a}" for a in ats])
    rows_ndcg, rows_mrr, in_place = [], [], []
    with torch.no_grad():
        for xb, yb, ib in dl:
            s = model.score(xb, yb == -1, ib)
            rows_ndcg.append(metrics.ndcg(s, yb, ats=ats))
            rows_mrr.append(metrics.mrr(s, yb, ats=ats))
    ref_ndcg = torch.mean(torch.cat(rows_ndcg), dim=0).cpu().numpy()       # train_utils.py:37-43
    ref_mrr = torch.mean(torch.cat(rows_mrr), dim=0).cpu().numpy()
    for a, v in zip(ats, ref_ndcg):
        assert got[f"ndcg_{a}"] == pytest.approx(v, rel=1e-6)
    for a, v in zip(ats, ref_mrr):
        assert got[f"mrr_{a}"] == pytest.approx(v, rel=1e-6)
    one = training.metric_on_epoch(lambda p, t: metrics.ndcg(p, t, ats=ats), model, dl, torch.device("cuda"))
    assert np.allclose(one, ref_ndcg, rtol=1e-6)
    # ranking a slate by descending position score reproduces its order: ndcg of ranked labels with decreasing scores
    pos_scores = torch.arange(ry.shape[1], 0, -1, dtype=torch.float32).expand_as(ry).cuda()
    again = metrics.ndcg(pos_scores, ry.cuda(), ats=ats).mean(dim=0).cpu().numpy()
    assert np.allclose(again, ref_ndcg, rtol=1e-6)


def test_rank_slates_accepts_a_map_style_dataset_like_the_reference(store):
    """What rank_and_click.py:86 passes after patch_allrank(): the reference's LibSVMDataset objects -- map-style
    Datasets of UN-batched (x[S,F], y[S], indices[S]) samples, which inference_utils.__create_data_loader (:33-34)
    batches with config.data.batch_size."""
    from allrank_b200 import inference
    from allrank_b200.data import DeviceSlateLoader
    from allrank_b200.model import make_model
    torch.manual_seed(2)
    model = make_model(fc_model={"sizes": [32], "input_norm": False, "activation": None, "dropout": 0.0},
                       transformer=None, post_model={"d_output": 1, "output_activation": None},
                       n_features=store.n_features).cuda().eval()
    S = store.longest_query_length
    batches = list(DeviceSlateLoader(store, batch_size=4, slate_length=S, shuffle=False))

    class PlainDataset(torch.utils.data.Dataset):      # same sample protocol as LibSVMDataset.__getitem__
        def __init__(self):
            self.x = torch.cat([b[0] for b in batches]).cpu()
            self.y = torch.cat([b[1] for b in batches]).cpu()
            self.i = torch.cat([b[2] for b in batches]).cpu()

        def __len__(self):
            return self.x.shape[0]

        def __getitem__(self, k):
            return self.x[k], self.y[k], self.i[k]

    class Cfg:
        class data:
            batch_size = 3
            num_workers = 0
    ds = PlainDataset()
    ranked = inference.rank_slates({"vali": ds}, model, Cfg)["vali"]
    direct = inference.rank_dataloader(batches, model)
    assert tuple(ranked[0].shape) == (len(ds), S, store.n_features)
    valid = direct[1] != -1
    assert torch.equal(ranked[1][valid], direct[1][valid]) and torch.equal(ranked[0][valid], direct[0][valid])


def test_config1_training_loop_matches_the_oracle(store):
    """BASELINE configs[0] in miniature (FC-only scorer + listNet + Adam + StepLR, the reference's CPU-runnable case,
    SURVEY 8c "end-to-end anchor"): the device pipeline (SlateStore loader -> fused scorer -> fused listNet -> Adam ->
    one-pass epoch metrics) follows the eager oracle (FixLength padding -> nn.Module -> listNet -> ndcg) step by step."""
    from oracle import losses_ref, metrics_ref
    from oracle.scorer_ref import make_ref_model
    from allrank_b200 import losses, training
    from allrank_b200.data import DeviceSlateLoader
    from allrank_b200.model import make_model
    F, S = store.n_features, store.longest_query_length        # every query padded (validation-style): no sampling
    torch.manual_seed(42)
    ref = make_ref_model(F, [64], 0, 0, 0).train()
    mine = make_model(fc_model={"sizes": [64], "input_norm": False, "activation": None, "dropout": 0.0},
                      transformer=None, post_model={"d_output": 1, "output_activation": None}, n_features=F)
    mine.load_state_dict(ref.state_dict())
    mine = mine.cuda().train()
    groups = host_groups(store)
    padded = [slates_ref.fix_length(xs, ys, S) if len(ys) < S else (xs, ys, np.arange(S)) for xs, ys in groups]
    X = torch.tensor(np.stack([p[0] for p in padded]), dtype=torch.float32)
    Y = torch.tensor(np.stack([p[1] for p in padded]), dtype=torch.float32)
    I = torch.tensor(np.stack([p[2] for p in padded]), dtype=torch.long)
    bs = 4
    dl = DeviceSlateLoader(store, batch_size=bs, slate_length=S, shuffle=False)
    # the longest query is "sampled" (a permutation) on both sides: take the device loader's order for it
    for bi, (xb, yb, ib) in enumerate(dl):
        for r in range(xb.shape[0]):
            q = bi * bs + r
            if len(groups[q][1]) == S:
                X[q], Y[q], I[q] = xb[r].cpu(), yb[r].cpu(), ib[r].cpu()
            else:
                assert torch.equal(xb[r].cpu(), X[q]) and torch.equal(yb[r].cpu(), Y[q]) and torch.equal(ib[r].cpu(), I[q])
    o_ref = torch.optim.Adam(ref.parameters(), lr=1e-3)
    o_mine = torch.optim.Adam(mine.parameters(), lr=1e-3)
    s_ref = torch.optim.lr_scheduler.StepLR(o_ref, step_size=3, gamma=0.5)
    s_mine = torch.optim.lr_scheduler.StepLR(o_mine, step_size=3, gamma=0.5)
    for epoch in range(4):
        mine.train(); ref.train()
        for bi in range(0, len(groups), bs):
            xb, yb, ib = X[bi:bi + bs], Y[bi:bi + bs], I[bi:bi + bs]
            l_ref = losses_ref.listNet(ref(xb, yb == -1, ib), yb)
            l_ref.backward(); o_ref.step(); o_ref.zero_grad()
            l_mine = losses.listNet(mine(xb.cuda(), (yb == -1).cuda(), ib.cuda()), yb.cuda())
            l_mine.backward(); o_mine.step(); o_mine.zero_grad()
            assert l_mine.item() == pytest.approx(l_ref.item(), rel=2e-3, abs=2e-3), (epoch, bi)
        s_ref.step(); s_mine.step()
    mine.eval(); ref.eval()
    with torch.no_grad():
        ref_ndcg = torch.mean(torch.cat([metrics_ref.ndcg(ref.score(X[b:b + bs], Y[b:b + bs] == -1, I[b:b + bs]),
                                                          Y[b:b + bs], ats=[5]) for b in range(0, len(groups), bs)]), dim=0)
        batches = [(X[b:b + bs].cuda(), Y[b:b + bs].cuda(), I[b:b + bs].cuda()) for b in range(0, len(groups), bs)]
        got = training.compute_metrics({"ndcg": [5]}, mine, batches, torch.device("cuda"))
    assert got["ndcg_5"] == pytest.approx(ref_ndcg.item(), abs=2e-2)     # small slates: one TF32-induced swap moves it
