"""Small end-to-end pass over every kernel for `compute-sanitizer --tool memcheck` (not collected by pytest):
    compute-sanitizer --tool memcheck python tests/sanitize_smoke.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from allrank_b200 import losses, metrics  # noqa: E402
from allrank_b200.model import make_model  # noqa: E402
from allrank_b200.optim import FlatAdam  # noqa: E402
from allrank_b200.synth import make_slates  # noqa: E402


def main():
    torch.manual_seed(0)
    for (F, d, N, h, dff, B, S, p) in [(136, 64, 1, 2, 128, 3, 37, 0.0), (20, 32, 2, 2, 64, 2, 130, 0.2),
                                       (136, 128, 1, 4, 256, 2, 300, 0.0)]:
        model = make_model(fc_model={"sizes": [d], "input_norm": False, "activation": None, "dropout": p},
                           transformer={"N": N, "d_ff": dff, "h": h, "positional_encoding": None, "dropout": p},
                           post_model={"d_output": 1, "output_activation": "Tanh"}, n_features=F).cuda().train()
        opt = FlatAdam(model)
        x, y, _ = make_slates(B, S, F, seed=1, mean_len=0.6 * S, std_len=0.3 * S)
        x, y = x.cuda(), y.cuda()
        for fn, kw in [(losses.listNet, {}), (losses.listMLE, {}), (losses.approxNDCGLoss, {}),
                       (losses.lambdaLoss, {"weighing_scheme": "ndcgLoss2PP_scheme"}),
                       (losses.lambdaLoss, {"weighing_scheme": "ndcgLoss1_scheme", "k": 5, "reduction": "mean"}),
                       (losses.neuralNDCG, {}), (losses.neuralNDCG, {"stochastic": True, "n_samples": 2})]:
            loss = fn(model(x, y == -1, None), y, **kw)
            loss.backward()
            opt.step()
            opt.zero_grad()
        with torch.no_grad():
            s = model.eval()(x, y == -1, None)
            metrics.all_metrics(s, y, [1, 5, 10])
            metrics.ndcg(s, y)
            metrics.ranking(s, y)
        torch.cuda.synchronize()
        print("ok", (F, d, N, h, dff, B, S, p), float(loss))
    packed_rows()
    next_rows()


def packed_rows():
    """The packed-rows layout (default for calls without dropout): ragged / full / single-item / empty slates, slate
    lengths that are not multiples of 16, row counts that are not multiples of 128, bf16 mode, both widths of the
    several-rows-per-warp row kernels (128 and 256), an FC block with an activation and input norm; train and eval."""
    from allrank_b200 import _lib
    assert _lib.lib().arb_get_pack_rows() == 1
    for (F, d, N, h, dff, B, S, kw) in [(136, 128, 2, 4, 256, 9, 240, {}), (136, 128, 1, 4, 256, 5, 120, {}),
                                        (20, 64, 1, 4, 64, 7, 50, {}), (136, 256, 2, 8, 512, 6, 240, {"compute_dtype": "bf16"}),
                                        (136, 128, 1, 8, 128, 33, 48, {"sizes": [96, 128], "act": "ReLU", "norm": True})]:
        sizes, act, norm = kw.pop("sizes", [d]), kw.pop("act", None), kw.pop("norm", False)
        model = make_model(fc_model={"sizes": sizes, "input_norm": norm, "activation": act, "dropout": 0.0},
                           transformer={"N": N, "d_ff": dff, "h": h, "positional_encoding": None, "dropout": 0.0},
                           post_model={"d_output": 1, "output_activation": None}, n_features=F, **kw).cuda().train()
        opt = FlatAdam(model)
        x, y, _ = make_slates(B, S, F, seed=3, mean_len=0.5 * S, std_len=0.3 * S)
        y[0] = 1.0                       # a full slate
        y[1] = -1.0
        y[1, 0] = 2.0                    # a single item
        x, y = x.cuda(), y.cuda()
        for fn in (losses.approxNDCGLoss, losses.listNet):
            loss = fn(model(x, y == -1, None), y)
            loss.backward()
            opt.step()
            opt.zero_grad()
        y[2] = -1.0                      # an empty slate (scored, not trained on: the reference's losses turn it into NaN)
        model(x, y == -1, None).sum().backward()
        opt.zero_grad()
        with torch.no_grad():
            s = model.eval()(x, y == -1, None)
            assert torch.isfinite(s).all()
            metrics.ndcg(s, y)
        torch.cuda.synchronize()
        print("ok packed", (F, d, N, h, dff, B, S), float(loss))


def next_rows():
    """The SURVEY 8(f) kernels: remaining losses, multi-output head + ordinal, positional encodings, slate movers."""
    import numpy as np
    from allrank_b200 import data, inference
    x, y, idx = make_slates(3, 50, 20, seed=2, mean_len=30, std_len=12)
    x, y, idx = x.cuda(), y.cuda(), idx.cuda()
    for n_out, act, pe in [(4, "Sigmoid", None), (3, None, {"strategy": "learned", "max_indices": 40}),
                           (1, None, {"strategy": "fixed", "max_indices": 40})]:
        for tr in ({"N": 1, "d_ff": 64, "h": 2, "positional_encoding": pe, "dropout": 0.1}, None):
            model = make_model(fc_model={"sizes": [32], "input_norm": False, "activation": None, "dropout": 0.1},
                               transformer=tr, post_model={"d_output": n_out, "output_activation": act},
                               n_features=20).cuda().train()
            out = model(x, y == -1, idx)
            if n_out > 1 and act == "Sigmoid":
                losses.ordinal(out, y, n=n_out).backward()
            else:
                out.sum().backward()
            with torch.no_grad():
                model.eval().score(x, y == -1, idx)
    s = torch.randn(y.shape, device="cuda", requires_grad=True)
    for fn, kw in [(losses.rankNet, {}), (losses.rankNet_weightByGTDiff_pow, {}), (losses.binary_listNet, {}),
                   (losses.pointwise_rmse, {"no_of_levels": 4}), (losses.neuralNDCG_transposed, {})]:
        fn(s, y, **kw).backward()
    losses.bce(torch.sigmoid(s), (y > 0).float()).backward()
    rng = np.random.RandomState(0)
    lens = [3, 40, 17, 64, 1]
    store = data.SlateStore(rng.randn(sum(lens), 20).astype(np.float32), rng.randint(0, 3, sum(lens)).astype(np.float32),
                            np.repeat(np.arange(len(lens)), lens), device="cuda")
    for S in (8, 17, 64, 100):
        xb, yb, ib = store.assemble(torch.tensor([0, 1, 2, 3, 4, 3]), S, seed=S)
        inference.reorder_slates(xb, yb, metrics.ranking(torch.randn(yb.shape, device="cuda"), yb))
    torch.cuda.synchronize()
    print("ok next rows")


if __name__ == "__main__":
    main()
