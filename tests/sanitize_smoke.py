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


if __name__ == "__main__":
    main()
