"""Debug tooling: fused vs unfused attention under dropout with the same masks (same call seed).
Modes: 0 unfused fwd+bwd, 1 fused fwd + unfused bwd, 2 fused fwd+bwd."""
import os
import sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from allrank_b200 import _lib
from allrank_b200.model import make_model
from allrank_b200.synth import make_slates

lib = _lib.lib()


def run(mode, d, h, dff, S, B, p, seed=31, full=False):
    torch.manual_seed(3)
    m = make_model(fc_model={"sizes": [d], "input_norm": False, "activation": None, "dropout": 0.0},
                   transformer={"N": 1, "d_ff": dff, "h": h, "positional_encoding": None, "dropout": p},
                   post_model={"d_output": 1, "output_activation": None}, n_features=20).cuda().train()
    x, y, _ = make_slates(B, S, n_features=20, seed=9, mean_len=0.66 * S, std_len=0.2 * S, full=full)
    lib.arb_set_attention_mode(mode)
    torch.manual_seed(seed)
    out = m(x.cuda(), (y == -1).cuda(), None)
    w = torch.randn(out.shape, generator=torch.Generator().manual_seed(6)).cuda()
    (out * w).sum().backward()
    lib.arb_set_attention_mode(2)
    return out.detach().cpu(), {k: q.grad.detach().cpu().clone() for k, q in m.named_parameters()}, y


for (d, h, dff, S, B, p, full) in [(32, 1, 64, 48, 6, 0.2, False), (32, 1, 64, 48, 6, 0.2, True), (32, 2, 64, 24, 6, 0.2, False),
                                   (64, 2, 128, 48, 5, 0.3, False), (128, 4, 256, 240, 4, 0.3, False), (32, 1, 64, 48, 6, 0.0, False)]:
    o0, g0, y = run(0, d, h, dff, S, B, p, full=full)
    valid = y != -1
    for mode in (1, 2):
        o, g, _ = run(mode, d, h, dff, S, B, p, full=full)
        errs = {k.replace("encoder.layers.0.", ""): float((g[k] - g0[k]).norm() / max(g0[k].norm(), 1e-12)) for k in g}
        worst = sorted(errs.items(), key=lambda kv: -kv[1])[:5]
        print(f"d={d} h={h} S={S} B={B} p={p} full={full} mode {mode} vs 0: score diff all {float((o - o0).abs().max()):.2e} "
              f"valid {float((o - o0)[valid].abs().max()):.2e} | " + " ".join(f"{k}={v:.4f}" for k, v in worst), flush=True)
