"""Debug tooling: localise the train-mode gradient discrepancy of tests/test_gpu_fc_block.py case 2 by varying one
thing at a time (shape, dropout sites, switches).  Prints the error of a few parameter gradients per variant."""
import os
import sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from allrank_b200 import _lib
from allrank_b200.model import make_model
from allrank_b200.synth import make_slates
from oracle.scorer_ref import make_ref_model
from oracle.tf32_emulation import scorer_forward
from tests.dropout_masks import scorer_masks

KEYS = ["feed_forward.w_2.weight", "feed_forward.w_1.weight", "sublayer.1.norm.a_2", "self_attn.linears.3.weight",
        "self_attn.linears.2.weight", "self_attn.linears.0.weight", "sublayer.0.norm.a_2", "input_layer.layers.0.weight"]


def run(tag, Fn=136, sizes=(64, 32), act="Sigmoid", inorm=True, N=1, h=2, dff=64, p=0.2, p_fc=0.25, B=6, S=24, mode="rna"):
    ref = make_ref_model(Fn, list(sizes), N, h, dff, dropout=0.0, fc_activation=act, seed=3, input_norm=inorm)
    mine = make_model(fc_model={"sizes": list(sizes), "input_norm": inorm, "activation": act, "dropout": p_fc},
                      transformer={"N": N, "d_ff": dff, "h": h, "positional_encoding": None, "dropout": p},
                      post_model={"d_output": 1, "output_activation": None}, n_features=Fn)
    mine.load_state_dict(ref.state_dict())
    mine = mine.cuda().train()
    x, y, _ = make_slates(B, S, n_features=Fn, seed=9, mean_len=0.66 * S, std_len=0.2 * S)
    mask = y == -1
    torch.manual_seed(31)
    seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if (p > 0 or p_fc > 0) else 0
    torch.manual_seed(31)
    out = mine(x.cuda(), mask.cuda(), None)
    w = torch.randn(out.shape, generator=torch.Generator().manual_seed(6))
    (out * w.cuda()).sum().backward()
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in mine.state_dict().items()}
    r = scorer_forward(sd, x, mask, N, h, None, mode, drop=scorer_masks(seed, B, S, list(sizes), N, h, dff, p, p_fc), fc_act=act)
    (r * w).sum().backward()
    floor = 1e-2 * max(v.grad.abs().max().item() for v in sd.values())
    errs = {}
    for k, q in mine.named_parameters():
        a, g = q.grad.cpu().double().numpy(), sd[k].grad.double().numpy()
        errs[k] = np.linalg.norm(a - g) / max(np.linalg.norm(g), floor * np.sqrt(g.size))
    line = " ".join(f"{kk.split('.')[-2][:6]}.{kk.split('.')[-1][:1]}={errs[k]:.4f}" for kk in KEYS for k in errs if k.endswith(kk))
    print(f"{tag:34s} score {float((r.detach() - out.detach().cpu())[~mask].abs().max()):.5f} | {line}", flush=True)


lib = _lib.lib()
run("case2")
run("case2 trunc-emulation", mode="trunc")
run("p=0 (fc drop only)", p=0.0)
run("p_fc=0 (encoder drop only)", p_fc=0.0)
run("no dropout", p=0.0, p_fc=0.0)
run("dff=128", dff=128)
run("d=64 (sizes 64,64) h=2", sizes=(64, 64))
run("h=1 (dk=32)", h=1)
run("h=4 (dk=8, unfused)", h=4)
run("no input_norm, act None", act=None, inorm=False)
run("S=48", S=48)
run("B=24", B=24)
lib.arb_set_pdl(0); run("PDL off"); lib.arb_set_pdl(1)
lib.arb_set_attention_skip_padding(0); run("skip off"); lib.arb_set_attention_skip_padding(1)
lib.arb_set_attention_mode(0); run("unfused attention"); lib.arb_set_attention_mode(2)
lib.arb_set_tf32_round_on_load(0); run("tensor-core truncation vs trunc emu", mode="trunc"); lib.arb_set_tf32_round_on_load(1)
