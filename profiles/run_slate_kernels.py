"""Measurement tooling (not product code): launch every loss / metric kernel once per repetition at the bench shapes,
so that `ncu` can capture them and CUDA events can time them alone.

    python profiles/run_slate_kernels.py [--batch 4096] [--reps 3] [--json out.json]

With --json the script times each kernel with CUDA events (20 launches after 3 warm-ups; the [B,S] inputs of all
kernels together are far smaller than L2, so an L2 flush buffer is written between launches) and writes achieved
algorithmic GB/s (12 S + 4 bytes per slate for fused loss fwd+bwd, 8 S + 4 n_ats for metrics) and
pairs/s next to the measured HBM copy peak.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from allrank_b200 import losses, metrics  # noqa: E402
from allrank_b200.synth import make_scores, make_slates  # noqa: E402


def cases(B):
    out = []
    for S in (240, 120):
        _, y, _ = make_slates(B, S, 4, seed=5)
        s = make_scores(B, S, seed=6)
        out.append((S, s.cuda().requires_grad_(True), y.cuda()))
    return out


KERNELS = [
    ("metrics_kernel", 240, lambda s, y: metrics.all_metrics(s.detach(), y, [1, 5, 10, 30]), "metrics"),
    ("listnet", 240, lambda s, y: losses.listNet(s, y).backward(), "loss"),
    ("listmle_kernel", 240, lambda s, y: losses.listMLE(s, y).backward(), "loss"),
    ("approx_ndcg_kernel", 240, lambda s, y: losses.approxNDCGLoss(s, y).backward(), "loss_pairs"),
    ("lambda_loss_kernel", 240,
     lambda s, y: losses.lambdaLoss(s, y, weighing_scheme="ndcgLoss2PP_scheme").backward(), "loss_pairs"),
    ("ranknet_kernel", 240, lambda s, y: losses.rankNet(s, y).backward(), "loss_pairs"),
    ("neural_ndcg_reg_kernel", 120, lambda s, y: losses.neuralNDCG(s, y).backward(), "loss_sinkhorn"),
    ("neural_ndcg_kernel", 240, lambda s, y: losses.neuralNDCG(s, y).backward(), "loss_sinkhorn"),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=1)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    data = {S: (s, y) for S, s, y in cases(a.batch)}
    peak = 6487.1
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk))["hbm_gbs"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    rows = []
    for name, S, fn, kind in KERNELS:
        s, y = data[S]
        B = a.batch if name != "neural_ndcg_kernel" else min(a.batch, 512)
        s, y = s[:B].detach().clone().requires_grad_(True), y[:B].contiguous()
        for _ in range(a.reps):
            fn(s, y)
            s.grad = None
        if a.json:
            n = 20
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
            for _ in range(3):
                fn(s, y)
            for e0, e1 in ev:
                flush.fill_(1)
                e0.record()
                fn(s, y)
                e1.record()
            torch.cuda.synchronize()
            ms = sorted(e0.elapsed_time(e1) for e0, e1 in ev)[n // 2]
            valid = (y != -1).sum(1).double()
            nbytes = B * ((8 * S + 16) if kind == "metrics" else (12 * S + 4))
            row = {"kernel": name, "B": B, "S": S, "ms_host_call": round(ms, 4),
                   "slates_per_s": round(B / ms * 1e3), "algorithmic_bytes": nbytes,
                   "achieved_gbs": round(nbytes / ms / 1e6, 2), "hbm_peak_gbs": peak,
                   "frac_of_measured_hbm_peak": round(nbytes / ms / 1e6 / peak, 5)}
            if kind == "loss_pairs":
                row["pairs_per_s"] = float((valid * valid).sum() / ms * 1e3)
            if kind == "loss_sinkhorn":
                row["matrix_element_updates_per_s"] = float((valid * valid).sum() * 100 / ms * 1e3)
            rows.append(row)
            print(json.dumps(row))
    torch.cuda.synchronize()
    if a.json:
        json.dump({"note": "CUDA-event time of the whole host call (kernel + finalize launch), median of 20, L2 flushed "
                           "between launches; bytes are algorithmic (SURVEY 8d)", "rows": rows}, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
