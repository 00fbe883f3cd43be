#!/usr/bin/env python
"""Benchmark of the allRank hot path on B200: slates/sec of a full training step
(scorer forward + listwise loss + backward + Adam) on synthetic MSLR-WEB30K-shaped slates.

    python bench.py --gpus 1 --steps 20 --warmup 5                  # this repo's CUDA path
    python bench.py --impl reference --steps 5 --warmup 3           # the reference's eager CPU path (oracle port)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0).  Contract: see the task statement; fields are documented in DESIGN.md section 6.
Workloads (BASELINE.json configs):
    cfg2  Transformer(N=2,h=4,d=128,d_ff=512) + approxNDCGLoss, S=240, F=136      <- the metric's configuration
    cfg3  Transformer(N=4,h=8,d=256,d_ff=1024) + lambdaLoss(ndcgLoss2PP), S=240   (TF32 here; bf16 not built yet)
    cfg4  Transformer(N=2,h=4,d=128,d_ff=512) + neuralNDCG, S=120
    cfg5  Transformer(N=4,h=8,d=256,d_ff=1024) + listMLE, S=240
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "cfg2": dict(N=2, h=4, d=128, dff=512, S=240, loss="approxNDCGLoss", loss_args={"alpha": 1.0}),
    "cfg3": dict(N=4, h=8, d=256, dff=1024, S=240, loss="lambdaLoss",
                 loss_args={"weighing_scheme": "ndcgLoss2PP_scheme", "k": None, "mu": 10.0, "sigma": 1.0}),
    "cfg4": dict(N=2, h=4, d=128, dff=512, S=120, loss="neuralNDCG",
                 loss_args={"temperature": 1.0, "k": None, "powered_relevancies": True}),
    "cfg5": dict(N=4, h=8, d=256, dff=1024, S=240, loss="listMLE", loss_args={}),
}
F = 136
PAD = -1


def flops_per_slate_step(w):
    """Algorithmic FLOPs of one training step per slate: 3 x forward."""
    S, d, dff, N = w["S"], w["d"], w["dff"], w["N"]
    fwd = 2 * S * F * d + N * (8 * S * d * d + 4 * S * S * d + 4 * S * d * dff) + 2 * S * d
    return 3 * fwd


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p["bf16_tflops"],
                    bf16_tflops_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx = max(mx, float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return None
        busy = [v for v in sm if v > 0.5 * mx] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_steps(w, batch, steps, warmup, threads):
    """The reference's eager PyTorch path on the host cores: oracle port of make_model + loss + torch Adam
    (train_utils.loss_batch semantics, allrank/training/train_utils.py:18-29)."""
    from oracle import losses_ref
    from oracle.scorer_ref import make_ref_model
    from allrank_b200.synth import make_slates
    torch.manual_seed(42)
    model = make_ref_model(F, [w["d"]], w["N"], w["h"], w["dff"]).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    loss_fn = losses_ref.LOSSES[w["loss"]]
    x, y, idx = make_slates(batch, w["S"], F, seed=1234)

    def one_step():
        t0 = time.perf_counter()
        mask = y == PAD
        loss = loss_fn(model(x, mask, idx), y, **w["loss_args"])
        loss.backward()
        opt.step()
        opt.zero_grad()
        _ = loss.item()
        return time.perf_counter() - t0

    # "all the host threads it can use": intra-op thread counts above the box's real core budget make eager
    # PyTorch slower, so pick the fastest of a few candidates (one untimed + one timed step each) -- the
    # baseline reported is the best the host achieves.
    best, best_t = threads, None
    for cand in sorted({c for c in (8, 16, 32, 64, threads) if c <= threads}):
        torch.set_num_threads(cand)
        one_step()
        t = one_step()
        if best_t is None or t < best_t:
            best, best_t = cand, t
    torch.set_num_threads(best)
    times = [one_step() for _ in range(warmup + steps)][warmup:]
    total = sum(times)
    return batch * steps / total, 1e3 * total / steps, best


def run_reference(args, w):
    rank, _, world = dist_env()
    if rank != 0:
        return
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    batch = args.ref_batch
    sps, ms, used = cpu_reference_steps(w, batch, args.steps, args.warmup, threads)
    sample = (f"{batch} slates/step x {args.steps} steps (S={w['S']}, F={F}), eager PyTorch fp32, {used} intra-op threads "
              f"(fastest of 8..{threads} on a {threads}-thread host)")
    out = {
        "impl": "reference", "metric": "slates/sec", "value": sps, "unit": "slates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, w, batch),
        "cpu_baseline": {"value": sps, "unit": "slates/s", "cores": used, "kind": "port", "sample": sample},
        "e2e": {"value": sps, "unit": "slates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def workload_config(args, w, batch):
    return {"workload": f"{args.workload}: Transformer(N={w['N']},h={w['h']},d_model={w['d']},d_ff={w['dff']}) + "
                        f"{w['loss']}, slate_len={w['S']}, {F} features, full training step (fwd+loss+bwd+Adam)",
            "batch_per_gpu": batch, "slate_len": w["S"], "n_features": F, "loss": w["loss"],
            "optimizer": "Adam(lr=1e-3)", "l2": "inputs (x alone is batch*S*F*4 bytes) and activations exceed the "
                                                "126 MB L2; no explicit flush"}


# ------------------------------------------------------------------------------------------------ GPU arm
def run_b200(args, w):
    rank, local_rank, world = dist_env()
    import torch.distributed as dist
    from allrank_b200 import _lib, losses
    from allrank_b200.ddp import FlatDDP
    from allrank_b200.model import make_model
    from allrank_b200.optim import FlatAdam
    from allrank_b200.synth import make_slates

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (allrank_b200 has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, S = args.batch, w["S"]
    torch.manual_seed(42)
    model = make_model(fc_model={"sizes": [w["d"]], "input_norm": False, "activation": None, "dropout": 0.0},
                       transformer={"N": w["N"], "d_ff": w["dff"], "h": w["h"], "positional_encoding": None,
                                    "dropout": 0.0},
                       post_model={"d_output": 1, "output_activation": None}, n_features=F).to(dev).train()
    loss_fn = getattr(losses, w["loss"])
    x_host, y_host, _ = make_slates(B, S, F, seed=1234 + rank)
    x_host, y_host = x_host.pin_memory(), y_host.pin_memory()
    x_dev, y_dev = x_host.to(dev), y_host.to(dev)
    opt = FlatAdam(model, lr=1e-3)
    ddp = FlatDDP(model) if world > 1 else None
    average = w["loss"] != "lambdaLoss"      # lambdaLoss(reduction="sum") gradients are summed across ranks
    if ddp:
        ddp.average = average

    def step(x, y):
        mask = y == PAD                                    # train_utils.py:19
        loss = loss_fn(model(x, mask, None), y, **w["loss_args"])
        loss.backward()
        scale = ddp.reduce_gradients(fold_average_into_optimizer=True) if ddp else 1.0
        opt.step(grad_scale=scale)
        opt.zero_grad()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    if world > 1:                     # pack + broadcast rank-0 weights before the first step
        model._ensure_packed(dev)
        ddp.sync_parameters()
    for _ in range(args.warmup):
        step(x_dev, y_dev)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()

    # ---- (1) device-resident: inputs already in HBM
    barrier()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step(x_dev, y_dev)
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = _lib.launch_count() - l0
    final_loss = loss.item()

    # ---- (2) end to end: every step's inputs come from pinned host memory (H2D inside the timed region, issued on
    #      a copy stream one step ahead, the way a training loop with a prefetching loader runs) and the loss is
    #      read back to the host every step (train_utils.loss_batch returns loss.item(), train_utils.py:29)
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)

    def prefetch():
        with torch.cuda.stream(copy_stream):
            xb = x_host.to(dev, non_blocking=True)
            yb = y_host.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return xb, yb, ev

    def e2e_loop(n):
        nxt = prefetch()
        for i in range(n):
            xb, yb, ev = nxt
            main_stream.wait_event(ev)
            xb.record_stream(main_stream)
            yb.record_stream(main_stream)
            if i + 1 < n:
                nxt = prefetch()
            _ = step(xb, yb).item()

    e2e_loop(2)
    barrier()
    e0.record()
    e2e_loop(args.steps)
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if sampler else None

    # ---- (3) per-launch device timing of every kernel class (roofline), a few extra steps
    prof = {}
    lib = _lib.lib()
    lib.arb_prof_enable.argtypes = [ctypes.c_int32]
    lib.arb_prof_collect.argtypes = [ctypes.c_int32, ctypes.POINTER(ctypes.c_double),
                                     ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64)]
    if rank == 0:
        lib.arb_prof_enable(1)
    psteps = 3
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record()
    for _ in range(psteps):          # every rank runs these steps (they contain the gradient all-reduce)
        step(x_dev, y_dev)
    pe1.record()
    torch.cuda.synchronize()
    if rank == 0:
        names = ["gemm_tf32", "scorer_simt", "loss", "metrics", "adam"]
        for cls, nm in enumerate(names):
            ms, work, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_int64()
            lib.arb_prof_collect(cls, ctypes.byref(ms), ctypes.byref(work), ctypes.byref(n))
            prof[nm] = {"ms_per_step": ms.value / psteps, "work_per_step": work.value / psteps,
                        "launches_per_step": n.value / psteps}
            if cls == 0:
                lib.arb_prof_last_bytes.restype = ctypes.c_double
                lib.arb_prof_last_bytes.argtypes = [ctypes.c_int32]
                prof[nm]["algorithmic_bytes_per_step"] = lib.arb_prof_last_bytes(0) / psteps
        prof["step_ms_profiled"] = pe0.elapsed_time(pe1) / psteps
        lib.arb_prof_enable(0)
    if world > 1:
        dist.barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = measured_peaks()
    slates = B * world * args.steps
    value = slates / (ms_total / 1e3)
    e2e_value = slates / (ms_e2e / 1e3)
    gemm = prof["gemm_tf32"]
    achieved_tflops = gemm["work_per_step"] / (gemm["ms_per_step"] * 1e-3) / 1e12 if gemm["ms_per_step"] > 0 else 0.0
    tf32_peak = peaks["bf16_tflops_sustained"] / 2.0
    out = {
        "metric": "slates/sec", "value": value, "unit": "slates/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "tf32", "data": "synthetic",
        "config": workload_config(args, w, B),
        "e2e": {"value": e2e_value, "unit": "slates/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(x_host.numel() * 4 + y_host.numel() * 4), "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {
            "bound": "tensor", "kernel": "gemm_tf32_kernel (tcgen05.mma kind::tf32)",
            "achieved": achieved_tflops, "peak": tf32_peak, "unit": "TFLOP/s",
            "frac": achieved_tflops / tf32_peak if tf32_peak else None,
            "traffic": measured_traffic(args, B),
            "traffic_note": "DRAM read+write bytes of the class's launches in one step (ncu dram__bytes_{read,write}.sum, "
                            "profiles/r1_dram_traffic_b4096.json: 66.8 GB vs 67.7 GB algorithmic at batch 4096); null when "
                            "this run's workload/batch has no committed capture",
            "peak_source": f"{peaks['source']}: MEASURED_PEAKS bf16_tflops_sustained={peaks['bf16_tflops_sustained']} / 2 "
                           "(kind::tf32 issues at half the bf16 rate)",
            "frac_of_bf16_peak": achieved_tflops / peaks["bf16_tflops_sustained"],
            "algorithmic_flops_per_step": gemm["work_per_step"],
            "kernel_ms_per_step": gemm["ms_per_step"], "launches_per_step": gemm["launches_per_step"],
            "share_of_step": gemm["ms_per_step"] / prof["step_ms_profiled"] if prof.get("step_ms_profiled") else None,
            "model_flops_per_step": flops_per_slate_step(w) * B,
            "whole_step_tflops": flops_per_slate_step(w) * B * world / (ms_total / args.steps * 1e-3) / 1e12,
            # the same kernels against the HBM roofline: with K = 128..512 the linears move 4(MK+NK+MN) bytes for
            # 2MNK flops (arithmetic intensity 32..100 flop/B, below the ~110 flop/B tf32 ridge), so HBM is the
            # binding roof; ncu (profiles/) shows 51-64 % of DRAM peak on the individual launches
            "hbm_view": {"achieved_gbs": gemm.get("algorithmic_bytes_per_step", 0.0) / (gemm["ms_per_step"] * 1e-3) / 1e9
                         if gemm["ms_per_step"] > 0 else 0.0,
                         "peak_gbs": peaks["hbm_gbs"],
                         "frac": (gemm.get("algorithmic_bytes_per_step", 0.0) / (gemm["ms_per_step"] * 1e-3) / 1e9) /
                                 peaks["hbm_gbs"] if gemm["ms_per_step"] > 0 else None,
                         "algorithmic_bytes_per_step": gemm.get("algorithmic_bytes_per_step", 0.0),
                         "asymmetric": asymmetric_hbm_view(args, B, gemm["ms_per_step"])},
        },
        "kernel_classes": prof,
        "final_loss": final_loss,
    }
    if world == 1 and not args.no_cpu_baseline:
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        cb, csteps = args.ref_batch, 3
        t0 = time.time()
        sps, ms, used = cpu_reference_steps(w, cb, csteps, 1, threads)
        out["cpu_baseline"] = {"value": sps, "unit": "slates/s", "cores": used, "kind": "port",
                               "sample": f"{cb} slates/step x {csteps} steps after 1 warm-up (same shapes), eager "
                                         f"PyTorch fp32, {used} intra-op threads (fastest of 8..{threads}), "
                                         f"{time.time() - t0:.1f}s of CPU work"}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _traffic_capture(args, batch):
    """The committed ncu capture of this exact workload / batch (profiles/r1_dram_traffic_b4096.json), else None."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1_dram_traffic_b4096.json")
    try:
        with open(path) as f:
            cap = json.load(f)
        if args.workload == "cfg2" and int(batch) == int(cap.get("batch", -1)):
            return cap
    except (OSError, ValueError, TypeError):
        pass
    return None


def measured_traffic(args, batch):
    """DRAM bytes per step of the tcgen05 class from the committed ncu capture."""
    cap = _traffic_capture(args, batch)
    return cap.get("tcgen05_dram_bytes_per_step") if cap else None


def asymmetric_hbm_view(args, batch, kernel_ms):
    """HBM reads and writes do not cost the same on this part (profiles/README.md): the class's measured read / write
    bytes against t = read/6.9 TB/s + write/3.25 TB/s, as a fraction of the class's measured time.  None without a
    capture of this workload."""
    try:
        cap = _traffic_capture(args, batch)
        if not cap or kernel_ms <= 0:
            return None
        rd, wr = float(cap["tcgen05_dram_read_bytes_per_step"]), float(cap["tcgen05_dram_write_bytes_per_step"])
        bound_ms = (rd / (float(cap["read_gbs_model"]) * 1e9) + wr / (float(cap["write_gbs_model"]) * 1e9)) * 1e3
        return {"read_bytes_per_step": rd, "write_bytes_per_step": wr, "bound_ms": bound_ms, "frac": bound_ms / kernel_ms,
                "model": cap.get("model")}
    except Exception:       # reporting extra only: never lose the bench line over it
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=4096,
                    help="slates per step per GPU (saturating batch; 64 = allRank's default batch_size, see profiles/)")
    ap.add_argument("--ref-batch", type=int, default=64, help="slates per CPU step (reference default batch_size)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_b200(args, w)


if __name__ == "__main__":
    main()
