"""torchrun worker of tests/test_gpu_ddp.py (not collected by pytest): W ranks x b slates must reproduce the
single-process flat gradient of the W*b batch through the CUDA scorer + loss + FlatDDP."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    from allrank_b200 import losses
    from allrank_b200.ddp import FlatDDP, loss_weight
    from allrank_b200.model import make_model
    from allrank_b200.synth import make_slates
    dev = torch.device("cuda")
    b, S, F = 6, 48, 20
    x, y, _ = make_slates(world * b, S, F, seed=17, mean_len=30, std_len=12)
    y[1] = torch.where(y[1] >= 0, torch.zeros_like(y[1]), y[1])       # slates without a relevant item, uneven
    y[2] = torch.where(y[2] >= 0, torch.zeros_like(y[2]), y[2])       # across ranks: rank 0 has two, rank 1 one
    y[b + 3] = torch.where(y[b + 3] >= 0, torch.zeros_like(y[b + 3]), y[b + 3])
    x, y = x.to(dev), y.to(dev)

    def build():
        torch.manual_seed(5)
        return make_model(fc_model={"sizes": [32], "input_norm": False, "activation": None, "dropout": 0.0},
                          transformer={"N": 1, "d_ff": 64, "h": 2, "positional_encoding": None, "dropout": 0.0},
                          post_model={"d_output": 1, "output_activation": None}, n_features=F).to(dev).train()

    cases = [("approxNDCGLoss", {}, "mean"), ("listNet", {}, "mean"),
             ("lambdaLoss", {"weighing_scheme": "ndcgLoss2PP_scheme"}, "sum"),
             ("neuralNDCG", {}, "weighted"), ("neuralNDCG", {}, "naive_mean")]
    report = {}
    for name, kw, mode in cases:
        fn = getattr(losses, name)
        single = build()
        fn(single(x, y == -1, None), y, **kw).backward()
        want = single.flat_gradients.clone()
        model = build()
        ddp = FlatDDP(model, average=(mode != "sum"))
        sl = slice(rank * b, (rank + 1) * b)
        fn(model(x[sl], y[sl] == -1, None), y[sl], **kw).backward()
        ddp.reduce_gradients(local_weight=loss_weight(name, y[sl]) if mode == "weighted" else None)
        got = model.flat_gradients
        report[name + ":" + mode] = ((got - want).norm() / want.norm()).item()
    if rank == 0:
        print("DDP_PARITY " + json.dumps(report))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
