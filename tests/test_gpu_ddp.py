"""Two-GPU data-parallel parity through the CUDA path (needs >= 2 devices: `gpurun --gpus 2`): 2 ranks x b slates
reproduce the single-process flat gradient of the 2b batch -- mean losses averaged, lambdaLoss(reduction="sum")
summed, neuralNDCG with the (numerator, count) reduction; the naive per-rank mean is shown NOT to."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_ranks_reproduce_the_single_process_gradient():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "ddp_parity_worker.py")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    rep = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("DDP_PARITY ")][-1][len("DDP_PARITY "):])
    print(rep)
    for key in ("approxNDCGLoss:mean", "listNet:mean", "lambdaLoss:sum", "neuralNDCG:weighted"):
        assert rep[key] < 2e-3, (key, rep[key])          # TF32 products, different reduction order
    assert rep["neuralNDCG:naive_mean"] > 5 * rep["neuralNDCG:weighted"] + 5e-3          # per-rank means are NOT the global mean when counts differ
