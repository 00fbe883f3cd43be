"""FlatAdam (one launch over the flat buffers) against torch.optim.Adam fed the same gradient sequence."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_flat_adam_matches_torch_adam():
    from allrank_b200.model import make_model
    from allrank_b200.optim import FlatAdam
    from allrank_b200.synth import make_slates
    torch.manual_seed(3)
    model = make_model(fc_model={"sizes": [64], "input_norm": False, "activation": None, "dropout": 0.0},
                       transformer={"N": 1, "d_ff": 128, "h": 2, "positional_encoding": None, "dropout": 0.0},
                       post_model={"d_output": 1, "output_activation": None}, n_features=136).cuda().train()
    x, y, _ = make_slates(4, 30, seed=2, mean_len=20, std_len=5)
    model(x.cuda(), (y == -1).cuda(), None).sum().backward()      # packs parameters, attaches flat gradients
    flat = model.flat_parameters
    twin = flat.clone().requires_grad_(True)                      # torch.optim.Adam steps an independent copy
    ref = torch.optim.Adam([twin], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4)
    mine = FlatAdam(model, lr=1e-3, weight_decay=1e-4)
    g = torch.Generator(device="cuda").manual_seed(1)
    for _ in range(7):
        grad = torch.randn(flat.shape, device="cuda", generator=g) * 0.1
        model.flat_gradients.copy_(grad)
        twin.grad = grad.clone()
        mine.step()
        ref.step()
    assert torch.allclose(flat, twin.detach(), rtol=1e-5, atol=1e-7)
    # folded 1/world scaling: step(grad_scale=0.5) on 2g == step on g
    a = flat.clone()
    model.flat_gradients.copy_(2 * grad)
    m2 = FlatAdam(model, lr=1e-3); m2.step(grad_scale=0.5)
    after_scaled = flat.clone()
    flat.copy_(a)
    model.flat_gradients.copy_(grad)
    m3 = FlatAdam(model, lr=1e-3); m3.step()
    assert torch.allclose(flat, after_scaled, rtol=1e-6, atol=1e-8)
