"""CUDA-graph replay of a training step (allrank_b200.graph.GraphedTrainStep) and the optimiser entry it needs
(arb_adam_step_dev: step counter on the device).  The captured sequence is the reference's loss_batch
(allrank/training/train_utils.py:18-29) followed by the optimiser step of fit()."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _model(seed=3):
    from allrank_b200.model import make_model
    torch.manual_seed(seed)
    return make_model(fc_model={"sizes": [64], "input_norm": False, "activation": None, "dropout": 0.0},
                      transformer={"N": 1, "d_ff": 128, "h": 2, "positional_encoding": None, "dropout": 0.0},
                      post_model={"d_output": 1, "output_activation": None}, n_features=136).cuda().train()


def test_device_side_step_counter_matches_the_host_side_one():
    """Same gradients, five steps: FlatAdam(capturable=True) -- bias corrections computed on the device from a counter
    it increments itself -- must give bit-identical parameters to the host-counter path."""
    from allrank_b200.optim import FlatAdam
    from allrank_b200.synth import make_slates
    x, y, _ = make_slates(4, 32, 136, seed=5)
    x, y = x.cuda(), y.cuda()
    outs = []
    for capturable in (False, True):
        model = _model()
        opt = FlatAdam(model, lr=1e-2, capturable=capturable)
        model(x, y == -1, None)                              # packs the parameters
        g = torch.Generator(device="cuda").manual_seed(1)
        for _ in range(5):
            model.flat_gradients.copy_(torch.randn(model.flat_gradients.shape, device="cuda", generator=g))
            opt.step()
        outs.append(model.flat_parameters.clone())
        if capturable:
            assert opt._dev_state[0].item() == 5.0
    assert torch.equal(outs[0], outs[1])


def test_graphed_training_step_follows_the_eager_steps():
    from allrank_b200.graph import GraphedTrainStep
    from allrank_b200.losses import approxNDCGLoss
    from allrank_b200.optim import FlatAdam
    from allrank_b200.synth import make_slates
    batches = [make_slates(16, 48, 136, seed=10 + k) for k in range(4)]
    batches = [(x.cuda(), y.cuda()) for x, y, _ in batches]

    eager = _model()
    opt = FlatAdam(eager, lr=1e-3, capturable=True)
    eager_losses = []
    for x, y in batches * 2:
        loss = approxNDCGLoss(eager(x, y == -1, None), y)
        opt.zero_grad()
        loss.backward()
        opt.step()
        eager_losses.append(loss.item())

    graphed = _model()
    gopt = FlatAdam(graphed, lr=1e-3, capturable=True)
    init = {k: v.clone() for k, v in graphed.state_dict().items()}
    step = GraphedTrainStep(graphed, approxNDCGLoss, gopt, *batches[0], warmup=2)
    # construction ran warm-up steps: rewind the model and the optimiser state, then replay the same eight batches
    graphed.load_state_dict(init)
    gopt.exp_avg.zero_(); gopt.exp_avg_sq.zero_(); gopt._dev_state.zero_()
    graph_losses = [step(x, y).item() for x, y in batches * 2]
    assert gopt._dev_state[0].item() == 8.0
    for a, b in zip(eager_losses, graph_losses):
        assert abs(a - b) <= 2e-4 * max(1.0, abs(a)), (eager_losses, graph_losses)
    assert eager_losses[0] != eager_losses[-1]
    # Adam turns summation-order noise of gradient entries that are mathematically zero (the key-projection bias) into
    # +-lr moves per step: two runs can differ by up to 2 * lr per step on such entries -- a loose bound
    assert (eager.flat_parameters - graphed.flat_parameters).abs().max().item() <= 2 * 8 * 1e-3 + 4e-3


def test_graphed_step_refuses_what_it_cannot_capture():
    from allrank_b200.graph import GraphedTrainStep
    from allrank_b200.losses import listNet
    from allrank_b200.model import make_model
    from allrank_b200.optim import FlatAdam
    from allrank_b200.synth import make_slates
    x, y, _ = make_slates(4, 32, 136, seed=5)
    x, y = x.cuda(), y.cuda()
    m = _model()
    with pytest.raises(ValueError):
        GraphedTrainStep(m, listNet, FlatAdam(m), x, y)                       # host-side step counter
    torch.manual_seed(1)
    d = make_model(fc_model={"sizes": [64], "input_norm": False, "activation": None, "dropout": 0.0},
                   transformer={"N": 1, "d_ff": 128, "h": 2, "positional_encoding": None, "dropout": 0.2},
                   post_model={"d_output": 1, "output_activation": None}, n_features=136).cuda().train()
    with pytest.raises(ValueError):
        GraphedTrainStep(d, listNet, FlatAdam(d, capturable=True), x, y)      # dropout seed drawn on the host
