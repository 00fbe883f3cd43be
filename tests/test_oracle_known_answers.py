"""Pin the CPU oracle against every known-answer value the reference's own tests hold for the path."""
import math

import pytest
import torch

from oracle import losses_ref, metrics_ref
from tests import cases


@pytest.mark.parametrize("name,kw,yp,yt,expected", cases.LOSS_KNOWN)
def test_loss_known_answers(name, kw, yp, yt, expected):
    val = losses_ref.LOSSES[name](torch.tensor([yp]), torch.tensor([yt]), **kw).item()
    assert math.isfinite(val)
    assert val == pytest.approx(expected)


@pytest.mark.parametrize("yp,yt,n,expected", cases.ORDINAL_KNOWN)
def test_ordinal_known_answers(yp, yt, n, expected):
    val = losses_ref.ordinal(torch.tensor([yp]), torch.tensor([yt]), n).item()
    assert math.isfinite(val)
    assert val == pytest.approx(expected)


def test_with_ordinals_known_answer():
    y, n, expected = cases.WITH_ORDINALS_KNOWN
    assert losses_ref.with_ordinals(torch.tensor([y]), n).tolist() == [expected]
    padded = losses_ref.with_ordinals(torch.tensor([[1.0, cases.PAD]]), 2).tolist()
    assert padded == [[[1.0, 0.0], [cases.PAD, cases.PAD]]]


@pytest.mark.parametrize("yp,yt,eps", cases.LISTNET_KNOWN)
def test_listnet_closed_form(yp, yt, eps):
    val = losses_ref.listNet(torch.tensor([yp]), torch.tensor([yt]), eps).item()
    assert math.isfinite(val)
    assert val == pytest.approx(cases.listnet_closed_form(yp, yt, eps))


@pytest.mark.parametrize("yp,yt,ats,expected,exact", cases.NDCG_KNOWN)
def test_ndcg_known(yp, yt, ats, expected, exact):
    out = metrics_ref.ndcg(torch.tensor([yp]), torch.tensor([yt]), ats=ats).numpy()[0]
    if exact:
        assert out[0] == torch.tensor(expected[0], dtype=torch.float32).item() or out[0] == expected[0]
    assert out == pytest.approx(expected)


@pytest.mark.parametrize("yp,yt,ats,expected", cases.MRR_KNOWN)
def test_mrr_known(yp, yt, ats, expected):
    out = metrics_ref.mrr(torch.tensor(yp), torch.tensor(yt), ats=ats).numpy()
    assert (out == torch.tensor(expected).numpy()).all()


@pytest.mark.parametrize("yp,yt,kw", cases.NEURALNDCG_EQUIV)
def test_neuralndcg_matches_ndcg_at_low_temperature(yp, yt, kw):
    p, t = torch.tensor([yp]), torch.tensor([yt])
    val = losses_ref.neuralNDCG(p, t, **kw).item()
    k = kw.get("k")
    expected = metrics_ref.ndcg(p, t, ats=None if k is None else [k]).mean().item()
    assert math.isfinite(val)
    assert -val == pytest.approx(expected)


def test_lambdaloss_error_conventions():
    p, t = torch.tensor([[0.5, 0.3]]), torch.tensor([[1.0, 0.0]])
    with pytest.raises(ValueError):
        losses_ref.lambdaLoss(p, t, reduction="median")
    with pytest.raises(ValueError):
        losses_ref.lambdaLoss(p, t, reduction_log="decimal")
    with pytest.raises(KeyError):
        losses_ref.lambdaLoss(p, t, weighing_scheme="nope")
