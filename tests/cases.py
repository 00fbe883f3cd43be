"""Known-answer cases restated from the reference's own unit tests.

Each entry cites the reference test it comes from.  The same tables drive the CPU oracle tests
(tests/test_oracle_known_answers.py) and the CUDA parity tests (tests/test_gpu_*.py).
"""
import math

PAD = -1.0

# (name, kwargs, y_pred, y_true, expected)      reference: tests/losses/test_*.py
LOSS_KNOWN = [
    # tests/losses/test_listmle.py:14-22
    ("listMLE", {}, [0.5, 0.3, 0.5], [1.0, 0.0, PAD], 0.5981389284133911),
    # tests/losses/test_approxndcg.py:10-20
    ("approxNDCGLoss", {"alpha": 1.0}, [0.5, 0.3, 0.5], [0.5, 0.3, 0.5], -0.8499219417),
    ("approxNDCGLoss", {"alpha": 1.0}, [0.5, 0.3, 0.5, 1.0], [0.5, 0.3, 0.5, PAD], -0.8499219417),
    # tests/losses/test_lambdaloss.py:10-46
    ("lambdaLoss", {"weighing_scheme": "ndcgLoss1_scheme", "reduction_log": "binary"},
     [0.5, 0.3, 0.5], [0.5, 0.3, 0.5], 2.9272110462),
    ("lambdaLoss", {"weighing_scheme": "ndcgLoss1_scheme", "reduction_log": "binary"},
     [0.5, 0.3, 0.5, 1.0], [0.5, 0.3, 0.5, PAD], 2.9272110462),
    ("lambdaLoss", {"weighing_scheme": "ndcgLoss2PP_scheme", "reduction_log": "binary"},
     [0.5, 0.3, 0.5], [0.5, 0.3, 0.5], 1.1244146823),
    ("lambdaLoss", {"weighing_scheme": "ndcgLoss2PP_scheme", "reduction_log": "binary"},
     [0.5, 0.3, 0.5, 1.0], [0.5, 0.3, 0.5, PAD], 1.1244146823),
    ("lambdaLoss", {"weighing_scheme": "rankNet_scheme", "reduction_log": "natural"},
     [0.5, 0.3, 0.5], [0.5, 0.3, 0.5], 1.1962778568),
    ("lambdaLoss", {"weighing_scheme": "rankNet_scheme", "reduction_log": "natural"},
     [0.5, 0.3, 0.5, 1.0], [0.5, 0.3, 0.5, PAD], 1.1962778568),
]


def _xe(true, pred):
    return -true * math.log(pred) - (1 - true) * math.log(1 - pred)


# (y_pred [S][n], y_true [S], n, expected)      reference: tests/losses/test_loss_ordinal.py:26-57
# (the padded case is the one the reference itself cannot run on torch >= 2; its closed form still pins the semantics)
ORDINAL_KNOWN = [
    ([[0.8, 0.6]], [1.0], 2, _xe(1, 0.8) + _xe(0, 0.6)),
    ([[0.8, 0.7], [0.4, 0.3], [0.2, 0.1]], [2.0, 1.0, 0.0], 2,
     (_xe(1, 0.8) + _xe(1, 0.7) + _xe(1, 0.4) + _xe(0, 0.3) + _xe(0, 0.2) + _xe(0, 0.1)) / 3.0),
    ([[0.8, 0.6], [0.2, 0.1]], [1.0, PAD], 2, _xe(1, 0.8) + _xe(0, 0.6)),
]
# tests/losses/test_loss_ordinal.py:20-24: with_ordinals([2,1,0], n=2)
WITH_ORDINALS_KNOWN = ([2.0, 1.0, 0.0], 2, [[1.0, 1.0], [1.0, 0.0], [0.0, 0.0]])


def _softmax(v):
    m = max(v)
    e = [math.exp(a - m) for a in v]
    z = sum(e)
    return [a / z for a in e]


def listnet_closed_form(y_pred, y_true, eps):
    valid = [(p, t) for p, t in zip(y_pred, y_true) if t != PAD]
    ps = _softmax([p for p, _ in valid])
    ts = _softmax([t for _, t in valid])
    return -sum(t * math.log(p + eps) for p, t in zip(ps, ts))


# tests/losses/test_listnet.py:16-46   (name, y_pred, y_true, eps)
LISTNET_KNOWN = [
    ([0.5, 0.2], [1.0, 0.0], 0.0),
    ([0.5, -1e30], [1.0, 0.0], 1e-10),
    ([0.5, 0.2, 0.5], [1.0, 0.0, PAD], 1e-10),
]

# tests/losses/test_ndcg.py:14-69    (y_pred, y_true, ats, expected [len(ats)] , exact?)
NDCG_KNOWN = [
    ([0.5, 0.2], [1.0, 0.0], None, [1.0], True),
    ([0.5, 0.2], [0.0, 1.0], None, [1 / math.log2(3)], True),
    # test_ndcg_zero_when_no_relevant FAILS in the reference (expects 0.0); the code returns
    # filler_value = 1.0 (metrics.py:8,24, reproducibility/HOWTO.md:32) -- we pin the CODE's behaviour.
    ([0.5, 0.2], [0.0, 0.0], None, [1.0], True),
    ([0.5, 0.2, 0.1], [1.0, 0.0, 1.0], [1, 2], [1.0, 1.0 / (1.0 + 1 / math.log2(3))], False),
    ([0.5, 0.2, 1.0], [1.0, 0.0, PAD], None, [1.0], True),
    ([0.5, 0.2, 1.0], [0.0, 1.0, PAD], None, [1 / math.log2(3)], True),
]

# tests/losses/test_mrr.py:19-104    (y_pred [B,S], y_true [B,S], ats, expected [B,len(ats)])
MRR_KNOWN = [
    ([[0.5, 0.2]], [[1.0, 0.0]], [10], [[1.0]]),
    ([[0.5, 0.2]], [[1.0, 0.0]], None, [[1.0]]),
    ([[0.5, 0.2]], [[0.0, 1.0]], [10], [[0.5]]),
    ([[0.2, 0.5], [0.5, 0.2]], [[0.0, 1.0], [0.0, 1.0]], [10], [[1.0], [0.5]]),
    ([[0.5, 0.2]], [[0.0, 1.0]], [1, 2], [[0.0, 0.5]]),
    ([[0.2, 0.5], [0.5, 0.2]], [[0.0, 1.0], [0.0, 1.0]], [1, 2], [[1.0, 1.0], [0.0, 0.5]]),
    ([[0.5, 0.2]], [[0.0, 0.0]], [10], [[0.0]]),
    ([[0.5, 0.2, 1.0]], [[1.0, 0.0, PAD]], [10], [[1.0]]),
    ([[0.5, 0.2, 1.0]], [[0.0, 1.0, PAD]], [10], [[0.5]]),
]

# tests/losses/test_neuralndcg.py:10-94 : -neuralNDCG(tau -> 0) == ndcg     (y_pred, y_true, kwargs)
NEURALNDCG_EQUIV = [
    ([0.5, 0.2], [1.0, 0.0], {"temperature": 1e-4}),
    ([0.5, 0.2, 0.1, 0.4, 1.0, -1.0, 0.63], [1.0, 2.0, 2.0, 4.0, 1.0, 4.0, 3.0], {"temperature": 1e-4}),
    ([0.5, -1e30], [1.0, 0.0], {"temperature": 1e-4}),
    ([0.5, 0.2, 0.1, 0.4, 1.0, -1.0, 0.63, 1.0, 0.5, 0.3],
     [1.0, 2.0, 2.0, 4.0, 1.0, 4.0, 3.0, PAD, PAD, PAD], {"temperature": 1e-3}),
    ([0.5, 0.2, 0.1, 0.4, 1.0, -1.0, 0.63], [1.0, 2.0, 2.0, 4.0, 1.0, 4.0, 3.0], {"temperature": 1e-4, "k": 3}),
]
