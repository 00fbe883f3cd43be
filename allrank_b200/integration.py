"""Plug allrank_b200 into an unmodified allRank checkout.

allRank resolves its hot-path callables by NAME at run time:
    loss   : getattr(allrank.models.losses, config.loss.name)        allrank/main.py:83
    metric : getattr(allrank.models.metrics, name)                   allrank/training/train_utils.py:50
    model  : allrank.models.model.make_model(...)                    allrank/main.py:75
so `patch_allrank()` only has to rebind those attributes; `allrank/main.py` and `train_utils.fit` then run
unchanged on the B200 kernels.  (allRank must be importable; nothing here imports it otherwise.)
"""
from . import losses as _losses
from . import metrics as _metrics
from . import model as _model

LOSS_NAMES = ("listNet", "listMLE", "approxNDCGLoss", "lambdaLoss", "neuralNDCG", "neuralNDCG_transposed", "rankNet",
              "rankNet_weightByGTDiff", "rankNet_weightByGTDiff_pow", "binary_listNet", "pointwise_rmse", "bce",
              "ordinal", "with_ordinals")
METRIC_NAMES = ("ndcg", "dcg", "mrr")


def patch_allrank(patch_model=True):
    """Rebind allrank.models.{losses,metrics}.<name> (and make_model) to the B200 implementations.
    Returns the dict of original callables so the caller can restore them."""
    import allrank.models.losses as ref_losses
    import allrank.models.metrics as ref_metrics
    import allrank.models.model as ref_model
    saved = {}
    for name in LOSS_NAMES:
        saved["losses." + name] = getattr(ref_losses, name)
        setattr(ref_losses, name, getattr(_losses, name))
    for name in METRIC_NAMES:
        saved["metrics." + name] = getattr(ref_metrics, name)
        setattr(ref_metrics, name, getattr(_metrics, name))
    if patch_model:
        saved["model.make_model"] = ref_model.make_model
        ref_model.make_model = _model.make_model
        try:                                   # allrank/main.py does `from allrank.models.model import make_model`
            import allrank.main as ref_main
            saved["main.make_model"] = ref_main.make_model
            ref_main.make_model = _model.make_model
        except Exception:
            pass
    return saved


def unpatch_allrank(saved):
    import importlib
    for key, fn in saved.items():
        mod, name = key.split(".")
        target = importlib.import_module({"losses": "allrank.models.losses", "metrics": "allrank.models.metrics",
                                          "model": "allrank.models.model", "main": "allrank.main"}[mod])
        setattr(target, name, fn)
