"""Plug allrank_b200 into an unmodified allRank checkout.

allRank resolves its hot-path callables by NAME at run time:
    loss   : getattr(allrank.models.losses, config.loss.name)        allrank/main.py:83
    metric : getattr(allrank.models.metrics, name)                   allrank/training/train_utils.py:50
    model  : allrank.models.model.make_model(...)                    allrank/main.py:75
so `patch_allrank()` only has to rebind those attributes; `allrank/main.py` and `train_utils.fit` then run
unchanged on the B200 kernels.  (allRank must be importable; nothing here imports it otherwise.)

The callers either side of the path are rebound the same way:
    epoch metrics : allrank.training.train_utils.{metric_on_epoch, compute_metrics}    train_utils.py:37-56
    rank_slates   : allrank.inference.inference_utils.rank_slates                      inference_utils.py:12-30
    slate loading : allrank.data.dataset_loading.{load_libsvm_dataset, create_data_loaders} (+ the names main.py
                    imported from it, main.py:8) -- opt-in (`patch_data=True`): the corpus then lives in HBM.
"""
from . import data as _data
from . import inference as _inference
from . import losses as _losses
from . import metrics as _metrics
from . import model as _model
from . import training as _training

LOSS_NAMES = ("listNet", "listMLE", "approxNDCGLoss", "lambdaLoss", "neuralNDCG", "neuralNDCG_transposed", "rankNet",
              "rankNet_weightByGTDiff", "rankNet_weightByGTDiff_pow", "binary_listNet", "pointwise_rmse", "bce",
              "ordinal", "with_ordinals")
METRIC_NAMES = ("ndcg", "dcg", "mrr")


_MODULES = {"losses": "allrank.models.losses", "metrics": "allrank.models.metrics", "model": "allrank.models.model",
            "main": "allrank.main", "train_utils": "allrank.training.train_utils",
            "inference_utils": "allrank.inference.inference_utils", "dataset_loading": "allrank.data.dataset_loading"}


def single_process_model(model):
    """Stands in for allrank.models.model_utils.CustomDataParallel when allrank/main.py:76-78 runs on a box with
    several visible GPUs.  The B200 path scales as ONE PROCESS PER GPU (torchrun + allrank_b200.ddp.FlatDDP, a single
    NCCL all-reduce of the flat gradient); single-process nn.DataParallel would replicate a model whose parameters
    are views of one flat device buffer, which cannot work.  The model is returned unwrapped and trains on the device
    allRank selects (cuda:0 of the visible devices): launch one process per GPU with CUDA_VISIBLE_DEVICES / torchrun."""
    return model


def _rebind(saved, mod_key, name, fn):
    import importlib
    target = importlib.import_module(_MODULES[mod_key])
    saved[mod_key + "." + name] = getattr(target, name)
    setattr(target, name, fn)


def patch_allrank(patch_model=True, patch_eval=True, patch_data=False):
    """Rebind allrank.models.{losses,metrics}.<name> (and make_model, the epoch-metric helpers, rank_slates and --
    opt-in -- the slate loaders) to the B200 implementations.
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
            # main.py:76-78 wraps the model in CustomDataParallel when device_count() > 1: neutralised (see above)
            saved["main.CustomDataParallel"] = ref_main.CustomDataParallel
            ref_main.CustomDataParallel = single_process_model
        except Exception:
            pass
    if patch_eval:
        _rebind(saved, "train_utils", "metric_on_epoch", _training.metric_on_epoch)
        _rebind(saved, "train_utils", "compute_metrics", _training.compute_metrics)
        _rebind(saved, "inference_utils", "rank_slates", _inference.rank_slates)
    if patch_data:
        for name in ("load_libsvm_dataset", "load_libsvm_dataset_role", "load_libsvm_role", "create_data_loaders"):
            _rebind(saved, "dataset_loading", name, getattr(_data, name))
        try:
            _rebind(saved, "main", "load_libsvm_dataset", _data.load_libsvm_dataset)
            _rebind(saved, "main", "create_data_loaders", _data.create_data_loaders)
        except Exception:
            pass
    return saved


def unpatch_allrank(saved):
    import importlib
    for key, fn in saved.items():
        mod, name = key.split(".")
        setattr(importlib.import_module(_MODULES[mod]), name, fn)
