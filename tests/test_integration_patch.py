"""patch_allrank() against the real reference package (only where /root/reference exists: the build container).
Checks the rebinding mechanics INTEGRATION.md describes; the kernels themselves are covered by the GPU tests."""
import os
import sys

import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "allrank")), reason="reference tree not present on this box")
def test_patch_and_unpatch_rebinds_the_reference_names():
    sys.path[:0] = [os.path.join(ROOT, "oracle", "_stubs"), REF]
    try:
        import allrank.models.losses as ref_losses
        import allrank.models.metrics as ref_metrics
        import allrank.models.model as ref_model
        import inspect
        from allrank_b200 import integration, losses, metrics, model
        # same call surface: every patched callable accepts the reference's parameter names in the same order
        for name in integration.LOSS_NAMES:
            ref_params = list(inspect.signature(getattr(ref_losses, name)).parameters)
            mine = list(inspect.signature(getattr(losses, name)).parameters)
            assert mine[:len(ref_params)] == ref_params, (name, ref_params, mine)
        for name in integration.METRIC_NAMES:
            ref_params = list(inspect.signature(getattr(ref_metrics, name)).parameters)
            assert list(inspect.signature(getattr(metrics, name)).parameters) == ref_params, name
        ref_params = list(inspect.signature(ref_model.make_model).parameters)
        mine = list(inspect.signature(model.make_model).parameters)
        assert mine[:len(ref_params)] == ref_params and mine[len(ref_params):] == ["compute_dtype"]   # one extension
        # the callers either side of the path: same parameter names as the reference functions they replace
        import allrank.training.train_utils as ref_tu
        import allrank.inference.inference_utils as ref_iu
        import allrank.data.dataset_loading as ref_dl
        from allrank_b200 import data, inference, training
        for ref_mod, mine_mod, names in ((ref_tu, training, ("metric_on_batch", "metric_on_epoch", "compute_metrics")),
                                         (ref_iu, inference, ("rank_slates",)),
                                         (ref_dl, data, ("load_libsvm_dataset", "load_libsvm_dataset_role",
                                                         "load_libsvm_role", "create_data_loaders"))):
            for name in names:
                ref_params = list(inspect.signature(getattr(ref_mod, name)).parameters)
                mine = list(inspect.signature(getattr(mine_mod, name)).parameters)
                assert mine[:len(ref_params)] == ref_params, (name, ref_params, mine)
        assert data.PADDED_Y_VALUE == ref_dl.PADDED_Y_VALUE and data.PADDED_INDEX_VALUE == ref_dl.PADDED_INDEX_VALUE
        originals = {n: getattr(ref_losses, n) for n in integration.LOSS_NAMES}
        orig_cm, orig_loader = ref_tu.compute_metrics, ref_dl.create_data_loaders
        saved = integration.patch_allrank(patch_data=True)
        try:
            assert ref_losses.lambdaLoss is losses.lambdaLoss
            assert ref_losses.ordinal is losses.ordinal
            assert ref_metrics.ndcg is metrics.ndcg
            assert ref_model.make_model is model.make_model
            assert ref_tu.compute_metrics is training.compute_metrics
            assert ref_iu.rank_slates is inference.rank_slates
            assert ref_dl.create_data_loaders is data.create_data_loaders
            import allrank.main as ref_main
            assert ref_main.load_libsvm_dataset is data.load_libsvm_dataset
        finally:
            integration.unpatch_allrank(saved)
        for n, fn in originals.items():
            assert getattr(ref_losses, n) is fn
        assert ref_tu.compute_metrics is orig_cm and ref_dl.create_data_loaders is orig_loader
    finally:
        for p in (os.path.join(ROOT, "oracle", "_stubs"), REF):
            if p in sys.path:
                sys.path.remove(p)
