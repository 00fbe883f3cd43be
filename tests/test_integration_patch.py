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
        assert (list(inspect.signature(model.make_model).parameters) ==
                list(inspect.signature(ref_model.make_model).parameters))
        originals = {n: getattr(ref_losses, n) for n in integration.LOSS_NAMES}
        saved = integration.patch_allrank()
        try:
            assert ref_losses.lambdaLoss is losses.lambdaLoss
            assert ref_metrics.ndcg is metrics.ndcg
            assert ref_model.make_model is model.make_model
        finally:
            integration.unpatch_allrank(saved)
        for n, fn in originals.items():
            assert getattr(ref_losses, n) is fn
    finally:
        for p in (os.path.join(ROOT, "oracle", "_stubs"), REF):
            if p in sys.path:
                sys.path.remove(p)
