"""Host-side logic of the scorer module that needs no GPU: state_dict surface, seeded init order,
error conventions, and that the C-ABI library loads and exports every symbol include/allrank_b200.h declares."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make(n_features=20, d=32, N=2, h=2, dff=64, act=None, dropout=0.0):
    from allrank_b200.model import make_model
    return make_model(fc_model={"sizes": [d], "input_norm": False, "activation": None, "dropout": 0.0},
                      transformer={"N": N, "d_ff": dff, "h": h, "positional_encoding": None, "dropout": dropout},
                      post_model={"d_output": 1, "output_activation": act}, n_features=n_features)


def test_state_dict_keys_and_shapes_equal_reference(golden):
    g = golden("init_seed123")
    ref = {k[2:]: g[k] for k in g.files}
    sd = make().state_dict()
    assert list(sd.keys()) == list(ref.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == ref[k].shape, k


def test_seeded_init_reproduces_reference(golden):
    g = golden("init_seed123")
    torch.manual_seed(123)
    sd = make().state_dict()
    for k in g.files:
        assert np.array_equal(sd[k[2:]].numpy(), g[k]), k


def test_reference_state_dict_loads(golden):
    g = golden("scorer_tiny")
    F, d, N, h, dff, B, S = [int(v) for v in g["meta"]]
    model = make(F, d, N, h, dff)
    sd = {k[2:]: torch.tensor(g[k]) for k in g.files if k.startswith("p:")}
    missing = model.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys


def test_multi_output_head_has_the_reference_state_dict_layout(golden):
    """post_model.d_output = n > 1 (the ordinal configuration): same keys and shapes as the reference's module tree."""
    from allrank_b200.model import make_model
    g = golden("scorer_dout4")
    F, d, N, h, dff, B, S, n_out = [int(v) for v in g["meta"]]
    model = make_model(fc_model={"sizes": [d], "input_norm": False, "activation": None, "dropout": 0.0},
                       transformer={"N": N, "d_ff": dff, "h": h, "positional_encoding": None, "dropout": 0.0},
                       post_model={"d_output": n_out, "output_activation": "Sigmoid"}, n_features=F)
    sd = {k[2:]: torch.tensor(g[k]) for k in g.files if k.startswith("p:")}
    assert {k: tuple(v.shape) for k, v in model.state_dict().items()} == {k: tuple(v.shape) for k, v in sd.items()}
    assert tuple(model.output_layer.w_1.weight.shape) == (n_out, d) and model.output_layer.d_output == n_out
    model.load_state_dict(sd, strict=True)


def test_unsupported_configs_raise_not_fallback():
    from allrank_b200.model import make_model
    post = {"d_output": 1, "output_activation": None}
    with pytest.raises(NotImplementedError):      # no input FC block at all
        make_model(fc_model=None, transformer=None, post_model=post, n_features=20)
    with pytest.raises(NotImplementedError):      # an activation class the kernels do not implement
        make_model(fc_model={"sizes": [32, 32], "input_norm": False, "activation": "GELU", "dropout": 0.0},
                   transformer=None, post_model=post, n_features=20)
    with pytest.raises(NotImplementedError):      # input_norm over a feature count that needs padding
        make_model(fc_model={"sizes": [32], "input_norm": True, "activation": None, "dropout": 0.0},
                   transformer=None, post_model=post, n_features=21)
    with pytest.raises(NotImplementedError):
        make_model(fc_model={"sizes": [8] * 9, "input_norm": False, "activation": None, "dropout": 0.0},
                   transformer=None, post_model=post, n_features=20)
    m = make()
    with pytest.raises(Exception):   # CPU tensors: no eager fallback
        m(torch.zeros(1, 4, 20), torch.zeros(1, 4, dtype=torch.bool), None)


def test_cabi_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "allrank_b200.h")).read()
    declared = set(re.findall(r"\b(arb_[a-z0-9_]+)\s*\(", header))
    declared -= {"arb_scorer_config"}
    assert len(declared) >= 16
    lib = ctypes.CDLL(os.path.join(ROOT, "allrank_b200", "liballrank_b200.so"))
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} is declared in include/allrank_b200.h but not exported"
    lib.arb_abi_version.restype = ctypes.c_int32
    assert lib.arb_abi_version() == 4


def test_param_count_matches_reference_models():
    from allrank_b200 import _lib
    from allrank_b200.model import ScorerConfig
    import allrank_b200.model  # noqa: F401  (registers signatures)
    # parameter counts the survey measured for the two BASELINE shapes
    for (F, d, N, h, dff, expect) in [(136, 128, 2, 4, 512, 414465), (136, 256, 4, 8, 1024, 3194881)]:
        cfg = ScorerConfig(F, d, N, h, dff, 0, 1e-6, 0.0, 0.0, 0, 0)
        assert _lib.lib().arb_scorer_param_count(ctypes.byref(cfg)) == expect


@pytest.mark.parametrize("strategy", ["fixed", "learned"])
def test_positional_encoding_state_dict_surface(golden, strategy):
    from allrank_b200.model import make_model
    g = golden("scorer_pe_" + strategy)
    F, d, N, h, dff, B, S, max_idx = [int(v) for v in g["meta"]]
    model = make_model(fc_model={"sizes": [d], "input_norm": False, "activation": None, "dropout": 0.0},
                       transformer={"N": N, "d_ff": dff, "h": h, "dropout": 0.0,
                                    "positional_encoding": {"strategy": strategy, "max_indices": max_idx}},
                       post_model={"d_output": 1, "output_activation": None}, n_features=F)
    ref = {k[2:]: g[k] for k in g.files if k.startswith("p:")}
    sd = model.state_dict()
    assert list(sd.keys()) == list(ref.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == ref[k].shape, k
    if strategy == "fixed":
        assert np.array_equal(sd["encoder.position.pe"].numpy(), ref["encoder.position.pe"])
    with pytest.raises(ValueError):
        make_model(fc_model={"sizes": [d], "input_norm": False, "activation": None, "dropout": 0.0},
                   transformer={"N": 1, "d_ff": 8, "h": 1, "dropout": 0.0,
                                "positional_encoding": {"strategy": "rotary", "max_indices": 4}},
                   post_model={"d_output": 1, "output_activation": None}, n_features=F)


def test_slate_movers_have_no_cpu_fallback():
    """The callers either side of the path are CUDA-only like the path itself: CPU use raises, nothing falls back."""
    import numpy as np
    from allrank_b200 import _lib, data, inference, losses
    with pytest.raises(_lib.ArbError):
        data.SlateStore(np.zeros((3, 4), dtype=np.float32), np.zeros(3, dtype=np.float32), np.zeros(3), device="cpu")
    with pytest.raises(Exception):
        inference.reorder_slates(torch.zeros(1, 2, 4), torch.zeros(1, 2), torch.zeros(1, 2, dtype=torch.int32))
    with pytest.raises(Exception):
        losses.ordinal(torch.full((1, 2, 3), 0.5), torch.zeros(1, 2), n=3)
    with pytest.raises(Exception):
        losses.with_ordinals(torch.zeros(1, 2), 3)
    with pytest.raises(ValueError):
        data.DeviceSlateLoader(object(), batch_size=0, slate_length=4)


def test_header_is_plain_c(tmp_path):
    """include/allrank_b200.h is the drop-in boundary: it must compile as C99 (and C++) on its own and link against
    the library -- no torch types, no C++-isms."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "t.c"
    src.write_text('#include "allrank_b200.h"\n'
                   'int main(void) { arb_scorer_config c; (void)c; return arb_abi_version() == 4 ? 0 : 1; }\n')
    inc = os.path.join(root, "include")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, "-c", str(src), "-o",
                    str(tmp_path / "t.o")], check=True)
    if shutil.which("g++") is not None:
        subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-I", inc, "-x", "c++", "-c", str(src), "-o",
                        str(tmp_path / "t2.o")], check=True)


def test_graphed_train_step_refuses_cpu_tensors():
    """allrank_b200 has no CPU path: the CUDA-graph helper checks its inputs before touching the device."""
    from allrank_b200.graph import GraphedTrainStep
    with pytest.raises(ValueError):
        GraphedTrainStep(object(), None, None, torch.zeros(2, 4, 8), torch.zeros(2, 4))
