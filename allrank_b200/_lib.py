"""ctypes binding of liballrank_b200.so (the C ABI in include/allrank_b200.h).

The product path has NO fallback: if the library is missing or a call fails, we raise.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liballrank_b200.so")

c_f = ctypes.c_float
c_i = ctypes.c_int32
c_p = ctypes.c_void_p
c_sz = ctypes.c_size_t

_SIGS = {
    "arb_last_error": (ctypes.c_char_p, []),
    "arb_abi_version": (c_i, []),
    "arb_launch_count": (ctypes.c_int64, []),
    "arb_rank_metrics": (c_i, [c_p, c_p, c_i, c_i, c_p, c_p, c_p, c_i, c_i, c_f, c_f, c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
    "arb_listnet": (c_i, [c_p, c_p, c_i, c_i, c_f, c_f, c_p, c_p, c_p, c_p]),
    "arb_listmle": (c_i, [c_p, c_p, c_i, c_i, c_f, c_f, c_p, c_p, c_p, c_p, c_p, c_p]),
    "arb_approx_ndcg": (c_i, [c_p, c_p, c_i, c_i, c_f, c_f, c_f, c_p, c_p, c_p, c_p]),
    "arb_lambda_loss": (c_i, [c_p, c_p, c_i, c_i, c_f, c_f, c_i, c_i, c_f, c_f, c_i, c_i, c_p, c_p, c_p, c_p]),
    "arb_ranknet": (c_i, [c_p, c_p, c_i, c_i, c_f, c_i, c_p, c_p, c_p, c_p]),
    "arb_pointwise_loss": (c_i, [c_p, c_p, c_i, c_i, c_f, c_i, c_f, c_f, c_p, c_p, c_p, c_p]),
    "arb_ordinal": (c_i, [c_p, c_p, c_i, c_i, c_i, c_f, c_p, c_p, c_p, c_p]),
    "arb_neural_ndcg_workspace_bytes": (c_sz, [c_i, c_i, c_i]),
    "arb_neural_ndcg": (c_i, [c_p, c_p, c_i, c_i, c_p, c_f, c_f, c_i, c_i, c_i, c_f, c_p, c_p, c_p, c_p, c_sz, c_p]),
}

_lib = None
_pack_default = 0
_relu_bits_default = 0


class ArbError(RuntimeError):
    pass


def lib():
    """Load the library once.  Raises (never falls back) when it is missing."""
    global _lib, _pack_default, _relu_bits_default
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ArbError(
                f"{LIB_PATH} is missing: build it with `python -m allrank_b200.build` "
                "(allrank_b200 has no CPU/eager fallback)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
        # A/B switches for measurements (include/allrank_b200.h); the defaults are the measured-fastest settings
        if os.environ.get("ARB_GEMM_PERSISTENT") in ("0", "1", "2", "3"):
            handle.arb_set_gemm_persistent(int(os.environ["ARB_GEMM_PERSISTENT"]))
        if os.environ.get("ARB_PDL") in ("0", "1"):
            handle.arb_set_pdl(int(os.environ["ARB_PDL"]))
        if os.environ.get("ARB_ATTN_SKIP_PADDING") in ("0", "1"):
            handle.arb_set_attention_skip_padding(int(os.environ["ARB_ATTN_SKIP_PADDING"]))
        if os.environ.get("ARB_ATTN_BWD_PERSISTENT") in ("0", "1"):
            handle.arb_set_attention_bwd_persistent(int(os.environ["ARB_ATTN_BWD_PERSISTENT"]))
        if os.environ.get("ARB_RELU_BITS") in ("0", "1"):
            handle.arb_set_relu_bits(int(os.environ["ARB_RELU_BITS"]))
        if os.environ.get("ARB_PACK_ROWS") in ("0", "1"):
            handle.arb_set_pack_rows(int(os.environ["ARB_PACK_ROWS"]))
        _pack_default = int(handle.arb_get_pack_rows())
        _relu_bits_default = int(handle.arb_get_relu_bits())
    return _lib


def default_pack_rows():
    """The packed-rows setting this process started with (library default or ARB_PACK_ROWS): what tests restore."""
    lib()
    return _pack_default


def default_relu_bits():
    """The ReLU-bit-mask setting this process started with (library default or ARB_RELU_BITS)."""
    lib()
    return _relu_bits_default


def register(name, restype, argtypes):
    _SIGS[name] = (restype, argtypes)
    if _lib is not None:
        fn = getattr(_lib, name)
        fn.restype = restype
        fn.argtypes = argtypes


def check(rc, what):
    if rc != 0:
        msg = lib().arb_last_error().decode()
        if rc == -1:
            raise ValueError(f"{what}: {msg}")
        raise ArbError(f"{what} failed (code {rc}): {msg}")


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise ArbError("allrank_b200 kernels need CUDA tensors (there is no CPU fallback)")


def launch_count():
    return int(lib().arb_launch_count())
