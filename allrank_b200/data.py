"""Device-side replacement for allRank's slate pipeline (allrank/data/dataset_loading.py).

The reference keeps the corpus as per-query numpy arrays and builds every slate in Python
(`LibSVMDataset.__getitem__` -> `FixLength` -> `ToTensor`, then DataLoader collation; dataset_loading.py:32-165,
:230-247).  Here the corpus is uploaded once and stays in HBM as query-grouped rows with CSR offsets (MSLR-WEB30K's
train fold is 2.27 M x 136 fp32 = 1.2 GB of the 180 GB), and one kernel launch (csrc/slates.cu,
`arb_assemble_slates`) builds a whole `(x, y, indices)` batch: padding is bit-identical to the reference, sampling
follows the same rules with a counter-hash stream instead of numpy's.

Same call surface as the reference module:

    train_ds, val_ds = load_libsvm_dataset(input_path, slate_length, validation_ds_role)    # :206-217
    n_features = train_ds.shape[-1]                                                         # main.py:63
    train_dl, val_dl = create_data_loaders(train_ds, val_ds, num_workers, batch_size)       # :230-247
    for xb, yb, indices in train_dl: ...                                                    # train_utils.py:93-96

No CPU fallback: the loaders need a CUDA device.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib

PADDED_Y_VALUE = -1        # dataset_loading.py:15
PADDED_INDEX_VALUE = -1    # dataset_loading.py:16

c_p, c_i = ctypes.c_void_p, ctypes.c_int32
_lib.register("arb_assemble_slates_smem_bytes", ctypes.c_size_t, [c_i, c_i])
_lib.register("arb_assemble_slates", c_i, [c_p, c_p, c_p, ctypes.c_int64, c_p, c_i, c_i, c_i, c_i, ctypes.c_uint64, c_p,
                                           c_p, c_p, c_p])


def group_offsets(query_ids):
    """CSR offsets of the query groups exactly as LibSVMDataset splits them (dataset_loading.py:106-111): groups are
    taken in order of first appearance and are `count` consecutive rows long."""
    query_ids = np.asarray(query_ids)
    _, first, counts = np.unique(query_ids, return_index=True, return_counts=True)
    ordered = counts[np.argsort(first)]
    return np.concatenate([[0], np.cumsum(ordered)]).astype(np.int64)


class SlateStore:
    """Query-grouped corpus resident on one CUDA device; the counterpart of LibSVMDataset (dataset_loading.py:96-165)."""

    def __init__(self, X, y, query_ids, device=None, slate_length=None):
        if hasattr(X, "toarray"):
            X = X.toarray()                                   # :104 (the reference densifies too)
        X = np.ascontiguousarray(np.asarray(X, dtype=np.float32))
        y = np.ascontiguousarray(np.asarray(y, dtype=np.float32))
        offsets = group_offsets(query_ids)
        if offsets[-1] != X.shape[0] or y.shape[0] != X.shape[0]:
            raise ValueError("X, y and query_ids must have one row per document")
        self._init_from_groups(X, y, offsets, device, slate_length)

    def _init_from_groups(self, X, y, offsets, device, slate_length):
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if device.type != "cuda":
            raise _lib.ArbError("SlateStore needs a CUDA device (allrank_b200 has no CPU fallback)")
        self.device = device
        self.offsets_host = np.asarray(offsets, dtype=np.int64)
        lengths = np.diff(self.offsets_host)
        self.n_queries = int(lengths.shape[0])
        self.n_features = int(X.shape[1])
        self.longest_query_length = int(lengths.max()) if self.n_queries else 0     # :112
        self.docs_x = torch.from_numpy(X).to(device)
        self.docs_y = torch.from_numpy(y).to(device)
        self.offsets = torch.from_numpy(self.offsets_host).to(device)
        self.slate_length = slate_length

    @classmethod
    def from_groups(cls, X_by_qid, y_by_qid, device=None, slate_length=None):
        """From per-query arrays (e.g. an allrank LibSVMDataset's X_by_qid / y_by_qid)."""
        self = cls.__new__(cls)
        lengths = [len(v) for v in y_by_qid]
        offsets = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
        X = np.ascontiguousarray(np.concatenate([np.asarray(v, dtype=np.float32) for v in X_by_qid], axis=0))
        y = np.ascontiguousarray(np.concatenate([np.asarray(v, dtype=np.float32) for v in y_by_qid], axis=0))
        self._init_from_groups(X, y, offsets, device, slate_length)
        return self

    @classmethod
    def from_svm_file(cls, svm_file_path, device=None, slate_length=None):
        """dataset_loading.py:118-130 (same sklearn parser as the reference; file or file-like object)."""
        from sklearn.datasets import load_svmlight_file
        x, y, query_ids = load_svmlight_file(svm_file_path, query_id=True)
        return cls(x, y, query_ids, device=device, slate_length=slate_length)

    def __len__(self):
        return self.n_queries

    @property
    def shape(self):
        """[queries, longest query, features] -- dataset_loading.py:149-165."""
        return [self.n_queries, self.longest_query_length, self.n_features]

    def assemble(self, queries, slate_length=None, seed=0):
        """One batch of slates for the given query numbers: (x [B,S,F] fp32, y [B,S] fp32, indices [B,S] int64)."""
        S = int(slate_length if slate_length is not None else self.slate_length)
        if S <= 0:
            raise ValueError("slate_length must be positive")
        queries = torch.as_tensor(queries, dtype=torch.int64)
        if queries.device.type == "cpu" and queries.numel() and (
                int(queries.min()) < 0 or int(queries.max()) >= self.n_queries):
            raise IndexError("query number out of range")     # (device tensors are range-checked by the kernel)
        queries = queries.to(self.device).contiguous()
        B, F = int(queries.numel()), self.n_features
        x = torch.empty((B, S, F), dtype=torch.float32, device=self.device)
        y = torch.empty((B, S), dtype=torch.float32, device=self.device)
        idx = torch.empty((B, S), dtype=torch.int64, device=self.device)
        if B == 0:
            return x, y, idx
        with torch.cuda.device(self.device):
            rc = _lib.lib().arb_assemble_slates(_lib.ptr(self.docs_x), _lib.ptr(self.docs_y), _lib.ptr(self.offsets),
                                                self.n_queries, _lib.ptr(queries), B, S, F,
                                                self.longest_query_length,
                                                ctypes.c_uint64(int(seed) & (2 ** 64 - 1)), _lib.ptr(x), _lib.ptr(y),
                                                _lib.ptr(idx), _lib.stream_ptr(self.device))
        _lib.check(rc, "arb_assemble_slates")
        return x, y, idx


class DeviceSlateLoader:
    """Iterates (xb, yb, indices) batches like the reference's DataLoader over a LibSVMDataset (:242-245), built on
    the device.  Shuffling and the per-batch sampling seed come from torch's global CPU generator, so
    torch.manual_seed makes an epoch repeatable."""

    def __init__(self, store, batch_size, slate_length=None, shuffle=False):
        self.dataset = store
        self.batch_size = int(batch_size)
        self.slate_length = int(slate_length if slate_length is not None else store.slate_length)
        self.shuffle = bool(shuffle)
        if self.batch_size <= 0:
            raise ValueError("batch_size must be positive")

    def __len__(self):
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        n = len(self.dataset)
        order = torch.randperm(n) if self.shuffle else torch.arange(n)
        order = order.to(self.dataset.device)
        for start in range(0, n, self.batch_size):
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
            yield self.dataset.assemble(order[start:start + self.batch_size], self.slate_length, seed)


def load_libsvm_role(input_path, role, device=None):
    """dataset_loading.py:168-182 (local files only: the gs:// path is the reference's I/O layer, out of scope)."""
    path = os.path.join(input_path, "{}.txt".format(role))
    with open(path, "rb") as stream:
        return SlateStore.from_svm_file(stream, device=device)


def load_libsvm_dataset_role(role, input_path, slate_length, device=None):
    """Train slates are fixed to `slate_length`, every other role to its longest query (:219-227)."""
    ds = load_libsvm_role(input_path, role, device=device)
    ds.slate_length = int(slate_length) if role == "train" else int(ds.longest_query_length)
    return ds


def load_libsvm_dataset(input_path, slate_length, validation_ds_role, device=None):
    """dataset_loading.py:206-217."""
    train_ds = load_libsvm_dataset_role("train", input_path, slate_length, device=device)
    val_ds = load_libsvm_dataset_role(validation_ds_role, input_path, slate_length, device=device)
    return train_ds, val_ds


def _as_store(ds, device=None):
    if isinstance(ds, SlateStore):
        return ds
    # an allrank LibSVMDataset: per-query arrays + Compose([FixLength(dim_given), ToTensor()])
    slate_length = None
    transform = getattr(ds, "transform", None)
    for t in getattr(transform, "transforms", []) or []:
        if hasattr(t, "dim_given"):
            slate_length = int(t.dim_given)
    return SlateStore.from_groups(ds.X_by_qid, ds.y_by_qid, device=device, slate_length=slate_length)


def create_data_loaders(train_ds, val_ds, num_workers, batch_size, device=None):
    """dataset_loading.py:230-247.  Accepts SlateStores or the reference's LibSVMDataset objects (uploaded once).
    `num_workers` is accepted for signature parity and unused (there are no loader processes).  One process drives
    one GPU here, so `batch_size` is the per-process batch (the reference multiplies it by the visible GPU count for
    nn.DataParallel, :237-238)."""
    train_store, val_store = _as_store(train_ds, device), _as_store(val_ds, device)
    train_dl = DeviceSlateLoader(train_store, batch_size, shuffle=True)
    val_dl = DeviceSlateLoader(val_store, batch_size, shuffle=False)
    return train_dl, val_dl


__all__ = ["SlateStore", "DeviceSlateLoader", "group_offsets", "load_libsvm_role", "load_libsvm_dataset_role",
           "load_libsvm_dataset", "create_data_loaders", "PADDED_Y_VALUE", "PADDED_INDEX_VALUE"]
