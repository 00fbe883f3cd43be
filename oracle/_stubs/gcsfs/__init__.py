"""Stub so the read-only allRank reference can be imported in the build container
(only used by oracle/make_golden.py; never on the GPU box, never by the product)."""


class GCSFileSystem:
    def open(self, *a, **k):
        raise RuntimeError("gcsfs stub: gs:// paths are not available")
