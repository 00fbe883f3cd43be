"""Stub for the reference's flatten_dict import (oracle tooling only)."""


def flatten(d, reducer="path"):
    out = {}

    def rec(prefix, v):
        if isinstance(v, dict) and v:
            for k, x in v.items():
                rec(prefix + [str(k)], x)
        else:
            out["/".join(prefix)] = v

    rec([], d)
    return out
