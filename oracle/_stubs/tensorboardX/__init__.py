"""Stub for the reference's tensorboardX import (oracle tooling only)."""


class SummaryWriter:
    def __init__(self, *a, **k):
        pass

    def add_scalar(self, *a, **k):
        pass

    def close(self):
        pass
