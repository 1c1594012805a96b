"""No-op stand-in for wandb (test infrastructure; see tests/shims/README.md)."""
import types
import uuid

util = types.SimpleNamespace(generate_id=lambda: uuid.uuid4().hex[:8])
logged = []


class Image:
    def __init__(self, data, caption=None):
        self.data, self.caption = data, caption


def init(*a, **k):
    return None


def log(values, step=None, **k):
    logged.append((values, step))


def finish(*a, **k):
    return None
