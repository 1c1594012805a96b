"""No-op stand-in for wandb (test infrastructure; see tests/shims/README.md)."""
import json
import os
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
    path = os.environ.get("MUSE_SHIM_WANDB_LOG")  # tests read what the script logged after its modules are gone
    if path:
        with open(path, "a") as f:
            f.write(json.dumps({"step": step, "keys": {k_: (len(v) if isinstance(v, (list, tuple)) else 1) for k_, v in values.items()},
                                "captions": [getattr(x, "caption", None) for v in values.values() if isinstance(v, (list, tuple)) for x in v],
                                "sizes": [list(getattr(getattr(x, "data", None), "size", ())) for v in values.values()
                                          if isinstance(v, (list, tuple)) for x in v]}) + "\n")


def finish(*a, **k):
    return None
