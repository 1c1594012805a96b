"""Minimal stand-in for omegaconf (test infrastructure; see tests/shims/README.md): nested attribute-access configs from a
YAML file merged with ``key.sub=value`` command-line overrides."""
import re
import sys

import yaml


_FLOAT = re.compile(r"^[+-]?(\d+\.?\d*|\.\d+)([eE][+-]?\d+)$")


def _wrap(v):
    if isinstance(v, str) and _FLOAT.match(v):  # YAML 1.1 reads "1e-4" as a string; OmegaConf reads a float
        return float(v)
    if isinstance(v, dict):
        return DictConfig(v)
    if isinstance(v, (list, tuple)):
        return ListConfig(v)
    return v


def _unwrap(v):
    if isinstance(v, DictConfig):
        return {k: _unwrap(x) for k, x in v._d.items()}
    if isinstance(v, ListConfig):
        return [_unwrap(x) for x in v._l]
    return v


class DictConfig:
    def __init__(self, d=None):
        object.__setattr__(self, "_d", {k: _wrap(v) for k, v in (d or {}).items()})

    def __getattr__(self, k):
        try:
            return self._d[k]
        except KeyError:
            raise AttributeError(f"Missing key {k}") from None

    def __setattr__(self, k, v):
        self._d[k] = _wrap(v)

    __setitem__ = __setattr__

    def __getitem__(self, k):
        return self._d[k]

    def __contains__(self, k):
        return k in self._d

    def __iter__(self):
        return iter(self._d)

    def __len__(self):
        return len(self._d)

    def keys(self):
        return self._d.keys()

    def values(self):
        return self._d.values()

    def items(self):
        return self._d.items()

    def items_ex(self, resolve=True):
        return self._d.items()

    def get(self, k, default=None):
        v = self._d.get(k, default)
        return default if v is None else v

    def pop(self, k, *a):
        return self._d.pop(k, *a)

    def __repr__(self):
        return repr(_unwrap(self))


class ListConfig:
    def __init__(self, l=()):
        self._l = [_wrap(v) for v in l]

    def __getitem__(self, i):
        return self._l[i]

    def __iter__(self):
        return iter(self._l)

    def __len__(self):
        return len(self._l)

    def _iter_ex(self, resolve=True):
        return iter(self._l)

    def __repr__(self):
        return repr(_unwrap(self))


def _parse_scalar(s):
    try:
        return yaml.safe_load(s)
    except Exception:
        return s


class OmegaConf:
    @staticmethod
    def create(d=None):
        return _wrap(d or {})

    @staticmethod
    def from_cli(args_list=None):
        args = sys.argv[1:] if args_list is None else args_list
        root = {}
        for a in args:
            if "=" not in a:
                continue
            key, val = a.split("=", 1)
            cur = root
            parts = key.split(".")
            for p in parts[:-1]:
                cur = cur.setdefault(p, {})
            cur[parts[-1]] = _parse_scalar(val)
        return DictConfig(root)

    @staticmethod
    def load(path):
        with open(str(path)) as f:
            return _wrap(yaml.safe_load(f))

    @staticmethod
    def merge(*cfgs):
        def rec(a, b):
            for k, v in b.items():
                if isinstance(v, dict) and isinstance(a.get(k), dict):
                    rec(a[k], v)
                else:
                    a[k] = v
            return a

        out = {}
        for c in cfgs:
            rec(out, _unwrap(c))
        return DictConfig(out)

    @staticmethod
    def save(config, f):
        with open(str(f), "w") as fh:
            yaml.safe_dump(_unwrap(config), fh)

    @staticmethod
    def to_container(cfg, resolve=True):
        return _unwrap(cfg)
