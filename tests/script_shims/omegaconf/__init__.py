"""Stand-in for omegaconf 2.3 (absent here): OmegaConf.load -> DictConfig / ListConfig containers with the access patterns the scripts use.
Like the real ones, ListConfig is a Sequence but NOT a list, and a missing key raises (the shipped configs lack train.fp16)."""
from collections.abc import MutableMapping, Sequence

import re

import yaml


class _Loader(yaml.SafeLoader):
    """omegaconf's YAML loader resolves 1e-4 as a float (PyYAML's YAML 1.1 resolver wants 1.0e-4)"""


_Loader.add_implicit_resolver('tag:yaml.org,2002:float', re.compile(
    r'^(?:[-+]?(?:[0-9][0-9_]*)\.[0-9_]*(?:[eE][-+]?[0-9]+)?|[-+]?(?:[0-9][0-9_]*)(?:[eE][-+]?[0-9]+)|\.[0-9_]+(?:[eE][-+][0-9]+)?|[-+]?\.(?:inf|Inf|INF)|\.(?:nan|NaN|NAN))$', re.X),
    list('-+0123456789.'))


class ListConfig(Sequence):
    def __init__(self, items):
        self._items = [_wrap(v) for v in items]

    def __getitem__(self, i):
        return self._items[i]

    def __len__(self):
        return len(self._items)

    def __repr__(self):
        return repr(self._items)


class DictConfig(MutableMapping):
    def __init__(self, d):
        self._d = {k: _wrap(v) for k, v in d.items()}

    def __getitem__(self, k):
        if k not in self._d:
            raise KeyError(f"Missing key {k}")      # omegaconf.errors.ConfigKeyError is a KeyError
        return self._d[k]

    def __setitem__(self, k, v):
        self._d[k] = _wrap(v)

    def __delitem__(self, k):
        del self._d[k]

    def __iter__(self):
        return iter(self._d)

    def __len__(self):
        return len(self._d)

    def __getattr__(self, k):
        if k.startswith('_'):
            raise AttributeError(k)
        return self[k]

    def __repr__(self):
        return repr(self._d)


def _wrap(v):
    if isinstance(v, dict):
        return DictConfig(v)
    if isinstance(v, (list, tuple)):
        return ListConfig(v)
    return v


def _unwrap(v):
    if isinstance(v, DictConfig):
        return {k: _unwrap(x) for k, x in v.items()}
    if isinstance(v, ListConfig):
        return [_unwrap(x) for x in v]
    return v


class OmegaConf:
    @staticmethod
    def load(path):
        with open(path) as f:
            return DictConfig(yaml.load(f, Loader=_Loader))

    @staticmethod
    def to_container(cfg, resolve=True):
        return _unwrap(cfg)
