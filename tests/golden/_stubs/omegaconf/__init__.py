"""Minimal stand-in for omegaconf, used ONLY by tests/golden/make_golden.py to import the
reference modules in the build container (omegaconf itself is not installed). Test infrastructure."""
import copy


class DictConfig(dict):
    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        for k, v in list(self.items()):
            self[k] = _wrap(v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def __setitem__(self, k, v):
        super().__setitem__(k, _wrap(v))

    def __deepcopy__(self, memo):
        return DictConfig({k: copy.deepcopy(v, memo) for k, v in self.items()})


class ListConfig(list):
    pass


def _wrap(v):
    if isinstance(v, dict) and not isinstance(v, DictConfig):
        return DictConfig(v)
    return v
