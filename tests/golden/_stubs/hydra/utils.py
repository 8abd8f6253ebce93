import importlib


def get_class(path):
    mod, name = path.rsplit(".", 1)
    return getattr(importlib.import_module(mod), name)


def instantiate(cfg, *args, **kwargs):
    cfg = dict(cfg)
    target = cfg.pop("_target_")
    cfg.pop("_recursive_", None)
    cfg.pop("_convert_", None)
    cfg.update(kwargs)
    return get_class(target)(*args, **cfg)
