"""Minimal recursive `_target_` resolver standing in for hydra.utils.instantiate.

Test infrastructure only.  Follows hydra's documented behaviour for the subset the
reference's YAMLs use: nested dicts with `_target_` are instantiated depth-first and
passed as keyword arguments; lists are converted element-wise."""
import importlib


def _resolve(path):
    mod, _, name = path.rpartition(".")
    return getattr(importlib.import_module(mod), name)


def instantiate(cfg, **overrides):
    if isinstance(cfg, dict):
        kwargs = {k: instantiate(v) for k, v in cfg.items() if k != "_target_"}
        if "_target_" in cfg:
            kwargs.update(overrides)
            return _resolve(cfg["_target_"])(**kwargs)
        return kwargs
    if isinstance(cfg, (list, tuple)):
        return [instantiate(v) for v in cfg]
    return cfg
