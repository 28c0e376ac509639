"""Minimal stand-in for the Hydra/OmegaConf ``DictConfig`` the reference passes to ``Net``
(/root/reference/config/images_config.yaml, neighborhood_config.yaml): attribute and item access
over a nested dict, ``key=value`` overrides.  Hydra itself is out of scope."""
from __future__ import annotations

import copy
from typing import Any, Iterable, Mapping


class Config(dict):
    """dict with attribute access, recursively (``cfg.mlp.hidden.width``)."""

    def __init__(self, data: Mapping = (), **kw):
        super().__init__()
        for k, v in dict(data, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, Mapping) and not isinstance(v, Config):
            v = Config(v)
        super().__setitem__(k, v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def to_dict(self) -> dict:
        return {k: (v.to_dict() if isinstance(v, Config) else copy.deepcopy(v)) for k, v in self.items()}

    def __deepcopy__(self, memo):
        return Config(self.to_dict())


def to_cfg(obj: Any) -> Any:
    """dict -> Config; objects that already offer attribute access (DictConfig, Config) pass."""
    if isinstance(obj, Config):
        return obj
    if isinstance(obj, Mapping):
        return Config(obj)
    return obj


def cfg_to_dict(cfg: Any) -> dict:
    if isinstance(cfg, Config):
        return cfg.to_dict()
    if isinstance(cfg, Mapping):
        return {k: cfg_to_dict(v) if isinstance(v, Mapping) else v for k, v in cfg.items()}
    try:  # OmegaConf, if it is ever installed
        from omegaconf import OmegaConf
        return OmegaConf.to_container(cfg, resolve=True)
    except Exception:
        return dict(cfg)


def apply_overrides(cfg: Config, overrides: Iterable[str]) -> Config:
    """Hydra-style ``a.b.c=value`` overrides (ints, floats, bools, null, strings)."""
    for ov in overrides:
        key, _, raw = ov.partition("=")
        node = cfg
        parts = key.split(".")
        for p in parts[:-1]:
            if p not in node:
                node[p] = Config()
            node = node[p]
        node[parts[-1]] = _parse(raw)
    return cfg


def _parse(raw: str):
    low = raw.strip().lower()
    if low in ("true", "false"):
        return low == "true"
    if low in ("null", "none"):
        return None
    for cast in (int, float):
        try:
            return cast(raw)
        except ValueError:
            pass
    return raw


def images_config(**overrides) -> Config:
    """The values of /root/reference/config/images_config.yaml + config/optimizer/sparse_lion.yaml, verbatim
    (polynomial n=3, 2 -> 10 -> 10 -> 10 -> 3, one segment everywhere, rotations=1); the tuned README networks are
    explicit overrides in examples/implicit_images.py, as on the reference's command lines (README.md:24,38)."""
    cfg = Config(
        images=["images/newt.jpg", "images/jupiter.jpg"],
        mlp=dict(layer_type="polynomial", n=3, n_in=None, n_out=None, n_hidden=None,
                 periodicity=None, rescale_output=False, scale=2.0,
                 input=dict(segments=1, width=2), output=dict(segments=1, width=3),
                 hidden=dict(segments=1, layers=2, width=10)),
        max_epochs=50, gpus=1, accelerator="cuda", lr=1e-3, batch_size=256,
        train=True, checkpoint=None, rotations=1,
        optimizer=dict(name="sparse_lion", lr=1e-4, patience=5, factor=0.1, gamma=0.9, scheduler="plateau"))
    for k, v in overrides.items():
        cfg[k] = v
    return cfg


def neighborhood_config(**overrides) -> Config:
    """Defaults of /root/reference/config/neighborhood_config.yaml + optimizer/adam.yaml (keys that reach the
    hot path); ``mlp.input.width`` / ``mlp.output.width`` are derived from ``data.width`` / ``data.outside`` by the
    caller, as examples/implicit_neighborhood.py:36-41 does."""
    cfg = Config({
        "images": ["images/newt.jpg", "images/jupiter.jpg", "images/mountains.jpg"],
        "mlp": dict(layer_type="continuous", n=3, n_in=None, n_out=None, n_hidden=None,
                    periodicity=2.0, rescale_output=False, scale=2.0,
                    input=dict(segments=1, width=None), output=dict(segments=1, width=None),
                    hidden=dict(segments=2, layers=1, width=10)),
        "data": dict(width=3, outside=3),      # ("data" is also the name of Config's positional argument)
        "max_epochs": 50, "gpus": 1, "accelerator": "cuda", "lr": 1e-3, "batch_size": 256, "train": True,
        "checkpoint": None,
        "optimizer": dict(name="adam", lr=1e-4, scheduler="plateau", patience=5, factor=0.1, gamma=0.9)})
    for k, v in overrides.items():
        cfg[k] = v
    return cfg


def random_interpolation_config(**overrides) -> Config:
    """Defaults of /root/reference/config/random_interpolation.yaml + optimizer/adam.yaml (``mlp.input.segments``
    is the interpolation ``${image_size}``; ``mlp.input.width`` = (3 + 2*rotations) * num_feature_pixels)."""
    cfg = Config({
        "folder": "images", "num_workers": 10, "save_plots": False,
        "mlp": dict(layer_type="continuous", n=3, n_in=None, n_out=None, n_hidden=None,
                    periodicity=2.0, rescale_output=False, scale=2.0,
                    input=dict(segments=64, width=224), output=dict(segments=2, width=3),
                    hidden=dict(segments=2, layers=2, width=100)),
        "num_target_pixels": 1, "num_feature_pixels": 32, "split_frac": 0.5, "image_size": 64,
        "max_epochs": 50, "gpus": 1, "accelerator": "cuda", "batch_size": 256, "train": True, "checkpoint": None,
        "rotations": 2, "iterations": 10, "all_random": True, "patience": 10, "image": "images/mountains.jpg",
        "optimizer": dict(name="adam", lr=1e-4, scheduler="plateau", patience=5, factor=0.1, gamma=0.9)})
    for k, v in overrides.items():
        cfg[k] = v
    return cfg


def generative_config(**overrides) -> Config:
    """Defaults of /root/reference/config/generative_config.yaml + optimizer/sparse_lion.yaml."""
    cfg = Config({
        "files": ["test_data/test.parquet"],
        "mlp": dict(periodicity=None, rescale_output=False, scale=2.0, width=100, layers=2, segments=2),
        "embedding_size": 384, "input_size": 2, "output_size": 3, "input_segments": 20, "layer_type": "continuous",
        "n": 3, "max_epochs": 50, "gpus": 1, "accelerator": "cuda", "lr": 1e-3, "batch_size": 256, "train": True,
        "checkpoint": None, "rotations": 1,
        "optimizer": dict(name="sparse_lion", lr=1e-4, patience=5, factor=0.1, gamma=0.9, scheduler="plateau")})
    for k, v in overrides.items():
        cfg[k] = v
    return cfg
