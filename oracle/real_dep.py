"""TEST INFRASTRUCTURE (like everything under oracle/): locate and drive a REAL install of the dependency
``high-order-layers-torch`` (^2.4.2, /root/reference/pyproject.toml:21), if one ever appears.

The package is not in /root/reference, not installed in this image, not in /opt/wheelhouse: today ``find()`` returns
None and every consumer (tests/test_real_dep_parity.py, scripts/verify_real_dep.py, ``bench.py --impl reference``)
skips or falls back to the restated oracle, saying so.  The in-tree alias (compat/high_order_layers_torch, version
``…+hoir_b200``) is never accepted as "real"."""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
from typing import Dict, Optional

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_EXTRA = [os.path.join(ROOT, "baseline", "_ref")]          # git-ignored install target (pip --target)


def find():
    """The real ``high_order_layers_torch`` module, or None.  The repository root and compat/ are taken off
    ``sys.path`` for the lookup; an already imported alias is ignored."""
    mod = sys.modules.get("high_order_layers_torch")
    if mod is not None and not _is_alias(mod):
        return mod
    saved_path, saved_mods = list(sys.path), {k: v for k, v in sys.modules.items()
                                              if k == "high_order_layers_torch" or k.startswith("high_order_layers_torch.")}
    try:
        for k in saved_mods:
            del sys.modules[k]
        banned = {ROOT, os.path.join(ROOT, "compat"), ""}
        sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") not in banned and p not in banned]
        sys.path[:0] = [p for p in _EXTRA if os.path.isdir(p)]
        if importlib.util.find_spec("high_order_layers_torch") is None:
            return None
        mod = importlib.import_module("high_order_layers_torch")
        if _is_alias(mod):
            return None
        for sub in ("layers", "networks", "utils", "sparse_optimizers"):
            importlib.import_module(f"high_order_layers_torch.{sub}")
        saved_mods = {}                      # keep the real modules registered
        return mod
    except Exception:                         # a broken install is "not found": callers say "parity unpinned"
        return None
    finally:
        sys.path[:] = saved_path
        sys.modules.update(saved_mods)


def _is_alias(mod) -> bool:
    origin = os.path.abspath(getattr(mod, "__file__", "") or "")
    return str(getattr(mod, "__version__", "")).endswith("+hoir_b200") or origin.startswith(ROOT + os.sep + "compat")


# layer cases of the pin: (name, layer_type, kwargs)
LAYER_CASES = [
    ("continuous_n3_s100", "continuous", dict(n=3, in_features=4, out_features=5, segments=100)),
    ("continuous_n2_s50", "continuous", dict(n=2, in_features=6, out_features=4, segments=50)),
    ("continuous_n3_s2", "continuous", dict(n=3, in_features=5, out_features=3, segments=2)),
    ("discontinuous_n3_s2", "discontinuous", dict(n=3, in_features=5, out_features=3, segments=2)),
    ("discontinuous_n3_s100", "discontinuous", dict(n=3, in_features=4, out_features=5, segments=100)),
    ("fourier_n31", "fourier", dict(n=31, in_features=4, out_features=5)),
    ("fourier_n3", "fourier", dict(n=3, in_features=5, out_features=3)),
    ("polynomial_n3", "polynomial", dict(n=3, in_features=4, out_features=5)),
    ("polynomial_n2", "polynomial", dict(n=2, in_features=7, out_features=2)),
]


def case_inputs(layer_type: str, kw: dict) -> torch.Tensor:
    g = torch.Generator().manual_seed(1234)
    x = torch.rand(64, kw["in_features"], generator=g) * 2.2 - 1.1
    seg = kw.get("segments", 1)
    b = (-1.0 + 2.0 * torch.arange(0, seg + 1, dtype=torch.float64) / seg).to(torch.float32)
    edge = torch.cat([b, torch.nextafter(b, torch.tensor(2.0)), torch.nextafter(b, torch.tensor(-2.0))])
    edge = edge[: (edge.numel() // kw["in_features"]) * kw["in_features"]].reshape(-1, kw["in_features"])
    return torch.cat([x, edge[:64]])


def run_layer(module_layers, layer_type: str, kw: dict, w: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """forward + backward of one layer built by ``module_layers.high_order_fc_layers`` (real package, oracle or
    hoir_b200 alike); ``w`` is copied in when given."""
    torch.manual_seed(0)
    layer = module_layers.high_order_fc_layers(layer_type=layer_type, **kw)
    if w is not None:
        with torch.no_grad():
            layer.w.copy_(w)
    x = case_inputs(layer_type, kw).requires_grad_(True)
    y = layer(x)
    g = torch.Generator().manual_seed(99)
    gy = torch.rand(y.shape, generator=g) * 2 - 1
    y.backward(gy)
    return {"x": x.detach(), "w": layer.w.detach().clone(), "y": y.detach(), "gy": gy, "gx": x.grad.clone(),
            "gw": layer.w.grad.clone()}
