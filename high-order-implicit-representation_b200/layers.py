"""Drop-in for ``high_order_layers_torch.layers`` as used by the reference
(/root/reference/high_order_implicit_representation/networks.py:7,24,54,148-161):
``high_order_fc_layers`` and ``MaxAbsNormalization``, backed by the sm_100a kernels of
libhoir_b200.so through ``torch.autograd.Function``s.  Parameter name (``w``), its layout
``[out_features, in_features, nodes]`` and the constructor kwargs follow the package.

The reference star-imports this module and relies on the names ``torch``, ``math``, ``Tensor``
and ``Any`` leaking out of it (SURVEY.md Quirk Q7) -- they are re-exported on purpose.
"""
from __future__ import annotations

import math  # noqa: F401  (re-exported, Quirk Q7)
from typing import Any, Optional  # noqa: F401  (re-exported, Quirk Q7)

import torch
from torch import Tensor, nn

from . import _lib, spec


def _require_cuda(x: Tensor) -> None:
    if not x.is_cuda:
        raise RuntimeError("hoir_b200 has no CPU path: expected a CUDA tensor, got a "
                           f"{x.device.type} tensor (build/ship libhoir_b200.so and move the "
                           "module and its inputs to a B200)")


def _as_2d(x: Tensor, in_features: int) -> Tensor:
    _require_cuda(x)
    if x.dtype != torch.float32:
        raise TypeError(f"hoir_b200 layers compute in fp32; got {x.dtype}")
    if x.shape[-1] != in_features and x.dim() == 2:
        raise ValueError(f"expected [B, {in_features}] input, got {tuple(x.shape)}")
    return x.reshape(-1, in_features).contiguous()


class _LayerFunction(torch.autograd.Function):
    """One high-order layer: hoir_layer_forward / hoir_layer_backward (include/hoir.h)."""

    @staticmethod
    def forward(ctx, x: Tensor, w: Tensor, desc: _lib.LayerDesc) -> Tensor:
        lib = _lib.lib()
        x2 = _as_2d(x, desc.in_features)
        wc = w.contiguous()
        y = torch.empty((x2.shape[0], desc.out_features), device=x2.device, dtype=torch.float32)
        with torch.cuda.device(x2.device):
            _lib.check(lib.hoir_layer_forward(desc, x2.shape[0], _lib.dptr(x2), _lib.dptr(wc),
                                              _lib.dptr(y), _lib.stream_ptr(x2.device)))
        ctx.save_for_backward(x2, wc)
        ctx.desc = desc
        ctx.x_shape = x.shape
        return y.reshape(*x.shape[:-1], desc.out_features) if x.dim() != 2 else y

    @staticmethod
    def backward(ctx, gy: Tensor):
        lib = _lib.lib()
        x2, w = ctx.saved_tensors
        desc = ctx.desc
        gy2 = gy.reshape(-1, desc.out_features).contiguous()
        gx = torch.empty_like(x2) if ctx.needs_input_grad[0] else None
        gw = torch.zeros_like(w) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(x2.device):
            _lib.check(lib.hoir_layer_backward(desc, x2.shape[0], _lib.dptr(x2), _lib.dptr(w),
                                               _lib.dptr(gy2), _lib.dptr(gx), _lib.dptr(gw),
                                               _lib.stream_ptr(x2.device)))
        if gx is not None:
            gx = gx.reshape(ctx.x_shape)
        return gx, gw, None


class _MaxAbsFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, eps: float) -> Tensor:
        lib = _lib.lib()
        xc = x.contiguous()
        y = torch.empty_like(xc)
        with torch.cuda.device(xc.device):
            _lib.check(lib.hoir_maxabs_forward(xc.shape[0], xc.shape[1], eps, _lib.dptr(xc),
                                               _lib.dptr(y), _lib.stream_ptr(xc.device)))
        ctx.save_for_backward(xc)
        ctx.eps = eps
        return y

    @staticmethod
    def backward(ctx, gy: Tensor):
        lib = _lib.lib()
        (x,) = ctx.saved_tensors
        gyc = gy.contiguous()
        gx = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.check(lib.hoir_maxabs_backward(x.shape[0], x.shape[1], ctx.eps, _lib.dptr(x),
                                                _lib.dptr(gyc), _lib.dptr(gx),
                                                _lib.stream_ptr(x.device)))
        return gx, None


class MaxAbsNormalization(nn.Module):
    """x / (max_i |x_i| + eps) over dim 1; no parameters.  Passed as a class at
    /root/reference/high_order_implicit_representation/networks.py:54."""

    def __init__(self, eps: float = spec.MAXABS_EPS, dim: int = 1):
        super().__init__()
        self._eps = eps
        self._dim = dim

    def forward(self, x: Tensor) -> Tensor:
        _require_cuda(x)
        if x.dim() != 2 or self._dim != 1:
            raise ValueError("MaxAbsNormalization (hoir_b200) normalises [B, width] over dim 1")
        if x.dtype != torch.float32:
            raise TypeError(f"hoir_b200 layers compute in fp32; got {x.dtype}")
        return _MaxAbsFunction.apply(x, self._eps)


class HighOrderLayer(nn.Module):
    """Common base of the four layer types; ``w`` is ``[out_features, in_features, nodes]``."""

    kind: str = ""

    def __init__(self, n: int, in_features: int, out_features: int, segments: Optional[int] = None,
                 length: float = 2.0, weight_magnitude: float = 1.0,
                 periodicity: Optional[float] = None, rescale_output: bool = False,
                 device: Optional[str] = None, **kwargs):
        super().__init__()
        if self.kind in ("continuous", "discontinuous") and not segments:
            raise ValueError(f"{self.kind} layers need segments >= 1")
        self.n = n
        self.in_features = in_features
        self.out_features = out_features
        self.segments = int(segments) if self.kind in ("continuous", "discontinuous") else 1
        self.length = float(length)
        self.periodicity = periodicity
        self.rescale_output = rescale_output
        self._desc = _lib.make_layer_desc(self.kind, n, self.segments, in_features, out_features,
                                          length=self.length, rescale_output=rescale_output,
                                          periodicity=periodicity)
        if n < 1 or n > _lib.HOIR_MAX_N:
            raise ValueError(f"n must be in [1, {_lib.HOIR_MAX_N}]")
        if self.kind == "continuous" and n < 2:
            raise ValueError("continuous layers need n >= 2")
        self.nodes = {"continuous": (n - 1) * self.segments + 1,
                      "discontinuous": n * self.segments}.get(self.kind, n)
        w = torch.empty(out_features, in_features, self.nodes, device=device)
        w.uniform_(-weight_magnitude / in_features, weight_magnitude / in_features)
        self.w = nn.Parameter(w)

    @property
    def desc(self) -> _lib.LayerDesc:
        return self._desc

    def forward(self, x: Tensor) -> Tensor:
        return _LayerFunction.apply(x, self.w, self._desc)

    def extra_repr(self) -> str:
        return (f"n={self.n}, in_features={self.in_features}, out_features={self.out_features}, "
                f"segments={self.segments}")


class PiecewisePolynomial(HighOrderLayer):
    kind = "continuous"


class PiecewiseDiscontinuousPolynomial(HighOrderLayer):
    kind = "discontinuous"


class Polynomial(HighOrderLayer):
    kind = "polynomial"


class FourierSeries(HighOrderLayer):
    kind = "fourier"


fc_layers = {
    "continuous": PiecewisePolynomial,
    "discontinuous": PiecewiseDiscontinuousPolynomial,
    "polynomial": Polynomial,
    "fourier": FourierSeries,
}


def high_order_fc_layers(layer_type: str, **kwargs) -> nn.Module:
    """Factory with the kwargs the reference passes
    (/root/reference/high_order_implicit_representation/networks.py:148-161 and, through
    HighOrderMLP, :41-55): n, in_features, out_features[, segments, rescale_output, scale,
    periodicity, device]."""
    if layer_type not in fc_layers:
        raise ValueError(f"Fully connected layer type {layer_type} not recognized. "
                         f"Must be one of {list(fc_layers.keys())}")
    kwargs = dict(kwargs)
    scale = kwargs.pop("scale", None)
    if scale is not None:
        kwargs["length"] = scale
    return fc_layers[layer_type](**kwargs)


def segment_index(layer: HighOrderLayer, x: Tensor):
    """Parity/debug entry: (id_min, wid_min, x_local) of a piecewise layer for x[B, in]."""
    lib = _lib.lib()
    x2 = _as_2d(x, layer.in_features)
    idm = torch.empty(x2.shape, device=x2.device, dtype=torch.int32)
    wid = torch.empty_like(idm)
    xl = torch.empty_like(x2)
    with torch.cuda.device(x2.device):
        _lib.check(lib.hoir_segment_index(layer.desc, x2.shape[0], _lib.dptr(x2),
                                          _lib.dptr(idm, torch.int32), _lib.dptr(wid, torch.int32),
                                          _lib.dptr(xl), _lib.stream_ptr(x2.device)))
    return idm, wid, xl
