"""Drop-in for ``high_order_layers_torch.networks`` (``HighOrderMLP``,
``initialize_network_polynomial_layers``) and for the reference's own model class ``Net``
(/root/reference/high_order_implicit_representation/networks.py:36-129).

``HighOrderMLP.forward`` runs the fused whole-network sm_100a kernel when libhoir_b200 has an
instantiation for the network (``hoir_mlp_fused_supported``) and the per-layer kernels
otherwise; both are CUDA-only.  ``state_dict`` keys are the package's: ``model.<k>.w`` with the
normalisation layers occupying the odd indices (``model.model.<k>.w`` under ``Net``).
"""
from __future__ import annotations

import logging
import math
from typing import Any, Callable, List, Optional

import torch
from torch import Tensor, nn

from . import _lib
from .config import cfg_to_dict, to_cfg
from .layers import (FourierSeries, HighOrderLayer, MaxAbsNormalization, PiecewiseDiscontinuousPolynomial,
                     PiecewisePolynomial, Polynomial, high_order_fc_layers)
from .sparse_optimizers import SparseLion

logger = logging.getLogger(__name__)


class _FusedMLPFunction(torch.autograd.Function):
    """Whole network in one launch: hoir_mlp_forward; backward = hoir_mlp_train_step with the
    upstream dL/dy (recomputes the forward, returns every dL/dw)."""

    @staticmethod
    def forward(ctx, x: Tensor, desc: _lib.MlpDesc, *ws: Tensor) -> Tensor:
        lib = _lib.lib()
        in_f = desc.layers[0].in_features
        out_f = desc.layers[desc.num_layers - 1].out_features
        if x.dtype != torch.float32:
            raise TypeError(f"hoir_b200 layers compute in fp32; got {x.dtype}")
        x2 = x.reshape(-1, in_f).contiguous()
        ws = [w.contiguous() for w in ws]
        with torch.cuda.device(x2.device):
            st = _lib.stream_ptr(x2.device)
            packed = torch.empty(lib.hoir_mlp_packed_size(desc), device=x2.device, dtype=torch.float32)
            _lib.check(lib.hoir_mlp_pack(desc, _lib.ptr_table(ws), _lib.dptr(packed), st))
            y = torch.empty((x2.shape[0], out_f), device=x2.device, dtype=torch.float32)
            _lib.check(lib.hoir_mlp_forward(desc, x2.shape[0], _lib.dptr(x2), None, _lib.dptr(packed),
                                            _lib.dptr(y), 0, st))
        ctx.save_for_backward(x2, packed, *ws)
        ctx.desc = desc
        return y

    @staticmethod
    def backward(ctx, gy: Tensor):
        lib = _lib.lib()
        x2, packed, *ws = ctx.saved_tensors
        desc = ctx.desc
        gy2 = gy.reshape(x2.shape[0], -1).contiguous()
        gws = [torch.zeros_like(w) for w in ws]
        with torch.cuda.device(x2.device):
            _lib.check(lib.hoir_mlp_train_step(desc, x2.shape[0], _lib.dptr(x2), None, 0, None,
                                               _lib.dptr(gy2), _lib.dptr(packed), _lib.ptr_table(gws),
                                               None, 0.0, _lib.stream_ptr(x2.device)))
        return (None, None, *gws)


class _InterpTrainFunction(torch.autograd.Function):
    """Training forward of a network whose forward alone is fused (the neighborhood interpolator family): ONE launch
    (hoir_mlp_forward_acts) that also stores the inputs of the layers 1 .. L-1; backward = the per-layer kernels
    (hoir_layer_backward: tcgen05 dL/dx / dL/dw of the 2-segment layers, the tensor-core dL/dw of the input layer;
    hoir_maxabs_backward) driven from here instead of from L + (L-1) autograd nodes."""

    @staticmethod
    def forward(ctx, x: Tensor, mlp: "HighOrderMLP", *ws: Tensor) -> Tensor:
        lib = _lib.lib()
        desc = mlp.mlp_desc()
        L = desc.num_layers
        hid = desc.layers[0].out_features
        x2 = x.contiguous()
        B = x2.shape[0]
        ws = [w.contiguous() for w in ws]
        y = torch.empty((B, desc.layers[L - 1].out_features), device=x2.device, dtype=torch.float32)
        acts = torch.empty((L - 1, 2, B, hid), device=x2.device, dtype=torch.float32)
        with torch.cuda.device(x2.device):
            _lib.check(lib.hoir_mlp_forward_acts(desc, B, _lib.dptr(x2), _lib.dptr(mlp.packed_weights()), _lib.dptr(y),
                                                 _lib.dptr(acts), _lib.stream_ptr(x2.device)))
        ctx.save_for_backward(x2, acts, *ws)
        ctx.desc, ctx.normalize, ctx.eps = desc, bool(desc.normalize), float(desc.norm_eps)
        return y

    @staticmethod
    def backward(ctx, gy: Tensor):
        lib = _lib.lib()
        x2, acts, *ws = ctx.saved_tensors
        desc = ctx.desc
        L, B = desc.num_layers, x2.shape[0]
        g = gy.contiguous()
        gws = [None] * L
        with torch.cuda.device(x2.device):
            st = _lib.stream_ptr(x2.device)
            for l in range(L - 1, 0, -1):
                xin = acts[l - 1, 1 if ctx.normalize else 0]
                gws[l] = torch.zeros_like(ws[l])
                gx = torch.empty_like(xin)
                _lib.check(lib.hoir_layer_backward(desc.layers[l], B, _lib.dptr(xin), _lib.dptr(ws[l]), _lib.dptr(g),
                                                   _lib.dptr(gx), _lib.dptr(gws[l]), st))
                if ctx.normalize:
                    gh = torch.empty_like(gx)
                    _lib.check(lib.hoir_maxabs_backward(B, gx.shape[1], ctx.eps, _lib.dptr(acts[l - 1, 0]), _lib.dptr(gx),
                                                        _lib.dptr(gh), st))
                    gx = gh
                g = gx
            gws[0] = torch.zeros_like(ws[0])
            gx0 = torch.empty_like(x2) if ctx.needs_input_grad[0] else None
            _lib.check(lib.hoir_layer_backward(desc.layers[0], B, _lib.dptr(x2), _lib.dptr(ws[0]), _lib.dptr(g),
                                               _lib.dptr(gx0), _lib.dptr(gws[0]), st))
        return (gx0, None, *gws)


class HighOrderMLP(nn.Module):
    """``nn.Sequential(L_in, (Norm, L_hid) x hidden_layers, Norm, L_out)`` with the kwargs used at
    /root/reference/high_order_implicit_representation/networks.py:41-55."""

    def __init__(self, layer_type: str, n: int, in_width: int, out_width: int, hidden_layers: int,
                 hidden_width: int, scale: float = 2.0, n_in: Optional[int] = None,
                 n_out: Optional[int] = None, n_hidden: Optional[int] = None,
                 rescale_output: bool = False, periodicity: Optional[float] = None,
                 non_linearity=None, in_segments: Optional[int] = None,
                 out_segments: Optional[int] = None, hidden_segments: Optional[int] = None,
                 normalization: Optional[Callable[[], nn.Module]] = None, resnet: bool = False,
                 device: Optional[str] = None, use_fused: bool = True):
        super().__init__()
        if resnet:
            raise ValueError("resnet=True is not used by the reference and is not implemented")
        n_in = n_in or n
        n_hidden = n_hidden or n
        n_out = n_out or n
        common = dict(rescale_output=rescale_output, scale=scale, periodicity=periodicity, device=device)
        layers: List[nn.Module] = [high_order_fc_layers(
            layer_type=layer_type, n=n_in, in_features=in_width, out_features=hidden_width,
            segments=in_segments, **common)]
        for _ in range(hidden_layers):
            if normalization is not None:
                layers.append(normalization())
            if non_linearity is not None:
                layers.append(non_linearity)
            layers.append(high_order_fc_layers(
                layer_type=layer_type, n=n_hidden, in_features=hidden_width,
                out_features=hidden_width, segments=hidden_segments, **common))
        if normalization is not None:
            layers.append(normalization())
        if non_linearity is not None:
            layers.append(non_linearity)
        layers.append(high_order_fc_layers(
            layer_type=layer_type, n=n_out, in_features=hidden_width, out_features=out_width,
            segments=out_segments, **common))
        self.model = nn.Sequential(*layers)
        self.use_fused = use_fused
        # batch sizes from which the forward-only-fused kernel is used.  It streams the whole input-layer table once per
        # 256-sample tile, yet measured faster than the 7 per-layer launches from batch 256 up when that table is the bulk
        # of the work (216 x 51 x 50: 0.17 against 0.38 ms eager, 0.38 against 0.40 ms per graphed training step at batch
        # 256); for a small input layer (the stock neighborhood_config.yaml: 3 nodes x 10 outputs per input) the tile's
        # 108 ring hand-offs cost more than the per-layer launches until the batch fills the GPU (65536: 0.42 / 1.11 ms).
        big_table = (in_segments or 1) * n_in * hidden_width >= 1024
        self.fused_train_min_batch = 1 if big_table else 16384
        self.fused_infer_min_batch = 1 if big_table else 16384
        self._mlp_desc: Optional[_lib.MlpDesc] = None
        self._desc_checked = False

    # -- description of the network for the C ABI --------------------------------------------
    def high_order_layers(self) -> List[HighOrderLayer]:
        return [m for m in self.model if isinstance(m, HighOrderLayer)]

    def mlp_desc(self) -> Optional[_lib.MlpDesc]:
        """hoir_mlp_desc of this network, or None when the Sequential is not the plain
        ``L (MaxAbs L)*`` / ``L L*`` stack the fused kernels implement."""
        if not self._desc_checked:
            self._desc_checked = True
            mods = list(self.model)
            hol = self.high_order_layers()
            norms = [m for m in mods if isinstance(m, MaxAbsNormalization)]
            others = [m for m in mods if not isinstance(m, (HighOrderLayer, MaxAbsNormalization))]
            ok = not others and len(hol) <= _lib.HOIR_MAX_LAYERS and len(norms) in (0, len(hol) - 1)
            if ok and norms:
                ok = all(isinstance(mods[k], HighOrderLayer) == (k % 2 == 0) for k in range(len(mods)))
                ok = ok and len({m._eps for m in norms}) == 1
            if ok:
                eps = norms[0]._eps if norms else 0.0
                self._mlp_desc = _lib.make_mlp_desc([l.desc for l in hol], bool(norms), eps)
        return self._mlp_desc

    def fused_supported(self) -> bool:
        desc = self.mlp_desc()
        return bool(desc is not None and _lib.lib().hoir_mlp_fused_supported(desc))

    def forward_fused_supported(self) -> bool:
        """The forward has a whole-network kernel (``hoir_mlp_forward``): every ``fused_supported`` network, plus the
        neighborhood-interpolator family whose TRAINING stays on the per-layer kernels."""
        desc = self.mlp_desc()
        return bool(desc is not None and _lib.lib().hoir_mlp_forward_supported(desc))

    def packed_weights(self) -> Tensor:
        """Kernel-side weights of ``hoir_mlp_forward`` / ``hoir_interp_frame``, re-packed only when a ``w`` changed
        (tensor version counters + the library's weight epoch, which the optimizers and graph replays of this package
        bump; the 101 frames of a generate_sequence pack once).  After writing weights through raw pointers from
        outside, call ``_lib.weights_changed()``."""
        ws = [l.w for l in self.high_order_layers()]
        key = (_lib.WEIGHT_EPOCH,) + tuple((w.data_ptr(), w._version) for w in ws)
        capturing = torch.cuda.is_current_stream_capturing()
        if getattr(self, "_packed_key", None) == key and not capturing:
            return self._packed
        lib, desc, dev = _lib.lib(), self.mlp_desc(), ws[0].device
        packed = torch.empty(lib.hoir_mlp_packed_size(desc), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(lib.hoir_mlp_pack(desc, _lib.ptr_table([w.detach().contiguous() for w in ws]), _lib.dptr(packed),
                                         _lib.stream_ptr(dev)))
        if capturing:
            # a CUDA graph records the pack: every replay re-packs the weights it then holds; the copy lives in the
            # graph's memory pool and is not valid outside a replay, so it never enters the cache
            return packed
        self._packed, self._packed_key = packed, key
        return self._packed

    def forward(self, x: Tensor) -> Tensor:
        if (self.use_fused and x.is_cuda and x.dim() == 2 and torch.is_grad_enabled() and x.dtype == torch.float32
                and x.shape[0] >= self.fused_train_min_batch and self.forward_fused_supported() and not self.fused_supported()):
            # training of a network whose forward alone is fused: one launch forward, per-layer backward kernels
            return _InterpTrainFunction.apply(x, self, *[l.w for l in self.high_order_layers()])
        if self.use_fused and x.is_cuda and x.dim() == 2 and not (x.requires_grad and torch.is_grad_enabled()):
            if self.fused_supported():
                return _FusedMLPFunction.apply(x, self.mlp_desc(), *[l.w for l in self.high_order_layers()])
            if (not torch.is_grad_enabled() and x.dtype == torch.float32 and x.shape[0] >= self.fused_infer_min_batch
                    and self.forward_fused_supported()):
                # inference of a network whose forward alone is fused (the interpolators): one launch, no autograd
                desc = self.mlp_desc()
                x2 = x.contiguous()
                y = torch.empty((x2.shape[0], desc.layers[desc.num_layers - 1].out_features), device=x2.device)
                with torch.cuda.device(x2.device):
                    _lib.check(_lib.lib().hoir_mlp_forward(desc, x2.shape[0], _lib.dptr(x2), None,
                                                           _lib.dptr(self.packed_weights()), _lib.dptr(y), 0,
                                                           _lib.stream_ptr(x2.device)))
                return y
        return self.model(x)


def node_positions(layer: HighOrderLayer) -> Tensor:
    """Global position in [-1, 1] of every weight node of a polynomial / piecewise layer."""
    n = layer.n
    if n == 1:
        X = torch.zeros(1)
    else:
        X = -torch.cos((torch.arange(0, n) * math.pi / (n - 1)).to(torch.float32))
    if isinstance(layer, Polynomial):
        return (layer.length / 2.0) * X
    seg = layer.segments
    pos = []
    for s in range(seg):
        lo, hi = -1.0 + 2.0 * s / seg, -1.0 + 2.0 * (s + 1) / seg
        seg_pos = lo + (X + 1.0) * 0.5 * (hi - lo)
        keep_all = isinstance(layer, PiecewiseDiscontinuousPolynomial) or s == seg - 1
        pos.append(seg_pos if keep_all else seg_pos[:-1])
    return torch.cat(pos)


def initialize_network_polynomial_layers(network: nn.Module, max_slope: float, max_offset: float,
                                         scale_slope: Callable[[int], float] = lambda in_features: 1.0,
                                         generator: Optional[torch.Generator] = None) -> None:
    """Re-initialise every polynomial / piecewise layer as a random straight line per (out, in)
    link (/root/reference/high_order_implicit_representation/networks.py:57).
    Fourier layers keep their constructor init.  One-time host work, not on the hot path."""
    for layer in network.modules():
        if isinstance(layer, (Polynomial, PiecewisePolynomial, PiecewiseDiscontinuousPolynomial)):
            out_f, in_f, _ = layer.w.shape
            slope = (torch.rand(out_f, in_f, 1, generator=generator) * 2 - 1) * max_slope * scale_slope(in_f)
            offset = (torch.rand(out_f, in_f, 1, generator=generator) * 2 - 1) * max_offset
            w = slope * node_positions(layer).view(1, 1, -1) + offset
            with torch.no_grad():
                layer.w.copy_(w.to(layer.w.device))


def _write_hyper(dev_tensor: Tensor, values) -> None:
    import ctypes as C
    arr = (C.c_float * 8)(*([float(v) for v in values] + [0.0] * (8 - len(values))))
    with torch.cuda.device(dev_tensor.device):
        _lib.check(_lib.lib().hoir_write_floats(_lib.dptr(dev_tensor), 8, arr, _lib.stream_ptr(dev_tensor.device)))


class Adam(torch.optim.Optimizer):
    """torch.optim.Adam semantics (lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0) on
    ``hoir_adam_step`` -- what /root/reference/.../networks.py:94-97 constructs.

    ``capturable=True``: the scalars (lr, betas, eps, weight decay, bias corrections) live in device memory
    (``hoir_adam_step_dev``), so ``step()`` can be captured in a CUDA graph; ``refresh_hyper()`` advances the step
    count and rewrites them (stream-ordered, outside the graph) before every replay -- the lr schedule and the bias
    corrections follow without re-capture.  Eager ``step()`` calls it itself."""

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, capturable: bool = False):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.capturable = capturable

    def refresh_hyper(self) -> None:
        for group in self.param_groups:
            dev = group["params"][0].device
            if "_hyper" not in group:
                group["_hyper"] = torch.zeros(8, device=dev)
                group["_step"] = 0
            group["_step"] += 1
            b1, b2 = group["betas"]
            bc1 = 1.0 - b1 ** group["_step"]
            bc2 = 1.0 - b2 ** group["_step"]
            _write_hyper(group["_hyper"], [group["lr"], b1, b2, group["eps"], group["weight_decay"], bc1, bc2 ** 0.5])

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.lib()
        _lib.weights_changed()               # (raw-pointer writes: invisible to the tensors' version counters)
        if self.capturable and not torch.cuda.is_current_stream_capturing():
            self.refresh_hyper()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                g = p.grad.contiguous()
                with torch.cuda.device(p.device):
                    if self.capturable:
                        _lib.check(lib.hoir_adam_step_dev(
                            p.numel(), _lib.dptr(p.data), _lib.dptr(g), _lib.dptr(st["exp_avg"]),
                            _lib.dptr(st["exp_avg_sq"]), _lib.dptr(group["_hyper"]), _lib.stream_ptr(p.device)))
                        continue
                    st["step"] += 1
                    _lib.check(lib.hoir_adam_step(
                        p.numel(), _lib.dptr(p.data), _lib.dptr(g), _lib.dptr(st["exp_avg"]),
                        _lib.dptr(st["exp_avg_sq"]), float(group["lr"]), float(b1), float(b2),
                        float(group["eps"]), float(group["weight_decay"]), st["step"],
                        _lib.stream_ptr(p.device)))
        return loss


class Net(nn.Module):
    """The reference's model class (/root/reference/high_order_implicit_representation/
    networks.py:36-129) without the Lightning base class (not installable here): same ``cfg``
    tree, same ``forward`` / ``eval_step`` / ``*_step`` / ``configure_optimizers``, same
    ``state_dict`` keys (``model.model.<k>.w``), Lightning-format checkpoints."""

    def __init__(self, cfg: Any):
        super().__init__()
        cfg = to_cfg(cfg)
        self.hparams = cfg          # save_hyperparameters(cfg), networks.py:39
        self.cfg = cfg
        self.model = HighOrderMLP(
            layer_type=cfg.mlp.layer_type,
            n=cfg.mlp.n,
            n_in=cfg.mlp.n_in,
            n_hidden=cfg.mlp.n_in,      # sic: the reference passes n_in here (Quirk Q1)
            n_out=cfg.mlp.n_out,
            in_width=cfg.mlp.input.width,
            in_segments=cfg.mlp.input.segments,
            out_width=cfg.mlp.output.width,
            out_segments=cfg.mlp.output.segments,
            hidden_width=cfg.mlp.hidden.width,
            hidden_layers=cfg.mlp.hidden.layers,
            hidden_segments=cfg.mlp.hidden.segments,
            normalization=MaxAbsNormalization,
        )
        initialize_network_polynomial_layers(self.model, max_slope=1.0, max_offset=0.0)
        self.loss = nn.MSELoss()
        self.logged = {}
        self.current_epoch = 0
        self.global_step = 0

    def forward(self, x):
        return self.model(x)

    def log(self, name: str, value, prog_bar: bool = False, **kwargs) -> None:
        self.logged[name] = value.detach() if isinstance(value, Tensor) else value

    def eval_step(self, batch, name: str):
        x, y = batch
        y_hat = self(x.flatten(1))
        loss = self.loss(y_hat.flatten(), y.flatten())
        self.log(f"{name}_loss", loss, prog_bar=True)
        return loss

    def training_step(self, batch, batch_idx):
        return self.eval_step(batch, "train")

    def validation_step(self, batch, batch_idx):
        return self.eval_step(batch, "val")

    def test_step(self, batch, batch_idx):
        return self.eval_step(batch, "test")

    def configure_optimizers(self):
        opt = self.cfg.optimizer
        if opt.name == "adahessian":
            raise ValueError("Optimizer adahessian needs torch_optimizer (Hessian-vector products); "
                             "out of scope for the B200 hot path")
        elif opt.name in ["adam", "sparse_lion"]:
            if opt.name == "adam":
                optimizer = Adam(params=self.parameters(), lr=opt.lr)
            else:
                optimizer = SparseLion(params=self.parameters(), lr=opt.lr)
            reduce_on_plateau = False
            scheduler = opt.get("scheduler") if hasattr(opt, "get") else getattr(opt, "scheduler", None)
            if scheduler == "plateau":
                logger.info("Reducing lr on plateau")
                lr_scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(
                    optimizer, patience=opt.patience, factor=opt.factor)
                reduce_on_plateau = True
            elif scheduler == "exponential":
                logger.info("Reducing lr exponentially")
                lr_scheduler = torch.optim.lr_scheduler.ExponentialLR(optimizer, gamma=opt.gamma)
            else:
                return optimizer
            return [optimizer], [{"scheduler": lr_scheduler, "reduce_on_plateau": reduce_on_plateau,
                                  "monitor": "train_loss"}]
        else:
            raise ValueError(f"Optimizer {opt.name} not recognized")

    # -- Lightning-format checkpoints -------------------------------------
    def checkpoint_dict(self) -> dict:
        return {"state_dict": {k: v.detach().cpu() for k, v in self.state_dict().items()},
                "hyper_parameters": cfg_to_dict(self.cfg), "epoch": self.current_epoch,
                "global_step": self.global_step, "pytorch-lightning_version": "hoir_b200"}

    def save_checkpoint(self, path: str) -> None:
        torch.save(self.checkpoint_dict(), path)

    @classmethod
    def load_from_checkpoint(cls, path: str, map_location: Any = None) -> "Net":
        ckpt = torch.load(path, map_location="cpu", weights_only=False)
        net = cls(ckpt["hyper_parameters"])
        net.load_state_dict(ckpt["state_dict"])
        net.current_epoch = ckpt.get("epoch", 0)
        net.global_step = ckpt.get("global_step", 0)
        if map_location is not None:
            net = net.to(map_location)
        return net


class GenerativeNetwork(nn.Module):
    """/root/reference/high_order_implicit_representation/networks.py:132-180: a polynomial (n=2) layer over the text
    embedding plus a piecewise layer over the pixel position, summed, then a HighOrderMLP.  Every layer runs through
    the C-ABI layer kernels; the MLP uses the fused whole-network kernel when it has an instantiation."""

    def __init__(self, embedding_size: int, input_size: int, output_size: int, mlp_width: int, mlp_layers: int,
                 input_segments: int, mlp_segments: int, n: int, normalization: Any = None,
                 layer_type: str = "continuous"):
        super().__init__()
        self.text_layer = high_order_fc_layers(layer_type="polynomial", n=2, in_features=embedding_size,
                                               out_features=mlp_width)
        self.position_layer = high_order_fc_layers(layer_type=layer_type, n=n, in_features=input_size,
                                                   out_features=mlp_width, segments=input_segments)
        self.mlp = HighOrderMLP(layer_type=layer_type, n=n, in_width=mlp_width, hidden_width=mlp_width,
                                hidden_layers=mlp_layers, out_width=output_size, in_segments=mlp_segments,
                                out_segments=mlp_segments, hidden_segments=mlp_segments, normalization=normalization)

    def forward(self, text_embedding: Tensor, position: Tensor) -> Tensor:
        text_out = self.text_layer(text_embedding)
        position_out = self.position_layer(position)
        return self.mlp(text_out + position_out)


class GenNet(Net):
    """networks.py:183-278: the text-to-image model (``batch = (caption embedding, position, colour)``); optimizer and
    checkpoint handling are Net's."""

    def __init__(self, cfg: Any):
        nn.Module.__init__(self)
        cfg = to_cfg(cfg)
        self.hparams = cfg
        self.cfg = cfg
        self.model = GenerativeNetwork(
            layer_type=cfg.layer_type, embedding_size=cfg.embedding_size, n=cfg.n, input_size=cfg.input_size,
            output_size=cfg.output_size, mlp_width=cfg.mlp.width, mlp_layers=cfg.mlp.layers,
            input_segments=cfg.input_segments, mlp_segments=cfg.mlp.segments, normalization=MaxAbsNormalization)
        initialize_network_polynomial_layers(self.model, max_slope=1.0, max_offset=0.0)
        self.loss = nn.MSELoss()
        self.logged = {}
        self.current_epoch = 0
        self.global_step = 0

    def forward(self, caption, x):
        return self.model(caption, x)

    def eval_step(self, batch, name: str):
        caption, x, color = batch
        y_hat = self(caption, x.flatten(1))
        loss = self.loss(y_hat.flatten(), color.flatten())
        self.log(f"{name}_loss", loss, prog_bar=True)
        return loss
