"""The reference's inference loops around the hot path
(/root/reference/high_order_implicit_representation/rendering.py): full-image rendering
(ImageGenerator, :195-213) and the neighborhood interpolator (neighborhood_sample_generator /
generate_sequence, :28-126).  TensorBoard / matplotlib output is out of scope."""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor, nn

from . import _lib
from .datasets import neighborhood_gather, neighborhood_patches
from .engine import FusedTrainer


def render_image(model: nn.Module, rows: int, cols: int, rotations: int, batch_size: Optional[int] = None) -> Tensor:
    """``[rows, cols, 3]`` in [0, 1]: the model evaluated at every pixel of the mesh, ``0.5 * (y + 1)``
    (rendering.py:199-213).  A network with a fused instantiation renders in ONE launch with in-kernel
    coordinates; others go through ``model(x)`` in ``batch_size`` chunks like the reference (including
    its empty trailing batch when ``N % batch_size == 0``)."""
    from .networks import HighOrderMLP
    from .utils import positions_from_mesh
    mlp = getattr(model, "model", model)
    dev = next(model.parameters()).device
    if isinstance(mlp, HighOrderMLP) and mlp.fused_supported() and mlp.high_order_layers()[0].in_features == 2 * rotations:
        lib = _lib.lib()
        ws = [l.w.detach().contiguous() for l in mlp.high_order_layers()]
        desc = mlp.mlp_desc()
        out = torch.empty((rows * cols, ws[-1].shape[0]), device=dev)
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            packed = torch.empty(lib.hoir_mlp_packed_size(desc), device=dev)
            _lib.check(lib.hoir_mlp_pack(desc, _lib.ptr_table(ws), _lib.dptr(packed), st))
            mesh = _lib.make_mesh_desc(rows, cols, rotations, True, 0)
            _lib.check(lib.hoir_mlp_forward(desc, rows * cols, None, mesh, _lib.dptr(packed), _lib.dptr(out), 1, st))
        return out.reshape(rows, cols, -1)
    with torch.no_grad():
        lines = positions_from_mesh(rows, cols, device=str(dev), rotations=rotations, normalize=True)
        inputs = torch.stack(lines, dim=2).reshape(-1, 2 * rotations)
        bs = batch_size or 1 << 20
        outs = [model(inputs[b * bs:(b + 1) * bs]) for b in range((len(inputs) + bs) // bs)]
        return 0.5 * (torch.cat(outs).reshape(rows, cols, -1) + 1.0)


def neighborhood_sample_generator(model: nn.Module, image: Tensor, width: int, outside: int,
                                  batch_size: int = 1 << 18, device: Optional[str] = None,
                                  alpha: float = 0.0, prev: Optional[Tensor] = None) -> Tensor:
    """One pass of the interpolator over ``image [C, H, W]``: reflection pad by ``2*outside + width``,
    stride-``width`` donut patches (gathered in-kernel), model forward, fold + crop (one kernel, optionally
    blended ``alpha * prev + (1 - alpha) * new``).  rendering.py:28-95."""
    if not image.is_cuda:
        raise RuntimeError("hoir_b200 has no CPU path: expected a CUDA image tensor")
    with torch.no_grad():
        img = image.contiguous().float()
        Cc, Hh, Ww = img.shape
        pad = 2 * outside + width
        from .networks import HighOrderMLP
        mlp = model if isinstance(model, HighOrderMLP) else getattr(model, "model", None)     # Net(cfg).model
        total = width + 2 * outside
        if (isinstance(mlp, HighOrderMLP) and mlp.use_fused and mlp.forward_fused_supported() and not mlp.fused_supported()
                and mlp.high_order_layers()[0].in_features == Cc * (total * total - width * width)
                and mlp.high_order_layers()[-1].out_features == Cc * width * width and pad < min(Hh, Ww)
                and next(mlp.parameters()).device == img.device):
            # ONE launch per frame: reflection pad + stride-w gather + whole-network forward + fold + crop + blend
            out = torch.empty_like(img)
            prev_c = prev.contiguous().float() if prev is not None else None
            with torch.cuda.device(img.device):
                _lib.check(_lib.lib().hoir_interp_frame(mlp.mlp_desc(), _lib.dptr(img), Cc, Hh, Ww, width, outside,
                                                        _lib.dptr(mlp.packed_weights()), float(alpha), _lib.dptr(prev_c),
                                                        _lib.dptr(out), _lib.stream_ptr(img.device)))
            return out
        feat, _, _ = neighborhood_gather(img, width, outside, stride=width, pad=pad, want_targets=False)
        outs = [model(feat[b * batch_size:(b + 1) * batch_size])
                for b in range((len(feat) + batch_size) // batch_size)]
        pred = torch.cat(outs).contiguous()
        out = torch.empty_like(img)
        with torch.cuda.device(img.device):
            _lib.check(_lib.lib().hoir_neighborhood_fold(_lib.dptr(pred), Cc, Hh, Ww, width, outside, float(alpha),
                                                         _lib.dptr(prev.contiguous()) if prev is not None else None,
                                                         _lib.dptr(out), _lib.stream_ptr(img.device)))
        return out


def generate_sequence(model: nn.Module, image: Tensor, frames: int, width: int, outside: int,
                      batch_size: int = 1 << 18, skip: int = 1, alpha: float = 0.75) -> Tensor:
    """rendering.py:98-126: iterate the interpolator, ``this = alpha*this + (1-alpha)*new`` (the reference's
    clamp discards its result, Quirk Q11); returns the kept frames scaled to [0, 1] on the host."""
    this_image = image.clone()
    image_list = [this_image.cpu()]
    for frame in range(frames):
        this_image = neighborhood_sample_generator(model, this_image, width, outside, batch_size,
                                                   alpha=alpha, prev=this_image)
        if frame % skip == 0:
            image_list.append(this_image.detach().cpu().clone())
    return 0.5 * (torch.stack(image_list, dim=0) + 1)
