"""The reference's image -> tensors transforms that sit either side of the hot path
(/root/reference/high_order_implicit_representation/single_image_dataset.py:22-51, 83-141), on the
GPU: coordinates come from ``hoir_mesh_positions`` and patches from ``hoir_neighborhood_gather``.
File decoding (PIL) is host-side I/O and not part of the path."""
from __future__ import annotations

import ctypes as C
from typing import Tuple, Union

import torch
from torch import Tensor

from . import _lib
from .utils import positions_from_mesh


def _load_image_u8(src: Union[str, Tensor]) -> Tensor:
    if isinstance(src, str):
        import numpy as np
        from PIL import Image
        return torch.from_numpy(np.array(Image.open(src).convert("RGB")))
    return src


def image_to_dataset(filename: Union[str, Tensor], peano: bool = False, rotations: int = 1,
                     device: str = "cuda") -> Tuple[Tensor, Tensor, Tensor]:
    """(flattened RGB in [-1, 1] ``[H*W, 3]``, positions ``[H*W, 2*rotations]``, uint8 image ``[H, W, 3]``);
    ``filename`` may also be an in-memory uint8 ``[H, W, 3]`` tensor."""
    img = _load_image_u8(filename).to(device)
    lines = positions_from_mesh(width=img.shape[0], height=img.shape[1], device=device,
                                rotations=rotations, normalize=True)
    pos = torch.stack(lines, dim=2).reshape(-1, 2 * rotations)
    # u*2/255-1 (single_image_dataset.py:49) with an IEEE division: torch's CUDA "divide by a Python scalar"
    # multiplies by the reciprocal (1 ulp off the reference's CPU result); a tensor divisor divides
    flat = (img.reshape(-1, 3) * 2.0) / torch.tensor(255.0, device=img.device) - 1
    return flat, pos, img


def neighborhood_patches(H: int, W: int, width: int, outside: int, stride: int, pad: int = 0) -> Tuple[int, int]:
    nr, nc = C.c_int(0), C.c_int(0)
    total = _lib.lib().hoir_neighborhood_patches(H, W, width, outside, stride, pad, C.byref(nr), C.byref(nc))
    if total < 0:
        raise ValueError("bad image / patch geometry")
    return (nr.value, nc.value) if total > 0 else (0, 0)


def neighborhood_gather(image: Tensor, width: int, outside: int, stride: int, pad: int = 0,
                        want_features: bool = True, want_targets: bool = True):
    """Patch gather on the device (optionally through a reflection padding of ``pad`` pixels)."""
    if not image.is_cuda:
        raise RuntimeError("hoir_b200 has no CPU path: expected a CUDA image tensor")
    img = image.contiguous().float()
    Cc, Hh, Ww = img.shape
    nr, nc = neighborhood_patches(Hh, Ww, width, outside, stride, pad)
    P = nr * nc
    T = width + 2 * outside
    feat = torch.empty((P, Cc * (T * T - width * width)), device=img.device) if want_features else None
    targ = torch.empty((P, Cc * width * width), device=img.device) if want_targets else None
    with torch.cuda.device(img.device):
        _lib.check(_lib.lib().hoir_neighborhood_gather(_lib.dptr(img), Cc, Hh, Ww, width, outside, stride, pad, 0, P,
                                                       _lib.dptr(feat), _lib.dptr(targ), _lib.stream_ptr(img.device)))
    return feat, targ, (nr, nc)


def image_neighborhood_dataset(image: Tensor, width: int = 3, outside: int = 1, stride: int = 1,
                               device: str = "cuda"):
    """image ``[C, H, W]`` in [-1, 1] -> (donut features ``[nH*nW, C*(T^2-w^2)]``, block targets
    ``[nH*nW, C*w^2]``, image, lastx, lasty) with the reference's element order."""
    img = image.to(device)
    feat, targ, _ = neighborhood_gather(img, width, outside, stride)
    total = width + 2 * outside
    return feat, targ, img, img.shape[1] - total, img.shape[2] - total
