"""Drop-in for ``high_order_layers_torch.utils.positions_from_mesh`` as called at
/root/reference/high_order_implicit_representation/single_image_dataset.py:39-45."""
from __future__ import annotations

from typing import List

import torch
from torch import Tensor

from . import _lib


def positions_from_mesh(width: int, height: int, device: str = "cpu", rotations: int = 1,
                        normalize: bool = False) -> List[Tensor]:
    """List of ``2*rotations`` tensors ``[width, height]`` ("width" is the ROW count, Quirk Q10).

    The coordinates are produced by ``hoir_mesh_positions`` (sm_100a kernel); a ``device="cpu"`` caller (the
    reference's dataset code, single_image_dataset.py:39-45) gets the kernel's result copied to the host whenever a CUDA
    device is present -- the same bits as the in-kernel mesh of the training step.  Only on a machine WITHOUT a GPU
    (dataset preparation on a login node, the CPU test suite) are they computed with stock torch ops."""
    dev = torch.device(device)
    if dev.type == "cpu" and torch.cuda.is_available():
        return [t.cpu() for t in positions_from_mesh(width, height, device="cuda", rotations=rotations,
                                                     normalize=normalize)]
    if dev.type == "cuda":
        lib = _lib.lib()
        mesh = _lib.make_mesh_desc(width, height, rotations, normalize)
        n = width * height
        out = torch.empty((n, 2 * rotations), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(lib.hoir_mesh_positions(mesh, n, _lib.dptr(out), _lib.stream_ptr(dev)))
        return [out[:, k].reshape(width, height) for k in range(2 * rotations)]
    import math
    from . import spec
    max_size = max(width, height)
    xv, yv = torch.meshgrid([torch.arange(width), torch.arange(height)], indexing="ij")
    if normalize:
        if spec.MESH_NORMALIZE == "max_size":
            xv = (xv / max_size - 0.5) * 2.0
            yv = (yv / max_size - 0.5) * 2.0
        else:
            xv = xv / (width - 1) * 2.0 - 1.0
            yv = yv / (height - 1) * 2.0 - 1.0
    lines = []
    for i in range(rotations):
        theta = (math.pi / 2.0) * (i / rotations)
        rx, ry = math.cos(theta), math.sin(theta)
        rs = math.fabs(rx) + math.fabs(ry)
        lines.append((rx * xv + ry * yv) / rs)
        lines.append((rx * yv - ry * xv) / rs)
    return lines
