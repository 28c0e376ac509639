"""Generates tests/golden/layers_v1.npz: small seeded input/weight/output vectors of the ORACLE
(oracle/hol_oracle.py) for every layer kind and for the C1-shaped network.

The real package (high-order-layers-torch ^2.4.2) cannot be imported in this sandbox, so these
vectors pin the oracle against silent regressions -- they do NOT pin it against the package
("parity unpinned").  If the package ever becomes importable, regenerate with
`--package` to diff.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import hol_oracle as O  # noqa: E402

CASES = [("continuous", 3, 2, 3, 4), ("continuous", 2, 3, 2, 5), ("discontinuous", 3, 2, 3, 4),
         ("polynomial", 4, 2, 3, 1), ("fourier", 5, 2, 3, 1)]


def main():
    out = {}
    g = torch.Generator().manual_seed(20261019)
    for kind, n, in_f, out_f, seg in CASES:
        layer = O.high_order_fc_layers(kind, n=n, in_features=in_f, out_features=out_f, segments=seg)
        with torch.no_grad():
            layer.w.copy_(torch.randn(layer.w.shape, generator=g) * 0.5)
        x = (torch.rand(16, in_f, generator=g) * 2.4 - 1.2).requires_grad_(True)
        gy = torch.randn(16, out_f, generator=g)
        y = layer(x)
        y.backward(gy)
        key = f"{kind}_n{n}"
        out[key + "_w"] = layer.w.detach().numpy()
        out[key + "_x"] = x.detach().numpy()
        out[key + "_gy"] = gy.numpy()
        out[key + "_y"] = y.detach().numpy()
        out[key + "_gx"] = x.grad.numpy()
        out[key + "_gw"] = layer.w.grad.numpy()
        if kind in ("continuous", "discontinuous"):
            out[key + "_id"] = layer.segment_index(x.detach()).numpy()
    kw = dict(layer_type="continuous", n=3, in_width=2, out_width=3, hidden_layers=2, hidden_width=10,
              in_segments=100, out_segments=2, hidden_segments=2)
    torch.manual_seed(5)
    net = O.HighOrderMLP(normalization=O.MaxAbsNormalization, **kw)
    O.initialize_network_polynomial_layers(net, 1.0, 0.0, generator=g)
    x = torch.rand(32, 2, generator=g) * 2 - 1
    y = torch.rand(32, 3, generator=g) * 2 - 1
    loss = O.mse_training_step(net, x, y)
    loss.backward()
    for k, p in net.named_parameters():
        out["c1_" + k] = p.detach().numpy()
        out["c1_grad_" + k] = p.grad.numpy()
    out["c1_x"], out["c1_y"] = x.numpy(), y.numpy()
    out["c1_out"] = net(x).detach().numpy()
    out["c1_loss"] = np.array(loss.item(), dtype=np.float64)
    mesh = O.positions_from_mesh(4, 3, rotations=2, normalize=True)
    out["mesh_4x3_r2"] = torch.stack(mesh).numpy()
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "layers_v1.npz"), **out)
    print("wrote layers_v1.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
