"""Shared helpers of the parity tests: build the same network in the oracle (CPU, torch ops) and
in hoir_b200 (CUDA kernels) with COPIED weights (never by seed)."""
import torch

from oracle import hol_oracle as O

MLP_KW = {
    # BASELINE.json configs: C1, C2, C3, C4-shaped
    "C1": dict(layer_type="continuous", n=3, in_width=2, out_width=3, hidden_layers=2, hidden_width=10,
               in_segments=100, out_segments=2, hidden_segments=2),
    "C1d": dict(layer_type="discontinuous", n=3, in_width=2, out_width=3, hidden_layers=2, hidden_width=10,
                in_segments=100, out_segments=2, hidden_segments=2),
    "C1r2": dict(layer_type="continuous", n=3, in_width=4, out_width=3, hidden_layers=2, hidden_width=10,
                 in_segments=100, out_segments=2, hidden_segments=2),          # README.md:24 with the default rotations=2
    "C1n2": dict(layer_type="continuous", n=2, in_width=2, out_width=3, hidden_layers=2, hidden_width=10,
                 in_segments=100, out_segments=2, hidden_segments=2),
    "C2": dict(layer_type="continuous", n=3, in_width=4, out_width=3, hidden_layers=1, hidden_width=40,
               in_segments=100, out_segments=2, hidden_segments=2),
    "C2r1": dict(layer_type="continuous", n=3, in_width=2, out_width=3, hidden_layers=1, hidden_width=40,
                 in_segments=100, out_segments=2, hidden_segments=2),
    "C3d": dict(layer_type="discontinuous", n=3, in_width=4, out_width=3, hidden_layers=1, hidden_width=40,
                in_segments=100, out_segments=2, hidden_segments=2),
    "C3f": dict(layer_type="fourier", n=3, n_in=31, n_hidden=31, in_width=4, out_width=3, hidden_layers=1,
                hidden_width=40, in_segments=100, out_segments=2, hidden_segments=2),
    "C4s": dict(layer_type="continuous", n=2, in_width=48, out_width=27, hidden_layers=2, hidden_width=50,
                in_segments=50, out_segments=2, hidden_segments=2),
    "poly": dict(layer_type="polynomial", n=3, in_width=4, out_width=3, hidden_layers=2, hidden_width=10,
                 in_segments=100, out_segments=2, hidden_segments=2),
}


def rel_err(a: torch.Tensor, ref: torch.Tensor) -> float:
    a = a.detach().double().cpu()
    ref = ref.detach().double().cpu()
    denom = ref.abs().max().item()
    return (a - ref).abs().max().item() / (denom if denom > 0 else 1.0)


def make_pair(name: str, seed: int = 0, randomize: bool = True):
    """(oracle HighOrderMLP on CPU, hoir_b200 HighOrderMLP on cuda:0) with identical weights."""
    import hoir_b200 as H
    kw = MLP_KW[name]
    torch.manual_seed(seed)
    ref = O.HighOrderMLP(normalization=O.MaxAbsNormalization, **kw)
    O.initialize_network_polynomial_layers(ref, max_slope=1.0, max_offset=0.0)
    if randomize:  # straight lines have zero curvature: perturb so every node matters
        with torch.no_grad():
            for p in ref.parameters():
                p.add_(0.3 * torch.randn_like(p) / p.shape[1] ** 0.5)
    net = H.HighOrderMLP(normalization=H.MaxAbsNormalization, **kw)
    net.load_state_dict(ref.state_dict())
    return ref, net.cuda()


def adversarial_inputs(segments: int) -> torch.Tensor:
    """Exact segment boundaries, +-1, values just outside, +-0, and nextafter neighbours (8d)."""
    b = -1.0 + 2.0 * torch.arange(0, segments + 1, dtype=torch.float64) / segments
    b = b.to(torch.float32)
    inf = torch.tensor(float("inf"))
    vals = [b, torch.nextafter(b, inf), torch.nextafter(b, -inf),
            torch.tensor([-1.5, -1.2, -1.0000001, 1.0000001, 1.2, 1.5, 0.0, -0.0, 1e-30, -1e-30, 1e-8, -1e-8])]
    return torch.cat(vals)


class FixedModel(torch.nn.Module):
    """The deterministic stand-in network of tests/golden/make_reference_golden.py (same definition)."""

    def __init__(self, f_in: int, f_out: int):
        super().__init__()
        i = torch.arange(f_out, dtype=torch.float32).view(-1, 1)
        j = torch.arange(f_in, dtype=torch.float32).view(1, -1)
        self.register_buffer("w", torch.cos(0.37 * i + 0.11 * j) / f_in ** 0.5)

    def forward(self, x):
        return torch.tanh(x @ self.w.t())


def reference_golden():
    import os
    import numpy as np
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_own_v1.npz"))
