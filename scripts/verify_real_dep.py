#!/usr/bin/env python
"""SURVEY.md A.9 -- run this the day ``high-order-layers-torch`` becomes importable (a wheel in /opt/wheelhouse,
``pip install --target baseline/_ref``, ...).  It

  1. finds the REAL package (never the compat/ alias), prints its version and location;
  2. diffs it against the restated oracle (oracle/hol_oracle.py) in the order of A.9: ``w.shape`` per layer type,
     ``HighOrderMLP.state_dict()`` keys, outputs / gradients on fixed inputs with COPIED weights for all four layer
     types (for Fourier and the mesh normalisation it also tries the alternative settings of ``spec.py`` and names the
     one that matches), ``MaxAbsNormalization`` eps, ``positions_from_mesh(4, 3, rotations=2, normalize=True)``,
     one ``SparseLion`` step on a sparse gradient, ``initialize_network_polynomial_layers`` statistics;
  3. writes REAL golden vectors to tests/golden/real_dep_v1.npz (tests/test_real_dep_parity.py then pins the oracle
     on the CPU and the CUDA kernels on the GPU box, where the package need not exist).

Exit code 0: everything matches (parity pinned); 1: mismatches listed; 3: package not found (today's state:
"parity unpinned")."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import hol_oracle as O      # noqa: E402
from oracle import real_dep             # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "real_dep_v1.npz")


def rel(a, b):
    d = b.abs().max().item()
    return (a - b).abs().max().item() / (d if d > 0 else 1.0)


def main() -> int:
    real = real_dep.find()
    if real is None:
        print("high_order_layers_torch: NOT FOUND (repo root and compat/ excluded from the search; baseline/_ref "
              "checked) -> layer math stays 'parity unpinned'")
        return 3
    from high_order_layers_torch import layers as RL, networks as RN, sparse_optimizers as RS, utils as RU
    print(f"real package: version {getattr(real, '__version__', '?')} at {real.__file__}")
    bad, gold = [], {}

    # -- layers: shapes, outputs, gradients with copied weights ----------------------------------------------------
    for name, lt, kw in real_dep.LAYER_CASES:
        r = real_dep.run_layer(RL, lt, kw)
        for k, v in r.items():
            gold[f"layer/{name}/{k}"] = v.numpy()
        try:
            o = real_dep.run_layer(O, lt, kw, w=r["w"])
        except Exception as e:  # noqa: BLE001
            bad.append(f"{name}: oracle cannot take the real weights ({e}); real w.shape={tuple(r['w'].shape)}")
            continue
        tol = 1e-4 if lt == "fourier" else 1e-5
        errs = {k: rel(o[k], r[k]) for k in ("y", "gx", "gw")}
        ok = all(e <= tol for e in errs.values())
        print(f"  {name:24s} w{tuple(r['w'].shape)}  " + "  ".join(f"{k} {e:.2e}" for k, e in errs.items()),
              "ok" if ok else "MISMATCH")
        if not ok and lt == "fourier":
            for scale in (2.0, 1.0):
                for odd_sin in (True, False):
                    O.FOURIER_ARG_SCALE, O.FOURIER_ODD_IS_SIN = scale, odd_sin
                    o2 = real_dep.run_layer(O, lt, kw, w=r["w"])
                    if rel(o2["y"], r["y"]) <= tol:
                        print(f"    -> matches with FOURIER_ARG_SCALE={scale}, FOURIER_ODD_IS_SIN={odd_sin}: set both in "
                              "high-order-implicit-representation_b200/spec.py and at the top of oracle/hol_oracle.py")
        if not ok:
            bad.append(f"{name}: {errs}")

    # -- HighOrderMLP: state_dict keys, forward ------------------------------------------------------------------------
    kw = dict(layer_type="continuous", n=3, in_width=4, out_width=3, hidden_layers=1, hidden_width=40,
              in_segments=100, out_segments=2, hidden_segments=2)
    torch.manual_seed(0)
    rn = RN.HighOrderMLP(normalization=RL.MaxAbsNormalization, **kw)
    on = O.HighOrderMLP(normalization=O.MaxAbsNormalization, **kw)
    rk, ok_ = list(rn.state_dict().keys()), list(on.state_dict().keys())
    print("  HighOrderMLP state_dict keys:", rk, "ok" if rk == ok_ else f"MISMATCH (oracle {ok_})")
    if rk != ok_:
        bad.append("state_dict keys")
    else:
        on.load_state_dict(rn.state_dict())
        x = torch.rand(256, 4, generator=torch.Generator().manual_seed(5)) * 2 - 1
        e = rel(on(x), rn(x))
        gold["mlp/x"], gold["mlp/y"] = x.numpy(), rn(x).detach().numpy()
        for k, v in rn.state_dict().items():
            gold[f"mlp/w/{k}"] = v.numpy()
        print(f"  HighOrderMLP C2 forward: {e:.2e}", "ok" if e <= 1e-5 else "MISMATCH")
        if e > 1e-5:
            bad.append(f"HighOrderMLP forward {e}")
    eps = getattr(RL.MaxAbsNormalization(), "_eps", None)
    print("  MaxAbsNormalization eps:", eps, "ok" if eps == O.MAXABS_EPS else f"MISMATCH (oracle {O.MAXABS_EPS})")
    if eps != O.MAXABS_EPS:
        bad.append("MaxAbs eps")

    # -- positions_from_mesh -------------------------------------------------------------------------------------------
    rp = torch.stack(RU.positions_from_mesh(4, 3, rotations=2, normalize=True))
    gold["mesh/4x3r2"] = rp.numpy()
    match = None
    for mode in ("max_size", "size_minus_1"):
        O.MESH_NORMALIZE = mode
        if torch.equal(torch.stack(O.positions_from_mesh(4, 3, rotations=2, normalize=True)), rp):
            match = mode
    print("  positions_from_mesh(4,3,rotations=2,normalize=True): matches MESH_NORMALIZE =", match)
    if match is None:
        bad.append("positions_from_mesh: neither normalisation matches")

    # -- SparseLion one step on a sparse gradient ----------------------------------------------------------------------
    p0 = torch.rand(50, generator=torch.Generator().manual_seed(7))
    g0 = torch.rand(50, generator=torch.Generator().manual_seed(8)) - 0.5
    g0[::3] = 0.0
    outs = []
    for mod in (RS, O):
        p = torch.nn.Parameter(p0.clone())
        opt = mod.SparseLion([p], lr=1e-2)
        for _ in range(2):
            p.grad = g0.clone()
            opt.step()
        outs.append(p.detach().clone())
    gold["lion/p0"], gold["lion/g"], gold["lion/p2"] = p0.numpy(), g0.numpy(), outs[0].numpy()
    e = (outs[0] - outs[1]).abs().max().item()
    print(f"  SparseLion, two steps on a sparse gradient: max diff {e:.2e}", "ok" if e <= 1e-7 else "MISMATCH")
    if e > 1e-7:
        bad.append("SparseLion")

    np.savez_compressed(OUT, **gold)
    print("wrote", OUT, f"({len(gold)} arrays)")
    print("PARITY PINNED" if not bad else "MISMATCHES:\n  " + "\n  ".join(bad))
    return 0 if not bad else 1


if __name__ == "__main__":
    sys.exit(main())
