"""CPU-side checks of the C ABI: the library loads, exports every symbol include/hoir.h declares,
and its argument validation / error mapping works without touching a GPU."""
import ctypes as C
import os
import re

import pytest

import hoir_b200 as H
from hoir_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "hoir.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hoir_[a-z0-9_]+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 18
    for name in names:
        assert hasattr(lib, name), f"libhoir_b200.so does not export {name}"
    assert set(names) == set(_lib.EXPORTED_SYMBOLS)
    assert lib.hoir_version() == 106


def test_struct_layouts_match_header():
    assert C.sizeof(_lib.LayerDesc) == 40
    assert C.sizeof(_lib.MlpDesc) == 4 + 8 * 40 + 8
    assert C.sizeof(_lib.MeshDesc) == 16 + 8 + 3 * 8 * 4


@pytest.mark.parametrize("kind,n,seg,nodes", [("continuous", 3, 100, 201), ("discontinuous", 3, 100, 300),
                                              ("polynomial", 5, 7, 5), ("fourier", 31, 1, 31),
                                              ("continuous", 2, 50, 51)])
def test_layer_nodes(kind, n, seg, nodes):
    d = _lib.make_layer_desc(kind, n, seg, 4, 40)
    assert _lib.lib().hoir_layer_nodes(d) == nodes


def test_packed_layout_sizes_c2():
    net = H.HighOrderMLP(layer_type="continuous", n=3, in_width=4, out_width=3, hidden_layers=1,
                         hidden_width=40, in_segments=100, out_segments=2, hidden_segments=2,
                         normalization=H.MaxAbsNormalization)
    assert sum(p.numel() for p in net.parameters()) == 40760          # README.md:35
    desc = net.mlp_desc()
    lib = _lib.lib()
    assert lib.hoir_mlp_packed_offset(desc, 0) == 0
    assert lib.hoir_mlp_packed_offset(desc, 1) == 4 * 201 * 40
    assert lib.hoir_mlp_packed_offset(desc, 2) == 4 * 201 * 40 + 40 * 5 * 40
    # + the half-major second copy of the layer-0 weights the tensor-core training kernel gathers from
    assert lib.hoir_mlp_packed_size(desc) == 4 * 201 * 40 + 40 * 5 * 40 + 40 * 5 * 4 + 4 * 201 * 40
    assert lib.hoir_mlp_fused_supported(desc) == 1


def test_fused_supported_matrix():
    def sup(**kw):
        base = dict(layer_type="continuous", n=3, in_width=4, out_width=3, hidden_layers=1, hidden_width=40,
                    in_segments=100, out_segments=2, hidden_segments=2, normalization=H.MaxAbsNormalization)
        base.update(kw)
        return H.HighOrderMLP(**base).fused_supported()
    assert sup()
    assert sup(layer_type="discontinuous")
    assert sup(in_width=2, hidden_width=10, hidden_layers=2)
    assert sup(in_segments=37)                      # layer-0 segments are a runtime value
    assert sup(layer_type="fourier")                # the fused Fourier family (csrc/fused_fourier.cu)
    assert sup(layer_type="fourier", n_in=31, n_hidden=31)
    assert not sup(layer_type="fourier", hidden_width=41)
    assert not sup(layer_type="fourier", hidden_layers=2)
    assert not sup(layer_type="fourier", n_out=7)
    assert not sup(hidden_width=41)
    assert not sup(hidden_segments=3)
    assert not sup(periodicity=2.0)


def test_forward_supported_matrix():
    """hoir_mlp_forward_supported: 2 = whole training step fused, 1 = forward only (the neighborhood-interpolator family,
    csrc/fused_interp.cu), 0 = per-layer kernels.  Host logic: no GPU needed."""
    lib = _lib.lib()

    def sup(**kw):
        base = dict(layer_type="continuous", n=3, in_width=216, out_width=27, hidden_layers=2, hidden_width=50,
                    in_segments=50, out_segments=2, hidden_segments=2, normalization=H.MaxAbsNormalization)
        base.update(kw)
        return lib.hoir_mlp_forward_supported(H.HighOrderMLP(**base).mlp_desc())
    assert sup() == 1 and sup(n=2) == 1 and sup(layer_type="discontinuous", in_segments=32) == 1
    assert sup(hidden_layers=3) == 1 and sup(hidden_width=64) == 1 and sup(out_width=32) == 1
    assert sup(hidden_segments=1) == 1 and sup(in_segments=1, out_segments=1, hidden_width=10, hidden_layers=1) == 1   # stock yaml
    assert sup(hidden_width=65) == 0 and sup(out_width=33) == 0 and sup(hidden_layers=4) == 0
    assert sup(hidden_segments=3) == 0 and sup(n=4) == 0 and sup(layer_type="fourier") == 0 and sup(in_segments=200) == 0
    assert sup(in_width=4, out_width=3, hidden_layers=1, hidden_width=40, in_segments=100) == 2                  # C2
    # the thresholds of the fused forward follow the size of the input-layer table
    big = H.HighOrderMLP(layer_type="continuous", n=2, in_width=216, out_width=27, hidden_layers=2, hidden_width=50,
                         in_segments=50, out_segments=2, hidden_segments=2)
    stock = H.HighOrderMLP(layer_type="continuous", n=3, in_width=216, out_width=27, hidden_layers=1, hidden_width=10,
                           in_segments=1, out_segments=1, hidden_segments=2)
    assert big.fused_train_min_batch == 1 and stock.fused_train_min_batch == 16384
    e0 = _lib.WEIGHT_EPOCH
    _lib.weights_changed()
    assert _lib.WEIGHT_EPOCH == e0 + 1


def test_errors_map_to_python_exceptions():
    lib = _lib.lib()
    bad = _lib.LayerDesc(7, 3, 2, 4, 4, 2.0, 1.0, 0.0, 2.0, 1)
    with pytest.raises(ValueError):
        _lib.check(lib.hoir_layer_forward(bad, 0, None, None, None, None))
    ok = _lib.make_layer_desc("continuous", 3, 2, 4, 4)
    with pytest.raises(ValueError):
        _lib.check(lib.hoir_layer_forward(ok, -1, None, None, None, None))
    # B == 0 is legal and touches no pointer (the reference's render loops feed an empty batch)
    _lib.check(lib.hoir_layer_forward(ok, 0, None, None, None, None))
    _lib.check(lib.hoir_layer_backward(ok, 0, None, None, None, None, None, None))
    _lib.check(lib.hoir_maxabs_forward(0, 4, 1e-6, None, None, None))
    with pytest.raises(ValueError, match="not recognized"):
        H.high_order_fc_layers("no_such_layer", n=3, in_features=2, out_features=2)
    unfused = H.HighOrderMLP(layer_type="fourier", n=3, in_width=4, out_width=3, hidden_layers=2,
                             hidden_width=40, normalization=H.MaxAbsNormalization)
    with pytest.raises(H.HoirUnsupported):
        _lib.check(lib.hoir_mlp_forward(unfused.mlp_desc(), 0, None, None, None, None, 0, None))


def test_no_cpu_fallback():
    import torch
    layer = H.high_order_fc_layers("continuous", n=3, in_features=2, out_features=2, segments=4)
    with pytest.raises(RuntimeError, match="no CPU path"):
        layer(torch.zeros(3, 2))


def test_dependency_alias_package_exports_what_the_reference_imports():
    """The reference reaches the hot path through `high_order_layers_torch.*` imports; the alias
    package under compat/ must satisfy every one of them, star imports included (Quirk Q7).  It is NOT importable
    from the repository root alone (a real install of the dependency must never be shadowed by accident)."""
    import importlib.util
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    assert not os.path.exists(os.path.join(root, "high_order_layers_torch"))
    spec = importlib.util.find_spec("high_order_layers_torch")
    if spec is not None and not (spec.origin or "").startswith(os.path.join(root, "compat")):
        pytest.skip("a real high_order_layers_torch is installed: the alias is not exercised here")
    sys.path.insert(0, os.path.join(root, "compat"))
    ns = {}
    exec("from high_order_layers_torch.layers import *\n"
         "from high_order_layers_torch.networks import *\n"
         "from high_order_layers_torch.sparse_optimizers import SparseLion\n"
         "from high_order_layers_torch.layers import high_order_fc_layers\n"
         "from high_order_layers_torch.utils import positions_from_mesh\n", ns)
    for name in ("high_order_fc_layers", "MaxAbsNormalization", "HighOrderMLP", "initialize_network_polynomial_layers",
                 "SparseLion", "positions_from_mesh", "torch", "math", "Tensor", "Any"):
        assert name in ns, name
    import hoir_b200 as H
    assert ns["HighOrderMLP"] is H.HighOrderMLP and ns["high_order_fc_layers"] is H.high_order_fc_layers
    # host-side use exactly as single_image_dataset.py:39-45 does (CPU tensors: dataset preparation)
    lines = ns["positions_from_mesh"](width=4, height=3, device="cpu", rotations=2, normalize=True)
    assert len(lines) == 4 and lines[0].shape == (4, 3)
