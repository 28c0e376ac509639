"""compute-sanitizer is not available on the GPU pool: instead the SAME sources are built with -DHOIR_DEBUG_BOUNDS
(csrc/Makefile `debug` -> libhoir_b200_dbg.so: device asserts on every swizzled shared-memory offset and TMEM column range
of the tensor-core kernels) and the tensor-core paths are run once through that build in a subprocess
(HOIR_LIB=...): a violated assert prints its file:line and traps, which fails this test."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DBG = os.path.join(ROOT, "high-order-implicit-representation_b200", "libhoir_b200_dbg.so")

SCRIPT = r'''
import sys
sys.path.insert(0, %(root)r)
sys.path.insert(0, %(root)r + "/tests")
import torch
import hoir_b200 as H
from hoir_b200 import _lib
from oracle import hol_oracle as O
from helpers import MLP_KW, make_pair, rel_err
assert _lib.LIB_PATH.endswith("libhoir_b200_dbg.so"), _lib.LIB_PATH
for name in ("C2", "C3d", "C2r1"):
    ref, net = make_pair(name)
    kw = MLP_KW[name]
    g = torch.Generator().manual_seed(1)
    x = torch.rand(3000, kw["in_width"], generator=g) * 2.2 - 1.1
    y = torch.rand(3000, 3, generator=g) * 2 - 1
    tr = H.FusedTrainer(net, optimizer="adam", lr=1e-3)
    loss = tr.train_step(x.cuda(), y.cuda()).item()                 # tensor-core training kernel, tail in-kernel
    want = O.mse_training_step(ref, x, y).item()
    assert abs(loss - want) <= 1e-5 * abs(want), (name, loss, want)
    tr.sort_min_batch, tr.batch_order = 0, "sorted"
    tr.train_step(x.cuda(), y.cuda())                               # cell-sorted visiting order
    out = tr.render(40, 52, kw["in_width"] // 2)                     # tensor-core forward kernel (2 CTAs / SM)
    assert torch.isfinite(out).all()
    xl = torch.rand(70000, kw["in_width"], generator=g) * 2 - 1
    tr.batch_order = "two_pass"
    tr.train_step(xl.cuda(), torch.zeros(70000, 3).cuda())          # two-pass workspace path
# Fourier whole-network kernel (tcgen05 forward / dL/dx / dL/dw phases)
kw = MLP_KW["C3f"]
torch.manual_seed(0)
ref = O.HighOrderMLP(normalization=O.MaxAbsNormalization, **kw)
net = H.HighOrderMLP(normalization=H.MaxAbsNormalization, **kw)
net.load_state_dict(ref.state_dict())
tr = H.FusedTrainer(net.cuda(), optimizer="adam", lr=1e-3)
g = torch.Generator().manual_seed(2)
x = torch.rand(70000, 4, generator=g) * 2 - 1
y = torch.rand(70000, 3, generator=g) * 2 - 1
loss = tr.train_step(x[:3000].cuda(), y[:3000].cuda()).item()
want = O.mse_training_step(ref, x[:3000], y[:3000]).item()
assert abs(loss - want) <= 1e-4 * abs(want), (loss, want)
tr.train_step(x.cuda(), y.cuda())                                   # > 64 units per CTA slice: the mid-slice flush path
assert torch.isfinite(tr.forward(x.cuda())).all()
# per-layer tensor-core kernels (dense_tc.cu) under autograd
net2 = H.HighOrderMLP(normalization=H.MaxAbsNormalization, **MLP_KW["C4s"]).cuda()
xs = (torch.rand(5000, 48, generator=g) * 2 - 1).cuda()
net2(xs).square().mean().backward()
# fused interpolator forward (layer-0 ring + tcgen05 layers, csrc/fused_interp.cu): explicit rows and the one-launch frame
with torch.no_grad():
    fused = net2(xs)
    net2.use_fused = False
    assert rel_err(fused, net2(xs)) <= 1e-5
    net2.use_fused = True
    img = (torch.rand(3, 70, 45, generator=g) * 2 - 1).cuda()
    fr = H.neighborhood_sample_generator(net2, img, 3, 1, alpha=0.5, prev=img)
    net2.use_fused = False
    assert rel_err(fr, H.neighborhood_sample_generator(net2, img, 3, 1, alpha=0.5, prev=img)) <= 1e-5
# dL/dw of a many-segment layer at large batch on the tensor cores (csrc/gw0_tc.cu)
lay = H.high_order_fc_layers("continuous", n=2, in_features=48, out_features=27, segments=50).cuda()
xg = (torch.rand(9000, 48, generator=g) * 2.2 - 1.1).cuda().requires_grad_(True)
lay(xg).square().mean().backward()
assert torch.isfinite(lay.w.grad).all()
torch.cuda.synchronize()
print("DEBUG_BOUNDS_OK")
'''


def test_tensor_core_paths_under_debug_bounds_build():
    if not os.path.exists(DBG):
        pytest.skip("libhoir_b200_dbg.so not built (python -c 'import __graft_entry__ as g; g.build()')")
    env = dict(os.environ, HOIR_LIB=DBG)
    res = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}], env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "DEBUG_BOUNDS_OK" in res.stdout, (res.stdout[-2000:], res.stderr[-3000:])
    assert "HOIR_DEBUG_BOUNDS violated" not in res.stdout
