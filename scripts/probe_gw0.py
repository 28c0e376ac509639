"""dL/dw of the interpolator's input layer: timing and error against the CPU oracle (fp64) of whichever kernel
hoir_layer_backward dispatches to (HOIR_NO_GW0_TC=1: slab_gw_kernel; default: the tcgen05 kernel gw0_tc.cu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hoir_b200 as H
from hoir_b200 import _lib
from oracle import hol_oracle as O
dev = torch.device("cuda:0")
lib = _lib.lib()
def timeit(fn, reps=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for kind, n, in_f, out_f, seg in (("continuous", 2, 216, 50, 50), ("continuous", 3, 216, 50, 50), ("discontinuous", 2, 48, 27, 20),
                                  ("discontinuous", 3, 100, 64, 40), ("continuous", 3, 224, 52, 64)):
    torch.manual_seed(0)
    ref = O.high_order_fc_layers(kind, n=n, in_features=in_f, out_features=out_f, segments=seg).double()
    layer = H.high_order_fc_layers(kind, n=n, in_features=in_f, out_features=out_f, segments=seg).to(dev)
    with torch.no_grad():
        layer.w.copy_(ref.w.float())
    for B in (8192 + 17, 1 << 16):
        g = torch.Generator().manual_seed(B)
        x = torch.rand(B, in_f, generator=g) * 2.2 - 1.1
        gy = torch.randn(B, out_f, generator=g)
        xd, gyd = x.to(dev), gy.to(dev)
        gw = torch.zeros_like(layer.w)
        def f():
            _lib.check(lib.hoir_layer_backward(layer.desc, B, _lib.dptr(xd), _lib.dptr(layer.w), _lib.dptr(gyd), None, _lib.dptr(gw), _lib.stream_ptr(dev)))
        f(); torch.cuda.synchronize()
        err = None
        if B < 10000:
            xr = x.double().requires_grad_(False)
            ref.zero_grad()
            (ref(xr) * gy.double()).sum().backward()
            want = ref.w.grad
            err = ((gw.double().cpu() - want).abs().max() / want.abs().max()).item()
        ms = timeit(f)
        print(f"{kind} n={n} {in_f}->{out_f} seg {seg} B={B}: {ms:.3f} ms" + (f"  rel err vs fp64 oracle {err:.2e}" if err is not None else ""))
