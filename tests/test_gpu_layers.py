"""GPU parity of the single-layer kernels (through the C ABI) against the CPU oracle on the same
seeded inputs and copied weights.  Bars (BASELINE.json): segment indices / node windows
bit-exact; outputs and gradients <= 1e-5 relative (max|delta| / max|ref|); Fourier <= 1e-4."""
import pytest
import torch

import hoir_b200 as H
from oracle import hol_oracle as O
from helpers import adversarial_inputs, rel_err

pytestmark = pytest.mark.gpu

TOL = {"continuous": 1e-5, "discontinuous": 1e-5, "polynomial": 1e-5, "fourier": 1e-4}

LAYER_CASES = [
    # kind, n, in, out, segments
    ("continuous", 3, 4, 40, 100), ("continuous", 3, 40, 40, 2), ("continuous", 3, 40, 3, 2),
    ("continuous", 2, 216, 50, 50), ("continuous", 2, 50, 27, 2), ("continuous", 5, 7, 9, 13),
    ("discontinuous", 3, 4, 40, 100), ("discontinuous", 3, 40, 40, 2), ("discontinuous", 2, 2, 10, 100),
    ("discontinuous", 1, 3, 5, 8),
    ("polynomial", 3, 4, 10, 1), ("polynomial", 6, 10, 10, 1), ("polynomial", 2, 384, 16, 1),
    ("fourier", 31, 4, 40, 1), ("fourier", 31, 40, 40, 1), ("fourier", 3, 40, 3, 1), ("fourier", 1, 2, 2, 1),
]


def make_layers(kind, n, in_f, out_f, seg, **kw):
    torch.manual_seed(1234)
    ref = O.high_order_fc_layers(kind, n=n, in_features=in_f, out_features=out_f, segments=seg, **kw)
    with torch.no_grad():
        ref.w.normal_(0, 0.5)
    lay = H.high_order_fc_layers(kind, n=n, in_features=in_f, out_features=out_f, segments=seg, **kw)
    lay.load_state_dict(ref.state_dict())
    return ref, lay.cuda()


@pytest.mark.parametrize("kind,n,in_f,out_f,seg", LAYER_CASES)
@pytest.mark.parametrize("B", [1, 257])
def test_layer_forward_backward_parity(kind, n, in_f, out_f, seg, B):
    ref, lay = make_layers(kind, n, in_f, out_f, seg)
    g = torch.Generator().manual_seed(B)
    x = (torch.rand(B, in_f, generator=g) * 2.4 - 1.2)
    gy = torch.randn(B, out_f, generator=g)
    xr = x.clone().requires_grad_(True)
    yr = ref(xr)
    yr.backward(gy)
    xc = x.cuda().requires_grad_(True)
    yc = lay(xc)
    yc.backward(gy.cuda())
    tol = TOL[kind]
    assert yc.shape == yr.shape
    assert rel_err(yc, yr) <= tol
    # n == 1: the basis is constant, the oracle's autograd leaves x.grad unset -> zero gradient
    gx_ref = xr.grad if xr.grad is not None else torch.zeros_like(x)
    assert rel_err(xc.grad, gx_ref) <= tol
    assert rel_err(lay.w.grad, ref.w.grad) <= tol


@pytest.mark.parametrize("kind", ["continuous", "discontinuous"])
@pytest.mark.parametrize("n,seg", [(3, 100), (3, 2), (2, 50), (4, 7), (3, 64)])
def test_segment_index_bit_exact(kind, n, seg):
    ref, lay = make_layers(kind, n, 1, 3, seg)
    g = torch.Generator().manual_seed(7)
    x = torch.cat([adversarial_inputs(seg), torch.rand(20000, generator=g) * 2.6 - 1.3]).view(-1, 1)
    id_ref = ref.segment_index(x)
    wid_ref = id_ref * ref._stride
    xl_ref = ref.local_coordinate(x, id_ref)
    idm, wid, xl = H.segment_index(lay, x.cuda())
    assert torch.equal(idm.cpu().long(), id_ref)
    assert torch.equal(wid.cpu().long(), wid_ref)
    assert torch.equal(xl.cpu(), xl_ref)            # same roundings -> bit-identical local coordinate


@pytest.mark.parametrize("kind,n,seg", [("continuous", 3, 100), ("discontinuous", 3, 100), ("continuous", 2, 50)])
def test_layer_adversarial_inputs(kind, n, seg):
    ref, lay = make_layers(kind, n, 2, 5, seg)
    a = adversarial_inputs(seg)
    x = torch.stack([a, a.flip(0)], dim=1)
    y_ref = ref(x)
    y = lay(x.cuda())
    # extrapolation outside [-1, 1] grows the values; compare relative to the reference scale
    assert rel_err(y, y_ref) <= 1e-5


def test_layer_options_rescale_periodicity_scale():
    for kw in (dict(rescale_output=True), dict(periodicity=2.0), dict(scale=4.0)):
        for kind in ("continuous", "fourier", "polynomial"):
            ref, lay = make_layers(kind, 3, 5, 6, 4, **kw)
            x = torch.rand(300, 5, generator=torch.Generator().manual_seed(3)) * 6 - 3
            xr = x.clone().requires_grad_(True)
            yr = ref(xr)
            yr.sum().backward()
            xc = x.cuda().requires_grad_(True)
            yc = lay(xc)
            yc.sum().backward()
            tol = 1e-4 if kind == "fourier" else 2e-5
            assert rel_err(yc, yr) <= tol, (kw, kind)
            assert rel_err(lay.w.grad, ref.w.grad) <= tol, (kw, kind)
            assert rel_err(xc.grad, xr.grad) <= tol, (kw, kind)


def test_empty_batch():
    _, lay = make_layers("continuous", 3, 4, 40, 100)
    y = lay(torch.empty(0, 4, device="cuda"))
    assert y.shape == (0, 40)
    norm = H.MaxAbsNormalization()
    assert norm(torch.empty(0, 40, device="cuda")).shape == (0, 40)


@pytest.mark.parametrize("B,width", [(1, 40), (1000, 40), (33, 10), (64, 50), (5, 3), (4099, 40), (2048, 50), (1500, 64),
                                     (1024, 10), (3000, 27)])
def test_maxabs_parity(B, width):
    g = torch.Generator().manual_seed(B + width)
    x = torch.randn(B, width, generator=g)
    x[0, :] = 0.0                                        # all-zero row: y = 0 / eps-denominator
    if B > 2:
        x[1, 1] = x[1, 0] = 3.0                          # tie: gradient goes to the first index
        x[2, 2] = -5.0
    gy = torch.randn(B, width, generator=g)
    ref, mod = O.MaxAbsNormalization(), H.MaxAbsNormalization()
    xr = x.clone().requires_grad_(True)
    yr = ref(xr)
    yr.backward(gy)
    xc = x.cuda().requires_grad_(True)
    yc = mod(xc)
    yc.backward(gy.cuda())
    assert torch.equal(yc.cpu(), yr.detach())            # a division: bit-exact
    assert rel_err(xc.grad, xr.grad) <= 1e-5


@pytest.mark.parametrize("rows,cols,rot", [(4, 3, 2), (64, 48, 1), (37, 91, 2), (16, 16, 4)])
def test_positions_from_mesh_bit_exact(rows, cols, rot):
    for normalize in (True, False):
        ref = O.positions_from_mesh(rows, cols, rotations=rot, normalize=normalize)
        got = H.positions_from_mesh(rows, cols, device="cuda", rotations=rot, normalize=normalize)
        host = H.positions_from_mesh(rows, cols, rotations=rot, normalize=normalize)
        assert len(got) == len(ref) == 2 * rot
        for a, b, c in zip(got, ref, host):
            assert a.shape == (rows, cols)
            assert torch.equal(a.cpu(), b.to(torch.float32))
            assert torch.equal(c, b)


@pytest.mark.parametrize("kind,n,in_f,out_f,seg,B", [
    ("continuous", 2, 24, 50, 50, 5000), ("continuous", 3, 4, 40, 100, 5000), ("discontinuous", 3, 5, 33, 17, 5000),
    ("continuous", 2, 216, 50, 50, 5000), ("discontinuous", 2, 7, 64, 9, 5000),
    # > 64 outputs: output blocks through the slab kernels
    ("continuous", 3, 56, 100, 64, 5000), ("continuous", 2, 9, 70, 16, 5000), ("discontinuous", 3, 6, 131, 12, 5000),
    # >= 2^14 samples: the forward also goes through the streamed-slab kernel
    ("continuous", 3, 4, 40, 100, 17000), ("discontinuous", 2, 7, 64, 9, 16384), ("continuous", 3, 12, 100, 64, 17001),
    ("continuous", 2, 9, 70, 16, 16999),
    # slabs so large that only 1-2 inputs fit per CTA (odd segment count, discontinuous)
    ("continuous", 2, 7, 64, 151, 5000), ("discontinuous", 3, 5, 40, 150, 5000), ("continuous", 3, 3, 50, 200, 4500)])
def test_large_batch_many_segments_gradient(kind, n, in_f, out_f, seg, B):
    """B >= 4096 with many segments routes dL/dw through the shared-memory slab kernel and, from 2^14 samples, the
    forward through the streamed-slab kernel (slab_kernels.cu)."""
    ref, lay = make_layers(kind, n, in_f, out_f, seg)
    g = torch.Generator().manual_seed(B)
    x = torch.rand(B, in_f, generator=g) * 2.2 - 1.1
    gy = torch.randn(B, out_f, generator=g)
    xr = x.clone().requires_grad_(True)
    yr = ref(xr)
    yr.backward(gy)
    xc = x.cuda().requires_grad_(True)
    yc = lay(xc)
    yc.backward(gy.cuda())
    assert rel_err(yc, yr) <= 1e-5
    assert rel_err(xc.grad, xr.grad) <= 1e-5
    assert rel_err(lay.w.grad, ref.w.grad) <= 1e-5


DENSE_CASES = [
    # layers whose densified basis is a dense contraction: tcgen05 path (B >= 512, out <= 64, length 2)
    ("fourier", 31, 4, 40, 1), ("fourier", 31, 40, 40, 1), ("fourier", 3, 40, 3, 1), ("fourier", 7, 9, 5, 1),
    ("fourier", 16, 3, 64, 1),
    ("continuous", 2, 50, 50, 2), ("continuous", 2, 50, 27, 2), ("continuous", 3, 40, 40, 2), ("continuous", 3, 40, 3, 2),
    ("discontinuous", 3, 40, 40, 2), ("discontinuous", 2, 11, 17, 2),
    # wider than 64 outputs: cut into output blocks (row-strided y / dL/dy, accumulated dL/dx)
    ("continuous", 3, 100, 100, 2), ("continuous", 3, 40, 65, 2), ("fourier", 7, 9, 130, 1), ("discontinuous", 2, 30, 101, 2),
]


@pytest.mark.parametrize("kind,n,in_f,out_f,seg", DENSE_CASES)
@pytest.mark.parametrize("B", [512, 1531])
def test_dense_tensor_core_layer_parity(kind, n, in_f, out_f, seg, B):
    """The tcgen05 single-layer kernels (dense_tc.cu: forward, dL/dx, dL/dw; tf32 x 3 passes) against the oracle;
    ragged batch (not a multiple of the 256-sample tile), inputs beyond [-1, 1]."""
    ref, lay = make_layers(kind, n, in_f, out_f, seg)
    g = torch.Generator().manual_seed(B + n)
    x = torch.rand(B, in_f, generator=g) * 2.4 - 1.2
    gy = torch.randn(B, out_f, generator=g)
    xr = x.clone().requires_grad_(True)
    yr = ref(xr)
    yr.backward(gy)
    xc = x.cuda().requires_grad_(True)
    yc = lay(xc)
    yc.backward(gy.cuda())
    tol = TOL[kind]
    assert rel_err(yc, yr) <= tol
    assert rel_err(xc.grad, xr.grad) <= tol
    assert rel_err(lay.w.grad, ref.w.grad) <= tol


def test_dense_tensor_core_matches_cuda_core_kernels_at_full_size():
    """2^16 samples (too many for the oracle): the tensor-core path (B >= 512) equals the CUDA-core per-layer
    kernels run on sub-batches below the dispatch threshold, for outputs and summed gradients."""
    _, lay = make_layers("fourier", 31, 40, 40, 1)
    B, sub = 1 << 16, 500
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.rand(B, 40, device="cuda", generator=g) * 2 - 1
    gy = torch.randn(B, 40, device="cuda", generator=g)
    xc = x.clone().requires_grad_(True)
    y = lay(xc)
    y.backward(gy)
    gw_tc, gx_tc = lay.w.grad.clone(), xc.grad.clone()
    lay.w.grad = None
    idx = torch.arange(0, B, 4099, device="cuda")[:sub]          # a strided sample of rows
    xs = x[idx].clone().requires_grad_(True)
    ys = lay(xs)                                                   # 16 rows < 512: CUDA-core kernels
    ys.backward(gy[idx])
    assert rel_err(y[idx], ys) <= 2e-5
    assert rel_err(gx_tc[idx], xs.grad) <= 2e-5
    # linearity: gradient of the full batch = sum over its halves
    lay.w.grad = None
    for h in (slice(0, B // 2), slice(B // 2, B)):
        xh = x[h].clone().requires_grad_(True)
        lay(xh).backward(gy[h])
    assert rel_err(lay.w.grad, gw_tc) <= 2e-5


def _fuzz_cases(count=48, seed=20261020):
    """Random layer shapes around the dispatch boundaries of hoir_layer_forward/backward: batch 512 / 4096 / 2^14
    (tensor-core, slab dL/dw, slab forward), outputs 64 / 65 (output blocks), segments 2 / 8 / 9 (dense vs gather)."""
    import random
    rnd = random.Random(seed)
    cases = []
    for _ in range(count):
        kind = rnd.choice(["continuous", "continuous", "discontinuous", "fourier", "polynomial"])
        n = rnd.choice([2, 3, 3, 4, 6]) if kind != "fourier" else rnd.choice([1, 3, 5, 8, 17, 31])
        seg = 1 if kind in ("fourier", "polynomial") else rnd.choice([1, 2, 2, 3, 8, 9, 17, 64, 100])
        in_f = rnd.choice([1, 2, 3, 4, 7, 16, 33, 50])
        out_f = rnd.choice([1, 3, 10, 31, 40, 63, 64, 65, 72, 100, 129])
        B = rnd.choice([2, 33, 511, 512, 513, 1000, 4095, 4096, 4100, 16383, 16384, 16390])
        nodes = n if kind in ("fourier", "polynomial") else ((n - 1) * seg + 1 if kind == "continuous" else n * seg)
        if B * in_f * out_f * n > 60_000_000:          # the oracle materialises the gathered weights [B, in, out, n]
            B = max(2, 60_000_000 // (in_f * out_f * n))
        cases.append((kind, n, in_f, out_f, seg, B))
    return cases


@pytest.mark.parametrize("kind,n,in_f,out_f,seg,B", _fuzz_cases())
def test_layer_parity_fuzz(kind, n, in_f, out_f, seg, B):
    ref, lay = make_layers(kind, n, in_f, out_f, seg)
    g = torch.Generator().manual_seed(B * 31 + in_f)
    x = torch.rand(B, in_f, generator=g) * 2.3 - 1.15
    gy = torch.randn(B, out_f, generator=g)
    xr = x.clone().requires_grad_(True)
    yr = ref(xr)
    yr.backward(gy)
    xc = x.cuda().requires_grad_(True)
    yc = lay(xc)
    yc.backward(gy.cuda())
    tol = TOL[kind]
    assert rel_err(yc, yr) <= tol
    gx_ref = xr.grad if xr.grad is not None else torch.zeros_like(x)
    assert rel_err(xc.grad, gx_ref) <= tol
    assert rel_err(lay.w.grad, ref.w.grad) <= tol
