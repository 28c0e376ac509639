"""Explicit batches in arbitrary sample order -- the reference's DataLoader(shuffle=True) over image_to_dataset rows
(/root/reference/high_order_implicit_representation/single_image_dataset.py:22-51, 312-320) -- through the
cell-sorted path: hoir_mlp_batch_sort + the *_order training entries (include/hoir.h).  The batch sort only changes
the ORDER in which the kernel visits the samples; MSE(mean) and the summed gradients must equal the oracle's."""
import pytest
import torch

import hoir_b200 as H
from hoir_b200 import _lib
from oracle import hol_oracle as O
from helpers import make_pair, rel_err

pytestmark = pytest.mark.gpu


def pixel_batch(rows, cols, B, seed=0, rotations=2):
    """B shuffled pixels of a random uint8 image as the reference's dataset delivers them: (x[B, 2 rot], y[B, 3], idx)."""
    g = torch.Generator().manual_seed(seed)
    img = torch.randint(0, 256, (rows, cols, 3), generator=g, dtype=torch.uint8)
    y_all, x_all, _ = O.image_to_dataset(img, rotations=rotations)
    idx = torch.randperm(rows * cols, generator=g)[:B]
    return img, x_all[idx].contiguous(), y_all[idx].contiguous(), idx.to(torch.int32)


def cell_keys(x, seg):
    s = ((x.double() + 1.0) / 2.0 * seg).to(torch.float32)       # coherence check only: need not be bit-exact
    return s.long().clamp_(0, seg - 1)


@pytest.mark.parametrize("name,B", [("C2", 8192), ("C2", 40000), ("C2r1", 20000), ("C3d", 9000)])
def test_batch_sort_returns_a_permutation_in_cell_order(name, B):
    _, net = make_pair(name)
    tr = H.FusedTrainer(net, lr=0.0, keep_grads=True)
    rot = net.model[0].in_features // 2
    _, x, _, _ = pixel_batch(200, 300, B, seed=1, rotations=rot)
    order = torch.empty(B, dtype=torch.int32, device="cuda")
    probe = _lib.C.c_float(-1.0)
    _lib.check(_lib.lib().hoir_mlp_batch_sort(tr.desc, B, _lib.dptr(x.cuda()), None, None, _lib.dptr(order, torch.int32),
                                              _lib.C.byref(probe), _lib.stream_ptr(torch.device("cuda"))))
    o = order.cpu().long()
    assert torch.equal(o.sort().values, torch.arange(B))                      # a permutation of the rows
    k = cell_keys(x[o], 100)
    cell = k[:, 0] * 100 + k[:, 1]
    assert bool((cell[1:] >= cell[:-1]).all())                                # (segment of x0, segment of x1) ascending
    assert 0.0 <= probe.value <= 1.0


def test_coherence_probe_separates_pixel_batches_from_independent_inputs():
    """2^18 shuffled pixels of a 1024^2 image (~26 samples per layer-0 cell): coherent after the sort; independent random
    inputs of the same size stay incoherent in the rotated coordinates x2, x3."""
    _, net = make_pair("C2")
    tr = H.FusedTrainer(net, lr=0.0, keep_grads=True)
    B, n = 1 << 18, 1024
    g = torch.Generator(device="cuda").manual_seed(1)
    idx = torch.randint(0, n * n, (B,), device="cuda", generator=g)
    pos = torch.stack(H.positions_from_mesh(n, n, device="cuda", rotations=2, normalize=True), dim=2).reshape(-1, 4)
    order = torch.empty(B, dtype=torch.int32, device="cuda")
    probe = _lib.C.c_float(-1.0)
    for x, lo, hi in ((pos[idx].contiguous(), 0.0, 0.15), (torch.rand(B, 4, device="cuda", generator=g) * 2 - 1, 0.4, 1.0)):
        _lib.check(_lib.lib().hoir_mlp_batch_sort(tr.desc, B, _lib.dptr(x), None, None, _lib.dptr(order, torch.int32),
                                                  _lib.C.byref(probe), _lib.stream_ptr(torch.device("cuda"))))
        assert lo <= probe.value <= hi, probe.value
        assert torch.equal(order.long().sort().values, torch.arange(B, device="cuda"))


@pytest.mark.parametrize("name", ["C2", "C3d", "C2r1"])
@pytest.mark.parametrize("B", [8192, 33001])
@pytest.mark.parametrize("data", ["pixels", "random"])
def test_sorted_step_matches_unsorted_step_and_oracle(name, B, data):
    """One fused step visiting the samples in cell order: (a) loss and every dL/dw equal the same kernel's walk in the
    given order (same per-sample arithmetic, only the summation order differs), (b) they equal the oracle on a
    4096-sample sub-batch."""
    ref, net = make_pair(name)
    rot = net.model[0].in_features // 2
    if data == "pixels":
        _, x, y, _ = pixel_batch(256, 256, B, seed=3, rotations=rot)
    else:
        g = torch.Generator().manual_seed(4)
        x = torch.rand(B, 2 * rot, generator=g) * 2.2 - 1.1
        y = torch.rand(B, 3, generator=g) * 2 - 1
    tr = H.FusedTrainer(net, lr=0.0, keep_grads=True)
    tr.sort_min_batch = 0
    tr.batch_order = "sorted"
    loss = tr.train_step(x.cuda(), y.cuda()).item()
    assert tr.last_launches == 2                                              # the sort pre-pass + ONE training launch
    g_sorted = [v.clone() for v in tr.g_views]
    # (a) the given order: chunks below the two-pass threshold through the in-kernel walk, same global scale
    tr.batch_order = "two_pass"
    acc = [torch.zeros_like(v) for v in tr.g_views]
    loss_u = 0.0
    for k in range(0, B, 1 << 14):
        loss_u += tr.train_step(x[k:k + (1 << 14)].cuda(), y[k:k + (1 << 14)].cuda(), global_batch=B).item()
        assert tr.last_launches == 1
        for a, v in zip(acc, tr.g_views):
            a += v
    assert abs(loss - loss_u) <= 2e-6 * abs(loss_u)
    for a, v in zip(acc, g_sorted):
        assert rel_err(v, a) <= 2e-5
    # (b) the oracle (fp32 or fp64: a sample within round-off of a hidden segment boundary / MaxAbs tie may take the
    # other branch than the fp32 oracle) on the first 4096 samples, sorted path forced
    tr.batch_order = "sorted"
    n = 4096
    loss_s = tr.train_step(x[:n].cuda(), y[:n].cuda()).item()
    assert tr.last_launches == 2
    loss_ref = O.mse_training_step(ref, x[:n], y[:n])
    loss_ref.backward()
    assert abs(loss_s - loss_ref.item()) <= 1e-5 * abs(loss_ref.item())
    from helpers import MLP_KW
    ref64 = O.HighOrderMLP(normalization=O.MaxAbsNormalization, **MLP_KW[name]).double()
    ref64.load_state_dict({k: v.double() for k, v in ref.state_dict().items()})
    O.mse_training_step(ref64, x[:n].double(), y[:n].double()).backward()
    for gv, q, q64 in zip(tr.g_views, ref.parameters(), ref64.parameters()):
        assert min(rel_err(gv, q.grad), rel_err(gv, q64.grad)) <= 2e-5


def test_auto_mode_decides_once_from_the_probe():
    _, net = make_pair("C2")
    tr = H.FusedTrainer(net, lr=0.0, keep_grads=True)
    assert tr.batch_order == "auto"
    B, n = 1 << 17, 1024
    g = torch.Generator(device="cuda").manual_seed(5)
    idx = torch.randint(0, n * n, (B,), device="cuda", generator=g)
    pos = torch.stack(H.positions_from_mesh(n, n, device="cuda", rotations=2, normalize=True), dim=2).reshape(-1, 4)
    y = torch.rand(B, 3, device="cuda", generator=g) * 2 - 1
    tr.train_step(pos[idx].contiguous(), y)
    assert tr.batch_order == "sorted" and tr.batch_coherence < 0.25
    _, net2 = make_pair("C2")
    tr2 = H.FusedTrainer(net2, lr=0.0, keep_grads=True)
    tr2.train_step(torch.rand(B, 4, device="cuda", generator=g) * 2 - 1, y)
    assert tr2.batch_order == "two_pass" and tr2.batch_coherence > 0.25


@pytest.mark.parametrize("name", ["C2", "C3d"])
def test_pixel_index_step_equals_explicit_batch_step(name):
    """train_step_pixels (indices only; coordinates and uint8 targets produced in-kernel) == the explicit (x, y) batch of
    the same pixels == the oracle."""
    ref, net = make_pair(name)
    B = 20000
    img, x, y, idx = pixel_batch(192, 160, B, seed=7)
    tr = H.FusedTrainer(net, lr=0.0, keep_grads=True)
    loss_p = tr.train_step_pixels(img.cuda(), 2, idx.cuda()).item()
    g_p = tr.flat_g[:-1].clone()
    tr.batch_order = "sorted"
    loss_x = tr.train_step(x.cuda(), y.cuda()).item()
    g_x = tr.flat_g[:-1].clone()
    assert abs(loss_p - loss_x) <= 2e-6 * abs(loss_x)
    assert rel_err(g_p, g_x) <= 1e-5
    loss_ref = O.mse_training_step(ref, x, y)
    assert abs(loss_p - loss_ref.item()) <= 1e-5 * abs(loss_ref.item())


def test_full_size_sorted_equals_two_pass_on_shuffled_pixels():
    """BASELINE size (2^20 shuffled pixels of a 4096^2 image, too big for the oracle): the cell-sorted single-kernel path
    and the two-pass workspace path (oracle-checked at small sizes) give the same loss and gradients."""
    _, net = make_pair("C2")
    tr = H.FusedTrainer(net, lr=0.0, keep_grads=True)
    B, rows, cols = 1 << 20, 4096, 4096
    g = torch.Generator(device="cuda").manual_seed(8)
    idx = torch.randint(0, rows * cols, (B,), device="cuda", generator=g)
    pos = torch.stack(H.positions_from_mesh(rows, cols, device="cuda", rotations=2, normalize=True), dim=2).reshape(-1, 4)
    x = pos[idx].contiguous()
    del pos
    img = torch.randint(0, 256, (rows, cols, 3), device="cuda", dtype=torch.uint8, generator=g)
    y = (img.reshape(-1, 3)[idx].float() * 2 / 255 - 1).contiguous()
    tr.batch_order = "sorted"
    l_s = tr.train_step(x, y).item()
    g_s = tr.flat_g[:-1].clone()
    tr.batch_order = "two_pass"
    l_t = tr.train_step(x, y).item()
    g_t = tr.flat_g[:-1].clone()
    assert abs(l_s - l_t) <= 2e-6 * abs(l_t)
    assert rel_err(g_s, g_t) <= 1e-4                                           # fp32 atomics: summation order differs
    l_p = tr.train_step_pixels(img, 2, idx.to(torch.int32)).item()
    assert abs(l_p - l_t) <= 2e-6 * abs(l_t)
    assert rel_err(tr.flat_g[:-1], g_t) <= 1e-4


@pytest.mark.parametrize("seg0,in_width,B", [(300, 4, 9001), (129, 2, 8193), (1000, 2, 12345), (7, 4, 8200)])
def test_sort_edge_geometries(seg0, in_width, B):
    """Layer-0 segment counts that force coarser cells (more than 2^18 bins otherwise), two-input networks (no minor key),
    ragged batch sizes, inputs outside [-1, 1] and non-finite inputs: the order is always a permutation and the sorted
    step equals the unsorted one."""
    import hoir_b200 as H
    kw = dict(layer_type="continuous", n=3, in_width=in_width, out_width=3, hidden_layers=1, hidden_width=40,
              in_segments=seg0, out_segments=2, hidden_segments=2)
    torch.manual_seed(1)
    net = H.HighOrderMLP(normalization=H.MaxAbsNormalization, **kw)
    H.initialize_network_polynomial_layers(net, 1.0, 0.0, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        for p in net.parameters():
            p.add_(0.2 * torch.randn(p.shape, generator=torch.Generator().manual_seed(2)) / p.shape[1] ** 0.5)
    tr = H.FusedTrainer(net.cuda(), lr=0.0, keep_grads=True)
    g = torch.Generator().manual_seed(3)
    x = torch.rand(B, in_width, generator=g) * 2.6 - 1.3                 # 13 % of the samples outside the domain
    y = torch.rand(B, 3, generator=g) * 2 - 1
    order = torch.empty(B, dtype=torch.int32, device="cuda")
    _lib.check(_lib.lib().hoir_mlp_batch_sort(tr.desc, B, _lib.dptr(x.cuda()), None, None, _lib.dptr(order, torch.int32),
                                              None, _lib.stream_ptr(torch.device("cuda"))))
    assert torch.equal(order.long().sort().values.cpu(), torch.arange(B))
    tr.sort_min_batch, tr.batch_order = 0, "sorted"
    l_s = tr.train_step(x.cuda(), y.cuda()).item()
    g_s = tr.flat_g[:-1].clone()
    tr.batch_order = "two_pass"                                          # below 32768 samples: the unsorted in-kernel walk
    l_u = tr.train_step(x.cuda(), y.cuda()).item()
    assert abs(l_s - l_u) <= 2e-6 * abs(l_u)
    assert rel_err(g_s, tr.flat_g[:-1]) <= 2e-5
    # non-finite coordinates must not crash the sort (segment_index maps NaN to segment 0, +-inf to the end segments)
    xb = x.clone()
    xb[::97, 0] = float("nan")
    xb[5::101, 1] = float("inf")
    xb[7::103, 0] = float("-inf")
    _lib.check(_lib.lib().hoir_mlp_batch_sort(tr.desc, B, _lib.dptr(xb.cuda()), None, None, _lib.dptr(order, torch.int32),
                                              None, _lib.stream_ptr(torch.device("cuda"))))
    assert torch.equal(order.long().sort().values.cpu(), torch.arange(B))


def test_sort_rejects_bad_arguments():
    _, net = make_pair("C2")
    tr = H.FusedTrainer(net, lr=0.0)
    lib = _lib.lib()
    with pytest.raises(ValueError):
        _lib.check(lib.hoir_mlp_batch_sort(tr.desc, 16, None, None, None, None, None, None))            # neither x nor mesh
    _lib.check(lib.hoir_mlp_batch_sort(tr.desc, 0, None, _lib.make_mesh_desc(8, 8, 2, True, 0), None, None, None, None))   # B == 0
    _, net10 = make_pair("C1")
    tr10 = H.FusedTrainer(net10, lr=0.0)
    assert not lib.hoir_mlp_order_supported(tr10.desc)                  # FFMA-only network: no *_order entries
    x = torch.zeros(64, 2, device="cuda")
    order = torch.zeros(64, dtype=torch.int32, device="cuda")
    _lib.check(lib.hoir_mlp_batch_sort(tr10.desc, 64, _lib.dptr(x), None, None, _lib.dptr(order, torch.int32), None, None))
    with pytest.raises(H.HoirUnsupported):
        _lib.check(lib.hoir_mlp_train_step_order(tr10.desc, 64, _lib.dptr(x), None, _lib.dptr(order, torch.int32), 0,
                                                 _lib.dptr(torch.zeros(64, 3, device="cuda")), None, _lib.dptr(tr10.packed),
                                                 tr10._g_table, tr10._loss_ptr, 1.0, None))
