"""BASELINE.json sizes the CPU oracle cannot run in seconds: size-independent properties of the large-batch kernel
paths (streamed-slab forward, slab dL/dw, tensor-core layers) against the small-batch kernels that ARE checked
against the oracle in test_gpu_layers.py -- agreement on strided sub-samples and linearity of the summed gradients."""
import pytest
import torch

import hoir_b200 as H
from helpers import MLP_KW, rel_err

pytestmark = pytest.mark.gpu

C4 = dict(layer_type="continuous", n=2, in_width=216, out_width=27, hidden_layers=2, hidden_width=50,
          in_segments=50, out_segments=2, hidden_segments=2)


def _net(kw, seed=0):
    torch.manual_seed(seed)
    net = H.HighOrderMLP(normalization=H.MaxAbsNormalization, **kw)
    H.initialize_network_polynomial_layers(net, 1.0, 0.0, generator=torch.Generator().manual_seed(seed))
    with torch.no_grad():
        for p in net.parameters():
            p.add_(0.2 * torch.randn(p.shape, generator=torch.Generator().manual_seed(seed + 1)) / p.shape[1] ** 0.5)
    return net.cuda()


def test_c4_interpolator_full_batch_matches_small_batch_kernels():
    """C4: 2048^2 synthetic image, 3x3 blocks with a 3-pixel donut, 2^16 patches per step."""
    net = _net(C4)
    g = torch.Generator().manual_seed(0)
    img = (torch.randint(0, 256, (3, 2048, 2048), generator=g).float() / 127.5 - 1).cuda()
    B = 1 << 16
    feat, targ, _ = H.neighborhood_gather(img, 3, 3, stride=3)
    feat, targ = feat[:B].contiguous(), targ[:B].contiguous()
    assert feat.shape == (B, 216) and targ.shape == (B, 27)
    y = net(feat)
    loss = torch.nn.functional.mse_loss(y, targ)
    loss.backward()
    g_full = [p.grad.clone() for p in net.parameters()]
    # forward: a strided sub-sample through the small-batch (oracle-checked) kernels
    idx = torch.arange(0, B, 131, device="cuda")[:500]
    with torch.no_grad():
        y_small = net(feat[idx])
    assert rel_err(y[idx], y_small) <= 1e-5
    # gradients: the same loss accumulated over chunks below the large-batch dispatch thresholds
    for p in net.parameters():
        p.grad = None
    chunk = 4000
    for k in range(0, B, chunk):
        yk = net(feat[k:k + chunk])
        (torch.nn.functional.mse_loss(yk, targ[k:k + chunk], reduction="sum") / (B * 27)).backward()
    for a, p in zip(g_full, net.parameters()):
        assert rel_err(p.grad, a) <= 2e-5


def test_c3_fourier_network_full_batch_linearity():
    """C3 Fourier (n_in = 31): 2^18 samples through the tensor-core layers; loss and gradients are additive over halves."""
    net = _net(MLP_KW["C3f"])
    B = 1 << 18
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand(B, 4, device="cuda", generator=g) * 2 - 1
    t = torch.rand(B, 3, device="cuda", generator=g) * 2 - 1

    def run(sl):
        for p in net.parameters():
            p.grad = None
        l = torch.nn.functional.mse_loss(net(x[sl]), t[sl], reduction="sum") / (B * 3)
        l.backward()
        return l.detach(), [p.grad.clone() for p in net.parameters()]

    l_all, g_all = run(slice(0, B))
    l0, g0 = run(slice(0, B // 2))
    l1, g1 = run(slice(B // 2, B))
    assert abs((l0 + l1).item() - l_all.item()) <= 1e-5 * l_all.item()
    for a, b, c in zip(g_all, g0, g1):
        assert rel_err(b + c, a) <= 1e-4
    idx = torch.arange(0, B, 523, device="cuda")[:400]
    with torch.no_grad():
        assert rel_err(net(x)[idx], net(x[idx])) <= 1e-4          # 400 rows: CUDA-core kernels


def test_c5_render_tiles_are_independent_of_the_sharding():
    """C5: image-tile sharded rendering (no collective): rendering [first, first+count) ranges separately gives the
    same pixels, bit for bit, as one launch over the whole 2048 x 1024 mesh (rotations = 2)."""
    net = _net(MLP_KW["C2"])
    tr = H.FusedTrainer(net, lr=0.0)
    rows, cols = 2048, 1024
    whole = tr.render(rows, cols, 2)
    parts = []
    for r in range(8):
        first, count = H.parallel.render_shard(rows * cols, r, 8)
        parts.append(tr.render(rows, cols, 2, first_pixel=first, count=count))
    assert torch.equal(torch.cat(parts), whole)
    assert torch.isfinite(whole).all()


@pytest.mark.parametrize("name", ["C2", "C3d"])
def test_c2_full_step_mesh_path_equals_explicit_batch_path(name):
    """C2 / C3 at the BASELINE size: 2^20 pixels of a 4096 x 4096 image.  The in-kernel coordinate / uint8-target path
    with the run-merged layer-0 owner walk must give the same loss and gradients as the explicit (x, y) batch of the
    same pixels, which takes the two-pass layer-0 path (workspace stream + segment-sorted second pass)."""
    net = _net(MLP_KW[name])
    tr = H.FusedTrainer(net, lr=0.0, keep_grads=True)
    g = torch.Generator().manual_seed(4)
    img = torch.randint(0, 256, (4096, 4096, 3), generator=g, dtype=torch.uint8).cuda()
    B, first = 1 << 20, 5 * (1 << 20) + 4096 * 3 + 17           # not tile aligned
    loss_mesh = tr.train_step_mesh(img, 2, first, B).clone()
    g_mesh = tr.flat_g[:-1].clone()
    flat, pos, _ = H.image_to_dataset(img, rotations=2)
    x, y = pos[first:first + B].contiguous(), flat[first:first + B].contiguous()
    loss_x = tr.train_step(x, y).clone()
    g_x = tr.flat_g[:-1].clone()
    assert abs(loss_mesh.item() - loss_x.item()) <= 1e-6 * abs(loss_x.item())
    assert rel_err(g_mesh, g_x) <= 2e-5                          # fp32 atomics: summation order differs
