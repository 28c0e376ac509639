"""The fused whole-network kernel for Fourier-series implicit MLPs (csrc/fused_fourier.cu; BASELINE config C3 "fourier
n_in=31": /root/reference/README.md:44, /root/reference/high_order_implicit_representation/networks.py:41-55 with Quirk
Q1) against the CPU oracle: forward, loss, every dL/dw, optimizer trajectories, the in-kernel weight-image refresh.
Tolerance: 1e-4 relative (BASELINE.json: Fourier layers), tf32 x 3 tensor-core passes."""
import pytest
import torch

import hoir_b200 as H
from oracle import hol_oracle as O
from helpers import MLP_KW, make_pair, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4

VARIANTS = {
    "C3f": MLP_KW["C3f"],
    "small": dict(layer_type="fourier", n=3, n_in=5, n_hidden=7, n_out=2, in_width=2, out_width=1, hidden_layers=1,
                  hidden_width=16),
    "rot1": dict(layer_type="fourier", n=3, n_in=31, n_hidden=31, in_width=2, out_width=3, hidden_layers=1,
                 hidden_width=40),
    "wide24": dict(layer_type="fourier", n=4, n_in=32, n_hidden=12, n_out=4, in_width=6, out_width=4, hidden_layers=1,
                   hidden_width=24),
}


def pair(name, normalization=True, seed=0):
    kw = VARIANTS[name]
    torch.manual_seed(seed)
    norm_o = O.MaxAbsNormalization if normalization else None
    norm_h = H.MaxAbsNormalization if normalization else None
    ref = O.HighOrderMLP(normalization=norm_o, **kw)
    if normalization:                          # (without MaxAbs the activations would leave [-1, 1]: the fp32 argument
        with torch.no_grad():                  # of the 15th harmonic then carries less than 1e-4 itself)
            for p in ref.parameters():
                p.mul_(3.0)                   # constructor init is U(-1/in, 1/in): make the harmonics matter
    net = H.HighOrderMLP(normalization=norm_h, **kw)
    net.load_state_dict(ref.state_dict())
    return ref, net.cuda()


def batch(kw, B, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, kw["in_width"], generator=g) * 2.2 - 1.1, torch.rand(B, kw["out_width"], generator=g) * 2 - 1


@pytest.mark.parametrize("name", list(VARIANTS))
@pytest.mark.parametrize("norm", [True, False])
def test_fused_supported_and_forward(name, norm):
    ref, net = pair(name, norm)
    assert net.fused_supported()
    with torch.no_grad():
        scale = ref(batch(VARIANTS[name], 3000, seed=3000)[0]).abs().max().item()     # max |ref| over a batch
    for B in (1, 255, 256, 257, 3000):
        x, _ = batch(VARIANTS[name], B, seed=B)
        with torch.no_grad():
            assert (net(x.cuda()).cpu() - ref(x)).abs().max().item() <= TOL * scale, B
    assert net(torch.empty(0, VARIANTS[name]["in_width"], device="cuda")).shape == (0, VARIANTS[name]["out_width"])


@pytest.mark.parametrize("name", list(VARIANTS))
@pytest.mark.parametrize("B", [1, 300, 5000])
def test_fused_step_loss_and_gradients_match_oracle(name, B):
    ref, net = pair(name)
    x, y = batch(VARIANTS[name], B, seed=7)
    tr = H.FusedTrainer(net, lr=0.0, keep_grads=True)
    loss = tr.train_step(x.cuda(), y.cuda()).item()
    assert tr.last_launches == 1
    loss_ref = O.mse_training_step(ref, x, y)
    loss_ref.backward()
    assert abs(loss - loss_ref.item()) <= TOL * abs(loss_ref.item())
    for gk, (k, p) in zip(tr.g_views, ref.named_parameters()):
        assert rel_err(gk, p.grad) <= TOL, k


def test_autograd_through_the_fused_kernels():
    """HighOrderMLP.forward under autograd: forward = hoir_mlp_forward, backward = hoir_mlp_train_step with the upstream
    dL/dy (the same kernel, no optimizer tail)."""
    ref, net = pair("C3f")
    x, y = batch(VARIANTS["C3f"], 1000, seed=2)
    loss_ref = O.mse_training_step(ref, x, y)
    loss_ref.backward()
    loss = torch.nn.functional.mse_loss(net(x.cuda()).flatten(), y.cuda().flatten())
    loss.backward()
    assert abs(loss.item() - loss_ref.item()) <= TOL * abs(loss_ref.item())
    for (k, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
        assert rel_err(p.grad, q.grad) <= TOL, k
    # and against the per-layer tensor-core kernels of this library
    net.use_fused = False
    net.zero_grad()
    loss2 = torch.nn.functional.mse_loss(net(x.cuda()).flatten(), y.cuda().flatten())
    loss2.backward()
    assert abs(loss2.item() - loss.item()) <= 1e-5 * abs(loss.item())


@pytest.mark.parametrize("opt", ["adam", "sparse_lion"])
def test_training_trajectory_and_weight_images(opt):
    """Four optimizer steps in-kernel == the oracle's torch loop; the kernel-side weight images written by the tail are
    what the next forward uses (render after training == oracle forward with the trained weights)."""
    ref, net = pair("C3f")
    tr = H.FusedTrainer(net, optimizer=opt, lr=1e-3)
    ref_opt = torch.optim.Adam(ref.parameters(), lr=1e-3) if opt == "adam" else O.SparseLion(ref.parameters(), lr=1e-3)
    for step in range(4):
        x, y = batch(VARIANTS["C3f"], 2000, seed=30 + step)
        ref_opt.zero_grad()
        loss_ref = O.mse_training_step(ref, x, y)
        loss_ref.backward()
        ref_opt.step()
        loss = tr.train_step(x.cuda(), y.cuda()).item()
        assert abs(loss - loss_ref.item()) <= 2 * TOL * abs(loss_ref.item()), step
        assert float(tr.flat_g.abs().max()) == 0.0
    for (k, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
        frac_off = ((p.detach().cpu() - q.detach()).abs() > 1e-4).float().mean().item()
        assert frac_off <= 5e-3, (k, frac_off)
    x, _ = batch(VARIANTS["C3f"], 777, seed=99)
    ref.load_state_dict({k: v.cpu() for k, v in net.state_dict().items()})
    assert rel_err(tr.forward(x.cuda()), ref(x)) <= TOL          # images in sync with flat_w
    packed_by_tail = tr.packed.clone()
    tr.pack()
    assert torch.equal(packed_by_tail, tr.packed)                 # bit-identical to a fresh pack of the same weights


def test_mesh_mode_with_uint8_targets():
    ref, net = pair("C3f")
    rows, cols = 40, 56
    img = torch.randint(0, 256, (rows, cols, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(3))
    y_flat, x_flat, _ = O.image_to_dataset(img, rotations=2)
    tr = H.FusedTrainer(net, lr=0.0, keep_grads=True)
    first, count = 300, 1500
    loss = tr.train_step_mesh(img.cuda(), 2, first, count).item()
    loss_ref = O.mse_training_step(ref, x_flat[first:first + count], y_flat[first:first + count])
    loss_ref.backward()
    assert abs(loss - loss_ref.item()) <= TOL * abs(loss_ref.item())
    for gk, p in zip(tr.g_views, ref.parameters()):
        assert rel_err(gk, p.grad) <= TOL
    out = tr.render(rows, cols, 2, post_scale=False)
    assert rel_err(out, ref(x_flat)) <= TOL


def test_full_size_linearity():
    """2^18 samples (too many for the oracle): loss and gradients are additive over halves."""
    _, net = pair("C3f")
    tr = H.FusedTrainer(net, lr=0.0, keep_grads=True)
    B = 1 << 18
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand(B, 4, device="cuda", generator=g) * 2 - 1
    y = torch.rand(B, 3, device="cuda", generator=g) * 2 - 1
    loss = tr.train_step(x, y).item()
    gw = tr.flat_g[:-1].clone()
    l0 = tr.train_step(x[: B // 2], y[: B // 2], global_batch=B).item()
    g0 = tr.flat_g[:-1].clone()
    l1 = tr.train_step(x[B // 2:], y[B // 2:], global_batch=B).item()
    g1 = tr.flat_g[:-1].clone()
    assert abs(l0 + l1 - loss) <= 1e-5 * loss
    assert rel_err(g0 + g1, gw) <= 1e-4
