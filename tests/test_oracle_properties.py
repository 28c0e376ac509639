"""CPU tests of the oracle: the property tests of SURVEY.md A.8 that stand in for the missing
golden vectors of high-order-layers-torch, the reference's own shape pins
(/root/reference/tests/test_single_image_dataset.py:13-37, test_rendering.py:13-74) and the
committed golden fixture (tests/golden/layers_v1.npz)."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import hol_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "layers_v1.npz")


def mlp(layer_type, n=3, n_in=None, n_hidden=None, n_out=None, in_width=4, hidden=40, layers=1, in_seg=100):
    return O.HighOrderMLP(layer_type=layer_type, n=n, n_in=n_in, n_hidden=n_hidden, n_out=n_out,
                          in_width=in_width, out_width=3, hidden_layers=layers, hidden_width=hidden,
                          in_segments=in_seg, out_segments=2, hidden_segments=2,
                          normalization=O.MaxAbsNormalization)


def count(m):
    return sum(p.numel() for p in m.parameters())


def test_parameter_counts():                                   # A.8 #1, README.md:35
    assert count(mlp("continuous")) == 40760
    assert count(mlp("continuous", in_width=2, hidden=10, layers=2)) == 5170
    assert count(mlp("discontinuous")) == 58320
    assert count(mlp("fourier", n=3, n_in=31, n_hidden=31, n_out=3)) == 54920   # Quirk Q1
    assert list(mlp("continuous").state_dict().keys()) == ["model.0.w", "model.2.w", "model.4.w"]


@pytest.mark.parametrize("kind,n,seg", [("continuous", 3, 4), ("continuous", 2, 5), ("discontinuous", 3, 4),
                                        ("continuous", 5, 3)])
def test_interpolation_at_nodes(kind, n, seg):                 # A.8 #2
    layer = O.high_order_fc_layers(kind, n=n, in_features=1, out_features=3, segments=seg).double()
    pos = O.node_positions(layer).double()
    y = layer(pos.view(-1, 1) * (1 - 1e-12))
    w = layer.w.detach()[:, 0, :].t()
    if kind == "continuous":
        assert torch.allclose(y, w, atol=1e-6)
    else:  # interior of each segment's node set
        idx = [k for k in range(pos.numel()) if k % n not in (0, n - 1)]
        assert torch.allclose(y[idx], w[idx], atol=1e-6)


def test_continuity_and_jumps():                               # A.8 #3
    eps = 1e-6
    b = torch.tensor([-0.5, 0.0, 0.5], dtype=torch.float64).view(-1, 1)
    cont = O.high_order_fc_layers("continuous", n=3, in_features=1, out_features=2, segments=4).double()
    assert (cont(b - eps) - cont(b + eps)).abs().max() < 1e-4
    disc = O.high_order_fc_layers("discontinuous", n=3, in_features=1, out_features=2, segments=4).double()
    jump = disc(b + eps) - disc(b - eps)
    w = disc.w.detach()[:, 0, :]
    want = torch.stack([w[:, 3 * s] - w[:, 3 * s - 1] for s in (1, 2, 3)])
    assert torch.allclose(jump, want, atol=1e-4)


@pytest.mark.parametrize("kind", ["continuous", "discontinuous"])
def test_polynomial_exactness_and_linear_init(kind):           # A.8 #4
    n, seg = 4, 5
    layer = O.high_order_fc_layers(kind, n=n, in_features=1, out_features=1, segments=seg).double()
    pos = O.node_positions(layer).double()
    f = lambda t: 0.3 * t ** 3 - 0.7 * t ** 2 + 0.2 * t - 0.1   # degree n-1
    with torch.no_grad():
        layer.w.copy_(f(pos).view(1, 1, -1))
    x = torch.linspace(-0.999, 0.999, 333, dtype=torch.float64).view(-1, 1)
    assert torch.allclose(layer(x), f(x), atol=1e-5)
    net = O.HighOrderMLP(layer_type=kind, n=3, in_width=2, out_width=1, hidden_layers=0, hidden_width=3,
                         in_segments=7, out_segments=2)
    O.initialize_network_polynomial_layers(net, max_slope=1.0, max_offset=0.5)
    first = net.model[0]
    xs = torch.rand(50, 2) * 1.9 - 0.95
    pos = O.node_positions(first)
    slope = (first.w[:, :, -1] - first.w[:, :, 0]) / (pos[-1] - pos[0])
    offset = first.w[:, :, 0] - slope * pos[0]
    want = xs @ slope.t() + offset.sum(1)
    assert torch.allclose(first(xs), want, atol=1e-5)


def test_partition_of_unity():                                 # A.8 #5
    for n in (2, 3, 5, 8):
        poly = O.LagrangePoly(n)
        x = torch.linspace(-1, 1, 101, dtype=torch.float64)
        assert torch.allclose(sum(poly(x, j) for j in range(n)), torch.ones_like(x), atol=1e-9)


def test_edge_semantics():                                     # A.8 #6
    layer = O.high_order_fc_layers("continuous", n=3, in_features=1, out_features=1, segments=100)
    x = torch.tensor([-1.5, -1.0, -1.0 + 1e-6, 0.0, 1.0 - 1e-6, 1.0, 1.2]).view(-1, 1)
    assert layer.segment_index(x).flatten().tolist() == [0, 0, 0, 50, 99, 99, 99]
    assert abs(O.chebyshev_lobatto(3)[1].item() - 4.3711388e-08) < 1e-12   # not snapped to 0


@pytest.mark.parametrize("kind", ["continuous", "discontinuous", "polynomial", "fourier"])
def test_gradcheck_fp64(kind):                                 # A.8 #7
    layer = O.high_order_fc_layers(kind, n=3, in_features=3, out_features=2, segments=4).double()
    x = (torch.rand(5, 3, dtype=torch.float64) * 1.6 - 0.8 + 0.013).requires_grad_(True)
    assert torch.autograd.gradcheck(lambda t: layer(t), (x,), eps=1e-7, atol=1e-5)
    norm = O.MaxAbsNormalization()
    xn = torch.randn(4, 6, dtype=torch.float64, requires_grad=True)
    assert torch.autograd.gradcheck(norm, (xn,), eps=1e-7, atol=1e-6)


def test_maxabs_properties():                                  # A.8 #8
    x = torch.randn(64, 40)
    y = O.MaxAbsNormalization()(x)
    m = x.abs().max(1)[0]
    assert torch.allclose(y.abs().max(1)[0], m / (m + 1e-6))
    assert torch.allclose(O.MaxAbsNormalization()(7.5 * x), y, atol=1e-5)


def test_fourier_properties():                                 # A.8 #9
    layer = O.high_order_fc_layers("fourier", n=7, in_features=3, out_features=2).double()
    x = torch.rand(20, 3, dtype=torch.float64) * 2 - 1
    assert torch.allclose(layer(x), layer(x + 2.0), atol=1e-9)
    with torch.no_grad():
        layer.w.zero_()
        layer.w[:, :, 0] = 1.0
    assert torch.allclose(layer(x), torch.full((20, 2), 1.5, dtype=torch.float64))


def test_neighborhood_shape_pins():                            # A.8 #10
    img = torch.arange(99 * 99 * 3, dtype=torch.float32).reshape(3, 99, 99)
    feat, targ, _, _, _ = O.image_neighborhood_dataset(img, width=3, outside=3, stride=3)
    assert feat.shape == (961, 216) and targ.shape == (961, 27)
    img = torch.rand(3, 32, 32)
    feat, targ, im, lastx, lasty = O.image_neighborhood_dataset(img, width=3, outside=1, stride=1)
    assert targ.shape == (784, 27) and feat.shape == (784, 48) and lastx == lasty == 27
    # feature / target ordering: channel-major, then window cells row-major
    f0 = feat[0].view(3, -1)
    assert f0[1, 0] == img[1, 0, 0] and f0[1, 1] == img[1, 0, 1]
    t0 = targ[0].view(3, 3, 3)
    assert torch.equal(t0, img[:, 1:4, 1:4])


@pytest.mark.parametrize("h,w,width,outside", [(21, 32, 1, 1), (32, 21, 3, 3), (21, 21, 5, 1), (32, 32, 3, 5)])
def test_neighborhood_generator_shape(h, w, width, outside):   # test_rendering.py:13-74
    in_w = ((width + 2 * outside) ** 2 - width ** 2) * 3
    net = O.HighOrderMLP(layer_type="discontinuous", n=2, in_width=in_w, out_width=3 * width * width,
                         hidden_layers=2, hidden_width=10, in_segments=100, out_segments=2,
                         hidden_segments=2, normalization=O.MaxAbsNormalization)
    img = torch.rand(3, h, w) * 2 - 1
    out = O.neighborhood_sample_generator(net, img, width=width, outside=outside, batch_size=64)
    assert out.shape == img.shape


def test_render_loop_feeds_empty_batch():                      # Quirk Q5
    seen = []

    class Probe(torch.nn.Module):
        def forward(self, x):
            seen.append(x.shape[0])
            return torch.zeros(x.shape[0], 3)

    O.render_image(Probe(), torch.zeros(64, 4), (8, 8, 3), batch_size=16)
    assert seen == [16, 16, 16, 16, 0]


def test_sparse_lion_touches_only_nonzero_grads():
    p = torch.nn.Parameter(torch.ones(6))
    opt = O.SparseLion([p], lr=0.1)
    p.grad = torch.tensor([0.0, 1.0, -2.0, 0.0, 3.0, 0.0])
    opt.step()
    assert p.detach().tolist() == pytest.approx([1.0, 0.9, 1.1, 1.0, 0.9, 1.0])
    assert opt.state[p]["exp_avg"].tolist() == pytest.approx([0, 0.01, -0.02, 0, 0.03, 0])


def test_positions_from_mesh_definition():
    lines = O.positions_from_mesh(4, 3, rotations=2, normalize=True)
    assert len(lines) == 4 and lines[0].shape == (4, 3)
    xv = (torch.arange(4).view(4, 1).expand(4, 3) / 4 - 0.5) * 2
    yv = (torch.arange(3).view(1, 3).expand(4, 3) / 4 - 0.5) * 2
    assert torch.allclose(lines[0], xv) and torch.allclose(lines[1], yv)
    assert torch.allclose(lines[2], (xv + yv) / 2, atol=1e-7) and torch.allclose(lines[3], (yv - xv) / 2, atol=1e-7)
    flat, pos, _ = O.image_to_dataset(torch.randint(0, 256, (4, 3, 3), dtype=torch.uint8), rotations=2)
    assert pos.shape == (12, 4) and flat.shape == (12, 3) and flat.min() >= -1 and flat.max() <= 1


def test_oracle_matches_golden_fixture():
    gold = np.load(GOLDEN)
    for kind, n, in_f, out_f, seg in [("continuous", 3, 2, 3, 4), ("continuous", 2, 3, 2, 5),
                                      ("discontinuous", 3, 2, 3, 4), ("polynomial", 4, 2, 3, 1),
                                      ("fourier", 5, 2, 3, 1)]:
        key = f"{kind}_n{n}"
        layer = O.high_order_fc_layers(kind, n=n, in_features=in_f, out_features=out_f, segments=seg)
        with torch.no_grad():
            layer.w.copy_(torch.from_numpy(gold[key + "_w"]))
        x = torch.from_numpy(gold[key + "_x"]).requires_grad_(True)
        y = layer(x)
        y.backward(torch.from_numpy(gold[key + "_gy"]))
        np.testing.assert_allclose(y.detach().numpy(), gold[key + "_y"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(x.grad.numpy(), gold[key + "_gx"], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(layer.w.grad.numpy(), gold[key + "_gw"], rtol=1e-5, atol=1e-6)
        if key + "_id" in gold:
            assert np.array_equal(layer.segment_index(x.detach()).numpy(), gold[key + "_id"])
    mesh = torch.stack(O.positions_from_mesh(4, 3, rotations=2, normalize=True)).numpy()
    assert np.array_equal(mesh, gold["mesh_4x3_r2"])
