"""GPU parity of the neighborhood interpolator pieces (patch gather, fold + crop + blend, rendering loops)
against the oracle's restatement of the reference functions; shape pins of
/root/reference/tests/test_single_image_dataset.py and test_rendering.py."""
import pytest
import torch

import hoir_b200 as H
from oracle import hol_oracle as O
from helpers import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("h,w,width,outside,stride", [(99, 99, 3, 3, 3), (32, 32, 3, 1, 1), (21, 32, 1, 1, 1),
                                                      (32, 21, 5, 3, 2), (40, 37, 3, 3, 1)])
def test_image_neighborhood_dataset_bit_exact(h, w, width, outside, stride):
    img = torch.rand(3, h, w, generator=torch.Generator().manual_seed(h + w)) * 2 - 1
    f_ref, t_ref, _, lx, ly = O.image_neighborhood_dataset(img, width=width, outside=outside, stride=stride)
    f, t, im, lastx, lasty = H.image_neighborhood_dataset(img, width=width, outside=outside, stride=stride)
    assert (lastx, lasty) == (lx, ly)
    assert torch.equal(f.cpu(), f_ref) and torch.equal(t.cpu(), t_ref)


def test_reference_shape_pins():
    img = torch.arange(99 * 99 * 3, dtype=torch.float32).reshape(3, 99, 99)
    f, t, _, _, _ = H.image_neighborhood_dataset(img, width=3, outside=3, stride=3)
    assert f.shape == (961, 216) and t.shape == (961, 27)
    f, t, im, lastx, lasty = H.image_neighborhood_dataset(torch.rand(3, 32, 32), width=3, outside=1)
    assert t.shape == (784, 27) and f.shape == (784, 48) and lastx == lasty == 27 and im.shape == (3, 32, 32)


def _nets(width, outside):
    in_w = ((width + 2 * outside) ** 2 - width ** 2) * 3
    kw = dict(layer_type="discontinuous", n=2, in_width=in_w, out_width=3 * width * width, hidden_layers=2,
              hidden_width=10, in_segments=100, out_segments=2, hidden_segments=2)
    torch.manual_seed(0)
    ref = O.HighOrderMLP(normalization=O.MaxAbsNormalization, **kw)
    net = H.HighOrderMLP(normalization=H.MaxAbsNormalization, **kw)
    net.load_state_dict(ref.state_dict())
    return ref, net.cuda()


@pytest.mark.parametrize("h,w", [(21, 21), (21, 32), (32, 21)])
@pytest.mark.parametrize("width,outside", [(1, 1), (3, 3), (5, 1), (3, 5)])
def test_neighborhood_sample_generator(h, w, width, outside):
    ref, net = _nets(width, outside)
    img = torch.rand(3, h, w, generator=torch.Generator().manual_seed(7)) * 2 - 1
    want = O.neighborhood_sample_generator(ref, img, width=width, outside=outside, batch_size=64)
    got = H.neighborhood_sample_generator(net, img.cuda(), width=width, outside=outside, batch_size=64)
    assert got.shape == img.shape
    assert rel_err(got, want) <= 1e-5


def test_generate_sequence_blend():
    ref, net = _nets(3, 3)
    img = torch.rand(3, 24, 24, generator=torch.Generator().manual_seed(3)) * 2 - 1
    seq = H.generate_sequence(net, img.cuda(), frames=3, width=3, outside=3, batch_size=128, alpha=0.75)
    assert seq.shape == (4, 3, 24, 24)
    this = img.clone()
    for _ in range(3):
        this = 0.75 * this + 0.25 * O.neighborhood_sample_generator(ref, this, 3, 3, 128)
    assert rel_err(seq[-1], 0.5 * (this + 1)) <= 1e-5


def test_render_image_paths():
    kw = dict(layer_type="continuous", n=3, in_width=4, out_width=3, hidden_layers=1, hidden_width=40,
              in_segments=100, out_segments=2, hidden_segments=2)
    torch.manual_seed(1)
    ref = O.HighOrderMLP(normalization=O.MaxAbsNormalization, **kw)
    O.initialize_network_polynomial_layers(ref, 1.0, 0.0)
    net = H.HighOrderMLP(normalization=H.MaxAbsNormalization, **kw)
    net.load_state_dict(ref.state_dict())
    net = net.cuda()
    img = torch.zeros(24, 36, 3, dtype=torch.uint8)
    _, x_flat, _ = O.image_to_dataset(img, rotations=2)
    want = O.render_image(ref, x_flat, (24, 36, 3), batch_size=24)
    assert rel_err(H.render_image(net, 24, 36, 2), want) <= 1e-5
    net.use_fused = False
    assert rel_err(H.render_image(net, 24, 36, 2, batch_size=24 * 9), want) <= 1e-5
    flat, pos, im = H.image_to_dataset(img, rotations=2)
    assert torch.equal(pos.cpu(), x_flat) and flat.shape == (24 * 36, 3)
