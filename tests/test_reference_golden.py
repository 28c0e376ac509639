"""Golden vectors produced by RUNNING THE REFERENCE'S OWN CODE (tests/golden/make_reference_golden.py imports
/root/reference with stubs for the absent third-party packages): the parts of the path that live in the reference
itself -- patch gather order, reflect-pad / fold / crop, the alpha-blended sequence, the uint8 target scaling.
CPU tests pin the oracle to them bit for bit; GPU tests pin the CUDA kernels (through the C ABI) to them."""
import numpy as np
import pytest
import torch

from oracle import hol_oracle as O
from helpers import FixedModel, reference_golden

G = reference_golden()


@pytest.mark.parametrize("k", range(4))
def test_oracle_neighborhood_dataset_equals_reference(k):
    h, w, width, outside, stride, lastx, lasty = [int(v) for v in G[f"nbr{k}_geom"]]
    img = torch.from_numpy(G[f"nbr{k}_image"])
    feat, targ, _, lx, ly = O.image_neighborhood_dataset(img, width=width, outside=outside, stride=stride)
    assert (lx, ly) == (lastx, lasty)
    assert np.array_equal(feat.numpy(), G[f"nbr{k}_features"])
    assert np.array_equal(targ.numpy(), G[f"nbr{k}_targets"])


@pytest.mark.parametrize("k", range(3))
def test_oracle_sample_generator_equals_reference(k):
    h, w, width, outside = [int(v) for v in G[f"gen{k}_geom"]]
    img = torch.from_numpy(G[f"gen{k}_image"])
    total = width + 2 * outside
    model = FixedModel(3 * (total * total - width * width), 3 * width * width)
    res = O.neighborhood_sample_generator(model, img, width=width, outside=outside, batch_size=7)
    assert np.array_equal(res.numpy(), G[f"gen{k}_result"])
    if k == 0:
        this, frames = img.clone(), [img.clone()]
        for _ in range(3):
            this = 0.75 * this + (1 - 0.75) * O.neighborhood_sample_generator(model, this, width, outside, 50)
            frames.append(this.clone())
        assert np.array_equal((0.5 * (torch.stack(frames) + 1)).numpy(), G["seq0_frames"])


def test_oracle_target_scaling_equals_reference():
    u8 = torch.from_numpy(G["jupiter_u8"])
    flat, pos, _ = O.image_to_dataset(u8.view(-1, 1, 3), rotations=2)
    assert np.array_equal(flat.numpy(), G["jupiter_flat"])
    assert list(G["jupiter_shape"]) == [591, 960, 3] and list(G["jupiter_pos_shape"]) == [591 * 960, 4]


# ---------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("k", range(4))
def test_cuda_neighborhood_gather_equals_reference(k):
    import hoir_b200 as H
    h, w, width, outside, stride, lastx, lasty = [int(v) for v in G[f"nbr{k}_geom"]]
    img = torch.from_numpy(G[f"nbr{k}_image"])
    feat, targ, _, lx, ly = H.image_neighborhood_dataset(img, width=width, outside=outside, stride=stride)
    assert (lx, ly) == (lastx, lasty)
    assert np.array_equal(feat.cpu().numpy(), G[f"nbr{k}_features"])        # a gather: bit-exact
    assert np.array_equal(targ.cpu().numpy(), G[f"nbr{k}_targets"])


@pytest.mark.gpu
@pytest.mark.parametrize("k", range(3))
def test_cuda_sample_generator_equals_reference(k):
    import hoir_b200 as H
    h, w, width, outside = [int(v) for v in G[f"gen{k}_geom"]]
    img = torch.from_numpy(G[f"gen{k}_image"]).cuda()
    total = width + 2 * outside
    model = FixedModel(3 * (total * total - width * width), 3 * width * width).cuda()
    res = H.neighborhood_sample_generator(model, img, width=width, outside=outside, batch_size=7)
    # gather / fold are exact; the stand-in model is a cuBLAS matmul + tanh, hence the tolerance
    assert np.abs(res.cpu().numpy() - G[f"gen{k}_result"]).max() <= 2e-6
    if k == 0:
        seq = H.generate_sequence(model, img, frames=3, width=width, outside=outside, batch_size=50, skip=1, alpha=0.75)
        assert seq.shape == G["seq0_frames"].shape
        assert np.abs(seq.numpy() - G["seq0_frames"]).max() <= 5e-6


@pytest.mark.gpu
def test_cuda_target_scaling_equals_reference():
    """The fused training kernel scales uint8 targets in-kernel (u*2/255-1, single_image_dataset.py:49): with zero
    weights the prediction is 0 and loss = mean(t^2), which must equal the reference's scaled values exactly."""
    import hoir_b200 as H
    u8 = torch.from_numpy(G["jupiter_u8"]).cuda()
    flat, _, _ = H.image_to_dataset(u8.view(-1, 1, 3).contiguous(), rotations=2)
    assert np.array_equal(flat.cpu().numpy(), G["jupiter_flat"])
    net = H.HighOrderMLP(layer_type="continuous", n=3, in_width=4, out_width=3, hidden_layers=1, hidden_width=40,
                         in_segments=100, out_segments=2, hidden_segments=2, normalization=H.MaxAbsNormalization).cuda()
    with torch.no_grad():
        for p in net.parameters():
            p.zero_()
    tr = H.FusedTrainer(net, lr=0.0)
    img = u8.view(-1, 1, 3).contiguous()
    loss = tr.train_step_mesh(img, 2, 0, img.shape[0]).item()
    want = float((torch.from_numpy(G["jupiter_flat"]).double() ** 2).mean())
    assert abs(loss - want) <= 1e-6 * want
