"""The random-interpolation path: radial sampler kernels against golden vectors made
by RUNNING THE REFERENCE (tests/golden/make_reference_radial_golden.py) and against the oracle restatement;
GenerativeNetwork (8f rank 4) and the 224-wide interpolation network against the oracle.

Bar: the uniforms (r, theta) are inputs, so pixel indices and gathered values are bit-exact -- except where the fp32
product (0.5 r) cos(theta) S lies within 1e-4 of an integer (the reference's CPU libm and CUDA's cosf may differ by
an ulp there and the truncation flips); such points are counted and must be rarer than 1 in 10 000."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import hol_oracle as O
from helpers import FixedModel, rel_err

R = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_radial_v1.npz"))


def _t(name):
    return torch.from_numpy(R[name])


def _near_integer(r, theta, size):
    """points whose offset product is within 1e-4 of an integer (in double): truncation may legitimately flip"""
    r, theta = r.double(), theta.double()
    px, py = 0.5 * r * torch.cos(theta) * size, 0.5 * r * torch.sin(theta) * size
    return ((px - px.round()).abs() < 1e-4) | ((py - py.round()).abs() < 1e-4)


# ---- CPU: the oracle restatement equals the reference's outputs ------------------------------------------
@pytest.mark.parametrize("k", range(3))
def test_oracle_symmetric_sample_equals_reference(k):
    size, f = [int(v) for v in R[f"sym{k}_geom"]]
    xy = O.random_symmetric_sample(size, _t(f"sym{k}_samples").long(), _t(f"sym{k}_r"), _t(f"sym{k}_theta"))
    assert np.array_equal(xy.numpy(), R[f"sym{k}_xy"])
    assert xy.min() >= 0 and xy.max() < size          # the reference test's own assertion


@pytest.mark.parametrize("k", range(3))
def test_oracle_radial_samples_equal_reference(k):
    feat, targ = O.random_radial_samples_from_image(_t(f"rad{k}_image"), _t(f"rad{k}_stripes"), _t(f"rad{k}_r"),
                                                    _t(f"rad{k}_theta"))
    assert np.array_equal(feat.numpy(), R[f"rad{k}_features"])
    assert np.array_equal(targ.numpy(), R[f"rad{k}_targets"])


def test_reference_shape_pins():
    """/root/reference/tests/test_random_sample_dataset.py:12-36: 32 x 32 images, 25 points, 1 rotation ->
    1024 examples per image, 25 x 5 features, 1 x 3 targets."""
    assert R["rad1_features"].shape == (1024, 25, 5) and R["rad1_targets"].shape == (1024, 1, 3)


# ---- GPU: the kernels -------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("k", range(3))
def test_cuda_symmetric_sample_equals_reference(k):
    import hoir_b200 as H
    size, f = [int(v) for v in R[f"sym{k}_geom"]]
    r, theta = _t(f"sym{k}_r"), _t(f"sym{k}_theta")
    xy = H.random_symmetric_sample(size, f, _t(f"sym{k}_samples"), device="cuda", r=r, theta=theta).cpu()
    want = torch.from_numpy(R[f"sym{k}_xy"]).long()
    bad = (xy != want).any(dim=2)
    assert not (bad & ~_near_integer(r, theta, size)).any()
    assert bad.float().mean().item() <= 1e-4
    assert xy.dtype == torch.int64 and xy.min() >= 0 and xy.max() < size


@pytest.mark.gpu
@pytest.mark.parametrize("k", range(3))
def test_cuda_radial_samples_equal_reference(k):
    import hoir_b200 as H
    size, f, rot = [int(v) for v in R[f"rad{k}_geom"]]
    r, theta = _t(f"rad{k}_r"), _t(f"rad{k}_theta")
    feat, targ = H.random_radial_samples_from_image(_t(f"rad{k}_image").cuda(), _t(f"rad{k}_stripes").cuda(), size, f,
                                                    r=r, theta=theta)
    assert np.array_equal(targ.cpu().numpy(), R[f"rad{k}_targets"])
    bad = (feat.cpu() != torch.from_numpy(R[f"rad{k}_features"])).any(dim=2)
    assert not (bad & ~_near_integer(r, theta, size)).any()
    assert bad.float().mean().item() <= 1e-4


@pytest.mark.gpu
def test_cuda_radial_batch_and_generic_width():
    """A batch of images in one launch equals image-by-image calls; 4 channels / 3 rotations take the generic kernel."""
    import hoir_b200 as H
    g = torch.Generator().manual_seed(5)
    for (M, C, S, F, rot) in [(3, 3, 16, 6, 2), (2, 4, 9, 5, 3), (2, 3, 8, 1, 1)]:
        imgs = (torch.rand(M, C, S, S, generator=g) * 2 - 1).cuda()
        stripes = H.random_sample.stripe_planes(S, rot, "cuda")
        r = torch.rand(M * S * S, F, generator=g)
        theta = torch.rand(M * S * S, F, generator=g) * 2 * math.pi
        feat, targ = H.random_radial_samples_from_image(imgs, stripes, S, F, r=r, theta=theta)
        assert feat.shape == (M * S * S, F, C + 2 * rot) and targ.shape == (M * S * S, 1, C)
        for m in range(M):
            sl = slice(m * S * S, (m + 1) * S * S)
            f1, t1 = H.random_radial_samples_from_image(imgs[m], stripes, S, F, r=r[sl], theta=theta[sl])
            assert torch.equal(f1, feat[sl]) and torch.equal(t1, targ[sl])
            of, ot = O.random_radial_samples_from_image(imgs[m].cpu(), stripes.cpu(), r[sl], theta[sl])
            bad = (f1.cpu() != of).any(dim=2)
            assert not (bad & ~_near_integer(r[sl], theta[sl], S)).any()
            assert torch.equal(t1.cpu(), ot)


@pytest.mark.gpu
def test_cuda_in_kernel_random_stream():
    """The Philox stream: explicit replay == in-kernel draw, reproducible under (seed, offset), distinct across
    offsets, uniform on [0, 1) x [0, 2 pi), and torch.manual_seed makes the default stream reproducible."""
    import hoir_b200 as H
    S, F = 32, 25
    img = (torch.rand(3, S, S) * 2 - 1).cuda()
    stripes = H.random_sample.stripe_planes(S, 2, "cuda")
    r, theta = H.random_sample.radial_uniforms(S * S * F, seed=11, offset=3)
    a = H.random_radial_samples_from_image(img, stripes, S, F, seed=11, offset=3)
    b = H.random_radial_samples_from_image(img, stripes, S, F, r=r.view(-1, F), theta=theta.view(-1, F))
    c = H.random_radial_samples_from_image(img, stripes, S, F, seed=11, offset=4)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert not torch.equal(a[0], c[0])
    xy1 = H.random_symmetric_sample(S, F, H.indices_from_grid(S)[0], seed=11, offset=3)
    lin = (xy1[:, :, 0] + xy1[:, :, 1] * S).cpu()
    assert torch.equal(img.reshape(3, -1).t()[lin.cuda()], a[0][:, :, :3])          # same points as the gather
    n = 1 << 20
    r, theta = H.random_sample.radial_uniforms(n, seed=1, offset=0)
    assert 0 <= r.min() and r.max() < 1 and 0 <= theta.min() and theta.max() < 2 * math.pi + 1e-6
    assert abs(r.mean().item() - 0.5) < 2e-3 and abs(r.var().item() - 1 / 12) < 1e-3
    assert abs(theta.mean().item() - math.pi) < 1e-2
    hist = torch.histc(r, bins=64, min=0, max=1)
    assert (hist - n / 64).abs().max().item() < 6 * math.sqrt(n / 64)
    assert abs(torch.corrcoef(torch.stack([r, theta]))[0, 1].item()) < 5e-3
    torch.manual_seed(123)
    d1 = H.random_radial_samples_from_image(img, stripes, S, F, offset=0)
    torch.manual_seed(123)
    d2 = H.random_radial_samples_from_image(img, stripes, S, F, offset=0)
    assert torch.equal(d1[0], d2[0])


@pytest.mark.gpu
def test_cuda_radial_rejects_bad_arguments():
    import hoir_b200 as H
    img = torch.zeros(3, 8, 8, device="cuda")
    stripes = H.random_sample.stripe_planes(8, 1, "cuda")
    with pytest.raises(ValueError):
        H.random_radial_samples_from_image(img, stripes, 8, 4, r=torch.zeros(64, 4))
    with pytest.raises(ValueError):
        H.random_radial_samples_from_image(img, stripes, 9, 4)
    with pytest.raises(RuntimeError):
        H.random_radial_samples_from_image(img.cpu(), stripes, 8, 4)
    with pytest.raises(ValueError):
        H.random_radial_samples_from_image(img, stripes, 8, 4, r=torch.zeros(64, 5), theta=torch.zeros(64, 5))


@pytest.mark.gpu
def test_cuda_generate_sample_radial_equals_reference():
    """utils.py:20-97 replayed with the reference's uniforms and the fixed stand-in model (cuBLAS matmul + tanh:
    values to 1e-5; every application transposes the picture, as the reference does)."""
    import hoir_b200 as H
    size, f, rot, iters = [int(v) for v in R["gen_geom"]]
    model = FixedModel(f * (3 + 2 * rot), 3).cuda()
    uni = [(_t("gen_r")[i], _t("gen_theta")[i]) for i in range(iters)]
    frames = H.generate_sample_radial(model, features=f, iterations=iters, start_image=_t("gen_start").cuda(),
                                      image_size=size, rotations=rot, all_random=False, uniforms=uni)
    got = torch.stack(frames).cpu().numpy()
    assert got.shape == R["gen_frames"].shape
    assert np.abs(got - R["gen_frames"]).max() <= 1e-5
    # the other modes run and keep the shapes
    out = H.generate_sample_radial(model, features=f, iterations=2, image_size=size, rotations=rot, batch_size=50)
    assert len(out) == 2 and out[0].shape == (3, size, size)
    out = H.generate_sample(model, features=f, iterations=2, image_size=size, rotations=rot)
    assert len(out) == 2 and out[0].shape == (3, size, size) and torch.isfinite(out[1]).all()


@pytest.mark.gpu
def test_uniform_sampler_shapes_and_disjointness():
    """RandomImageSampleDataset (random_sample_dataset.py:41-119): 32 x 32, 25 + 1 pixels -> 39 examples per image
    (the reference test pins 78 for two images), features and targets never share a pixel."""
    import hoir_b200 as H
    imgs = torch.rand(2, 3, 32, 32) * 2 - 1
    ds = H.RandomImageSampleDataset(32, imgs, num_feature_pixels=25, num_target_pixels=1, rotations=1)
    feat, targ = ds.batch([0, 1])
    assert feat.shape == (78, 25, 5) and targ.shape == (78, 1, 3)
    stripes = H.random_sample.stripe_planes(32, 1, "cuda")
    perm = torch.randperm(1024, device="cuda")
    f1, t1 = H.random_sample.random_image_samples(imgs[0].cuda(), stripes, 25, 1, perm=perm)
    planes = torch.cat([imgs[0].cuda(), stripes]).reshape(5, -1).t()
    assert torch.equal(f1[:, :, :3], planes[perm[:975]].reshape(39, 25, 5)[:, :, :3])
    assert torch.equal(t1, planes[perm[975:1014]].reshape(39, 1, 5)[:, :, :3])
    rd = H.RadialRandomImageSampleDataset(32, imgs, num_feature_pixels=25, rotations=1)
    feat, targ = rd.batch([0, 1])
    assert feat.shape == (2048, 25, 5) and targ.shape == (2048, 1, 3)          # the reference test's 2048
    assert len(rd) == 2 and rd[1][0].shape == (1024, 25, 5)


# ---- the networks fed by these samplers ----------------------------------------------------------------------
@pytest.mark.gpu
def test_interpolation_network_step_matches_oracle():
    """Net(random_interpolation config, scaled down in image size only): 224 -> 100 -> 100 -> 100 -> 3 continuous
    n=3 network on radial samples; forward, loss and every dL/dw against the oracle (fp32, 1e-5 relative)."""
    import hoir_b200 as H
    S, F, rot = 16, 32, 2
    cfg = H.random_interpolation_config(image_size=S)
    cfg.mlp.input.segments = S
    torch.manual_seed(0)
    net = H.Net(cfg).cuda()
    ref = O.build_net(H.config.cfg_to_dict(cfg)["mlp"])
    ref.load_state_dict(net.model.state_dict())
    img = (torch.rand(2, 3, S, S) * 2 - 1).cuda()
    stripes = H.random_sample.stripe_planes(S, rot, "cuda")
    feat, targ = H.random_radial_samples_from_image(img, stripes, S, F, seed=3, offset=0)
    loss = net.training_step((feat, targ), 0)
    loss.backward()
    ref_loss = O.mse_training_step(ref, feat.cpu(), targ.cpu())
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= 1e-5 * abs(ref_loss.item())
    for (name, p), q in zip(net.model.named_parameters(), ref.parameters()):
        assert rel_err(p.grad, q.grad) <= 1e-5, name


@pytest.mark.gpu
@pytest.mark.parametrize("layer_type", ["continuous", "discontinuous"])
def test_generative_network_matches_oracle(layer_type):
    """GenerativeNetwork (networks.py:132-180) and GenNet.eval_step: forward, loss, all gradients vs the oracle."""
    import hoir_b200 as H
    cfg = H.generative_config(layer_type=layer_type)
    cfg.mlp.width = 40
    torch.manual_seed(1)
    gen = H.GenNet(cfg).cuda()
    ref = O.GenerativeNetwork(embedding_size=384, input_size=2, output_size=3, mlp_width=40, mlp_layers=2,
                              input_segments=20, mlp_segments=2, n=3, normalization=O.MaxAbsNormalization,
                              layer_type=layer_type)
    with torch.no_grad():
        for p in gen.parameters():
            p.add_(0.2 * torch.randn_like(p) / p.shape[1] ** 0.5)
    ref.load_state_dict(gen.model.state_dict())
    B = 700
    caption = (torch.rand(B, 384) * 2 - 1).cuda()
    pos = (torch.rand(B, 2) * 2 - 1).cuda()
    color = (torch.rand(B, 3) * 2 - 1).cuda()
    y = gen(caption, pos)
    y_ref = ref(caption.cpu(), pos.cpu())
    assert rel_err(y, y_ref) <= 1e-5
    loss = gen.training_step((caption, pos, color), 0)
    loss.backward()
    ref_loss = torch.nn.functional.mse_loss(y_ref.flatten(), color.cpu().flatten())
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= 1e-5 * abs(ref_loss.item())
    for (name, p), q in zip(gen.model.named_parameters(), ref.parameters()):
        assert rel_err(p.grad, q.grad) <= 2e-5, name
    assert sorted(gen.state_dict()) == sorted("model." + k for k in ref.state_dict())
    opt = gen.configure_optimizers()
    assert isinstance(opt[0][0], H.SparseLion)
