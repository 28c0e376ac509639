"""The caller of the hot path: epoch loop, lr schedule, render, Lightning-format checkpoint."""
import os

import pytest
import torch

import hoir_b200 as H

pytestmark = pytest.mark.gpu


def _cfg(**kw):
    cfg = H.images_config(mlp=dict(layer_type="continuous", n=3, n_in=None, n_out=None, n_hidden=None,
                                   periodicity=None, rescale_output=False, scale=2.0,
                                   input=dict(segments=20, width=4), output=dict(segments=2, width=3),
                                   hidden=dict(segments=2, layers=1, width=40)),
                          batch_size=4096, max_epochs=6, rotations=2,
                          optimizer=dict(name="adam", lr=3e-3, scheduler="plateau", patience=2, factor=0.5))
    for k, v in kw.items():
        cfg[k] = v
    return cfg


def _card(rows=96, cols=128):
    u = torch.linspace(0, 1, cols).view(1, cols, 1)
    v = torch.linspace(0, 1, rows).view(rows, 1, 1)
    img = 0.5 + 0.5 * torch.sin(6.0 * u + 4.0 * v + torch.tensor([0.0, 2.0, 4.0]).view(1, 1, 3))
    return (img * 255).round().to(torch.uint8)


@pytest.mark.parametrize("sampling", ["shuffle", "tiles"])
def test_image_trainer_fits_and_checkpoints(sampling, tmp_path):
    torch.manual_seed(0)
    tr = H.ImageTrainer(_cfg(), _card(), sampling=sampling)
    assert tr.fused is not None
    hist = tr.fit()
    assert len(hist) == 6 and hist[-1] < 0.5 * hist[0]          # the loss goes down
    img = tr.render()
    assert img.shape == (96, 128, 3) and torch.isfinite(img).all()
    path = os.path.join(tmp_path, "last.ckpt")
    tr.save_checkpoint(path)
    ckpt = torch.load(path, weights_only=False)
    assert set(ckpt["state_dict"]) == {"model.model.0.w", "model.model.2.w", "model.model.4.w"}
    assert ckpt["hyper_parameters"]["mlp"]["hidden"]["width"] == 40
    net = H.Net.load_from_checkpoint(path, map_location="cuda")
    again = H.render_image(net, 96, 128, 2)
    assert torch.allclose(again, img, atol=1e-6)


def test_image_trainer_autograd_path_for_fourier():
    cfg = _cfg(max_epochs=2)
    cfg.mlp.layer_type = "fourier"
    cfg.mlp.n = 5
    torch.manual_seed(0)
    tr = H.ImageTrainer(cfg, _card(48, 64), sampling="shuffle")
    assert tr.fused is None                                     # per-layer (tensor-core) kernels under autograd
    hist = tr.fit()
    assert len(hist) == 2 and all(h == h for h in hist) and hist[1] < hist[0]
    # the same run with every full batch replayed from ONE CUDA graph (GraphedStep): same trajectory
    torch.manual_seed(0)
    tg = H.ImageTrainer(cfg, _card(48, 64), sampling="shuffle", cuda_graph=True)
    assert tg._graphed is not None
    hist_g = tg.fit()
    for a, b in zip(hist, hist_g):
        assert abs(a - b) <= 2e-4 * abs(a)


def test_plateau_schedule_matches_torch():
    from hoir_b200.trainer import _LrSchedule
    opt = dict(name="adam", lr=1e-2, scheduler="plateau", patience=2, factor=0.5)
    p = torch.nn.Parameter(torch.zeros(1))
    ref_opt = torch.optim.SGD([p], lr=1e-2)
    ref = torch.optim.lr_scheduler.ReduceLROnPlateau(ref_opt, patience=2, factor=0.5)
    mine = _LrSchedule(H.Config(opt), 1e-2)
    for m in [1.0, 0.9, 0.95, 0.96, 0.97, 0.98, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0]:
        ref.step(m)
        assert abs(mine.step(m) - ref_opt.param_groups[0]["lr"]) < 1e-12


def test_resume_from_checkpoint_continues_the_same_trajectory(tmp_path):
    """Train 2 + 2 epochs with a save / load in between == train 4 epochs straight (same shuffles, same Adam moments,
    same bias-correction step): the checkpoint carries weights, optimizer state, lr schedule and counters."""
    cfg = _cfg(max_epochs=4)
    a = H.ImageTrainer(cfg, _card(64, 64), sampling="tiles", seed=3)
    b = H.ImageTrainer(cfg, _card(64, 64), sampling="tiles", seed=3)
    b.net.load_state_dict(a.net.state_dict())
    hist_a = a.fit(4)
    b.fit(2)
    path = os.path.join(tmp_path, "mid.ckpt")
    b.save_checkpoint(path)
    c = H.ImageTrainer(cfg, _card(64, 64), sampling="tiles", seed=3)
    c.load_checkpoint(path)
    c.gen.set_state(b.gen.get_state())                     # the data order is the caller's business
    assert c.fused.step_count == b.fused.step_count and c.net.current_epoch == 2
    hist_c = c.fit(2)
    assert c.net.current_epoch == 4
    # fp32 atomics make runs differ in the last bits; Adam keeps them close over a few epochs
    for x, y in zip(hist_a[2:], hist_c[2:]):
        assert abs(x - y) <= 2e-3 * abs(x)
    ck = torch.load(path, weights_only=False)
    assert set(ck["optimizer_states"][0]["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
    assert ck["optimizer_states"][0]["state"][0]["exp_avg"].shape == a.net.model.model[0].w.shape


def test_neighborhood_trainer_fits_and_generates():
    """examples/implicit_neighborhood.py flow: patches gathered per step from the resident image, per-layer kernels
    under autograd, then the iterated interpolator."""
    cfg = H.neighborhood_config(batch_size=4096, max_epochs=3)
    cfg.mlp.periodicity = None
    cfg.mlp.n = 2
    cfg.mlp.input.segments = 10
    cfg.mlp.hidden.width = 20
    u = torch.linspace(0, 1, 96).view(1, 1, 96)
    v = torch.linspace(0, 1, 80).view(1, 80, 1)
    img = torch.sin(9 * u + torch.tensor([0.0, 1.0, 2.0]).view(3, 1, 1)) * torch.cos(7 * v) * 0.8
    torch.manual_seed(0)
    tr = H.NeighborhoodTrainer(cfg, [img])
    assert cfg.mlp.input.width == 216 and cfg.mlp.output.width == 27
    hist = tr.fit()
    assert len(hist) == 3 and hist[-1] < hist[0]
    seq = tr.generate(image=img, frames=4, skip=2)
    assert seq.shape == (3, 3, 80, 96) and torch.isfinite(seq).all()


def test_random_interpolation_trainer_fits_and_generates(tmp_path):
    """examples/random_interpolation.py flow (config/random_interpolation.yaml scaled down): radial samples drawn
    in-kernel each step, loss decreases, Lightning-format checkpoint, iterated generation keeps the shapes."""
    g = torch.Generator().manual_seed(0)
    S = 16
    u = torch.linspace(0, 1, S).view(1, 1, 1, S)
    v = torch.linspace(0, 1, S).view(1, 1, S, 1)
    f = torch.rand(6, 3, 1, 1, generator=g) * 4 + 1
    imgs = torch.sin(f * u) * torch.cos(f * v) * 0.9
    cfg = H.random_interpolation_config(image_size=S, num_feature_pixels=8, max_epochs=6, batch_size=3)
    cfg.mlp.hidden.width = 20
    cfg.optimizer.lr = 2e-3
    tr = H.RandomInterpolationTrainer(cfg, imgs)
    assert tr.cfg.mlp.input.width == 8 * 7 and tr.cfg.mlp.input.segments == S
    hist = tr.fit()
    assert len(hist) == 6 and hist[-1] < 0.7 * hist[0]
    path = str(tmp_path / "ri.ckpt")
    tr.save_checkpoint(path)
    net = H.Net.load_from_checkpoint(path, map_location="cuda")
    for a, b in zip(net.state_dict().values(), tr.net.state_dict().values()):
        assert torch.equal(a, b)
    frames = tr.generate(iterations=3)
    assert len(frames) == 3 and frames[0].shape == (3, S, S) and all(torch.isfinite(f).all() for f in frames)
    frames = tr.generate(start_image=imgs[0].cuda(), iterations=2, all_random=False)
    assert len(frames) == 2


@pytest.mark.parametrize("kind,opt", [("fourier", "adam"), ("c4", "adam"), ("c4", "sparse_lion")])
def test_graphed_step_equals_eager_training(kind, opt):
    """GraphedStep: the whole autograd-path training step replayed from one CUDA graph follows the eager trajectory
    (same kernels, same order): weights after 4 steps agree to round-off (the atomics of dL/dw are unordered), the
    Adam bias corrections and a mid-run lr change reach the captured kernels through the device scalars, and
    building the graph does not move the weights."""
    kw = (dict(layer_type="fourier", n=3, n_in=7, n_hidden=7, in_width=4, out_width=3, hidden_layers=1, hidden_width=20)
          if kind == "fourier" else
          dict(layer_type="continuous", n=2, in_width=24, out_width=12, hidden_layers=2, hidden_width=20,
               in_segments=20, out_segments=2, hidden_segments=2))
    B = 600

    def make():
        torch.manual_seed(3)
        net = H.HighOrderMLP(normalization=H.MaxAbsNormalization, **kw)
        H.initialize_network_polynomial_layers(net, 1.0, 0.0, generator=torch.Generator().manual_seed(0))
        return net.cuda()
    g = torch.Generator().manual_seed(1)
    xs = [(torch.rand(B, kw["in_width"], generator=g) * 2 - 1).cuda() for _ in range(4)]
    ys = [(torch.rand(B, kw["out_width"], generator=g) * 2 - 1).cuda() for _ in range(4)]
    Opt = H.Adam if opt == "adam" else H.SparseLion
    a, b = make(), make()
    oa, ob = Opt(a.parameters(), lr=1e-3), Opt(b.parameters(), lr=1e-3, capturable=True)
    before = [p.detach().clone() for p in b.parameters()]
    gs = H.GraphedStep(b, ob, xs[0], ys[0])
    for p, q in zip(b.parameters(), before):
        assert torch.equal(p, q)
    for k in range(4):
        if k == 2:
            for grp in oa.param_groups + ob.param_groups:
                grp["lr"] = 3e-4
        oa.zero_grad(set_to_none=True)
        la = torch.nn.functional.mse_loss(a(xs[k]).flatten(), ys[k].flatten())
        la.backward()
        oa.step()
        lb = gs.step(xs[k], ys[k])
        assert abs(la.item() - lb.item()) <= 1e-5 * abs(la.item())
    for p, q in zip(a.parameters(), b.parameters()):
        if opt == "adam":
            assert (p - q).abs().max().item() <= 2e-5
        else:       # sign updates: entries with round-off-sized momenta may flip
            assert ((p - q).abs() > 1e-6).float().mean().item() <= 2e-3


def test_neighborhood_trainer_with_cuda_graph_matches_eager():
    """NeighborhoodTrainer(cuda_graph=True): full batches replay one CUDA graph fed by the gather kernel, the ragged
    last block runs eagerly; same loss history as the eager trainer (to round-off), lr schedule included."""
    g = torch.Generator().manual_seed(0)
    img = torch.rand(3, 40, 44, generator=g) * 2 - 1
    hist = []
    for graph in (False, True):
        cfg = H.neighborhood_config(batch_size=256, max_epochs=3)
        cfg.mlp.n, cfg.mlp.input.segments, cfg.mlp.output.segments = 2, 10, 2
        cfg.mlp.hidden.width, cfg.mlp.hidden.layers = 20, 2
        cfg.data.outside = 2
        cfg.optimizer.lr = 1e-3
        torch.manual_seed(5)
        tr = H.NeighborhoodTrainer(cfg, [img], seed=1, cuda_graph=graph)
        hist.append(tr.fit())
    assert len(hist[0]) == 3 and hist[0][-1] < hist[0][0]
    for a, b in zip(*hist):
        assert abs(a - b) <= 2e-4 * abs(a)
