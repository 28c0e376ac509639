"""GPU parity of the fused whole-network kernels and of the training engine against the CPU
oracle (copied weights, same seeded inputs): forward, dL/dw through autograd and through the
fused MSE step, optimizers, rendering, the Net drop-in and its checkpoints."""
import math

import pytest
import torch

import hoir_b200 as H
from oracle import hol_oracle as O
from helpers import MLP_KW, make_pair, rel_err

pytestmark = pytest.mark.gpu

FUSED = ["C1", "C1d", "C1r2", "C1n2", "C2", "C2r1", "C3d", "poly", "C3f"]
UNFUSED = ["C4s"]


def batch(name, B, seed=0):
    kw = MLP_KW[name]
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, kw["in_width"], generator=g) * 2.2 - 1.1
    y = torch.rand(B, kw["out_width"], generator=g) * 2 - 1
    return x, y


@pytest.mark.parametrize("name", FUSED + UNFUSED)
@pytest.mark.parametrize("B", [1, 255, 1000])
def test_mlp_forward_backward_parity(name, B):
    ref, net = make_pair(name)
    assert net.fused_supported() == (name in FUSED)
    x, y = batch(name, B)
    loss_ref = O.mse_training_step(ref, x, y)
    loss_ref.backward()
    y_hat = net(x.cuda())
    loss = torch.nn.functional.mse_loss(y_hat.flatten(), y.cuda().flatten())
    loss.backward()
    tol = 1e-4 if name == "C3f" else 1e-5
    assert rel_err(y_hat, ref(x)) <= tol
    assert abs(loss.item() - loss_ref.item()) <= tol * abs(loss_ref.item())
    for (k, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
        assert rel_err(p.grad, q.grad) <= 2 * tol, k


@pytest.mark.parametrize("name", FUSED)
def test_fused_matches_per_layer_kernels(name):
    _, net = make_pair(name)
    x, _ = batch(name, 777, seed=5)
    y_fused = net(x.cuda())
    net.use_fused = False
    y_layers = net(x.cuda())      # B >= 512: the hidden / output layers run on the tcgen05 per-layer kernels
    # two independent evaluation orders (FFMA chains vs tf32 x 3 tensor-core passes); a sample whose hidden
    # activation lands within round-off of a segment boundary may take the other branch;
    # Fourier: the fused kernel reduces the basis argument with sincospif, the per-layer kernels with sincosf
    assert rel_err(y_fused, y_layers) <= (1e-4 if name == "C3f" else 1e-5)


@pytest.mark.parametrize("name", ["C1", "C2", "C3d", "poly"])
@pytest.mark.parametrize("opt", ["adam", "sparse_lion"])
def test_trainer_steps_match_oracle_training(name, opt):
    ref, net = make_pair(name)
    tr = H.FusedTrainer(net, optimizer=opt, lr=1e-3, keep_grads=(opt == "adam"))
    ref_opt = torch.optim.Adam(ref.parameters(), lr=1e-3) if opt == "adam" else O.SparseLion(ref.parameters(), lr=1e-3)
    for step in range(3):
        x, y = batch(name, 600, seed=step)
        ref_opt.zero_grad()
        loss_ref = O.mse_training_step(ref, x, y)
        loss_ref.backward()
        ref_opt.step()
        loss = tr.train_step(x.cuda(), y.cuda())
        assert abs(loss.item() - loss_ref.item()) <= 1e-5 * abs(loss_ref.item())
        if step == 0 and tr.keep_grads:
            for gv, q in zip(tr.g_views, ref.parameters()):
                assert rel_err(gv, q.grad) <= 2e-5
    # sign-based Lion updates can flip on round-off-sized momenta: compare in absolute lr units
    for (k, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
        if opt == "adam":
            assert (p.detach().cpu() - q.detach()).abs().max().item() <= 5e-5, k
        else:
            frac_off = ((p.detach().cpu() - q.detach()).abs() > 1e-6).float().mean().item()
            assert frac_off <= 1e-3, k
    # state_dict still reads the live (flat-buffer) weights
    sd = net.state_dict()
    assert rel_err(sd["model.0.w"], ref.state_dict()["model.0.w"]) <= 1e-4


@pytest.mark.parametrize("rows,cols,rot,name", [(37, 53, 2, "C2"), (64, 48, 1, "C1"), (40, 40, 2, "C3d"), (33, 47, 2, "poly")])
def test_mesh_mode_training_and_render(rows, cols, rot, name):
    ref, net = make_pair(name)
    g = torch.Generator().manual_seed(11)
    img = torch.randint(0, 256, (rows, cols, 3), generator=g, dtype=torch.uint8)
    y_flat, x_flat, _ = O.image_to_dataset(img, rotations=rot)
    tr = H.FusedTrainer(net, optimizer="adam", lr=0.0, keep_grads=True)      # lr 0: weights stay put, grads visible
    first, count = 5, rows * cols - 9
    loss_ref = O.mse_training_step(ref, x_flat[first:first + count], y_flat[first:first + count])
    loss_ref.backward()
    loss = tr.train_step_mesh(img.cuda(), rot, first, count)
    assert abs(loss.item() - loss_ref.item()) <= 1e-5 * abs(loss_ref.item())
    for gv, q in zip(tr.g_views, ref.parameters()):
        assert rel_err(gv, q.grad) <= 2e-5
    # rendering.py:195-213 -- including the empty trailing batch when N % bs == 0
    want = O.render_image(ref, x_flat, (rows, cols, 3), batch_size=rows)
    got = tr.render(rows, cols, rot).reshape(rows, cols, 3)
    assert rel_err(got, want) <= 1e-5
    part = tr.render(rows, cols, rot, first_pixel=first, count=count, post_scale=False)
    assert rel_err(part, ref(x_flat[first:first + count])) <= 1e-5


@pytest.mark.parametrize("name", ["C1", "C2"])
def test_single_launch_step_equals_unfused_tail(name):
    """The in-kernel optimizer tail (grid barrier, one launch per step; gradients zeroed by the tail) must give
    the same weights as the inspection mode (memset + kernel + tail reading a kept gradient) -- both against
    the oracle's Adam and against each other."""
    ref, net_a = make_pair(name)
    _, net_b = make_pair(name)
    net_b.load_state_dict(net_a.state_dict())
    tr_a = H.FusedTrainer(net_a, optimizer="adam", lr=1e-3)
    tr_b = H.FusedTrainer(net_b, optimizer="adam", lr=1e-3, keep_grads=True)
    ref_opt = torch.optim.Adam(ref.parameters(), lr=1e-3)
    for step in range(4):
        x, y = batch(name, 3000, seed=20 + step)
        ref_opt.zero_grad()
        loss_ref = O.mse_training_step(ref, x, y)
        loss_ref.backward()
        ref_opt.step()
        la = tr_a.train_step(x.cuda(), y.cuda()).clone()
        lb = tr_b.train_step(x.cuda(), y.cuda()).clone()
        assert abs(la.item() - loss_ref.item()) <= 1e-5 * abs(loss_ref.item())
        assert abs(la.item() - lb.item()) <= 2e-6 * abs(lb.item())
        assert float(tr_a.flat_g.abs().max()) == 0.0          # left zeroed for the next step
    # Adam turns round-off-sized gradients (fp32 atomics: run-to-run summation order) into +-lr steps:
    # compare the bulk of the entries, not the worst one
    assert ((tr_a.flat_w - tr_b.flat_w).abs() > 2e-6).float().mean().item() <= 2e-3
    for (k, p), (_, q) in zip(net_a.named_parameters(), ref.named_parameters()):
        frac_off = ((p.detach().cpu() - q.detach()).abs() > 5e-5).float().mean().item()
        assert frac_off <= 2e-3, (k, frac_off)
    # weights written from outside (load_state_dict) are re-packed before the next step
    net_a.load_state_dict(net_b.state_dict())
    x, y = batch(name, 512, seed=99)
    assert abs(tr_a.train_step(x.cuda(), y.cuda()).item() - tr_b.train_step(x.cuda(), y.cuda()).item()) <= 1e-5


def test_load_state_dict_after_a_step_reaches_the_kernels():
    """after any step / render has packed the weights once, ``load_state_dict`` with CLEARLY
    different weights must reach the fused kernels (forward, render and the next step's loss)."""
    _, net = make_pair("C2", seed=0)
    ref2, _ = make_pair("C2", seed=7)
    assert max((a - b).abs().max().item() for a, b in zip(ref2.state_dict().values(),
                                                          net.cpu().state_dict().values())) > 1e-2
    net = net.cuda()
    tr = H.FusedTrainer(net, optimizer="adam", lr=1e-3)
    x, y = batch("C2", 2048, seed=3)
    tr.train_step(x.cuda(), y.cuda())
    tr.render(8, 8, 2)                                        # packed, and the tail has kept it in sync since
    net.load_state_dict(ref2.state_dict())
    assert rel_err(tr.forward(x.cuda()), ref2(x)) <= 1e-5
    pos = torch.stack([t.cpu() for t in H.positions_from_mesh(8, 8, device="cuda", rotations=2, normalize=True)],
                      dim=2).reshape(-1, 4)
    assert rel_err(tr.render(8, 8, 2, post_scale=False), ref2(pos)) <= 1e-5
    tr.lr = 0.0
    loss = tr.train_step(x.cuda(), y.cuda()).item()
    loss_ref = O.mse_training_step(ref2, x, y).item()
    assert abs(loss - loss_ref) <= 1e-5 * abs(loss_ref)
    # in-place edit through flat_w is seen as well
    with torch.no_grad():
        tr.flat_w.mul_(0.5)
        for p in ref2.parameters():
            p.mul_(0.5)
    assert rel_err(tr.forward(x.cuda()), ref2(x)) <= 1e-5


def test_fused_empty_and_edge_batches():
    ref, net = make_pair("C2")
    assert net(torch.empty(0, 4, device="cuda")).shape == (0, 3)
    tr = H.FusedTrainer(net, lr=0.0, keep_grads=True)
    assert tr.render(8, 8, 2, first_pixel=64, count=0).shape == (0, 3)
    for B in (1, 256, 257, 148 * 256 + 3):
        x, _ = batch("C2", B, seed=B)
        assert rel_err(tr.forward(x.cuda()), ref(x)) <= 1e-5


def test_large_batch_linearity_of_gradients():
    """Full-size property check (2^20 samples, too big for the oracle): dL/dw over the batch equals
    the sum over its halves, and the loss is the mean of the halves' losses."""
    _, net = make_pair("C2")
    tr = H.FusedTrainer(net, lr=0.0, keep_grads=True)
    B = 1 << 20
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.rand(B, 4, device="cuda", generator=g) * 2 - 1
    y = torch.rand(B, 3, device="cuda", generator=g) * 2 - 1
    loss = tr.train_step(x, y).clone()
    gw = tr.flat_g[:-1].clone()
    l0 = tr.train_step(x[: B // 2], y[: B // 2], global_batch=B).clone()
    g0 = tr.flat_g[:-1].clone()
    l1 = tr.train_step(x[B // 2:], y[B // 2:], global_batch=B).clone()
    g1 = tr.flat_g[:-1].clone()
    assert abs((l0 + l1).item() - loss.item()) <= 1e-5 * loss.item()
    assert rel_err(g0 + g1, gw) <= 1e-4          # fp32 atomics: summation order differs


def test_optimizer_kernels_match_torch():
    torch.manual_seed(0)
    w0 = torch.randn(1000)
    grads = [torch.randn(1000) * (torch.rand(1000) > 0.3) for _ in range(4)]
    for name in ("adam", "sparse_lion"):
        pr = torch.nn.Parameter(w0.clone())
        pc = torch.nn.Parameter(w0.clone().cuda())
        o_ref = torch.optim.Adam([pr], lr=1e-2) if name == "adam" else O.SparseLion([pr], lr=1e-2)
        o_gpu = H.Adam([pc], lr=1e-2) if name == "adam" else H.SparseLion([pc], lr=1e-2)
        for g in grads:
            pr.grad = g.clone()
            pc.grad = g.clone().cuda()
            o_ref.step()
            o_gpu.step()
        assert (pc.detach().cpu() - pr.detach()).abs().max().item() <= 1e-5, name


def test_net_drop_in_and_checkpoint(tmp_path):
    cfg = H.images_config(mlp=dict(layer_type="continuous", n=3, n_in=None, n_out=None,
                                   input=dict(segments=100, width=4), output=dict(segments=2, width=3),
                                   hidden=dict(segments=2, layers=1, width=40)),
                          optimizer=dict(name="adam", lr=1e-3, scheduler="plateau", patience=10, factor=0.1))
    net = H.Net(cfg).cuda()
    assert list(net.state_dict().keys()) == ["model.model.0.w", "model.model.2.w", "model.model.4.w"]
    assert sum(p.numel() for p in net.parameters()) == 40760
    (opt,), (sched,) = net.configure_optimizers()
    x, y = batch("C2", 256)
    losses = []
    for _ in range(5):
        opt.zero_grad()
        loss = net.training_step((x.cuda(), y.cuda()), 0)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0]
    assert "train_loss" in net.logged and sched["monitor"] == "train_loss"
    path = str(tmp_path / "epoch=0-step=5.ckpt")
    net.save_checkpoint(path)
    ckpt = torch.load(path, weights_only=False)
    assert set(ckpt) >= {"state_dict", "hyper_parameters", "epoch", "global_step"}
    net2 = H.Net.load_from_checkpoint(path).cuda()
    assert torch.equal(net2(x.cuda()), net(x.cuda()))
    with pytest.raises(ValueError, match="not recognized"):
        H.Net(H.images_config(mlp=cfg.mlp, optimizer=dict(name="sgd", lr=1.0))).configure_optimizers()


def test_ffma_peak_microbenchmark_runs():
    import ctypes as C
    ms, fl = C.c_float(), C.c_double()
    for form in (0, 1):
        H._lib.check(H._lib.lib().hoir_ffma_peak(2000, form, C.byref(ms), C.byref(fl), None))
        tflops = fl.value / (ms.value * 1e-3) / 1e12
        assert 5.0 < tflops < 120.0


@pytest.mark.parametrize("coherent,name", [(False, "C2"), (True, "C2"), (False, "C3d")])
def test_two_pass_layer0_gradient_matches_in_kernel_path(coherent, name, monkeypatch):
    """Large generic batches route the layer-0 dL/dw through the segment-owner kernel; it must agree with the
    oracle on a sub-sample and with full-batch linearity (the oracle cannot run 2^16 samples quickly)."""
    ref, net = make_pair(name)
    tr = H.FusedTrainer(net, lr=0.0, keep_grads=True)
    B = 1 << 16
    g = torch.Generator().manual_seed(9)
    x = torch.rand(B, 4, generator=g) * 2.2 - 1.1
    if coherent:
        x = x.sort(dim=0)[0]
    y = torch.rand(B, 3, generator=g) * 2 - 1
    tr.train_step(x.cuda(), y.cuda())
    g_big = [v.clone() for v in tr.g_views]
    # the same batch in chunks below the two-pass threshold (in-kernel atomics path), same global scale
    acc = [torch.zeros_like(v) for v in tr.g_views]
    for k in range(0, B, 1 << 14):
        tr.train_step(x[k:k + (1 << 14)].cuda(), y[k:k + (1 << 14)].cuda(), global_batch=B)
        for a, v in zip(acc, tr.g_views):
            a += v
    for a, v in zip(acc, g_big):
        assert rel_err(v, a) <= 2e-5
    # and against the oracle on the first 4096 samples.  The hidden layers see COMPUTED inputs, so a sample
    # whose activation lands within round-off of a segment boundary / MaxAbs tie can take the other branch
    # than the fp32 oracle (whose own fp64 evaluation differs by the same single-sample amount): accept the
    # fp32 or the fp64 evaluation of the oracle as the reference for the summed gradient.
    tr.train_step(x[:4096].cuda(), y[:4096].cuda())
    O.mse_training_step(ref, x[:4096], y[:4096]).backward()
    ref64 = O.HighOrderMLP(normalization=O.MaxAbsNormalization, **MLP_KW[name]).double()
    ref64.load_state_dict({k: v.double() for k, v in ref.state_dict().items()})
    O.mse_training_step(ref64, x[:4096].double(), y[:4096].double()).backward()
    for gv, q, q64 in zip(tr.g_views, ref.parameters(), ref64.parameters()):
        assert min(rel_err(gv, q.grad), rel_err(gv, q64.grad)) <= 2e-5


def test_host_batch_stepper_pipeline():
    ref, net = make_pair("C1")
    tr = H.FusedTrainer(net, optimizer="adam", lr=1e-3)
    st = H.HostBatchStepper(tr, 512)
    ref_opt = torch.optim.Adam(ref.parameters(), lr=1e-3)
    got, want = [], []
    for step in range(4):
        x, y = batch("C1", 512, seed=step)
        ref_opt.zero_grad()
        l = O.mse_training_step(ref, x, y)
        l.backward()
        ref_opt.step()
        want.append(l.item())
        r = st.step(x.pin_memory(), y.pin_memory())
        if r is not None:
            got.append(r)
    got.append(st.flush())
    assert len(got) == 4 and all(abs(a - b) <= 1e-4 * abs(b) for a, b in zip(got, want))


def test_quad_thread_variant_of_the_tensor_core_kernel():
    """HOIR_TC_QUAD=1 selects the four-threads-per-sample instantiation of the training kernel (measured slower than
    pairs, kept as an A/B switch): it must pass the same oracle parity.  The switch is read once per process."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, HOIR_TC_QUAD="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_fused.py"), "-m", "gpu", "-q", "-x",
                        "-k", "trainer_steps_match_oracle_training or mesh_mode_training_and_render or two_pass"],
                       env=env, cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("name", ["C2", "C3d", "C1"])
@pytest.mark.parametrize("mode", ["coherent", "two_pass"])
def test_packed_copies_follow_the_weights(name, mode):
    """The kernel-side packed buffer ([in][node][out_pad4] per layer, + for tensor-core networks the half-major second
    copy of layer 0, [in][half][node][out/2]) is written by hoir_mlp_pack and re-written by the optimizer tail (inside
    the kernel, or as its own launch after the second pass): after training steps every copy equals the live weights."""
    ref, net = make_pair(name)
    tr = H.FusedTrainer(net, optimizer="adam", lr=1e-2)
    B = 40000 if mode == "two_pass" else 700          # explicit batches >= 32768 take the two-pass path on TC networks
    for step in range(2):
        x, y = batch(name, B, seed=step)
        tr.train_step(x.cuda(), y.cuda())
    lib, desc = H._lib.lib(), tr.desc
    layers = net.high_order_layers()
    end = 0
    for l, lay in enumerate(layers):
        w = lay.w.detach()
        out_f, in_f, nodes = w.shape
        op = (out_f + 3) // 4 * 4
        off = lib.hoir_mlp_packed_offset(desc, l)
        got = tr.packed[off:off + in_f * nodes * op].view(in_f, nodes, op)
        assert torch.equal(got[:, :, :out_f], w.permute(1, 2, 0)), f"layer {l}"
        end = off + in_f * nodes * op
    assert end == lib.hoir_mlp_packed_offset(desc, len(layers))
    extra = lib.hoir_mlp_packed_size(desc) - end
    w0 = layers[0].w.detach()
    if name in ("C2", "C3d"):                          # tensor-core entries carry the second copy
        out_f, in_f, nodes = w0.shape
        assert extra == in_f * nodes * out_f
        alt = tr.packed[end:].view(in_f, 2, nodes, out_f // 2)
        want = w0.view(2, out_f // 2, in_f, nodes).permute(2, 0, 3, 1)
        assert torch.equal(alt, want)
    else:
        assert extra == 0
