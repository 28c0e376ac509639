"""Secondary BASELINE.json configs (C1, C3 variants, C4, C5) -- device-timed, one JSON line each.
usage: python scripts/bench_configs.py [names...]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import hoir_b200 as H

dev = torch.device("cuda", 0)
PEAK = 72.6  # TFLOP/s, measured FFMA peak (bench.py measures it live)


def timeit(fn, n, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def mlp(**kw):
    net = H.HighOrderMLP(normalization=H.MaxAbsNormalization, **kw)
    H.initialize_network_polynomial_layers(net, 1.0, 0.0, generator=torch.Generator().manual_seed(0))
    return net.to(dev)


def autograd_step(net, opt, x, y):
    opt.zero_grad(set_to_none=True)
    loss = torch.nn.functional.mse_loss(net(x).flatten(), y.flatten())
    loss.backward()
    opt.step()
    return loss


def report(name, samples, ms, flop, note):
    sps = samples / (ms * 1e-3)
    print(json.dumps({"config": name, "samples_per_s": sps, "ms_per_step": ms, "flop_per_sample": flop,
                      "achieved_tflops": sps * flop / 1e12, "frac_of_ffma_peak": sps * flop / 1e12 / PEAK, "note": note}), flush=True)


def c1():
    kw = dict(layer_type="continuous", n=3, in_width=2, out_width=3, hidden_layers=2, hidden_width=10,
              in_segments=100, out_segments=2, hidden_segments=2)
    net = mlp(**kw)
    tr = H.FusedTrainer(net, optimizer="sparse_lion", lr=1e-4)
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.rand(256, 2, device=dev, generator=g) * 2 - 1
    y = torch.rand(256, 3, device=dev, generator=g) * 2 - 1
    ms = timeit(lambda: tr.train_step(x, y), 200)
    report("C1 continuous 2-10-10-10-3, batch 256 (fused FFMA kernel, eager launches)", 256, ms, 4380, "launch-latency bound")
    gr = torch.cuda.CUDAGraph()
    tr.train_step(x, y); torch.cuda.synchronize()
    with torch.cuda.graph(gr):
        tr.train_step(x, y)
    ms = timeit(gr.replay, 500)
    report("C1 same, step captured in a CUDA graph", 256, ms, 4380, "one cooperative launch per replay (optimizer tail in-kernel)")


def cdefault():
    """config/images_config.yaml as shipped: polynomial n=3, 4-10-10-10-3, batch 256, sparse_lion."""
    kw = dict(layer_type="polynomial", n=3, in_width=4, out_width=3, hidden_layers=2, hidden_width=10,
              in_segments=100, out_segments=2, hidden_segments=2)
    net = mlp(**kw)
    assert net.fused_supported()
    tr = H.FusedTrainer(net, optimizer="sparse_lion", lr=1e-4)
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.rand(256, 4, device=dev, generator=g) * 2 - 1
    y = torch.rand(256, 3, device=dev, generator=g) * 2 - 1
    ms = timeit(lambda: tr.train_step(x, y), 200)
    flop = 3 * 2 * 3 * (4 * 10 + 10 * 10 * 2 + 10 * 3) - 2 * 3 * 4 * 10
    report("stock images_config.yaml: polynomial n=3 4-10-10-10-3, batch 256 (fused FFMA kernel, one launch per step)", 256, ms, flop, "launch-latency bound")
    img = torch.randint(0, 256, (744, 1000, 3), dtype=torch.uint8, device=dev)
    B = 1 << 18
    ms = timeit(lambda: tr.train_step_mesh(img, 2, 0, B), 20)
    report("same network, 2^18 pixels of a 1000x744 image per step (mesh mode)", B, ms, flop, "")


def fused(name, flop, **kw):
    net = mlp(**kw)
    tr = H.FusedTrainer(net, optimizer="adam", lr=1e-4)
    img = torch.randint(0, 256, (4096, 4096, 3), dtype=torch.uint8, device=dev)
    B = 1 << 20
    ms = timeit(lambda: tr.train_step_mesh(img, kw["in_width"] // 2, 0, B), 10)
    report(name, B, ms, flop, "fused whole-network kernel, mesh mode, 2^20 pixels/step")
    return tr


def c3f():
    kw = dict(layer_type="fourier", n=3, n_in=31, n_hidden=31, n_out=3, in_width=4, out_width=3, hidden_layers=1, hidden_width=40)
    net = mlp(**kw)
    opt = H.Adam(net.parameters(), lr=1e-4)
    B = 1 << 16
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.rand(B, 4, device=dev, generator=g) * 2 - 1
    y = torch.rand(B, 3, device=dev, generator=g) * 2 - 1
    ms = timeit(lambda: autograd_step(net, opt, x, y), 30)
    report("C3 fourier n_in=31 (hidden 31), 4-40-40-3, batch 2^16 (tcgen05 per-layer kernels + autograd)", B, ms, 319600, "dense_tc.cu: tf32x3 on the tensor pipe, 15 launches per step")
    for bs in (B, 256):
        net2 = mlp(**kw)
        gs = H.GraphedStep(net2, H.Adam(net2.parameters(), lr=1e-4, capturable=True), x[:bs], y[:bs])
        ms = timeit(gs.replay, 50)
        report("C3 fourier, batch %d, the whole step replayed from ONE CUDA graph (GraphedStep)" % bs, bs, ms, 319600, "")
    net3 = mlp(**kw)
    opt3 = H.Adam(net3.parameters(), lr=1e-4)
    ms = timeit(lambda: autograd_step(net3, opt3, x[:256], y[:256]), 50)
    report("C3 fourier, batch 256 (the reference's batch size), eager launches", 256, ms, 319600, "launch-bound")


def c4():
    kw = dict(layer_type="continuous", n=2, in_width=216, out_width=27, hidden_layers=2, hidden_width=50,
              in_segments=50, out_segments=2, hidden_segments=2)
    net = mlp(**kw)
    opt = H.Adam(net.parameters(), lr=1e-4)
    g = torch.Generator().manual_seed(0)
    img = (torch.randint(0, 256, (3, 2048, 2048), generator=g).float() / 127.5 - 1).to(dev)
    B = 1 << 16
    nr, nc = H.datasets.neighborhood_patches(2048, 2048, 3, 3, 1)
    feat = torch.empty(B, 216, device=dev)
    targ = torch.empty(B, 27, device=dev)
    lib, L = H._lib.lib(), H._lib

    def step():
        L.check(lib.hoir_neighborhood_gather(L.dptr(img), 3, 2048, 2048, 3, 3, 1, 0, 12345, B, L.dptr(feat), L.dptr(targ), L.stream_ptr(dev)))
        autograd_step(net, opt, feat, targ)
    ms = timeit(step, 20)
    report("C4 interpolator continuous n=2, 216-50-50-50-27, 2048^2 image, batch 2^16 patches (gather + slab dL/dw + tcgen05 hidden layers)",
           B, ms, 162600, "layer 0 = 2.2 MB gather table (CUDA cores); hidden and output layers on tcgen05")
    net2 = mlp(**kw)
    gs = H.GraphedStep(net2, H.Adam(net2.parameters(), lr=1e-4, capturable=True), feat, targ)

    def gstep():
        L.check(lib.hoir_neighborhood_gather(L.dptr(img), 3, 2048, 2048, 3, 3, 1, 0, 12345, B, L.dptr(gs.x), L.dptr(gs.y), L.stream_ptr(dev)))
        gs.replay()
    ms = timeit(gstep, 20)
    report("C4 same, gather + the whole step replayed from ONE CUDA graph (GraphedStep)", B, ms, 162600, "")
    ms = timeit(lambda: H.neighborhood_sample_generator(net, img, 3, 3), 3)
    report("C4 inference: one generate_sequence frame on 2048^2 (hoir_interp_frame: pad + gather + forward + fold + blend in ONE launch)", 2048 * 2048, ms, 68600 / 9, "per output pixel")


def c5(tr):
    rows = cols = 8192
    out = torch.empty((rows * cols, 3), device=dev)
    ms = timeit(lambda: tr.render(rows, cols, 2, out=out), 3)
    report("C5 render 8192x8192, C2 model, rotations=2, one GPU (tcgen05 forward kernel, 2 CTAs per SM)", rows * cols, ms, 11280, "12 B/pixel written")


def ri():
    """random_interpolation.yaml: 224-100-100-100-3 continuous n=3 on radial samples of 64 x 64 images."""
    S, F, rot, M = 64, 32, 2, 16
    kw = dict(layer_type="continuous", n=3, in_width=F * (3 + 2 * rot), out_width=3, hidden_layers=2, hidden_width=100,
              in_segments=S, out_segments=2, hidden_segments=2)
    net = mlp(**kw)
    opt = H.Adam(net.parameters(), lr=1e-4)
    imgs = (torch.rand(64, 3, S, S, generator=torch.Generator().manual_seed(0)) * 2 - 1).to(dev)
    stripes = H.random_sample.stripe_planes(S, rot, dev)
    ms = timeit(lambda: H.random_radial_samples_from_image(imgs, stripes, S, F), 20)
    pts = 64 * S * S * F
    print(json.dumps({"config": "radial sampler, 64 images 64x64, 32 points, rotations=2 (in-kernel Philox)", "points_per_s": pts / ms * 1e3,
                      "ms": ms, "GB_per_s_written": pts * 28 / ms / 1e6, "note": "28 B written per point; includes the output allocation"}), flush=True)
    B = M * S * S

    def step():
        feat, targ = H.random_radial_samples_from_image(imgs[:M], stripes, S, F)
        autograd_step(net, opt, feat.flatten(1), targ)
    ms = timeit(step, 10)
    report("random interpolation: radial sampling + 224-100-100-100-3 continuous n=3 (seg 64 / 2), batch 16 images = 2^16 samples",
           B, ms, 634200, "per-layer kernels + autograd")
    model = net
    ms = timeit(lambda: H.generate_sample_radial(model, features=F, iterations=1, image_size=S, rotations=rot), 10)
    report("random interpolation inference: one generate_sample_radial application on 64x64", S * S, ms, 256200, "latency-bound (4096 samples)")


if __name__ == "__main__":
    which = sys.argv[1:] or ["c1", "c2", "c3d", "c3f", "c4", "c5"]
    C2 = dict(layer_type="continuous", n=3, in_width=4, out_width=3, hidden_layers=1, hidden_width=40,
              in_segments=100, out_segments=2, hidden_segments=2)
    tr = None
    if "c1" in which: c1()
    if "default" in which: cdefault()
    if "c2" in which or "c5" in which: tr = fused("C2 continuous n=3 4-40-40-3 (tcgen05 kernel)", 32880, **C2)
    if "c3d" in which: fused("C3 discontinuous n=3 4-40-40-3 (tcgen05 kernel, 6 slots per input)", 32880, **dict(C2, layer_type="discontinuous"))
    if "c3f" in which: c3f()
    if "c4" in which: c4()
    if "c5" in which: c5(tr)
    if "ri" in which: ri()
