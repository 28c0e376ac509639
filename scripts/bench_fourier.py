"""C3 Fourier (n_in = 31): fused whole-network kernel, training step and render over batch sizes.
usage: python scripts/bench_fourier.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import hoir_b200 as H

dev = torch.device("cuda", 0)
KW = dict(layer_type="fourier", n=3, n_in=31, n_hidden=31, n_out=3, in_width=4, out_width=3, hidden_layers=1, hidden_width=40)


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


net = H.HighOrderMLP(normalization=H.MaxAbsNormalization, **KW).to(dev)
tr = H.FusedTrainer(net, optimizer="adam", lr=1e-4)
g = torch.Generator(device=dev).manual_seed(0)
for B in [256, 4096, 1 << 16, 1 << 18, 1 << 20]:
    x = torch.rand(B, 4, device=dev, generator=g) * 2 - 1
    y = torch.rand(B, 3, device=dev, generator=g) * 2 - 1
    ms = timeit(lambda: tr.train_step(x, y), 20 if B <= (1 << 18) else 5)
    fwd = timeit(lambda: tr.forward(x), 20 if B <= (1 << 18) else 5)
    print(json.dumps({"B": B, "train_ms": ms, "samples_per_s": B / ms * 1e3, "tflops_alg": B * 319600 / ms / 1e9,
                      "fwd_ms": fwd, "fwd_samples_per_s": B / fwd * 1e3}), flush=True)
