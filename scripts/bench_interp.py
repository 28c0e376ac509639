"""Times one frame of the neighborhood interpolator (hoir_interp_frame, one launch) against this package's unfused
gather -> per-layer kernels -> fold path, and the fused forward on explicit feature rows.  Usage: python scripts/bench_interp.py [n]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hoir_b200 as H  # noqa: E402


def timeit(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    dev = torch.device("cuda:0")
    kw = dict(layer_type="continuous", n=n, in_width=216, out_width=27, hidden_layers=2, hidden_width=50,
              in_segments=50, out_segments=2, hidden_segments=2)
    torch.manual_seed(0)
    net = H.HighOrderMLP(normalization=H.MaxAbsNormalization, **kw).to(dev)
    for size in (1024, 2048):
        im = (torch.randint(0, 256, (3, size, size), generator=torch.Generator().manual_seed(0)).float() / 127.5 - 1).to(dev)
        patches = ((size + 9) // 3 + 1) ** 2
        net.use_fused = True
        f = timeit(lambda: H.neighborhood_sample_generator(net, im, 3, 3), 10)
        net.use_fused = False
        u = timeit(lambda: H.neighborhood_sample_generator(net, im, 3, 3), 5)
        net.use_fused = True
        print(f"n={n} frame {size}^2 ({patches} patches): fused {f:.3f} ms ({patches / f * 1e-3:.1f} Mpatch/s), unfused {u:.3f} ms")
    B = 1 << 18
    x = torch.rand(B, 216, device=dev) * 2 - 1
    with torch.no_grad():
        f = timeit(lambda: net(x), 10)
        net.use_fused = False
        u = timeit(lambda: net(x), 5)
    print(f"n={n} forward B=2^18 rows: fused {f:.3f} ms ({B / f * 1e-3:.1f} Msample/s), per-layer {u:.3f} ms")


if __name__ == "__main__":
    main()
