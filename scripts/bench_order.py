"""Explicit shuffled-pixel batches of the 4096^2 image: step time of the cell-sorted path vs the two-pass / unsorted
path over batch sizes.  usage: python scripts/bench_order.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import hoir_b200 as H

dev = torch.device("cuda", 0)
KW = dict(layer_type="continuous", n=3, in_width=4, out_width=3, hidden_layers=1, hidden_width=40,
          in_segments=100, out_segments=2, hidden_segments=2)


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


net = H.HighOrderMLP(normalization=H.MaxAbsNormalization, **KW)
H.initialize_network_polynomial_layers(net, 1.0, 0.0, generator=torch.Generator().manual_seed(0))
tr = H.FusedTrainer(net.to(dev), optimizer="adam", lr=1e-4)
rows = cols = 4096
g = torch.Generator(device=dev).manual_seed(0)
img = torch.randint(0, 256, (rows, cols, 3), device=dev, dtype=torch.uint8, generator=g)
pos = torch.stack(H.positions_from_mesh(rows, cols, device="cuda", rotations=2, normalize=True), dim=2).reshape(-1, 4)
for B in [4096, 8192, 16384, 32768, 65536, 1 << 18, 1 << 20]:
    idx = torch.randint(0, rows * cols, (B,), device=dev, generator=g)
    x = pos[idx].contiguous()
    y = (img.reshape(-1, 3)[idx].float() * 2 / 255 - 1).contiguous()
    idx32 = idx.to(torch.int32)
    res = {"B": B}
    tr.sort_min_batch = 0
    for mode in ("sorted", "two_pass"):
        tr.batch_order = mode
        res[mode + "_ms"] = timeit(lambda: tr.train_step(x, y), 20)
    res["pixels_ms"] = timeit(lambda: tr.train_step_pixels(img, 2, idx32), 20)
    res["mesh_ms"] = timeit(lambda: tr.train_step_mesh(img, 2, 0, B), 20)
    order = torch.empty(B, dtype=torch.int32, device=dev)
    L = H._lib
    res["sort_only_ms"] = timeit(lambda: L.check(L.lib().hoir_mlp_batch_sort(tr.desc, B, L.dptr(x), None, None, L.dptr(order, torch.int32), None, L.stream_ptr(dev))), 20)
    print(json.dumps(res), flush=True)
