"""Where does the host-batch (e2e) step spend its time?  H2D alone, generic-x kernel alone, mesh kernel."""
import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hoir_b200 as H
import bench
dev = torch.device("cuda", 0)
net = H.HighOrderMLP(normalization=H.MaxAbsNormalization, **bench.MLP_KW)
H.initialize_network_polynomial_layers(net, 1.0, 0.0)
net = net.to(dev)
tr = H.FusedTrainer(net, lr=1e-4)
B = bench.BATCH
g = torch.Generator().manual_seed(0)
hx = (torch.rand(B, 4, generator=g) * 2 - 1).pin_memory()
hy = (torch.rand(B, 3, generator=g) * 2 - 1).pin_memory()
dx, dy = hx.to(dev), hy.to(dev)
img = bench.synthetic_image(dev)
coords = torch.stack(H.positions_from_mesh(4096, 4096, device="cuda", rotations=2, normalize=True), dim=2).reshape(-1, 4)
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
xbuf, ybuf = torch.empty_like(dx), torch.empty_like(dy)
print("H2D x+y (29.4 MB) ms:", timeit(lambda: (xbuf.copy_(hx, non_blocking=True), ybuf.copy_(hy, non_blocking=True))))
print("train_step random x (device) ms:", timeit(lambda: tr.train_step(dx, dy)))
cx = coords[:B].contiguous()
print("train_step coherent x (first 256 rows, device) ms:", timeit(lambda: tr.train_step(cx, dy)))
perm = torch.randperm(4096 * 4096, device=dev)[:B]
px = coords[perm].contiguous()
print("train_step shuffled image pixels (device) ms:", timeit(lambda: tr.train_step(px, dy)))
print("train_step_mesh ms:", timeit(lambda: tr.train_step_mesh(img, 2, 0, B)))
os.environ["X"] = "1"
