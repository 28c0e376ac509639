import sys, os, torch
sys.path.insert(0, os.getcwd())
import hoir_b200 as H
dev = torch.device("cuda", 0)
kw = dict(layer_type="continuous", n=3, in_width=2, out_width=3, hidden_layers=2, hidden_width=10, in_segments=100, out_segments=2, hidden_segments=2)
net = H.HighOrderMLP(normalization=H.MaxAbsNormalization, **kw)
H.initialize_network_polynomial_layers(net, 1.0, 0.0, generator=torch.Generator().manual_seed(0))
net = net.to(dev)
tr = H.FusedTrainer(net, lr=1e-4)
img = torch.randint(0, 256, (4096, 4096, 3), dtype=torch.uint8, device=dev)
B = 1 << 20
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ms = t(lambda: tr.train_step_mesh(img, 1, 0, B))
print("C1-shaped net, 2^20 pixels/step: %.3f ms, %.3e samples/s" % (ms, B / ms * 1e3))
out = torch.empty((4096 * 4096, 3), device=dev)
ms = t(lambda: tr.render(4096, 4096, 1, out=out), 5)
print("render 4096^2: %.3f ms, %.3e px/s" % (ms, 4096 * 4096 / ms * 1e3))
