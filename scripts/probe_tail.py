import sys, os, torch
sys.path.insert(0, os.getcwd())
import hoir_b200 as H
from bench import MLP_KW, synthetic_image
dev = torch.device("cuda", 0)
net = H.HighOrderMLP(normalization=H.MaxAbsNormalization, **MLP_KW)
H.initialize_network_polynomial_layers(net, 1.0, 0.0, generator=torch.Generator().manual_seed(0))
net = net.to(dev)
tr = H.FusedTrainer(net, lr=1e-4)
img = synthetic_image(dev)
L = H._lib
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print("step us", t(lambda: tr.train_step_mesh(img, 2, 0, 1 << 20)))
tr.step_count = 5
print("tail us", t(lambda: L.check(tr.lib.hoir_mlp_optimizer_tail(tr.desc, L.C.byref(tr._opt_desc()), L.stream_ptr(dev)))))
mesh = L.make_mesh_desc(4096, 4096, 2, True, 0)
print("kernel only us", t(lambda: L.check(tr.lib.hoir_mlp_train_step(tr.desc, 1 << 20, None, mesh, 1, L.dptr(img, torch.uint8), None, L.dptr(tr.packed), tr._g_table, tr._loss_ptr, 1e-6, L.stream_ptr(dev)))))
