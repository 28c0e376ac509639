"""A few fused training steps on an INCOHERENT batch (random coordinates): main kernel -> gw0_slab_kernel -> tail."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hoir_b200 as H
dev = torch.device("cuda", 0)
net = H.HighOrderMLP(normalization=H.MaxAbsNormalization, layer_type="continuous", n=3, in_width=4, out_width=3,
                     hidden_layers=1, hidden_width=40, in_segments=100, out_segments=2, hidden_segments=2)
H.initialize_network_polynomial_layers(net, 1.0, 0.0, generator=torch.Generator().manual_seed(0))
tr = H.FusedTrainer(net.to(dev), lr=1e-4)
B = 1 << 20
x = torch.rand(B, 4, device=dev) * 2 - 1
y = torch.rand(B, 3, device=dev) * 2 - 1
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
if len(sys.argv) > 2 and sys.argv[2] == "coherent":      # the same explicit-x path on a coherent batch (256 image rows)
    lines = H.positions_from_mesh(4096, 4096, device="cuda", rotations=2, normalize=True)
    x = torch.stack(lines, dim=2).reshape(-1, 4)[:B].contiguous()
for _ in range(3):
    tr.train_step(x, y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    tr.train_step(x, y)
e1.record()
torch.cuda.synchronize()
print("incoherent step: %.3f ms" % (e0.elapsed_time(e1) / n))
