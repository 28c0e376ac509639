"""Fixed cost of a training step of the tensor-core kernel: step time over k tiles per SM (mesh mode, C2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hoir_b200 as H
dev = torch.device("cuda", 0)
kw = dict(layer_type="continuous", n=3, in_width=4, out_width=3, hidden_layers=1, hidden_width=40, in_segments=100, out_segments=2, hidden_segments=2)
net = H.HighOrderMLP(normalization=H.MaxAbsNormalization, **kw)
H.initialize_network_polynomial_layers(net, 1.0, 0.0, generator=torch.Generator().manual_seed(0))
tr = H.FusedTrainer(net.to(dev), optimizer="adam", lr=1e-4)
img = torch.randint(0, 256, (4096, 4096, 3), dtype=torch.uint8, device=dev)
sms = torch.cuda.get_device_properties(0).multi_processor_count
for k in (1, 2, 4, 8, 16, 32, 55, 56):
    B = 128 * sms * k
    for _ in range(5): tr.train_step_mesh(img, 2, 0, B)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100): tr.train_step_mesh(img, 2, 0, B)
    e1.record(); torch.cuda.synchronize()
    print("tiles per SM %2d: %.1f us/step" % (k, e0.elapsed_time(e1) / 100 * 1e3), flush=True)
