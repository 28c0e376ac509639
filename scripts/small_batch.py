"""Step latency of the fused trainers at the reference's batch size (256): C1 (10-wide) and C2 (40-wide)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hoir_b200 as H
dev = torch.device("cuda", 0)
for name, kw in (("C1 2-10-10-10-3", dict(layer_type="continuous", n=3, in_width=2, out_width=3, hidden_layers=2, hidden_width=10, in_segments=100, out_segments=2, hidden_segments=2)),
                 ("C2 4-40-40-3", dict(layer_type="continuous", n=3, in_width=4, out_width=3, hidden_layers=1, hidden_width=40, in_segments=100, out_segments=2, hidden_segments=2))):
    net = H.HighOrderMLP(normalization=H.MaxAbsNormalization, **kw)
    H.initialize_network_polynomial_layers(net, 1.0, 0.0, generator=torch.Generator().manual_seed(0))
    tr = H.FusedTrainer(net.to(dev), optimizer="adam", lr=1e-4)
    for B in (256, 4096):
        x = torch.rand(B, kw["in_width"], device=dev) * 2 - 1
        y = torch.rand(B, 3, device=dev) * 2 - 1
        for _ in range(5): tr.train_step(x, y)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200): tr.train_step(x, y)
        e1.record(); torch.cuda.synchronize()
        print("%s batch %d: %.1f us/step" % (name, B, e0.elapsed_time(e1) / 200 * 1e3), flush=True)
