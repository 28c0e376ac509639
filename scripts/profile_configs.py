"""Kernel-time breakdown (torch.profiler / CUPTI) of the autograd-path configs: C4 interpolator, C3 Fourier."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
import hoir_b200 as H

dev = torch.device("cuda", 0)


def mlp(**kw):
    net = H.HighOrderMLP(normalization=H.MaxAbsNormalization, **kw)
    H.initialize_network_polynomial_layers(net, 1.0, 0.0, generator=torch.Generator().manual_seed(0))
    return net.to(dev)


def run(name, net, x, y):
    opt = H.Adam(net.parameters(), lr=1e-4)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.mse_loss(net(x).flatten(), y.flatten())
        loss.backward()
        opt.step()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            step()
        torch.cuda.synchronize()
    print("=====", name, "(5 steps)")
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=16, max_name_column_width=90))


B = 1 << 16
g = torch.Generator(device="cuda").manual_seed(0)
which = sys.argv[1:] or ["c4", "c3f"]
if "c4" in which:
    net = mlp(layer_type="continuous", n=2, in_width=216, out_width=27, hidden_layers=2, hidden_width=50,
              in_segments=50, out_segments=2, hidden_segments=2)
    run("C4", net, torch.rand(B, 216, device=dev, generator=g) * 2 - 1, torch.rand(B, 27, device=dev, generator=g) * 2 - 1)
if "c3f" in which:
    net = mlp(layer_type="fourier", n=3, n_in=31, n_hidden=31, n_out=3, in_width=4, out_width=3, hidden_layers=1, hidden_width=40)
    run("C3 fourier", net, torch.rand(B, 4, device=dev, generator=g) * 2 - 1, torch.rand(B, 3, device=dev, generator=g) * 2 - 1)
