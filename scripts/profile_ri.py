"""Kernel-time breakdown (torch.profiler / CUPTI) of the random-interpolation step and one inference application."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
import hoir_b200 as H

dev = torch.device("cuda", 0)
S, F, rot, M = 64, 32, 2, int(os.environ.get("M", "16"))
net = H.HighOrderMLP(normalization=H.MaxAbsNormalization, layer_type="continuous", n=3, in_width=F * 7, out_width=3,
                     hidden_layers=2, hidden_width=100, in_segments=S, out_segments=2, hidden_segments=2)
H.initialize_network_polynomial_layers(net, 1.0, 0.0, generator=torch.Generator().manual_seed(0))
net = net.to(dev)
opt = H.Adam(net.parameters(), lr=1e-4)
imgs = (torch.rand(M, 3, S, S) * 2 - 1).to(dev)
stripes = H.random_sample.stripe_planes(S, rot, dev)


def step():
    feat, targ = H.random_radial_samples_from_image(imgs, stripes, S, F)
    opt.zero_grad(set_to_none=True)
    loss = torch.nn.functional.mse_loss(net(feat.flatten(1)).flatten(), targ.flatten())
    loss.backward()
    opt.step()


def infer():
    H.generate_sample_radial(net, features=F, iterations=1, image_size=S, rotations=rot)


for name, fn in (("train step", step), ("inference", infer)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
    print("=====", name, "(3 iterations)")
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))
