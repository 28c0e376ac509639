import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import hoir_b200 as H
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
size = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
KW = dict(layer_type="continuous", n=n, in_width=216, out_width=27, hidden_layers=2, hidden_width=50,
          in_segments=50, out_segments=2, hidden_segments=2)
net = H.HighOrderMLP(normalization=H.MaxAbsNormalization, **KW).cuda()
im = (torch.randint(0, 256, (3, size, size)).float() / 127.5 - 1).cuda()
H.neighborhood_sample_generator(net, im, 3, 3); torch.cuda.synchronize()
print("frame", flush=True)
H.neighborhood_sample_generator(net, im, 3, 3); torch.cuda.synchronize()
