import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import hoir_b200 as H
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
KW = dict(layer_type="fourier", n=3, n_in=31, n_hidden=31, n_out=3, in_width=4, out_width=3, hidden_layers=1, hidden_width=40)
net = H.HighOrderMLP(normalization=H.MaxAbsNormalization, **KW).cuda()
tr = H.FusedTrainer(net, optimizer="adam", lr=1e-4)
x = torch.rand(B, 4, device="cuda") * 2 - 1
y = torch.rand(B, 3, device="cuda") * 2 - 1
tr.forward(x); torch.cuda.synchronize()
print("fwd"); tr.forward(x); torch.cuda.synchronize()
if len(sys.argv) > 2:
    print("train"); tr.train_step(x, y); torch.cuda.synchronize()
