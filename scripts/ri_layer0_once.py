import os, sys
sys.path.insert(0, os.getcwd())
import torch
import hoir_b200 as H
dev = torch.device("cuda", 0)
lay = H.high_order_fc_layers(layer_type="continuous", n=3, in_features=224, out_features=100, segments=64).to(dev)
B = 65536
x = (torch.rand(B, 224, device=dev) * 2 - 1)
gy = torch.randn(B, 100, device=dev)
for _ in range(2):
    y = lay(x)
    y.backward(gy)
torch.cuda.synchronize()
