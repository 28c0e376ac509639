import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import hoir_b200 as H
from oracle import hol_oracle as O
from helpers import make_pair, rel_err
ref, net = make_pair("C2")
tr = H.FusedTrainer(net, lr=0.0)
g = torch.Generator().manual_seed(9)
X = torch.rand(1 << 16, 4, generator=g) * 2.2 - 1.1
Y = torch.rand(1 << 16, 3, generator=g) * 2 - 1
for B in (1000, 2048, 4096, 8192, 16384):
    x, y = X[:B], Y[:B]
    for p in ref.parameters(): p.grad = None
    loss_ref = O.mse_training_step(ref, x, y); loss_ref.backward()
    loss = tr.train_step(x.cuda(), y.cuda())
    errs = [rel_err(gv, q.grad) for gv, q in zip(tr.g_views, ref.parameters())]
    # fp64 oracle for reference
    ref64 = O.HighOrderMLP(normalization=O.MaxAbsNormalization, **__import__("helpers").MLP_KW["C2"]).double()
    ref64.load_state_dict({k: v.double() for k, v in ref.state_dict().items()})
    O.mse_training_step(ref64, x.double(), y.double()).backward()
    e64 = [rel_err(gv, q.grad) for gv, q in zip(tr.g_views, ref64.parameters())]
    o64 = [rel_err(q32.grad, q.grad) for q32, q in zip(ref.parameters(), ref64.parameters())]
    print(B, "loss", abs(loss.item() - loss_ref.item()) / loss_ref.item(), "vs fp32 oracle", ["%.1e" % e for e in errs],
          "vs fp64 oracle", ["%.1e" % e for e in e64], "fp32 oracle vs fp64", ["%.1e" % e for e in o64])
