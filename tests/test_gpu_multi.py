"""2-GPU NCCL parity: two ranks training on disjoint halves of a global batch (one all-reduce of the flat
[gradient | loss] vector) end with the same weights and loss as one GPU training on the whole batch.
Skipped on single-GPU boxes (the CPU gloo test covers the host logic there)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q, no_peer):
    if no_peer:
        os.environ["HOIR_NO_PEER"] = "1"
    import torch.distributed as dist
    import hoir_b200 as H
    from helpers import MLP_KW
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    torch.manual_seed(0)
    net = H.HighOrderMLP(normalization=H.MaxAbsNormalization, **MLP_KW["C2"])
    H.initialize_network_polynomial_layers(net, 1.0, 0.0, generator=torch.Generator().manual_seed(0))
    net = net.cuda()
    tr = H.FusedTrainer(net, optimizer="adam", lr=1e-3)
    g = torch.Generator().manual_seed(1)
    img = torch.randint(0, 256, (64, 128, 3), generator=g, dtype=torch.uint8).cuda()
    per = 64 * 128 // world
    losses = []
    for step in range(3):
        loss = tr.train_step_mesh(img, 2, rank * per, per)
        losses.append(loss.item())
    q.put((rank, losses, tr.flat_w.cpu(), tr.peer is not None, tr.launches_per_step))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("no_peer", [False, True], ids=["nvlink-peer-exchange", "nccl-allreduce"])
def test_two_gpu_data_parallel_matches_single_gpu(no_peer):
    import hoir_b200 as H
    from helpers import MLP_KW
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, no_peer)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert torch.equal(res[0][2], res[1][2])          # replicated weights stay bit-identical
    if no_peer:
        assert not res[0][3] and res[0][4] == 2       # kernel + tail, NCCL all-reduce in between
    else:                                             # gradient exchange inside the kernel: one launch per step
        assert res[0][3] and res[1][3] and res[0][4] == 1, "peer-memory exchange was not active"
    assert res[0][1] == res[1][1]
    torch.manual_seed(0)
    net = H.HighOrderMLP(normalization=H.MaxAbsNormalization, **MLP_KW["C2"])
    H.initialize_network_polynomial_layers(net, 1.0, 0.0, generator=torch.Generator().manual_seed(0))
    net = net.cuda()
    tr = H.FusedTrainer(net, optimizer="adam", lr=1e-3)
    g = torch.Generator().manual_seed(1)
    img = torch.randint(0, 256, (64, 128, 3), generator=g, dtype=torch.uint8).cuda()
    losses = [tr.train_step_mesh(img, 2, 0, 64 * 128).item() for _ in range(3)]
    for a, b in zip(losses, res[0][1]):
        assert abs(a - b) <= 1e-5 * abs(a)
    assert (tr.flat_w.cpu() - res[0][2]).abs().max().item() <= 5e-5
