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


def _worker(rank, world, port, q, no_peer, name="C2"):
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
    net = H.HighOrderMLP(normalization=H.MaxAbsNormalization, **MLP_KW[name])
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


@pytest.mark.parametrize("world,name", [(2, "C2"), (4, "C2"), (8, "C2"), (2, "C3f")])
@pytest.mark.parametrize("no_peer", [False, True], ids=["nvlink-peer-exchange", "nccl-allreduce"])
def test_data_parallel_matches_single_gpu(no_peer, world, name):
    """2 / 4 / 8 ranks (skipped below that GPU count): the rank-ordered sum over N peers and the two-buffer reuse
    protocol of the in-kernel exchange (opt_tail.cuh) over three consecutive steps."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import hoir_b200 as H
    from helpers import MLP_KW
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, no_peer, name)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(1, world):
        assert torch.equal(res[0][2], res[r][2])      # replicated weights stay bit-identical
        assert res[0][1] == res[r][1]
    if no_peer:
        assert not res[0][3] and res[0][4] == 2       # kernel + tail, NCCL all-reduce in between
    else:                                             # gradient exchange inside the kernel: one launch per step
        assert all(r[3] for r in res) and res[0][4] == 1, "peer-memory exchange was not active"
    torch.manual_seed(0)
    net = H.HighOrderMLP(normalization=H.MaxAbsNormalization, **MLP_KW[name])
    H.initialize_network_polynomial_layers(net, 1.0, 0.0, generator=torch.Generator().manual_seed(0))
    net = net.cuda()
    tr = H.FusedTrainer(net, optimizer="adam", lr=1e-3)
    g = torch.Generator().manual_seed(1)
    img = torch.randint(0, 256, (64, 128, 3), generator=g, dtype=torch.uint8).cuda()
    losses = [tr.train_step_mesh(img, 2, 0, 64 * 128).item() for _ in range(3)]
    tol = 1e-4 if name == "C3f" else 1e-5          # (the fused Fourier kernel: tf32 x 3 tensor-core passes)
    for a, b in zip(losses, res[0][1]):
        assert abs(a - b) <= tol * abs(a)
    # Adam turns round-off-sized gradient differences into +-lr steps: compare the bulk, bound the worst
    diff = (tr.flat_w.cpu() - res[0][2]).abs()
    if name == "C3f":
        assert diff.max().item() <= 3 * 2e-3 and (diff > 5e-5).float().mean().item() <= 2e-2
    else:
        assert diff.max().item() <= 5e-5


def _trainer_worker(rank, world, port, q, no_peer, sampling):
    if no_peer:
        os.environ["HOIR_NO_PEER"] = "1"
    import torch.distributed as dist
    import hoir_b200 as H
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    cfg = H.images_config()
    H.apply_overrides(cfg, ["mlp.layer_type=continuous", "mlp.n=3", "mlp.n_in=3", "mlp.n_out=3", "mlp.input.width=4",
                            "mlp.input.segments=100", "mlp.hidden.width=40", "mlp.hidden.layers=1",
                            "mlp.hidden.segments=2", "mlp.output.segments=2", "rotations=2", "batch_size=256",
                            "optimizer.name=adam", "optimizer.lr=1e-3"])
    g = torch.Generator().manual_seed(1)
    img = torch.randint(0, 256, (28, 64, 3), generator=g, dtype=torch.uint8)       # 1792 px = 7 blocks of 256
    torch.manual_seed(0)
    tr = H.ImageTrainer(cfg, img, device=f"cuda:{rank}", sampling=sampling, seed=3, drop_last=False)
    losses = [tr.train_epoch() for _ in range(2)]
    torch.cuda.synchronize()
    q.put((rank, losses, tr.fused.flat_w.cpu(), tr.fused.step_count, tr.fused.peer is not None))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("sampling", ["tiles", "shuffle"])
@pytest.mark.parametrize("no_peer", [False, True], ids=["nvlink-peer-exchange", "nccl-allreduce"])
def test_image_trainer_two_ranks_ragged_last_group(no_peer, sampling):
    """``n_blocks % world != 0`` (7 blocks, 2 ranks) -- the surplus rank of the last group steps
    with an empty share instead of skipping the collective step (which trapped / hung the others)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_trainer_worker, args=(r, 2, port, q, no_peer, sampling)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res[0][3] == res[1][3] == 2 * 4              # 4 collective steps per epoch on BOTH ranks
    assert torch.equal(res[0][2], res[1][2])            # replicas stay bit-identical
    assert res[0][1] == res[1][1] and all(l == l and l > 0 for l in res[0][1])
    assert res[0][4] == (not no_peer)
