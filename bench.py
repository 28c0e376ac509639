#!/usr/bin/env python
"""bench.py -- the headline measurement of BASELINE.json: pixel-samples/s of one full training
step (forward + backward + optimizer) of the 40-wide continuous implicit-image MLP
(n=3, 4 -> 40 -> 40 -> 3, 100/2/2 segments) on a synthetic 4096x4096 image, 2^20 pixels per rank
per step (weak scaling over N ranks, one rank per GPU, data-parallel over disjoint pixel tiles,
one NCCL all-reduce of the flat gradient per step).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
    python bench.py --impl reference --steps K --warmup W     # reference CPU formulation (oracle)

Prints ONE JSON line (rank 0).  `value` is device-timed with inputs resident in HBM (coordinates
and targets are produced in-kernel from the uint8 image); `e2e` is the same step through the
reference-facing batch API with HOST (pinned) x/y buffers, H2D and the loss D2H inside the timed
region.  `roofline` is quoted against the FP32-FFMA peak as BASELINE.json asks:
algorithmic 32 880 FLOP/sample x samples per launch / the fused kernel's own launch duration.  The
hidden layer's contractions run on the tensor pipe (tcgen05, tf32 x 3 passes), so the fraction of
the FP32 peak can exceed what a CUDA-core-only kernel could reach; profiles/ holds the ncu pipe
utilisations (FMA, ALU, tensor) that explain it.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "pixel-samples/sec fwd+bwd+optimizer, 40-wide continuous implicit MLP (n=3, 100 input segments)"
UNIT = "samples/s"
ROWS = COLS = 4096
ROTATIONS = 2
BATCH = 1 << 20
FLOP_PER_SAMPLE = 32880          # fwd 11 280 + grad_w 11 280 + grad_x 10 320
MLP_KW = dict(layer_type="continuous", n=3, in_width=4, out_width=3, hidden_layers=1, hidden_width=40,
              in_segments=100, out_segments=2, hidden_segments=2)
CPU_BATCH = 8192


def synthetic_image(device) -> torch.Tensor:
    """smooth + noise so that training is non-degenerate; seed 0."""
    g = torch.Generator(device="cpu").manual_seed(0)
    noise = torch.randint(0, 32, (ROWS, COLS, 3), generator=g, dtype=torch.int16)
    u = (torch.arange(COLS, dtype=torch.float32) / COLS).view(1, COLS, 1)
    v = (torch.arange(ROWS, dtype=torch.float32) / ROWS).view(ROWS, 1, 1)
    base = 127.5 * (1.0 + torch.sin(2 * torch.pi * 3 * u) * torch.cos(2 * torch.pi * 5 * v))
    img = (base.to(torch.int16) + noise).clamp_(0, 255).to(torch.uint8)
    return img.to(device)


# --------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index: int):
        self.samples = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "50", "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self):
        if self.proc is not None:
            time.sleep(0.12)
            self.proc.terminate()
            self.proc = None

    def stats(self, windows):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, smax, reasons = [], [], set()
        for t, line in self.samples:
            if not any(a - 0.02 <= t <= b + 0.05 for a, b in windows):
                continue
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(self.NAMES, parts[3:7]):
                if val == "Active":
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(smax) if smax else None, "samples": len(sm),
                "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------
def cpu_reference_run(steps: int, warmup: int, batch: int = CPU_BATCH):
    """The reference's CPU PyTorch path of the same step (gather + einsum per layer, MaxAbs, MSE(mean), autograd
    backward, Adam) on all host cores: the REAL ``high_order_layers_torch`` when it is importable on this machine
    (kind "reference"; looked up with the repository's alias excluded, oracle/real_dep.py), otherwise the restated
    oracle (kind "port"; the package is not installable here -- DESIGN.md section 2)."""
    from oracle import hol_oracle as O
    from oracle import real_dep
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    real = real_dep.find()
    if real is not None:
        from high_order_layers_torch import layers as RL, networks as RN
        ref = RN.HighOrderMLP(normalization=RL.MaxAbsNormalization, **MLP_KW)
        RN.initialize_network_polynomial_layers(ref, max_slope=1.0, max_offset=0.0)
        kind, what = "reference", f"high_order_layers_torch {getattr(real, '__version__', '?')}"
    else:
        ref = O.HighOrderMLP(normalization=O.MaxAbsNormalization, **MLP_KW)
        O.initialize_network_polynomial_layers(ref, max_slope=1.0, max_offset=0.0)
        kind, what = "port", "restated oracle (reference package not installable)"
    opt = torch.optim.Adam(ref.parameters(), lr=1e-4)
    g = torch.Generator().manual_seed(1)
    x = torch.rand(batch, 4, generator=g) * 2 - 1
    y = torch.rand(batch, 3, generator=g) * 2 - 1
    loss_fn = torch.nn.MSELoss()
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        loss = loss_fn(ref(x).flatten(), y.flatten())      # Net.eval_step, networks.py:64-70
        loss.backward()
        opt.step()
        if s >= warmup:
            times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return {"value": batch / med, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{steps} steps (median) of {batch} random samples through the C2 model, "
                      f"fwd+bwd+Adam, {what}",
            "ms_per_step": med * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    res = cpu_reference_run(args.steps, max(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args.gpus, cpu=True),
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(n_gpus: int, cpu: bool = False):
    cfg = {"workload": "C2: implicit_images.py README config, continuous n=3, 4->40->40->3, segments 100/2/2, "
                       "synthetic 4096x4096 image, rotations=2",
           "batch_per_rank": CPU_BATCH if cpu else BATCH, "global_batch": (CPU_BATCH if cpu else BATCH * n_gpus),
           "optimizer": "adam", "parallelism": f"dp{n_gpus} over pixel tiles",
           "params": 40760}
    if not cpu:
        cfg["l2"] = "flushed between timed steps (256 MiB memset outside the per-step event pairs)"
    return cfg


# --------------------------------------------------------------------------------------------
# secondary workloads: the other BASELINE.json configs, measured in the same run (same clock sampling)
# --------------------------------------------------------------------------------------------
C2_FLOP, C3F_FLOP, C4_FLOP, C1_FLOP, C5_FLOP = 32880, 319600, 162600, 4380, 11280


def _mlp(H, dev, **kw):
    net = H.HighOrderMLP(normalization=H.MaxAbsNormalization, **kw)
    H.initialize_network_polynomial_layers(net, 1.0, 0.0, generator=torch.Generator().manual_seed(0))
    return net.to(dev)


def _timeit(fn, n, warm=3, flush=None):
    """mean ms per call over n calls, CUDA events per call (an L2 flush between calls stays outside the pairs)"""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        if flush is not None:
            flush.zero_()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev) / n


def run_secondary(H, dev, trainer, img, flush, peak_tf, rank, world, dist, which):
    """Every rank runs the single-GPU workloads (rank 0's numbers are reported); c5_render is sharded over ALL ranks."""
    out, windows = {}, {}

    def entry(name, samples, ms, flop, unit, **extra):
        sps = samples / (ms * 1e-3)
        n = extra.get("n_gpus", 1)
        e = {"value": sps, "unit": unit, "ms_per_step": ms, "flop_per_unit": flop,
             "achieved_tflops_per_gpu": sps * flop / 1e12 / n, "frac_of_ffma_peak": sps * flop / 1e12 / n / peak_tf}
        e.update(extra)
        out[name] = e

    def timed(name, fn):
        t0 = time.time()
        fn()
        torch.cuda.synchronize()
        windows[name] = (t0, time.time())

    def c2_incoherent():
        # the reference's sampling (DataLoader(shuffle=True) over image_to_dataset rows, Quirk Q9): one batch of 2^20
        # shuffled pixels of the synthetic image as explicit x[B,4] / y[B,3] tensors resident in HBM
        g = torch.Generator(device=dev).manual_seed(5)
        idx = torch.randint(0, ROWS * COLS, (BATCH,), device=dev, generator=g)
        mesh = H._lib.make_mesh_desc(ROWS, COLS, ROTATIONS, True, 0)
        pos = torch.stack(H.positions_from_mesh(ROWS, COLS, device=str(dev), rotations=ROTATIONS, normalize=True), dim=2)
        x = pos.reshape(-1, 2 * ROTATIONS)[idx].contiguous()
        del pos, mesh
        y = (img.reshape(-1, 3)[idx].float() * 2 / 255 - 1).contiguous()
        ms = _timeit(lambda: trainer.train_step(x, y), 10, flush=flush)
        entry("c2_incoherent", BATCH * world, ms, C2_FLOP, UNIT, n_gpus=world,
              workload="C2 model, explicit batch of 2^20 SHUFFLED pixels of the 4096^2 image per rank (x[B,4], y[B,3] in HBM)",
              launches_per_step=getattr(trainer, "last_launches", None))

    def c3_discontinuous():
        net = _mlp(H, dev, **dict(MLP_KW, layer_type="discontinuous"))
        tr = H.FusedTrainer(net, optimizer="adam", lr=1e-4, data_parallel=False)
        ms = _timeit(lambda: tr.train_step_mesh(img, ROTATIONS, 0, BATCH), 10, flush=flush)
        entry("c3_discontinuous", BATCH, ms, C2_FLOP, UNIT, workload="C3: discontinuous n=3, 4->40->40->3, 2^20 pixels/step (mesh)")

    def c3_fourier():
        kw = dict(layer_type="fourier", n=3, n_in=31, n_hidden=31, n_out=3, in_width=4, out_width=3,
                  hidden_layers=1, hidden_width=40)
        net = _mlp(H, dev, **kw)
        tr = H.FusedTrainer(net, optimizer="adam", lr=1e-4, data_parallel=False)
        ms = _timeit(lambda: tr.train_step_mesh(img, ROTATIONS, 0, BATCH), 5, flush=flush)
        import ctypes as C
        pms, pfl = C.c_float(), C.c_double()
        H._lib.check(H._lib.lib().hoir_tf32_peak(4000, C.byref(pms), C.byref(pfl), None))
        tf32_peak = pfl.value / (pms.value * 1e-3) / 1e12
        sps = BATCH / (ms * 1e-3)
        entry("c3_fourier", BATCH, ms, C3F_FLOP, UNIT,
              workload="C3: fourier n_in=31 (hidden 31, out 3), 4->40->40->3, 2^20 pixels/step (mesh)",
              path="fused whole-network kernel (fused_fourier.cu), one launch per step", launches_per_step=tr.last_launches,
              tensor={"bound": "tensor", "unit": "TFLOP/s", "peak": tf32_peak,
                      "peak_source": "hoir_tf32_peak micro-kernel (tcgen05 kind::tf32, 128x256x8, smem operands) measured in this run",
                      "achieved_tf32_equivalent": sps * C3F_FLOP * 3 / 1e12, "frac": sps * C3F_FLOP * 3 / 1e12 / tf32_peak,
                      "note": "every product runs as 3 tf32 passes (A*B + A_lo*B + A*B_lo) for fp32-level accuracy"})
        B16 = 1 << 16
        g = torch.Generator(device=dev).manual_seed(0)
        x = torch.rand(B16, 4, device=dev, generator=g) * 2 - 1
        y = torch.rand(B16, 3, device=dev, generator=g) * 2 - 1
        ms16 = _timeit(lambda: tr.train_step(x, y), 20)
        out["c3_fourier"]["batch_2^16"] = {"value": B16 / (ms16 * 1e-3), "unit": UNIT, "ms_per_step": ms16}

    def c4_train():
        kw = dict(layer_type="continuous", n=2, in_width=216, out_width=27, hidden_layers=2, hidden_width=50,
                  in_segments=50, out_segments=2, hidden_segments=2)
        g = torch.Generator().manual_seed(0)
        im = (torch.randint(0, 256, (3, 2048, 2048), generator=g).float() / 127.5 - 1).to(dev)
        B = 1 << 16
        net = _mlp(H, dev, **kw)
        lib, L = H._lib.lib(), H._lib
        gs = H.GraphedStep(net, H.Adam(net.parameters(), lr=1e-4, capturable=True),
                           torch.zeros(B, 216, device=dev), torch.zeros(B, 27, device=dev))

        def step():
            L.check(lib.hoir_neighborhood_gather(L.dptr(im), 3, 2048, 2048, 3, 3, 1, 0, 12345, B, L.dptr(gs.x),
                                                 L.dptr(gs.y), L.stream_ptr(dev)))
            gs.replay()
        ms = _timeit(step, 10)
        entry("c4_train", B, ms, C4_FLOP, "patches/s",
              workload="C4: interpolator continuous n=2, 216->50->50->50->27, 2048^2 image, 2^16 patches/step (gather + step)")
        ms = _timeit(lambda: H.neighborhood_sample_generator(net, im, 3, 3), 10, warm=2)
        net.use_fused = False
        ms_layers = _timeit(lambda: H.neighborhood_sample_generator(net, im, 3, 3), 3, warm=1)
        net.use_fused = True
        patches = ((2048 + 9) // 3 + 1) ** 2
        out["c4_frame"] = {"value": ms, "unit": "ms/frame", "patches_per_s": patches / (ms * 1e-3), "launches": 1,
                           "per_layer_path_ms": ms_layers,
                           "workload": "one generate_sequence frame on 2048^2 in ONE launch (hoir_interp_frame: reflect-pad + "
                                       "stride-3 gather + 216->50->50->50->27 forward + fold + crop + blend)"}

        # the reference's own operating point for this example: batch_size=256 (/root/reference/README.md:57)
        xs, ys = gs.x[:256].clone(), gs.y[:256].clone()
        net_s = _mlp(H, dev, **kw)
        gss = H.GraphedStep(net_s, H.Adam(net_s.parameters(), lr=1e-4, capturable=True), xs, ys)
        ms = _timeit(gss.replay, 100)
        out["c4_b256"] = {"value": 256 / (ms * 1e-3), "unit": "patches/s", "us_per_step": ms * 1e3,
                          "workload": "C4 network, batch 256 (the reference's batch size), whole step replayed from one CUDA graph"}

    def c1_b256():
        kw = dict(layer_type="continuous", n=3, in_width=2, out_width=3, hidden_layers=2, hidden_width=10,
                  in_segments=100, out_segments=2, hidden_segments=2)
        net = _mlp(H, dev, **kw)
        tr = H.FusedTrainer(net, optimizer="sparse_lion", lr=1e-4, data_parallel=False)
        g = torch.Generator(device=dev).manual_seed(0)
        x = torch.rand(256, 2, device=dev, generator=g) * 2 - 1
        y = torch.rand(256, 3, device=dev, generator=g) * 2 - 1
        ms = _timeit(lambda: tr.train_step(x, y), 200)
        entry("c1_b256", 256, ms, C1_FLOP, UNIT, workload="C1: continuous n=3, 2->10->10->10->3, batch 256 (the reference's batch size)",
              us_per_step=ms * 1e3)

    def c5_render():
        # C5: the C2 model rendered at 8192^2, rotations=2, the image sharded over ALL ranks by contiguous pixel
        # ranges (parallel.render_shard), no collective.  value = whole image / slowest rank's device time.
        rows = cols = 8192
        first, count = H.parallel.render_shard(rows * cols, rank, world)
        buf = torch.empty((count, 3), device=dev)
        reps = 3
        trainer.render(rows, cols, ROTATIONS, first_pixel=first, count=count, out=buf)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            trainer.render(rows, cols, ROTATIONS, first_pixel=first, count=count, out=buf)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = ms.item()
        entry("c5_render", rows * cols, ms, C5_FLOP, "pixels/s", n_gpus=world, scaling="strong",
              workload=f"C5: render 8192x8192 with the C2 model, rotations=2, {world} rank(s) x {count} pixels, "
                       "0.5*(y+1) in-kernel, 12 B/pixel written",
              hbm_write_gbs=rows * cols * 12 / (ms * 1e-3) / 1e9)

    table = {"c2_incoherent": c2_incoherent, "c3_discontinuous": c3_discontinuous, "c3_fourier": c3_fourier,
             "c4_train": c4_train, "c1_b256": c1_b256, "c5_render": c5_render}
    for name in which:
        try:
            timed(name, table[name])
        except Exception as e:  # noqa: BLE001 -- a secondary workload must never take the headline line down
            out[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
            torch.cuda.synchronize()
    return out, windows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="hoir_b200", choices=["hoir_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--secondary", default="all",
                    help="comma list of c2_incoherent,c3_discontinuous,c3_fourier,c4_train,c1_b256,c5_render | all | none")
    ap.add_argument("--strong", action="store_true",
                    help="strong scaling: the global batch stays 2^20 pixels, every rank takes 2^20 / N (default: weak)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (hoir_b200 has no CPU path); "
                         "use --impl reference for the CPU formulation")
    # the CPU baseline runs on rank 0 at N = 1 only, BEFORE any process group exists (at N > 1 the other ranks would
    # spin in a host barrier on the same cores; the reference arm `--impl reference` covers every N)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference_run(15, 3)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        torch.set_num_threads(max(1, (os.cpu_count() or 2) // 2))

    import torch.distributed as dist
    import hoir_b200 as H
    global BATCH
    if args.strong:
        BATCH = (1 << 20) // world
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    torch.manual_seed(0)
    net = H.HighOrderMLP(normalization=H.MaxAbsNormalization, **MLP_KW)
    H.initialize_network_polynomial_layers(net, max_slope=1.0, max_offset=0.0,
                                           generator=torch.Generator().manual_seed(0))
    net = net.to(dev)
    trainer = H.FusedTrainer(net, optimizer="adam", lr=1e-4)
    exchange = "none" if world == 1 else ("peer" if trainer.peer is not None else "nccl")
    img = synthetic_image(dev)
    n_blocks = ROWS * COLS // BATCH
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(s):
        block = H.parallel.training_block(s, rank, world, n_blocks)
        return trainer.train_step_mesh(img, ROTATIONS, block * BATCH, BATCH)

    for s in range(args.warmup):
        step(s)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None

    # ---- value: device-timed, inputs resident in HBM ------------------------------------------
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    trainer.kernel_events = []
    barrier()
    t_val0 = time.time()
    for s in range(args.steps):
        flush.zero_()
        ev[s][0].record()
        loss = step(args.warmup + s)
        ev[s][1].record()
    barrier()
    t_val1 = time.time()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    kern_ms = [a.elapsed_time(b) for a, b in trainer.kernel_events]
    trainer.kernel_events = None
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = total_ms.item()
    value = BATCH * world * args.steps / (total_ms * 1e-3)
    final_loss = float(loss.item())

    # ---- e2e: host (pinned) batches through the batch API --------------------------------------
    e2e = None
    windows = {"value": (t_val0, t_val1)}
    if not args.no_e2e:
        e2e, windows["e2e"] = run_e2e(H, dist, trainer, img, dev, rank, world, args, barrier)

    # ---- roofline denominators (rank 0 measures; every rank needs the number for the secondary fractions) ----------
    import ctypes as C
    peaks = {}
    for form, key in ((1, "ffma_rrr"), (0, "ffma_const")):
        ms, fl = C.c_float(), C.c_double()
        H._lib.check(H._lib.lib().hoir_ffma_peak(20000, form, C.byref(ms), C.byref(fl), None))
        peaks[key] = fl.value / (ms.value * 1e-3) / 1e12
    peak = max(peaks.values())

    which = [] if args.secondary == "none" or args.strong else (
        ["c2_incoherent", "c3_discontinuous", "c3_fourier", "c4_train", "c1_b256", "c5_render"]
        if args.secondary == "all" else [w for w in args.secondary.split(",") if w])
    barrier()
    secondary, sec_windows = run_secondary(H, dev, trainer, img, flush, peak, rank, world, dist, which)
    barrier()

    if rank == 0:
        sampler.stop()
        clocks = sampler.stats([windows["value"]] + ([windows["e2e"]] if "e2e" in windows else []))
        for name, win in sec_windows.items():
            if name in secondary and "error" not in secondary[name]:
                c = sampler.stats([win])
                secondary[name]["clocks"] = {"sm_mhz": c["sm_mhz"], "reasons": c["reasons"]}
        nominal = 148 * 128 * 2 * 1.965e9 / 1e12
        measured = {}
        try:
            measured = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        traffic = None
        try:   # dram bytes per launch of the same kernel, from the committed ncu capture (not measured in this run)
            t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["c2_mesh"]
            traffic = t["bytes"] * BATCH / t["samples_per_launch"] if not os.environ.get("HOIR_NO_TC") else None
            traffic_src = t["source"]
        except Exception:
            traffic_src = None
        k_ms = statistics.mean(kern_ms) if kern_ms else float("nan")
        achieved = BATCH * FLOP_PER_SAMPLE / (k_ms * 1e-3) / 1e12
        hbm_peak = measured.get("hbm_gbs", 6650.0)
        roofline = {"bound": "fp32_ffma", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": achieved / peak,
                    "traffic": traffic, "traffic_source": traffic_src,
                    "peak_source": "hoir_ffma_peak micro-kernel measured in this run (MEASURED_PEAKS.json has no "
                                   "FP32 entry); best of register-form / constant-form FFMA",
                    "peak_ffma_rrr": peaks["ffma_rrr"], "peak_ffma_const": peaks["ffma_const"],
                    "peak_nominal": nominal, "frac_of_nominal": achieved / nominal,
                    "kernel": ("fused_mlp_kernel<Shape<continuous,3,4,40,3,1,2,2>, TRAIN> (FFMA only)" if os.environ.get("HOIR_NO_TC")
                               else "tc::fused_mlp_tc_kernel<Shape<continuous,3,4,40,3,1,2,2>> (hidden layer on tcgen05 tf32x3, rest FFMA)"),
                    "kernel_ms": k_ms, "kernel_share_of_step": k_ms * len(kern_ms) / sum(step_ms) if kern_ms else None,
                    "flop_per_sample": FLOP_PER_SAMPLE, "samples_per_launch": BATCH,
                    "hbm": {"algorithmic_bytes_per_sample": 3, "achieved_gbs": BATCH * 3 / (k_ms * 1e-3) / 1e9,
                            "peak_gbs": hbm_peak,
                            "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in measured else "fallback"}}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
                "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(world), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
                "gpu_launches": trainer.launches_per_step * args.steps, "exchange": exchange, "clocks": clocks,
                "final_loss": final_loss, "secondary": secondary}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_e2e(H, dist, trainer, img, dev, rank, world, args, barrier):
    """The same step through the reference-facing batch API with HOST buffers: pinned x[B,4] / y[B,3] -> H2D -> fused
    step -> loss D2H, every step (HostBatchStepper: the copy of batch k+1 overlaps the compute of batch k).  The
    batches are what the reference's DataLoader(shuffle=True) delivers: shuffled rows of image_to_dataset
    (/root/reference/high_order_implicit_representation/single_image_dataset.py:22-51, 312-320) of the synthetic
    image -- pixel coordinates with rotations=2 and u*2/255-1 targets."""
    stepper = H.HostBatchStepper(trainer, BATCH)
    g = torch.Generator(device=dev).manual_seed(2 + rank)
    pos = torch.stack(H.positions_from_mesh(ROWS, COLS, device=str(dev), rotations=ROTATIONS, normalize=True),
                      dim=2).reshape(-1, 2 * ROTATIONS)
    hx, hy = [], []
    for _ in range(4):
        idx = torch.randint(0, ROWS * COLS, (BATCH,), device=dev, generator=g)
        hx.append(pos[idx].cpu().pin_memory())
        hy.append((img.reshape(-1, 3)[idx].float() * 2 / 255 - 1).cpu().pin_memory())
    del pos
    e_steps = max(3, min(args.steps, 50))
    for s in range(3):
        stepper.step(hx[s % 4], hy[s % 4])
    stepper.flush()
    barrier()
    t0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(e_steps + 1)]
    e0.record()
    marks[0].record()
    for s in range(e_steps):
        stepper.step(hx[s % 4], hy[s % 4])      # H2D of this batch overlaps the previous step's compute
        marks[s + 1].record()                   # (compute stream: the end of this step's kernel)
    e2e_loss = stepper.flush()                  # every step's loss has been read back by now
    e1.record()
    barrier()
    t1 = time.time()
    wall_e2e_ms = (t1 - t0) * 1e3
    wall_ms = torch.tensor([max(e0.elapsed_time(e1), wall_e2e_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(wall_ms, op=dist.ReduceOp.MAX)
    e2e = {"value": BATCH * world * e_steps / (wall_ms.item() * 1e-3), "unit": UNIT,
           "h2d_bytes_per_step": stepper.h2d_bytes, "d2h_bytes_per_step": stepper.d2h_bytes,
           "ms_per_step_median": statistics.median(a.elapsed_time(b) for a, b in zip(marks[:-1], marks[1:])),
           "steps": e_steps, "api": "HostBatchStepper.step(x_host[B,4], y_host[B,3]) -> loss of the previous step (2-deep H2D/compute pipeline)",
           "batches": "shuffled pixels of the synthetic image (image_to_dataset rows, rotations=2), pinned host memory",
           "last_loss": e2e_loss}
    return e2e, (t0, t1)


if __name__ == "__main__":
    main()
