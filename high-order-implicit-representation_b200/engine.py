"""Fused training / rendering engine over the whole-network sm_100a kernels.

This is the fast path behind ``Net``: all ``w`` tensors of a ``HighOrderMLP`` become views into
ONE flat fp32 parameter vector (boundary layout ``[out][in][nodes]`` per layer, so
``state_dict`` is unchanged), gradients and the scalar loss live in one flat vector that is
zeroed with a single memset, reduced across ranks with a single NCCL all-reduce (data-parallel
over disjoint pixel tiles) and consumed by a single fused Adam / SparseLion
launch.  One training step of ``Net.eval_step`` + ``backward`` + ``optimizer.step``
(/root/reference/high_order_implicit_representation/networks.py:64-70, 93-102) is
ONE launch on a single GPU (``hoir_mlp_train_step_fused``: the optimizer update, the weight re-pack
and the gradient zeroing run inside the whole-network kernel behind a grid barrier) and
``kernel, all_reduce, tail`` across GPUs -- no PyTorch autograd, no per-layer tensors in HBM.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
from torch import Tensor

from . import _lib, parallel
from .networks import HighOrderMLP


class FusedTrainer:
    """``Net.eval_step`` + ``backward`` + ``optimizer.step`` of a ``HighOrderMLP`` on the fused kernels.
    ``keep_grads=True`` leaves ``flat_g`` / ``g_views`` holding the step's gradient (parity tests,
    inspection) at the price of one extra memset per step; by default the fused tail zeroes it."""

    def __init__(self, mlp: HighOrderMLP, optimizer: str = "adam", lr: float = 1e-4,
                 betas: Optional[tuple] = None, eps: float = 1e-8, weight_decay: float = 0.0,
                 process_group=None, keep_grads: bool = False, data_parallel: bool = True):
        self.lib = _lib.lib()
        self.mlp = mlp
        self.layers = mlp.high_order_layers()
        self.desc = mlp.mlp_desc()
        if self.desc is None or not self.lib.hoir_mlp_fused_supported(self.desc):
            raise _lib.HoirUnsupported("no fused instantiation for this network; train it through "
                                       "the per-layer autograd path (Net.training_step)")
        if optimizer not in ("adam", "sparse_lion"):
            raise ValueError(f"Optimizer {optimizer} not recognized")
        dev = self.layers[0].w.device
        if dev.type != "cuda":
            raise RuntimeError("hoir_b200 has no CPU path: move the network to a CUDA device first")
        self.device = dev
        self.optimizer = optimizer
        self.lr = float(lr)
        self.betas = tuple(betas) if betas is not None else ((0.9, 0.999) if optimizer == "adam" else (0.9, 0.99))
        self.eps = float(eps)
        self.weight_decay = float(weight_decay)
        self.step_count = 0
        self.keep_grads = bool(keep_grads)
        self.group = process_group
        # data_parallel=False: a single-GPU trainer inside a multi-rank job (no gradient exchange)
        self.rank, self.world = parallel.world_info(process_group) if data_parallel else (0, 1)
        self.out_features = self.layers[-1].out_features
        self.in_features = self.layers[0].in_features
        sizes = [l.w.numel() for l in self.layers]
        self.n_params = sum(sizes)
        self.flat_w = torch.empty(self.n_params, device=dev, dtype=torch.float32)
        # multi-GPU: gradients live in a symmetric block mapped by every peer and are exchanged inside the
        # optimizer tail over NVLink (two alternating buffers); otherwise one flat vector (+ NCCL if world > 1)
        self.peer = None
        if self.world > 1 and not keep_grads:
            self.peer = parallel.try_peer_block(int(self.lib.hoir_peer_block_floats(self.desc)), dev, process_group)
            agree = torch.tensor([1 if self.peer is not None else 0], device=dev)
            torch.distributed.all_reduce(agree, op=torch.distributed.ReduceOp.MIN, group=process_group)
            if int(agree.item()) == 0:
                self.peer = None
        if self.peer is not None:
            stride = int(self.lib.hoir_peer_stride(self.desc))
            self._peer_bufs = [self.peer.tensor[k * stride:k * stride + self.n_params + 1] for k in range(2)]
            self._peer_desc = _lib.PeerDesc()
            self._peer_desc.world, self._peer_desc.rank = self.peer.world, self.peer.rank
            for r, ptr in enumerate(self.peer.ptrs):
                self._peer_desc.block[r] = ptr
            self.flat_g = self._peer_bufs[1]          # step 1 uses parity 1
        else:
            self.flat_g = torch.zeros(self.n_params + 1, device=dev, dtype=torch.float32)  # [-1] = loss accumulator
        self.loss_out = torch.zeros(1, device=dev, dtype=torch.float32)
        self.w_views, self.g_views = [], []
        off = 0
        with torch.no_grad():
            for l, sz in zip(self.layers, sizes):
                view = self.flat_w[off:off + sz].view_as(l.w)
                view.copy_(l.w)
                l.w.data = view                       # state_dict / checkpoints keep working
                self.w_views.append(view)
                self.g_views.append(self.flat_g[off:off + sz].view_as(l.w))
                off += sz
        self.exp_avg = torch.zeros(self.n_params, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(self.n_params, device=dev, dtype=torch.float32) if optimizer == "adam" else None
        self.packed = torch.zeros(self.lib.hoir_mlp_packed_size(self.desc), device=dev, dtype=torch.float32)
        self._w_table = _lib.ptr_table(self.w_views)
        self._g_table = _lib.ptr_table(self.g_views)
        self._loss_ptr = _lib.C.c_void_p(self.flat_g.data_ptr() + 4 * self.n_params)
        self._w_version = None
        self._opt = _lib.OptDesc()
        self._opt.kind = _lib.OPT_KIND[optimizer]
        self._opt.flat_w = self.flat_w.data_ptr()
        self._opt.flat_g = self.flat_g.data_ptr()
        self._opt.exp_avg = self.exp_avg.data_ptr()
        self._opt.exp_avg_sq = self.exp_avg_sq.data_ptr() if self.exp_avg_sq is not None else None
        self._opt.packed = self.packed.data_ptr()
        self._opt.loss_out = self.loss_out.data_ptr()
        #: kernels launched by the last train step
        self.launches_per_step = 1 if (self.world == 1 or self.peer is not None) else 2
        #: explicit batches x[B, in] in arbitrary order (DataLoader(shuffle=True)): "sorted" = visit the samples in
        #: layer-0 cell order (hoir_mlp_batch_sort; coordinates of image pixels), "two_pass" = the workspace path
        #: (independent random inputs), "auto" = decided ONCE from the coherence probe on the first large batch
        self.batch_order = os.environ.get("HOIR_BATCH_ORDER", "auto")
        self.sort_min_batch = 8192
        self._order_buf = None
        self._sort_ok = bool(self.lib.hoir_mlp_order_supported(self.desc))
        #: set to a list to collect (start, end) CUDA events around the fused training kernel
        self.kernel_events = None

    # ------------------------------------------------------------------------------------------
    def _versions(self) -> tuple:
        # ``l.w.data = view`` gives every Parameter its OWN version counter: ``load_state_dict`` / ``p.copy_`` bump
        # the Parameter's, in-place writes through ``flat_w`` bump flat_w's -- watch all of them
        return (self.flat_w._version,) + tuple(l.w._version for l in self.layers)

    def pack(self) -> None:
        """Refresh the kernel-side packed weights from ``flat_w``.  The fused tail keeps them in sync
        during training; this is only needed after the weights were written from outside
        (``load_state_dict``, in-place edits of ``w`` / ``flat_w``), which ``_ensure_packed`` detects by the
        version counters of ``flat_w`` and of every ``w`` Parameter.  Writes through ``w.data`` or raw pointers
        are invisible to the counters: call ``pack()`` after them."""
        _lib.check(self.lib.hoir_mlp_pack(self.desc, self._w_table, _lib.dptr(self.packed),
                                          _lib.stream_ptr(self.device)))
        self._w_version = self._versions()

    def _ensure_packed(self) -> None:
        if self._versions() != self._w_version:
            self.pack()

    def _loss_scale(self, B: int, global_batch: Optional[int]) -> float:
        return parallel.global_loss_scale(B, self.world, self.out_features, global_batch)

    def _opt_desc(self) -> "_lib.OptDesc":
        o = self._opt
        o.lr, o.beta1, o.beta2 = self.lr, self.betas[0], self.betas[1]
        o.eps, o.weight_decay = self.eps, self.weight_decay
        o.step = self.step_count
        o.keep_grad = 1 if self.keep_grads else 0
        return o

    def _order(self, B: int) -> Tensor:
        if self._order_buf is None or self._order_buf.numel() < B:
            self._order_buf = torch.empty(max(B, 1), device=self.device, dtype=torch.int32)
        return self._order_buf

    def _sorted(self, B: int, x2: Optional[Tensor], mesh, pixels: Optional[Tensor]):
        """order[B] (device int32) for this batch, or None when the batch should not be sorted."""
        if not self._sort_ok or B < self.sort_min_batch or self.batch_order == "two_pass":
            return None
        order = self._order(B)
        # below 32768 samples the alternative is the unsorted in-kernel walk (no two-pass mode): always sort
        probe = _lib.C.c_float(0.0) if (self.batch_order == "auto" and B >= 32768) else None
        _lib.check(self.lib.hoir_mlp_batch_sort(
            self.desc, B, _lib.dptr(x2) if x2 is not None else None, mesh,
            _lib.dptr(pixels, torch.int32) if pixels is not None else None, _lib.dptr(order, torch.int32),
            _lib.C.byref(probe) if probe is not None else None, _lib.stream_ptr(self.device)))
        if probe is not None:       # one synchronising probe per trainer: pixel coordinates ~0.03, random inputs ~0.5
            self.batch_coherence = float(probe.value)
            self.batch_order = "sorted" if probe.value < 0.25 else "two_pass"
            if self.batch_order == "two_pass":
                return None
        return order

    def _step(self, B: int, x_ptr, mesh, target_kind: int, target_ptr, global_batch: Optional[int],
              order: Optional[Tensor] = None) -> Tensor:
        """zero_grad + forward + MSE + backward + [all-reduce] + optimizer + weight re-pack.
        Single GPU: ONE launch (the optimizer tail runs inside the whole-network kernel behind a grid
        barrier), or kernel + second pass + tail for large incoherent batches.  Multi-GPU: kernel, one NCCL
        all-reduce of ``flat_g``, tail."""
        lib, st = self.lib, _lib.stream_ptr(self.device)
        self._ensure_packed()
        if self.keep_grads:
            self.flat_g.zero_()
        self.step_count += 1
        _lib.weights_changed()
        scale = self._loss_scale(B, global_batch)
        ev = None
        if self.kernel_events is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        o_ptr = _lib.dptr(order, torch.int32) if order is not None else None
        self.last_launches = (1 if (self.world == 1 or self.peer is not None) else 2) + (1 if order is not None else 0)
        if order is not None and (self.world == 1 or self.peer is not None):
            if self.peer is not None:
                self.flat_g = self._peer_bufs[self.step_count & 1]
            _lib.check(lib.hoir_mlp_train_step_fused_order(
                self.desc, B, x_ptr, mesh, o_ptr, target_kind, target_ptr, scale, _lib.C.byref(self._opt_desc()),
                _lib.C.byref(self._peer_desc) if self.peer is not None else None, st))
        elif order is not None:
            _lib.check(lib.hoir_mlp_train_step_order(self.desc, B, x_ptr, mesh, o_ptr, target_kind, target_ptr, None,
                                                     _lib.dptr(self.packed), self._g_table, self._loss_ptr, scale, st))
        elif self.world == 1:
            _lib.check(lib.hoir_mlp_train_step_fused(self.desc, B, x_ptr, mesh, target_kind, target_ptr,
                                                     scale, _lib.C.byref(self._opt_desc()), st))
        elif self.peer is not None:
            # ONE launch per step on every rank: the gradient exchange over NVLink peer memory and the update
            # run in the kernel's tail (no NCCL call on the step path)
            self.flat_g = self._peer_bufs[self.step_count & 1]
            _lib.check(lib.hoir_mlp_train_step_fused_peer(self.desc, B, x_ptr, mesh, target_kind, target_ptr,
                                                          scale, _lib.C.byref(self._opt_desc()),
                                                          _lib.C.byref(self._peer_desc), st))
        else:
            _lib.check(lib.hoir_mlp_train_step(self.desc, B, x_ptr, mesh, target_kind, target_ptr, None,
                                               _lib.dptr(self.packed), self._g_table, self._loss_ptr,
                                               scale, st))
        if ev is not None:
            ev[1].record()
            self.kernel_events.append(ev)
        if self.world > 1 and self.peer is None:
            parallel.allreduce_sum_(self.flat_g, self.group)
            _lib.check(lib.hoir_mlp_optimizer_tail(self.desc, _lib.C.byref(self._opt_desc()), st))
        return self.loss_out     # 1-element tensor: MSE(mean) of this step over all ranks

    def train_step(self, x: Tensor, y: Tensor, global_batch: Optional[int] = None) -> Tensor:
        """One optimisation step on explicit samples x[B, in], y[B, out] (the reference's batch).
        Returns a 1-element device tensor holding the loss; valid until the next step."""
        x2 = x if x.dim() == 2 else x.flatten(1)          # (an empty share of a global batch is legal: B == 0)
        y2 = y if y.dim() == 2 else y.flatten(1)
        x2, y2 = x2.contiguous(), y2.contiguous()
        if x2.shape[1] != self.in_features or y2.shape[1] != self.out_features:
            raise ValueError(f"expected x[B,{self.in_features}] / y[B,{self.out_features}], got "
                             f"{tuple(x.shape)} / {tuple(y.shape)}")
        with torch.cuda.device(self.device):
            B = x2.shape[0]
            order = self._sorted(B, x2, None, None)
            return self._step(B, _lib.dptr(x2), None, _lib.TARGET_F32, _lib.dptr(y2), global_batch, order)

    def train_step_pixels(self, image_u8: Tensor, rotations: int, pixels: Tensor,
                          global_batch: Optional[int] = None) -> Tensor:
        """One optimisation step on the pixels ``pixels[B]`` (int32 row-major indices, any order, e.g. a slice of a
        random permutation = the reference's DataLoader(shuffle=True) over image_to_dataset rows) of a uint8 image
        resident in HBM.  No dataset is materialised: the indices are sorted by layer-0 cell and the kernel produces
        coordinates (positions_from_mesh) and targets (u*2/255-1) from the pixel index -- 4 bytes per sample instead
        of the 28 of an explicit (x, y) batch."""
        if image_u8.dtype != torch.uint8 or image_u8.dim() != 3 or image_u8.shape[2] != 3:
            raise ValueError("expected a uint8 image [rows, cols, 3]")
        if pixels.dtype != torch.int32 or pixels.dim() != 1:
            raise TypeError("pixels must be a 1-D int32 tensor")
        if not self._sort_ok:
            raise _lib.HoirUnsupported("pixel-index batches need the tensor-core training kernel")
        B = int(pixels.shape[0])
        mesh = _lib.make_mesh_desc(image_u8.shape[0], image_u8.shape[1], rotations, True, 0)
        with torch.cuda.device(self.device):
            order = self._order(B)
            _lib.check(self.lib.hoir_mlp_batch_sort(self.desc, B, None, mesh, _lib.dptr(pixels.contiguous(), torch.int32),
                                                    _lib.dptr(order, torch.int32), None, _lib.stream_ptr(self.device)))
            return self._step(B, None, mesh, _lib.TARGET_U8_IMAGE, _lib.dptr(image_u8, torch.uint8), global_batch, order)

    def train_step_mesh(self, image_u8: Tensor, rotations: int, first_pixel: int, B: int,
                        global_batch: Optional[int] = None) -> Tensor:
        """One optimisation step on pixels first_pixel .. first_pixel+B-1 (row-major) of a uint8
        image [rows, cols, 3] resident in HBM: coordinates (positions_from_mesh) and targets
        (u*2/255-1) are produced in-kernel -- image_to_dataset of the reference,
        /root/reference/high_order_implicit_representation/single_image_dataset.py:22-51."""
        if image_u8.dtype != torch.uint8 or image_u8.dim() != 3 or image_u8.shape[2] != 3:
            raise ValueError("expected a uint8 image [rows, cols, 3]")
        mesh = _lib.make_mesh_desc(image_u8.shape[0], image_u8.shape[1], rotations, True, first_pixel)
        with torch.cuda.device(self.device):
            return self._step(B, None, mesh, _lib.TARGET_U8_IMAGE, _lib.dptr(image_u8, torch.uint8),
                              global_batch)

    # ------------------------------------------------------------------------------------------
    def optimizer_state_dict(self) -> dict:
        """The optimizer state in ``torch.optim.Optimizer.state_dict()`` layout (one entry per ``w`` tensor, in
        layer order), as a Lightning checkpoint stores it under ``optimizer_states[0]``."""
        state, off = {}, 0
        for k, l in enumerate(self.layers):
            n = l.w.numel()
            st = {"step": torch.tensor(float(self.step_count)),
                  "exp_avg": self.exp_avg[off:off + n].view_as(l.w).detach().cpu().clone()}
            if self.exp_avg_sq is not None:
                st["exp_avg_sq"] = self.exp_avg_sq[off:off + n].view_as(l.w).detach().cpu().clone()
            state[k] = st
            off += n
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.weight_decay,
                 "params": list(range(len(self.layers)))}
        return {"state": state, "param_groups": [group]}

    def load_optimizer_state_dict(self, sd: dict) -> None:
        off = 0
        with torch.no_grad():
            for k, l in enumerate(self.layers):
                n = l.w.numel()
                st = sd["state"].get(k, sd["state"].get(str(k)))
                if st is not None:
                    self.exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1))
                    if self.exp_avg_sq is not None and "exp_avg_sq" in st:
                        self.exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
                    self.step_count = int(float(st["step"]))
                off += n
        g = sd["param_groups"][0]
        self.lr = float(g["lr"])

    # ------------------------------------------------------------------------------------------
    def forward(self, x: Tensor) -> Tensor:
        x2 = (x if x.dim() == 2 else x.flatten(1)).contiguous()
        y = torch.empty((x2.shape[0], self.out_features), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            self._ensure_packed()
            _lib.check(self.lib.hoir_mlp_forward(self.desc, x2.shape[0], _lib.dptr(x2), None,
                                                 _lib.dptr(self.packed), _lib.dptr(y), 0,
                                                 _lib.stream_ptr(self.device)))
        return y

    def render(self, rows: int, cols: int, rotations: int, first_pixel: int = 0,
               count: Optional[int] = None, out: Optional[Tensor] = None,
               post_scale: bool = True) -> Tensor:
        """Forward over pixels first_pixel .. first_pixel+count-1 of a rows x cols mesh; with
        post_scale the 0.5*(y+1) of /root/reference/.../rendering.py:213 is applied in-kernel.
        Tile-sharded rendering = each rank calls this with its own pixel range (no collective)."""
        count = rows * cols - first_pixel if count is None else count
        if out is None:
            out = torch.empty((count, self.out_features), device=self.device, dtype=torch.float32)
        mesh = _lib.make_mesh_desc(rows, cols, rotations, True, first_pixel)
        with torch.cuda.device(self.device):
            self._ensure_packed()
            _lib.check(self.lib.hoir_mlp_forward(self.desc, count, None, mesh, _lib.dptr(self.packed),
                                                 _lib.dptr(out), 1 if post_scale else 0,
                                                 _lib.stream_ptr(self.device)))
        return out


class HostBatchStepper:
    """End-to-end training through host memory: pinned host (x, y) -> H2D -> fused step -> loss D2H.
    This is the reference-facing shape of a training step (a DataLoader batch arriving from host
    memory, /root/reference/high_order_implicit_representation/single_image_dataset.py:312-320).

    Two-deep pipeline: ``step(x, y)`` enqueues the H2D copy of this batch on a copy stream and its
    training step on the compute stream, and returns the loss of the PREVIOUS step (None for the
    first call); ``flush()`` returns the last one.  The copy of batch k+1 overlaps the compute of
    batch k, like the reference's pinned, non-blocking DataLoader feed."""

    def __init__(self, trainer: FusedTrainer, batch: int):
        self.trainer = trainer
        dev = trainer.device
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.bufs = [(torch.empty((batch, trainer.in_features), device=dev, dtype=torch.float32),
                      torch.empty((batch, trainer.out_features), device=dev, dtype=torch.float32))
                     for _ in range(2)]
        self.copied = [torch.cuda.Event() for _ in range(2)]
        self.consumed = [torch.cuda.Event() for _ in range(2)]
        self.loss_host = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]
        self.loss_ready = [torch.cuda.Event() for _ in range(2)]
        self.k = 0
        self.pending = None
        self.h2d_bytes = sum(t.numel() * 4 for t in self.bufs[0])
        self.d2h_bytes = 4

    def step(self, x_host: Tensor, y_host: Tensor) -> Optional[float]:
        slot = self.k & 1
        xd, yd = self.bufs[slot]
        compute = torch.cuda.current_stream(self.trainer.device)
        with torch.cuda.stream(self.copy_stream):
            if self.k >= 2:
                self.copy_stream.wait_event(self.consumed[slot])   # the step that last read this slot is done
            xd.copy_(x_host, non_blocking=True)
            yd.copy_(y_host, non_blocking=True)
            self.copied[slot].record(self.copy_stream)
        compute.wait_event(self.copied[slot])
        loss = self.trainer.train_step(xd, yd)
        self.consumed[slot].record(compute)
        self.loss_host[slot].copy_(loss, non_blocking=True)
        self.loss_ready[slot].record(compute)
        prev = self._collect()
        self.pending = slot
        self.k += 1
        return prev

    def _collect(self) -> Optional[float]:
        if self.pending is None:
            return None
        self.loss_ready[self.pending].synchronize()
        return float(self.loss_host[self.pending][0])

    def flush(self) -> Optional[float]:
        out = self._collect()
        self.pending = None
        return out


class GraphedStep:
    """``zero_grad + forward + loss + backward + optimizer.step`` of ANY hoir_b200 model (the per-layer kernels
    under autograd: Fourier networks, the neighborhood / random interpolators, GenerativeNetwork) captured in ONE
    CUDA graph for a fixed batch shape -- the launch-bound inner loop of the reference's training
    (/root/reference/high_order_implicit_representation/networks.py:64-73 under ``Trainer.fit``) replayed with a
    single ``cudaGraphLaunch``.  Networks with a fused whole-network kernel do not need it (``FusedTrainer`` is one
    launch already).

    Build it BEFORE running eager steps whose autograd graph is still referenced (e.g. a live ``loss`` tensor):
    such a graph keeps the parameters' AccumulateGrad nodes bound to the default stream and the capture fails with
    cudaErrorStreamCaptureImplicit.

    ``optimizer`` must be constructed with ``capturable=True`` (scalars in device memory).  ``x`` / ``y`` are the
    static input buffers: fill them (any stream-ordered producer, e.g. ``hoir_neighborhood_gather`` writing into
    them) and call ``replay()``, or pass tensors to ``step(x, y)``.  ``loss`` is the static loss tensor."""

    def __init__(self, model, optimizer, x_example: Tensor, y_example: Tensor, loss_fn=None, warmup: int = 3):
        if not getattr(optimizer, "capturable", False):
            raise ValueError("GraphedStep needs an optimizer constructed with capturable=True")
        self.model, self.optimizer = model, optimizer
        self.loss_fn = loss_fn or (lambda out, y: torch.nn.functional.mse_loss(out.flatten(), y.flatten()))
        self.x = x_example.detach().clone().contiguous()
        self.y = y_example.detach().clone().contiguous()
        params = [p for g in optimizer.param_groups for p in g["params"]]
        # warm-up on a side stream (allocator pools, library workspaces, lazily created optimizer state), then put
        # weights and optimizer state back: constructing a GraphedStep does not train
        w_backup = [p.detach().clone() for p in params]
        st_backup = {p: {k: (v.clone() if isinstance(v, Tensor) else v) for k, v in optimizer.state[p].items()}
                     for p in params if p in optimizer.state and len(optimizer.state[p])}
        steps_backup = [g.get("_step", 0) for g in optimizer.param_groups]
        side = torch.cuda.Stream(device=self.x.device)
        side.wait_stream(torch.cuda.current_stream(self.x.device))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._eager()
        torch.cuda.current_stream(self.x.device).wait_stream(side)
        with torch.no_grad():
            for p, w in zip(params, w_backup):
                p.copy_(w)
            for p in params:
                for k, v in optimizer.state[p].items():
                    if isinstance(v, Tensor):
                        v.copy_(st_backup[p][k]) if p in st_backup else v.zero_()
                    else:
                        optimizer.state[p][k] = st_backup[p][k] if p in st_backup else 0
        for g, n in zip(optimizer.param_groups, steps_backup):
            if "_step" in g:
                g["_step"] = n
        optimizer.zero_grad(set_to_none=True)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            loss = self.loss_fn(model(self.x), self.y)
            loss.backward()
            optimizer.step()
        self.loss = loss.detach()       # the static loss value; dropping `loss` frees the captured autograd graph
        del loss

    def _eager(self) -> Tensor:
        self.optimizer.zero_grad(set_to_none=True)
        loss = self.loss_fn(self.model(self.x), self.y)
        loss.backward()
        self.optimizer.step()
        return loss

    def replay(self) -> Tensor:
        _lib.weights_changed()
        self.optimizer.refresh_hyper()          # step count / lr -> device scalars, ordered before the replay
        self.graph.replay()
        return self.loss

    def step(self, x: Tensor, y: Tensor) -> Tensor:
        self.x.copy_(x, non_blocking=True)
        self.y.copy_(y, non_blocking=True)
        return self.replay()
