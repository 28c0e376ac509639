"""Drop-in for ``high_order_layers_torch.sparse_optimizers.SparseLion`` as constructed at
/root/reference/high_order_implicit_representation/networks.py:99-102, backed by
``hoir_sparse_lion_step`` (one launch per parameter tensor; ``hoir_b200.engine`` fuses the whole
network into a single launch on the flat parameter vector)."""
from __future__ import annotations

import torch

from . import _lib


class SparseLion(torch.optim.Optimizer):
    """Lion restricted to the entries whose gradient is non-zero."""

    def __init__(self, params, lr: float = 1e-4, betas=(0.9, 0.99), weight_decay: float = 0.0,
                 capturable: bool = False):
        if lr < 0.0:
            raise ValueError(f"Invalid learning rate: {lr}")
        super().__init__(params, dict(lr=lr, betas=betas, weight_decay=weight_decay))
        self.capturable = capturable         # scalars in device memory: step() can live in a CUDA graph (see Adam)

    def refresh_hyper(self) -> None:
        from .networks import _write_hyper
        for group in self.param_groups:
            if "_hyper" not in group:
                group["_hyper"] = torch.zeros(8, device=group["params"][0].device)
            b1, b2 = group["betas"]
            _write_hyper(group["_hyper"], [group["lr"], b1, b2, 0.0, group["weight_decay"]])

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.lib()
        _lib.weights_changed()               # (raw-pointer writes: invisible to the tensors' version counters)
        if self.capturable and not torch.cuda.is_current_stream_capturing():
            self.refresh_hyper()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                state = self.state[p]
                if len(state) == 0:
                    state["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                g = p.grad.contiguous()
                with torch.cuda.device(p.device):
                    if self.capturable:
                        _lib.check(lib.hoir_sparse_lion_step_dev(
                            p.numel(), _lib.dptr(p.data), _lib.dptr(g), _lib.dptr(state["exp_avg"]),
                            _lib.dptr(group["_hyper"]), _lib.stream_ptr(p.device)))
                        continue
                    _lib.check(lib.hoir_sparse_lion_step(
                        p.numel(), _lib.dptr(p.data), _lib.dptr(g), _lib.dptr(state["exp_avg"]),
                        float(group["lr"]), float(b1), float(b2), float(group["weight_decay"]),
                        _lib.stream_ptr(p.device)))
        return loss
