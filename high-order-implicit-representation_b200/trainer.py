"""The caller of the hot path: what ``examples/implicit_images.py`` does with Lightning's ``Trainer``,
``ImageDataModule`` and the ``ImageGenerator`` callback
(/root/reference/examples/implicit_images.py:31-52, /root/reference/high_order_implicit_representation/
single_image_dataset.py:286-333, rendering.py:166-222) as a plain loop over the fused engine:

    epochs x (shuffled pixel batches -> one fused training step each) -> epoch ``train_loss`` ->
    ReduceLROnPlateau / ExponentialLR on it (networks.py:104-118) -> render the image -> checkpoint

Lightning, Hydra and TensorBoard are not installable here and are host-side harness; this module keeps the
``cfg`` keys, the ``train_loss`` metric name and the Lightning checkpoint layout so runs are interchangeable.
SURVEY.md section 8f, row 1.
"""
from __future__ import annotations

import logging
import math
from typing import Callable, List, Optional, Union

import torch
from torch import Tensor

from . import _lib, parallel
from .config import to_cfg
from .datasets import _load_image_u8, image_to_dataset
from .engine import FusedTrainer
from .networks import Net
from .rendering import render_image

logger = logging.getLogger(__name__)


class _LrSchedule:
    """``optimizer.scheduler`` of the reference config on a bare learning rate (networks.py:104-118):
    'plateau' = ReduceLROnPlateau(patience, factor) on the epoch train_loss, 'exponential' = ExponentialLR(gamma)."""

    def __init__(self, opt_cfg, lr: float):
        get = opt_cfg.get if hasattr(opt_cfg, "get") else (lambda k, d=None: getattr(opt_cfg, k, d))
        self.kind = get("scheduler", None)
        self.lr = lr
        self.patience = int(get("patience", 10) or 10)
        self.factor = float(get("factor", 0.1) or 0.1)
        self.gamma = float(get("gamma", 0.99) or 0.99)
        self.best = math.inf
        self.bad = 0

    def step(self, metric: float) -> float:
        if self.kind == "plateau":          # torch defaults: mode='min', threshold=1e-4 'rel', cooldown=0
            if metric < self.best * (1.0 - 1e-4):
                self.best, self.bad = metric, 0
            else:
                self.bad += 1
            if self.bad > self.patience:
                self.lr *= self.factor
                self.bad = 0
        elif self.kind == "exponential":
            self.lr *= self.gamma
        return self.lr


class ImageTrainer:
    """Fit ``Net(cfg)`` to one image.  ``sampling='shuffle'`` draws the reference's shuffled pixel batches
    (DataLoader(shuffle=True) over image_to_dataset rows) on the device; ``sampling='tiles'`` walks contiguous
    pixel ranges in random order with coordinates and targets generated in-kernel (no dataset in HBM)."""

    def __init__(self, cfg, image: Union[str, Tensor], device: str = "cuda", sampling: str = "shuffle",
                 process_group=None, seed: int = 0, drop_last: bool = True, cuda_graph: bool = False):
        if sampling not in ("shuffle", "tiles"):
            raise ValueError(f"sampling {sampling} not recognized")
        self.cfg = to_cfg(cfg)
        self.device = torch.device(device)
        self.sampling = sampling
        self.drop_last = bool(drop_last)        # the reference's train DataLoader drops the ragged last batch
        self.rotations = int(self.cfg.rotations)
        self.batch_size = int(self.cfg.batch_size)
        self.image_u8 = _load_image_u8(image).to(self.device).contiguous()
        self.rows, self.cols = int(self.image_u8.shape[0]), int(self.image_u8.shape[1])
        self.net = Net(self.cfg).to(self.device)
        self.rank, self.world = parallel.world_info(process_group)
        self.gen = torch.Generator(device=self.device).manual_seed(seed)     # same stream on every rank
        opt = self.cfg.optimizer
        self.schedule = _LrSchedule(opt, float(opt.lr))
        self.fused: Optional[FusedTrainer] = None
        self.optimizer = None
        if self.net.model.fused_supported():
            self.fused = FusedTrainer(self.net.model, optimizer=opt.name, lr=float(opt.lr),
                                      process_group=process_group)
        else:                                   # per-layer kernels under autograd (Fourier, polynomial, ...)
            conf = self.net.configure_optimizers()
            self.optimizer = conf[0][0] if isinstance(conf, tuple) else conf
        self._graphed = None
        if cuda_graph and self.fused is None and self.world == 1:
            # networks without a fused kernel (Fourier, ...): full batches replay ONE CUDA graph of the whole step
            from .engine import GraphedStep
            self.optimizer.capturable = True
            self._graphed = GraphedStep(
                self.net, self.optimizer,
                torch.zeros((self.batch_size, int(self.cfg.mlp.input.width)), device=self.device),
                torch.zeros((self.batch_size, int(self.cfg.mlp.output.width)), device=self.device))
        self._pos = self._targets = None
        self.history: List[float] = []

    # ------------------------------------------------------------------------------------------------
    def _dataset(self):
        if self._pos is None:
            self._targets, self._pos, _ = image_to_dataset(self.image_u8, rotations=self.rotations,
                                                           device=str(self.device))
        return self._pos, self._targets

    def _set_lr(self, lr: float) -> None:
        if self.fused is not None:
            self.fused.lr = lr
        else:
            for g in self.optimizer.param_groups:
                g["lr"] = lr

    def _step(self, x: Optional[Tensor], y: Optional[Tensor], first: int, count: int, global_batch: int) -> Tensor:
        """One optimisation step on this rank's share of a global batch.  EVERY rank calls this for EVERY global
        batch, also with an empty share (count / x.shape[0] == 0): the gradient exchange (in-kernel peer rendezvous or
        NCCL all-reduce) is collective, a rank that skipped it would leave the others waiting
        (include/hoir.h: hoir_mlp_train_step_fused_peer)."""
        if self.fused is not None:
            if x is None:
                return self.fused.train_step_mesh(self.image_u8, self.rotations, first, count, global_batch=global_batch)
            return self.fused.train_step(x, y, global_batch=global_batch)
        if x is None:
            pos, tg = self._dataset()
            x, y = pos[first:first + count], tg[first:first + count]
        if self._graphed is not None and x.shape[0] == self.batch_size:
            return self._graphed.step(x, y).reshape(1)
        self.optimizer.zero_grad(set_to_none=True)
        local = int(x.shape[0])
        if local > 0:
            # MSE(mean) over the GLOBAL batch (Quirk Q12): scale this rank's mean by its share
            loss = self.net.training_step((x, y), 0) * (local / float(global_batch))
            loss.backward()
            loss = loss.detach().reshape(1)
        else:
            loss = torch.zeros(1, device=self.device)
        if self.world > 1:
            for p in self.net.parameters():
                if p.grad is None:
                    p.grad = torch.zeros_like(p)
                parallel.allreduce_sum_(p.grad, None)
            parallel.allreduce_sum_(loss, None)
        self.optimizer.step()
        return loss

    def train_epoch(self) -> float:
        """One pass over the image; every rank takes its share of each global batch.  Returns train_loss."""
        n = self.rows * self.cols
        bs = self.batch_size
        total = torch.zeros(1, device=self.device)
        steps = 0
        if self.sampling == "shuffle":
            perm = torch.randperm(n, device=self.device, generator=self.gen)
            # networks on the tensor-core training kernel take the shuffled pixel INDICES (4 B per sample; coordinates
            # and targets are produced in-kernel, no dataset in HBM); the others gather rows of image_to_dataset
            by_index = self.fused is not None and self.fused._sort_ok and n < 2 ** 31
            if by_index:
                perm = perm.to(torch.int32)
            else:
                pos, tg = self._dataset()
            for lo, hi, gb in parallel.shuffle_schedule(n, bs, self.rank, self.world, self.drop_last):
                mine = perm[lo:hi][self.rank::self.world]   # may be empty on the ragged last batch: still a step
                if by_index:
                    total += self.fused.train_step_pixels(self.image_u8, self.rotations, mine.contiguous(), global_batch=gb)
                else:
                    total += self._step(pos[mine], tg[mine], 0, 0, global_batch=gb)
                steps += 1
        else:
            n_blocks = (n + bs - 1) // bs
            order = torch.randperm(n_blocks, device=self.device, generator=self.gen).tolist()
            for first, count, gb in parallel.tile_schedule(order, n, bs, self.rank, self.world):
                total += self._step(None, None, first, count, global_batch=gb)
                steps += 1
        self.net.global_step += steps
        return float(total.item()) / max(steps, 1)

    def fit(self, max_epochs: Optional[int] = None,
            on_epoch_end: Optional[Callable[["ImageTrainer", int, float], None]] = None) -> List[float]:
        epochs = int(max_epochs if max_epochs is not None else self.cfg.max_epochs)
        first = int(self.net.current_epoch)                  # > 0 after load_checkpoint: epochs keep counting
        for epoch in range(first, first + epochs):
            loss = self.train_epoch()
            self.history.append(loss)
            self.net.current_epoch = epoch + 1
            self.net.log("train_loss", loss)
            self._set_lr(self.schedule.step(loss))
            logger.info("epoch %d train_loss %.6f lr %.3g", epoch, loss, self.schedule.lr)
            if on_epoch_end is not None:
                on_epoch_end(self, epoch, loss)
        return self.history

    # ------------------------------------------------------------------------------------------------
    def render(self) -> Tensor:
        """ImageGenerator.on_train_epoch_end (rendering.py:195-213): ``[rows, cols, 3]`` in [0, 1]."""
        if self.fused is not None:
            return self.fused.render(self.rows, self.cols, self.rotations).reshape(self.rows, self.cols, -1)
        return render_image(self.net, self.rows, self.cols, self.rotations, batch_size=max(self.batch_size, 1 << 16))

    def psnr(self) -> float:
        target = self.image_u8.float() / 255.0
        mse = torch.mean((self.render().clamp(0, 1) - target) ** 2).item()
        return 10.0 * math.log10(1.0 / max(mse, 1e-12))

    def save_checkpoint(self, path: str) -> None:
        """Lightning-format checkpoint: ``state_dict`` + ``hyper_parameters`` (what ``Net.load_from_checkpoint``
        reads) plus ``optimizer_states`` / ``lr_schedulers`` / ``epoch`` / ``global_step`` so a run can resume."""
        ckpt = self.net.checkpoint_dict()
        if self.fused is not None:
            ckpt["optimizer_states"] = [self.fused.optimizer_state_dict()]
        else:
            ckpt["optimizer_states"] = [self.optimizer.state_dict()]
        ckpt["lr_schedulers"] = [dict(kind=self.schedule.kind, lr=self.schedule.lr, best=self.schedule.best,
                                      num_bad_epochs=self.schedule.bad)]
        ckpt["train_loss_history"] = list(self.history)
        torch.save(ckpt, path)

    def load_checkpoint(self, path: str) -> None:
        """Resume: weights, optimizer moments and step, lr schedule, epoch counters."""
        ckpt = torch.load(path, map_location="cpu", weights_only=False)
        self.net.load_state_dict(ckpt["state_dict"])      # in-place into the flat parameter vector
        if self.fused is not None:
            self.fused.pack()                             # the kernel-side copy follows explicitly
        self.net.current_epoch = ckpt.get("epoch", 0)
        self.net.global_step = ckpt.get("global_step", 0)
        if ckpt.get("optimizer_states"):
            if self.fused is not None:
                self.fused.load_optimizer_state_dict(ckpt["optimizer_states"][0])
            else:
                self.optimizer.load_state_dict(ckpt["optimizer_states"][0])
        for sch in ckpt.get("lr_schedulers", [])[:1]:
            self.schedule.lr, self.schedule.best, self.schedule.bad = sch["lr"], sch["best"], sch["num_bad_epochs"]
            self._set_lr(self.schedule.lr)
        self.history = list(ckpt.get("train_loss_history", []))


class NeighborhoodTrainer:
    """``examples/implicit_neighborhood.py`` without Lightning: fit ``Net(cfg)`` to predict the inner
    ``width x width`` block of every patch of one or more images from its donut of ``outside`` pixels
    (ImageNeighborhoodDataModule, /root/reference/high_order_implicit_representation/
    single_image_dataset.py:336-389; config/neighborhood_config.yaml), then iterate the trained interpolator
    (generate_sequence, rendering.py:98-126).  Patches are never materialised as a dataset: every step gathers its
    batch of patches from the image resident in HBM (``hoir_neighborhood_gather`` with a first patch / count)."""

    def __init__(self, cfg, images, device: str = "cuda", seed: int = 0, cuda_graph: bool = False):
        self.cfg = to_cfg(cfg)
        self.device = torch.device(device)
        self.cuda_graph, self._graphed = bool(cuda_graph), None
        self.width, self.outside = int(self.cfg.data.width), int(self.cfg.data.outside)
        total = self.width + 2 * self.outside
        self.cfg.mlp.input.width = (total * total - self.width * self.width) * 3
        self.cfg.mlp.output.width = 3 * self.width * self.width
        self.batch_size = int(self.cfg.batch_size)
        if not isinstance(images, (list, tuple)):
            images = [images]
        self.images = []
        for im in images:                                   # [3, H, W] in [-1, 1] (ToTensor() * 2 - 1)
            if isinstance(im, str):
                im = _load_image_u8(im).permute(2, 0, 1).float() / 255.0 * 2 - 1
            self.images.append(im.to(self.device).contiguous().float())
        self.net = Net(self.cfg).to(self.device)
        conf = self.net.configure_optimizers()
        self.optimizer = conf[0][0] if isinstance(conf, tuple) else conf
        self.optimizer.capturable = self.cuda_graph          # scalars in device memory: the step can live in a graph
        self.schedule = _LrSchedule(self.cfg.optimizer, float(self.cfg.optimizer.lr))
        self.gen = torch.Generator(device=self.device).manual_seed(seed)
        self.history: List[float] = []

    def _patch_batches(self):
        """(image index, first patch, count) blocks of at most batch_size consecutive patches, shuffled."""
        from .datasets import neighborhood_patches
        blocks = []
        for k, im in enumerate(self.images):
            nr, nc = neighborhood_patches(im.shape[1], im.shape[2], self.width, self.outside, 1)
            total = nr * nc
            blocks += [(k, f, min(self.batch_size, total - f)) for f in range(0, total, self.batch_size)]
        order = torch.randperm(len(blocks), device=self.device, generator=self.gen).tolist()
        return [blocks[j] for j in order]

    def train_epoch(self) -> float:
        lib = _lib.lib()
        in_w, out_w = int(self.cfg.mlp.input.width), int(self.cfg.mlp.output.width)
        if self.cuda_graph and self._graphed is None:
            # captured before any eager step of this trainer: a live autograd graph from an eager iteration (its
            # `loss`) would pin the weights' AccumulateGrad nodes to the default stream and break the capture
            from .engine import GraphedStep
            self._graphed = GraphedStep(self.net, self.optimizer, torch.zeros((self.batch_size, in_w), device=self.device),
                                        torch.zeros((self.batch_size, out_w), device=self.device))
        if self._graphed is not None:
            feat, targ = self._graphed.x, self._graphed.y
        else:
            feat = torch.empty((self.batch_size, in_w), device=self.device)
            targ = torch.empty((self.batch_size, out_w), device=self.device)
        total = torch.zeros(1, device=self.device)
        steps = 0
        for k, first, count in self._patch_batches():
            im = self.images[k]
            with torch.cuda.device(self.device):
                _lib.check(lib.hoir_neighborhood_gather(_lib.dptr(im), 3, im.shape[1], im.shape[2], self.width,
                                                        self.outside, 1, 0, first, count, _lib.dptr(feat),
                                                        _lib.dptr(targ), _lib.stream_ptr(self.device)))
            if self.cuda_graph and count == self.batch_size:
                # full batches: zero_grad + forward + loss + backward + optimizer replayed from ONE CUDA graph whose
                # static inputs are the gather's output buffers (the ragged last block of an image runs eagerly)
                total += self._graphed.replay().detach()
                steps += 1
                continue
            self.optimizer.zero_grad(set_to_none=True)
            loss = self.net.training_step((feat[:count], targ[:count]), steps)
            loss.backward()
            self.optimizer.step()
            total += loss.detach()
            del loss
            steps += 1
        self.net.global_step += steps
        return float(total.item()) / max(steps, 1)

    def fit(self, max_epochs: Optional[int] = None) -> List[float]:
        epochs = int(max_epochs if max_epochs is not None else self.cfg.max_epochs)
        first = int(self.net.current_epoch)
        for epoch in range(first, first + epochs):
            loss = self.train_epoch()
            self.history.append(loss)
            self.net.current_epoch = epoch + 1
            self.net.log("train_loss", loss)
            lr = self.schedule.step(loss)
            for g in self.optimizer.param_groups:
                g["lr"] = lr
            logger.info("epoch %d train_loss %.6f lr %.3g", epoch, loss, lr)
        return self.history

    def generate(self, image: Optional[Tensor] = None, frames: int = 25, skip: int = 5, alpha: float = 0.75,
                 size=(180, 180)) -> Tensor:
        """NeighborGenerator / the evaluation branch of the example: iterate the interpolator from an image or
        from noise; returns the kept frames ``[k, 3, H, W]`` in [0, 1] on the host."""
        from .rendering import generate_sequence
        if image is None:
            image = torch.rand(3, size[0], size[1], device=self.device, generator=self.gen) * 2 - 1
        return generate_sequence(self.net, image.to(self.device), frames=frames, width=self.width,
                                 outside=self.outside, batch_size=max(self.batch_size, 1 << 16), skip=skip, alpha=alpha)

    def save_checkpoint(self, path: str) -> None:
        ckpt = self.net.checkpoint_dict()
        ckpt["optimizer_states"] = [self.optimizer.state_dict()]
        torch.save(ckpt, path)


class RandomInterpolationTrainer:
    """``examples/random_interpolation.py`` without Lightning: fit ``Net(cfg)`` to predict every pixel of a set of
    ``image_size x image_size`` images from ``num_feature_pixels`` points drawn at random radius / angle around it
    (RandomImageSampleDataModule + RadialRandomImageSampleDataset,
    /root/reference/high_order_implicit_representation/random_sample_dataset.py:122-173, 256-384;
    config/random_interpolation.yaml), then apply the trained interpolator iteratively (generate_sample_radial,
    utils.py:20-97).  A step = ``batch_size`` images (the reference's collate concatenates their S*S examples);
    the interpolation sets are drawn in-kernel, fresh every step, and never stored as a dataset."""

    def __init__(self, cfg, images, device: str = "cuda", seed: int = 0, images_per_step: Optional[int] = None):
        from .random_sample import RadialRandomImageSampleDataset
        self.cfg = to_cfg(cfg)
        self.device = torch.device(device)
        self.size = int(self.cfg.image_size)
        self.features = int(self.cfg.num_feature_pixels)
        self.rotations = int(self.cfg.rotations)
        self.cfg.mlp.input.width = (3 + 2 * self.rotations) * self.features
        self.cfg.mlp.input.segments = self.size                  # `segments: ${image_size}`
        self.dataset = RadialRandomImageSampleDataset(self.size, images, num_feature_pixels=self.features,
                                                      num_target_pixels=int(self.cfg.num_target_pixels),
                                                      rotations=self.rotations, device=device)
        self.images_per_step = int(images_per_step or min(int(self.cfg.batch_size), len(self.dataset)))
        self.net = Net(self.cfg).to(self.device)
        conf = self.net.configure_optimizers()
        self.optimizer = conf[0][0] if isinstance(conf, tuple) else conf
        self.schedule = _LrSchedule(self.cfg.optimizer, float(self.cfg.optimizer.lr))
        self.gen = torch.Generator(device=self.device).manual_seed(seed)
        self.seed, self.draws = seed, 0
        self.history: List[float] = []

    def train_epoch(self) -> float:
        order = torch.randperm(len(self.dataset), device=self.device, generator=self.gen).tolist()
        total = torch.zeros(1, device=self.device)
        steps = 0
        for b in range(0, len(order), self.images_per_step):
            feat, targ = self.dataset.batch(order[b:b + self.images_per_step], seed=self.seed, offset=self.draws)
            self.draws += 1
            self.optimizer.zero_grad(set_to_none=True)
            loss = self.net.training_step((feat, targ), steps)
            loss.backward()
            self.optimizer.step()
            total += loss.detach()
            steps += 1
        self.net.global_step += steps
        return float(total.item()) / max(steps, 1)

    fit = NeighborhoodTrainer.fit

    def generate(self, start_image: Optional[Tensor] = None, iterations: Optional[int] = None,
                 all_random: Optional[bool] = None) -> List[Tensor]:
        """ImageSampler / the evaluation branch of the example (utils.py:179-225, examples/random_interpolation.py:
        96-118): the estimates ``[3, S, S]`` in [-1, 1] after each application of the network."""
        from .random_sample import generate_sample_radial
        return generate_sample_radial(self.net, features=self.features, iterations=int(iterations or self.cfg.iterations),
                                      start_image=start_image, image_size=self.size, device=str(self.device),
                                      rotations=self.rotations,
                                      all_random=bool(self.cfg.all_random if all_random is None else all_random))

    save_checkpoint = NeighborhoodTrainer.save_checkpoint
