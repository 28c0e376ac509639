"""The random-interpolation path of the reference on the GPU: the radial sampler
(/root/reference/high_order_implicit_representation/random_sample_dataset.py:17-21, 122-244), the uniform
sampler (:41-119) and the iterated generators (utils.py:20-176).  The gathers run in ``hoir_radial_sample`` /
``hoir_radial_indices``; file decoding (PIL) is host-side I/O and not part of the path.

Random numbers: the reference draws ``r = torch.rand(N, F)`` then ``theta = torch.rand(N, F) * 2 * math.pi`` from
the global CPU generator.  Here every function takes ``r`` / ``theta`` explicitly (bit-exact parity with the
reference's stream) or, when they are omitted, draws them in the kernel from a counter-based Philox4x32-10
stream keyed by ``torch.initial_seed()`` with a process-wide call counter as the offset (``torch.manual_seed``
therefore makes runs reproducible, but the values are not torch's)."""
from __future__ import annotations

import itertools
from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor, nn

from . import _lib
from .utils import positions_from_mesh

_calls = itertools.count()


def _stream(seed: Optional[int], offset: Optional[int]) -> Tuple[int, int]:
    if seed is None:
        seed = torch.initial_seed()
    if offset is None:
        offset = next(_calls)
    return seed & 0xFFFFFFFFFFFFFFFF, offset & 0xFFFFFFFFFFFFFFFF


def _uniform_args(r: Optional[Tensor], theta: Optional[Tensor], shape, device):
    if (r is None) != (theta is None):
        raise ValueError("r and theta must be given together")
    if r is None:
        return None, None
    r = r.to(device=device, dtype=torch.float32).contiguous()
    theta = theta.to(device=device, dtype=torch.float32).contiguous()
    if tuple(r.shape) != tuple(shape) or tuple(theta.shape) != tuple(shape):
        raise ValueError(f"r / theta must have shape {tuple(shape)}, got {tuple(r.shape)} / {tuple(theta.shape)}")
    return r, theta


def indices_from_grid(image_size: int, device="cuda"):
    """``([S*S, 2] points (i, j) with k = i*S + j, linear pixel index i + j*S)`` -- random_sample_dataset.py:17-21."""
    k = torch.arange(image_size * image_size, device=device)
    indices = torch.stack([k // image_size, k % image_size], dim=1)
    return indices, indices[:, 0] + indices[:, 1] * image_size


def radial_uniforms(n: int, device="cuda", seed: Optional[int] = None, offset: Optional[int] = None):
    """The first ``n`` (r, theta) pairs of the in-kernel stream ``(seed, offset)``."""
    dev = torch.device(device)
    seed, offset = _stream(seed, offset)
    r = torch.empty(n, device=dev)
    theta = torch.empty(n, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().hoir_radial_uniforms(n, seed, offset, _lib.dptr(r), _lib.dptr(theta), _lib.stream_ptr(dev)))
    return r, theta


def random_symmetric_sample(image_size: int, interp_size: int, samples: Tensor, device="cuda",
                            r: Optional[Tensor] = None, theta: Optional[Tensor] = None,
                            seed: Optional[int] = None, offset: Optional[int] = None) -> Tensor:
    """``[num_samples, interp_size, 2]`` (x, y) grid indices of points at random radius / angle around every sample,
    reflected at the image borders (random_sample_dataset.py:211-244)."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("hoir_b200 has no CPU path: random_symmetric_sample needs a CUDA device")
    n = samples.shape[0]
    pts = samples.to(device=dev, dtype=torch.int32).contiguous()
    r, theta = _uniform_args(r, theta, (n, interp_size), dev)
    seed, offset = _stream(seed, offset)
    xy = torch.empty((n, interp_size, 2), device=dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().hoir_radial_indices(image_size, interp_size, n, _lib.dptr(pts, torch.int32), _lib.dptr(r),
                                                  _lib.dptr(theta), seed, offset, _lib.dptr(xy, torch.int32),
                                                  _lib.stream_ptr(dev)))
    return xy.long()


def stripe_planes(image_size: int, rotations: int, device="cuda") -> Tensor:
    """``[2*rotations, S, S]`` position planes, ``torch.cat([v.unsqueeze(0) for v in positions_from_mesh(...)])``."""
    lines = positions_from_mesh(width=image_size, height=image_size, device=str(device), rotations=rotations,
                                normalize=True)
    return torch.stack(lines, dim=0).contiguous()


def random_radial_samples_from_image(img: Tensor, stripe_list: Tensor, image_size: int, feature_pixels: int,
                                     indices: Optional[Tensor] = None, target_linear_indices: Optional[Tensor] = None,
                                     device=None, r: Optional[Tensor] = None, theta: Optional[Tensor] = None,
                                     seed: Optional[int] = None, offset: Optional[int] = None):
    """Random interpolation sets for EVERY pixel of ``img [C, S, S]`` (or a batch ``[M, C, S, S]``):
    ``features [M*S*S, F, C + P]`` (values, positions relative to the target) and ``targets [M*S*S, 1, C]``
    (random_sample_dataset.py:176-208).  ``indices`` / ``target_linear_indices`` are accepted for signature
    compatibility; the targets are always the full grid in ``indices_from_grid`` order, as in every call site."""
    if not img.is_cuda:
        raise RuntimeError("hoir_b200 has no CPU path: expected a CUDA image tensor")
    dev = img.device
    im = img.float().contiguous()
    if im.dim() == 3:
        im = im.unsqueeze(0)
    M, Cc, S, S2 = im.shape
    if S != image_size or S2 != image_size:
        raise ValueError(f"expected a square image of size {image_size}, got {S} x {S2}")
    st = stripe_list.to(device=dev, dtype=torch.float32).contiguous()
    P = st.shape[0]
    n = M * S * S
    r, theta = _uniform_args(r, theta, (n, feature_pixels), dev)
    seed, offset = _stream(seed, offset)
    features = torch.empty((n, feature_pixels, Cc + P), device=dev)
    targets = torch.empty((n, 1, Cc), device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().hoir_radial_sample(_lib.dptr(im), _lib.dptr(st), M, Cc, P, S, feature_pixels,
                                                 _lib.dptr(r), _lib.dptr(theta), seed, offset, _lib.dptr(features),
                                                 _lib.dptr(targets), _lib.stream_ptr(dev)))
    return features, targets


def random_image_samples(img: Tensor, stripe_list: Tensor, num_feature_pixels: int, num_target_pixels: int = 1,
                         perm: Optional[Tensor] = None):
    """The uniform sampler (RandomImageSampleDataset.__getitem__, random_sample_dataset.py:77-119): a random
    permutation of the pixels split into disjoint interpolation sets and targets.  Index bookkeeping on the device
    with torch ops (a permutation and two row gathers; no arithmetic of the layer path)."""
    planes = torch.cat([img.float(), stripe_list.to(img.device)]).reshape(img.shape[0] + stripe_list.shape[0], -1).permute(1, 0)
    size, channels = planes.shape
    frac = num_feature_pixels / (num_target_pixels + num_feature_pixels)
    feature_size = int(frac * size)
    examples = min(feature_size // num_feature_pixels, (size - feature_size) // num_target_pixels)
    if perm is None:
        perm = torch.randperm(size, device=img.device)
    feature_size, target_size = examples * num_feature_pixels, examples * num_target_pixels
    features = planes[perm[:feature_size], :].reshape(examples, num_feature_pixels, channels).clone()
    targets = planes[perm[feature_size:feature_size + target_size], :].reshape(examples, num_target_pixels, channels)
    features[:, :, 3:] = features[:, :, 3:] - targets[:, :, 3:]
    return features, targets[:, :, :3]


def load_image(path: str, image_size: int, flip: bool = False) -> Tensor:
    """``standard_transforms`` (random_sample_dataset.py:24-38) without torchvision: resize the short side to
    ``image_size`` (bilinear), optional horizontal flip, centre crop, scale to [-1, 1].  Host I/O."""
    import numpy as np
    from PIL import Image
    im = Image.open(path).convert("RGB")
    w, h = im.size
    if w <= h:
        nw, nh = image_size, max(image_size, int(image_size * h / w))
    else:
        nw, nh = max(image_size, int(image_size * w / h)), image_size
    im = im.resize((nw, nh), Image.BILINEAR)
    left, top = int(round((nw - image_size) / 2.0)), int(round((nh - image_size) / 2.0))
    im = im.crop((left, top, left + image_size, top + image_size))
    t = torch.from_numpy(np.array(im)).permute(2, 0, 1).float() / 255.0
    if flip:
        t = t.flip(2)
    return (t - 0.5) / 0.5


class RadialRandomImageSampleDataset:
    """random_sample_dataset.py:122-173 with the images resident on the device: item ``i`` is a fresh radial
    sampling of image ``i`` (``[S*S, F, 3 + 2*rotations]``, ``[S*S, 1, 3]``).  ``images`` may be paths or a tensor
    ``[M, 3, S, S]`` in [-1, 1]; ``batch(idx)`` samples several images in ONE launch (the collate of
    random_image_sample_collate_fn, :247-254)."""

    def __init__(self, image_size: int, path_list, num_feature_pixels: int = 25, num_target_pixels: int = 1,
                 rotations: int = 1, device: str = "cuda"):
        self._image_size = image_size
        self._num_feature_pixels = num_feature_pixels
        self._num_target_pixels = num_target_pixels
        if isinstance(path_list, Tensor):
            self.images = path_list.to(device).float().contiguous()
        else:
            flips = torch.rand(len(path_list)) < 0.5          # RandomHorizontalFlip, drawn once at load time
            self.images = torch.stack([load_image(p, image_size, bool(f)) for p, f in zip(path_list, flips)]).to(device)
        self._stripe_list = stripe_planes(image_size, rotations, device)

    def __len__(self) -> int:
        return self.images.shape[0]

    def batch(self, idx: Sequence[int], **uniforms):
        sel = self.images[torch.as_tensor(list(idx), device=self.images.device)]
        return random_radial_samples_from_image(sel, self._stripe_list, self._image_size, self._num_feature_pixels,
                                                **uniforms)

    def __getitem__(self, index: int):
        return self.batch([index])


class RandomImageSampleDataset(RadialRandomImageSampleDataset):
    """The uniform (permutation) sampler, random_sample_dataset.py:41-119."""

    def batch(self, idx: Sequence[int], **_):
        parts = [random_image_samples(self.images[i], self._stripe_list, self._num_feature_pixels, self._num_target_pixels)
                 for i in idx]
        return torch.cat([p[0] for p in parts]), torch.cat([p[1] for p in parts])


def _run_model(model: nn.Module, x: Tensor, batch_size: Optional[int]) -> Tensor:
    if not batch_size or batch_size >= x.shape[0]:
        return model(x)
    return torch.cat([model(x[b:b + batch_size]) for b in range(0, x.shape[0], batch_size)])


def generate_sample_radial(model: nn.Module, features: int = 10, targets: int = 1, iterations: int = 10,
                           start_image: Optional[Tensor] = None, image_size: int = 64, return_intermediate: bool = True,
                           device: str = "cuda", rotations: int = 2, batch_size: Optional[int] = None,
                           all_random: bool = True, uniforms: Optional[Sequence[Tuple[Tensor, Tensor]]] = None) -> List[Tensor]:
    """utils.py:20-97: apply the network ``iterations`` times to radial samplings of an image (a random one, the
    start image, or -- ``all_random=False`` -- the previous estimate); returns the ``[3, S, S]`` estimates.  As in
    the reference, estimate pixel ``[c, i, j]`` is the prediction for target ``k = i*S + j`` (whose pixel is
    ``[c, j, i]`` of the input: every application transposes the picture).  ``uniforms[it] = (r, theta)`` replays
    given random numbers."""
    model.eval()
    dev = torch.device(device)
    image = None if start_image is None else start_image.to(dev).float().clone()
    if image is None:
        image = torch.rand([3, image_size, image_size], device=dev) * 2 - 1
    stripes = stripe_planes(image_size, rotations, dev)
    results = []
    with torch.no_grad():
        for count in range(iterations):
            if all_random and start_image is None:
                image = torch.rand([3, image_size, image_size], device=dev) * 2 - 1
            elif all_random:
                image = start_image.to(dev).float().clone()
            kw = {} if uniforms is None else dict(r=uniforms[count][0], theta=uniforms[count][1])
            feat, _ = random_radial_samples_from_image(image, stripes, image_size, features, **kw)
            result = _run_model(model, feat.flatten(1), batch_size)
            image = result.reshape(image_size, image_size, 3).permute(2, 0, 1).contiguous()
            results.append(image)
    return results


def generate_sample(model: nn.Module, features: int = 10, targets: int = 1, iterations: int = 10,
                    image: Optional[Tensor] = None, image_size: int = 64, return_intermediate: bool = True,
                    device: str = "cuda", rotations: int = 2, batch_size: Optional[int] = None,
                    all_random: bool = True) -> List[Tensor]:
    """utils.py:100-176: the same iteration with uniformly random interpolation points,
    ``randperm(S*S*F) % (S*S)``; ``image`` is an UNnormalised [0, 255] picture."""
    model.eval()
    dev = torch.device(device)
    num_pixels = image_size * image_size
    if image is None:
        img = torch.rand([3, image_size, image_size], device=dev) * 2 - 1
    else:
        img = (image.to(dev).float() / 255) * 2 - 1
    stripes = stripe_planes(image_size, rotations, dev)
    results = []
    with torch.no_grad():
        for _ in range(iterations):
            if all_random:
                img = torch.rand([3, image_size, image_size], device=dev) * 2 - 1
            full = torch.cat([img, stripes]).reshape(3 + stripes.shape[0], -1).permute(1, 0)
            idx = torch.remainder(torch.randperm(num_pixels * features, device=dev), num_pixels)
            feat = full[idx].reshape(-1, features, full.shape[1]).clone()
            targ = full.reshape(-1, 1, full.shape[1])
            feat[:, :, 3:] = feat[:, :, 3:] - targ[:, :, 3:]
            result = _run_model(model, feat.flatten(1), batch_size)
            img = result.reshape(image_size, image_size, 3).permute(2, 0, 1).contiguous()
            results.append(img)
    return results
