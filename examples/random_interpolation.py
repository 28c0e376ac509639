#!/usr/bin/env python
"""The reference's ``examples/random_interpolation.py`` flow on the B200 engine, without Hydra / Lightning:

    python examples/random_interpolation.py folder=images image_size=64 num_feature_pixels=32 max_epochs=5

Same config keys as /root/reference/config/random_interpolation.yaml; ``train=false checkpoint=<file>`` applies a
saved interpolator iteratively to ``image=`` or to noise (/root/reference/examples/random_interpolation.py:46-128).
Without a folder of images a set of synthetic textures is used.
"""
import glob
import logging
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import hoir_b200 as H  # noqa: E402

logging.basicConfig(level=logging.INFO, format="%(message)s")
logger = logging.getLogger("random_interpolation")


def synthetic_textures(count: int, size: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(0)
    u = torch.linspace(0, 1, size).view(1, 1, 1, size)
    v = torch.linspace(0, 1, size).view(1, 1, size, 1)
    f = torch.rand(count, 3, 1, 1, generator=g) * 10 + 2
    ph = torch.rand(count, 3, 1, 1, generator=g) * 6.28
    return torch.sin(f * u + ph) * torch.cos(0.7 * f * v - ph) * 0.9


def main(argv):
    cfg = H.random_interpolation_config(out_dir="outputs")
    cfg.folder = None
    H.apply_overrides(cfg, [a for a in argv if "=" in a])
    os.makedirs(cfg.out_dir, exist_ok=True)
    size = int(cfg.image_size)
    if cfg.train:
        if cfg.folder:
            paths = sorted(p for ext in ("jpg", "jpeg", "png", "JPEG") for p in glob.glob(os.path.join(cfg.folder, "**", f"*.{ext}"), recursive=True))
            if not paths:
                raise ValueError(f"Folder {cfg.folder} holds no images.")
            images = paths[:int(len(paths) * float(cfg.split_frac)) or 1]
        else:
            images = synthetic_textures(64, size)
        tr = H.RandomInterpolationTrainer(cfg, images)
        logger.info("%d images, input %d, %d parameters", len(tr.dataset), cfg.mlp.input.width,
                    sum(p.numel() for p in tr.net.parameters()))
        tr.fit()
        tr.save_checkpoint(os.path.join(cfg.out_dir, "random_interpolation.ckpt"))
        frames = tr.generate()
    else:
        model = H.Net.load_from_checkpoint(cfg.checkpoint, map_location="cuda")
        start = H.random_sample.load_image(cfg.image, size).cuda() if cfg.image and os.path.exists(str(cfg.image)) else None
        frames = H.generate_sample_radial(model, features=int(cfg.num_feature_pixels), rotations=int(cfg.rotations),
                                          image_size=size, iterations=int(cfg.iterations), all_random=bool(cfg.all_random),
                                          start_image=start)
    from torchvision.utils import save_image
    save_image(0.5 * (torch.stack(frames).cpu() + 1), os.path.join(cfg.out_dir, "random_interpolation.png"), nrow=5)
    logger.info("%d estimates -> %s", len(frames), os.path.join(cfg.out_dir, "random_interpolation.png"))


if __name__ == "__main__":
    main(sys.argv[1:])
