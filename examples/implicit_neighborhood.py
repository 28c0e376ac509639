#!/usr/bin/env python
"""The reference's ``examples/implicit_neighborhood.py`` flow on the B200 engine, without Hydra / Lightning:

    python examples/implicit_neighborhood.py images=[a.jpg,b.jpg] mlp.n=2 mlp.input.segments=50 \
        mlp.hidden.width=50 mlp.hidden.layers=2 batch_size=65536 max_epochs=5

Same config keys as /root/reference/config/neighborhood_config.yaml; ``train=false checkpoint=<file>`` iterates a
saved interpolator from an image or from noise (/root/reference/examples/implicit_neighborhood.py:27-130).
Without ``images=`` a synthetic texture is used.
"""
import logging
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import hoir_b200 as H  # noqa: E402

logging.basicConfig(level=logging.INFO, format="%(message)s")
logger = logging.getLogger("implicit_neighborhood")


def synthetic_texture(h: int = 256, w: int = 256) -> torch.Tensor:
    u = torch.linspace(0, 1, w).view(1, 1, w)
    v = torch.linspace(0, 1, h).view(1, h, 1)
    ph = torch.tensor([0.0, 1.3, 2.6]).view(3, 1, 1)
    return torch.sin(12 * u + ph) * torch.cos(9 * v - ph) * 0.8


def main(argv):
    cfg = H.neighborhood_config(images=[], out_dir="outputs", frames=25, skip=5)
    cfg.mlp.periodicity = None
    for a in argv:
        if a.startswith("images="):
            cfg.images = [p.strip() for p in a[len("images="):].strip("[]").split(",") if p.strip()]
    H.apply_overrides(cfg, [a for a in argv if "=" in a and not a.startswith("images=")])
    os.makedirs(cfg.out_dir, exist_ok=True)
    images = list(cfg.images) if cfg.images else [synthetic_texture()]
    if cfg.train:
        tr = H.NeighborhoodTrainer(cfg, images)
        logger.info("input %d -> output %d, %d parameters", cfg.mlp.input.width, cfg.mlp.output.width,
                    sum(p.numel() for p in tr.net.parameters()))
        tr.fit()
        tr.save_checkpoint(os.path.join(cfg.out_dir, "neighborhood.ckpt"))
        seq = tr.generate(frames=int(cfg.frames), skip=int(cfg.skip))
    else:
        model = H.Net.load_from_checkpoint(cfg.checkpoint, map_location="cuda")
        image = torch.rand(3, 500, 500, device="cuda") * 2 - 1
        seq = H.generate_sequence(model, image, frames=int(cfg.frames), width=int(cfg.data.width),
                                  outside=int(cfg.data.outside), batch_size=1 << 16, skip=int(cfg.skip))
    from torchvision.utils import save_image
    save_image(seq, os.path.join(cfg.out_dir, "sequence.png"), nrow=8)
    logger.info("sequence of %d frames -> %s", seq.shape[0], os.path.join(cfg.out_dir, "sequence.png"))


if __name__ == "__main__":
    main(sys.argv[1:])
