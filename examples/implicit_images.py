#!/usr/bin/env python
"""The reference's ``examples/implicit_images.py`` flow on the B200 engine, without Hydra / Lightning:

    python examples/implicit_images.py images=[path.jpg] mlp.layer_type=continuous mlp.hidden.width=40 \
        mlp.hidden.layers=1 batch_size=1048576 max_epochs=20 optimizer.name=adam optimizer.lr=1e-3

Same config keys as /root/reference/config/images_config.yaml (``key=value`` overrides in Hydra's syntax);
``train=false checkpoint=<file>`` renders a saved model instead
(/root/reference/examples/implicit_images.py:31-87).  Without ``images=`` a synthetic test card is used.
"""
import logging
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import hoir_b200 as H  # noqa: E402

logging.basicConfig(level=logging.INFO, format="%(message)s")
logger = logging.getLogger("implicit_images")


def synthetic_card(rows: int = 512, cols: int = 512) -> torch.Tensor:
    u = torch.linspace(0, 1, cols).view(1, cols, 1)
    v = torch.linspace(0, 1, rows).view(rows, 1, 1)
    phase = torch.tensor([0.0, 2.1, 4.2]).view(1, 1, 3)
    img = 0.5 + 0.5 * torch.sin(2 * torch.pi * (3 * u + 2 * v) + phase) * torch.cos(2 * torch.pi * 4 * u * v)
    return (img * 255).round().clamp(0, 255).to(torch.uint8)


def save_image(img01: torch.Tensor, path: str) -> None:
    from PIL import Image
    arr = (img01.clamp(0, 1) * 255).round().to(torch.uint8).cpu().numpy()
    Image.fromarray(arr).save(path)


def main(argv):
    cfg = H.images_config(mlp=dict(layer_type="continuous", n=3, n_in=None, n_out=None, n_hidden=None,
                                   periodicity=None, rescale_output=False, scale=2.0,
                                   input=dict(segments=100, width=4), output=dict(segments=2, width=3),
                                   hidden=dict(segments=2, layers=1, width=40)),
                          images=[], batch_size=1 << 16, max_epochs=10, rotations=2, out_dir="outputs",
                          sampling="shuffle")
    overrides = [a for a in argv if "=" in a and not a.startswith("images=")]
    for a in argv:
        if a.startswith("images="):
            cfg.images = [p.strip() for p in a[len("images="):].strip("[]").split(",") if p.strip()]
    H.apply_overrides(cfg, overrides)
    cfg.mlp.input.width = 2 * int(cfg.rotations)
    os.makedirs(cfg.out_dir, exist_ok=True)
    image = cfg.images[0] if cfg.images else synthetic_card()

    if cfg.train:
        trainer = H.ImageTrainer(cfg, image, sampling=cfg.sampling)
        logger.info("network: %d parameters, fused kernel: %s", sum(p.numel() for p in trainer.net.parameters()),
                    trainer.fused is not None)

        def on_epoch_end(tr, epoch, loss):
            logger.info("epoch %3d  train_loss %.6f  psnr %.2f dB", epoch, loss, tr.psnr())

        trainer.fit(on_epoch_end=on_epoch_end)
        ckpt = os.path.join(cfg.out_dir, "last.ckpt")
        trainer.save_checkpoint(ckpt)
        save_image(trainer.render(), os.path.join(cfg.out_dir, "fit.png"))
        logger.info("checkpoint %s", ckpt)
    else:
        model = H.Net.load_from_checkpoint(cfg.checkpoint, map_location="cuda")
        src = H.datasets._load_image_u8(image)
        out = H.render_image(model, src.shape[0], src.shape[1], int(cfg.rotations))
        save_image(out, os.path.join(cfg.out_dir, "render.png"))
        logger.info("rendered %s", os.path.join(cfg.out_dir, "render.png"))


if __name__ == "__main__":
    main(sys.argv[1:])
