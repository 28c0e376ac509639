"""Generates tests/golden/reference_own_v1.npz by RUNNING THE REFERENCE'S OWN CODE (imported from
/root/reference, this container only) for the parts of the path that live in the reference itself rather than
in the missing dependency:

  * image_neighborhood_dataset      (single_image_dataset.py:83-141)   patch gather, element order
  * neighborhood_sample_generator   (rendering.py:28-95)               reflect pad -> patches -> model -> fold -> crop
  * generate_sequence               (rendering.py:98-126)              the alpha-blended iteration
  * image_to_dataset                (single_image_dataset.py:22-51)    uint8 -> u*2/255-1 target scaling
  * Net.eval_step's loss expression (networks.py:64-70)                MSELoss on flattened tensors

Third-party modules that are absent here and irrelevant to these functions (matplotlib, pytorch_lightning,
sentence_transformers, torch_optimizer, omegaconf, hydra) are replaced by empty stubs before the import;
`high_order_layers_torch` resolves to the alias package of this repository (only positions_from_mesh is
touched, by image_to_dataset -- the positions are therefore NOT a pin, only the targets are).
The "model" is a fixed deterministic torch module defined here, so the vectors pin the reference's data
movement (gather order, padding, fold, crop, blend), bit for bit.

    python tests/golden/make_reference_golden.py          # needs /root/reference
"""
import os
import sys
import types

import numpy as np
import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden", "reference_own_v1.npz")


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install_stubs():
    def imread(path):
        from PIL import Image
        return np.array(Image.open(path).convert("RGB"))
    mpl = _stub("matplotlib")
    mpl.image = _stub("matplotlib.image", imread=imread)
    mpl.pyplot = _stub("matplotlib.pyplot")
    pl = _stub("pytorch_lightning", LightningDataModule=object, LightningModule=nn.Module, Trainer=object)
    pl.utilities = _stub("pytorch_lightning.utilities")
    pl.utilities.rank_zero = _stub("pytorch_lightning.utilities.rank_zero", rank_zero_only=lambda f: f)
    pl.callbacks = _stub("pytorch_lightning.callbacks", Callback=object, LearningRateMonitor=object)
    _stub("sentence_transformers", SentenceTransformer=object)
    _stub("torch_optimizer")
    _stub("omegaconf", DictConfig=dict, OmegaConf=object)
    _stub("hydra")
    _stub("lion_pytorch", Lion=object)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "compat"))   # high_order_layers_torch -> alias package of this repository
    sys.path.insert(0, REF)


class FixedModel(nn.Module):
    """Deterministic stand-in for the network: features [P, F] -> [P, C*w*w], a fixed banded linear map + tanh."""

    def __init__(self, f_in: int, f_out: int):
        super().__init__()
        i = torch.arange(f_out, dtype=torch.float32).view(-1, 1)
        j = torch.arange(f_in, dtype=torch.float32).view(1, -1)
        self.register_buffer("w", torch.cos(0.37 * i + 0.11 * j) / f_in ** 0.5)
        self.device = "cpu"

    def forward(self, x):
        return torch.tanh(x @ self.w.t())


def main():
    install_stubs()
    from high_order_implicit_representation import rendering as R
    from high_order_implicit_representation import single_image_dataset as D

    out = {}
    g = torch.Generator().manual_seed(20261020)
    # ---- image_neighborhood_dataset: (H, W, width, outside, stride)
    for k, (h, w, width, outside, stride) in enumerate([(12, 10, 3, 1, 1), (9, 14, 3, 3, 2), (16, 16, 1, 2, 3), (11, 11, 2, 1, 1)]):
        img = torch.rand(3, h, w, generator=g) * 2 - 1
        feat, targ, _, lastx, lasty = D.image_neighborhood_dataset(img, width=width, outside=outside, stride=stride)
        out[f"nbr{k}_geom"] = np.array([h, w, width, outside, stride, lastx, lasty])
        out[f"nbr{k}_image"] = img.numpy()
        out[f"nbr{k}_features"] = feat.numpy()
        out[f"nbr{k}_targets"] = targ.numpy()
    # ---- neighborhood_sample_generator / generate_sequence with the fixed model
    for k, (h, w, width, outside) in enumerate([(12, 15, 3, 3), (10, 10, 2, 1), (7, 9, 3, 1)]):
        img = torch.rand(3, h, w, generator=g) * 2 - 1
        total = width + 2 * outside
        model = FixedModel(3 * (total * total - width * width), 3 * width * width)
        res = R.neighborhood_sample_generator(model, img, width=width, outside=outside, batch_size=7)
        out[f"gen{k}_geom"] = np.array([h, w, width, outside])
        out[f"gen{k}_image"] = img.numpy()
        out[f"gen{k}_result"] = res.numpy()
        if k == 0:
            seq = R.generate_sequence(model, img, frames=3, width=width, outside=outside, batch_size=50, skip=1, alpha=0.75)
            out["seq0_frames"] = seq.numpy()
    # ---- image_to_dataset on a real file of the reference (targets only; positions come from the dependency)
    path = os.path.join(REF, "images", "jupiter.jpg")
    flat, pos, img = D.image_to_dataset(path, rotations=2)
    rows = torch.tensor([0, 1, 591, 12345, flat.shape[0] // 2, flat.shape[0] - 1])
    out["jupiter_shape"] = np.array(list(img.shape))
    out["jupiter_rows"] = rows.numpy()
    out["jupiter_u8"] = img.reshape(-1, 3)[rows].numpy()
    out["jupiter_flat"] = flat[rows].numpy()
    out["jupiter_pos_shape"] = np.array(list(pos.shape))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
