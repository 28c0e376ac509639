"""Generates tests/golden/reference_radial_v1.npz by RUNNING THE REFERENCE'S OWN radial sampler (imported from
/root/reference, this container only; stubs as in make_reference_golden.py):

  * random_symmetric_sample            (random_sample_dataset.py:211-244)
  * random_radial_samples_from_image   (random_sample_dataset.py:176-208)
  * generate_sample_radial             (utils.py:20-97)       with the fixed stand-in model

The reference draws its uniforms with torch.rand from the global CPU generator: r = torch.rand(N, F) first, then
theta = torch.rand(N, F) * 2 * math.pi.  The script seeds the generator, runs the reference, then re-seeds and
draws the same (r, theta) to store them next to the outputs -- so kernels fed with these uniforms must reproduce the
reference's indices / features bit for bit.  Stripe planes come from positions_from_mesh of the alias package
(the dependency is absent): they are stored as INPUT, not as a pin.

    python tests/golden/make_reference_radial_golden.py          # needs /root/reference
"""
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_reference_golden import FixedModel, ROOT, install_stubs  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "reference_radial_v1.npz")


def draw(seed, n, f, times=1):
    torch.manual_seed(seed)
    out = []
    for _ in range(times):
        r = torch.rand(n, f)
        theta = torch.rand(n, f) * 2 * math.pi
        out.append((r, theta))
    return out


def main():
    install_stubs()
    from high_order_implicit_representation import random_sample_dataset as D
    from high_order_implicit_representation import utils as U
    from high_order_layers_torch.utils import positions_from_mesh
    sys.path.insert(0, ROOT)
    from oracle import hol_oracle as O

    out = {}
    g = torch.Generator().manual_seed(777)
    # ---- random_symmetric_sample: the reference test's five corner/centre points, and random points
    cases = [(11, 5, torch.tensor([[5, 5], [0, 0], [10, 0], [0, 10], [10, 10]])),
             (64, 32, torch.randint(0, 64, (300, 2), generator=g)),
             (7, 3, torch.randint(0, 7, (40, 2), generator=g))]
    for k, (size, f, samples) in enumerate(cases):
        seed = 100 + k
        torch.manual_seed(seed)
        xy = D.random_symmetric_sample(image_size=size, interp_size=f, samples=samples)
        (r, theta), = draw(seed, samples.shape[0], f)
        assert torch.equal(O.random_symmetric_sample(size, samples, r, theta), xy)      # the stream capture is right
        out[f"sym{k}_geom"] = np.array([size, f])
        out[f"sym{k}_samples"] = samples.numpy().astype(np.int32)
        out[f"sym{k}_r"], out[f"sym{k}_theta"] = r.numpy(), theta.numpy()
        out[f"sym{k}_xy"] = xy.numpy().astype(np.int32)
    # ---- random_radial_samples_from_image: (S, F, rotations)
    for k, (size, f, rot) in enumerate([(16, 6, 2), (32, 25, 1), (9, 4, 3)]):
        img = torch.rand(3, size, size, generator=g) * 2 - 1
        stripes = torch.cat([v.unsqueeze(0) for v in positions_from_mesh(width=size, height=size, device="cpu",
                                                                         rotations=rot, normalize=True)])
        indices, lin = D.indices_from_grid(size, device="cpu")
        assert torch.equal(indices, O.indices_from_grid(size)[0]) and torch.equal(lin, O.indices_from_grid(size)[1])
        seed = 200 + k
        torch.manual_seed(seed)
        feat, targ = D.random_radial_samples_from_image(img=img, stripe_list=stripes, image_size=size, feature_pixels=f,
                                                        indices=indices, target_linear_indices=lin)
        (r, theta), = draw(seed, size * size, f)
        of, ot = O.random_radial_samples_from_image(img, stripes, r, theta)
        assert torch.equal(of, feat) and torch.equal(ot, targ)
        out[f"rad{k}_geom"] = np.array([size, f, rot])
        out[f"rad{k}_image"], out[f"rad{k}_stripes"] = img.numpy(), stripes.numpy()
        out[f"rad{k}_r"], out[f"rad{k}_theta"] = r.numpy(), theta.numpy()
        out[f"rad{k}_features"], out[f"rad{k}_targets"] = feat.numpy(), targ.numpy()
    # ---- generate_sample_radial: iterate on the previous estimate (all_random=False) from a start image
    size, f, rot, iters = 12, 5, 2, 3
    start = torch.rand(3, size, size, generator=g) * 2 - 1
    model = FixedModel(f * (3 + 2 * rot), 3)
    seed = 300
    torch.manual_seed(seed)
    with torch.no_grad():
        frames = U.generate_sample_radial(model=model, features=f, iterations=iters, start_image=start, image_size=size,
                                          rotations=rot, all_random=False)
    uni = draw(seed, size * size, f, times=iters)
    out["gen_geom"] = np.array([size, f, rot, iters])
    out["gen_start"] = start.numpy()
    out["gen_r"] = torch.stack([u[0] for u in uni]).numpy()
    out["gen_theta"] = torch.stack([u[1] for u in uni]).numpy()
    out["gen_frames"] = torch.stack(frames).numpy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
