"""Writes tests/golden/reference_configs.json: the reference's Hydra config trees (config/*.yaml with the optimizer
group of their defaults list merged in), parsed here from /root/reference with PyYAML.  Run in the build container
(the reference does not travel to the GPU box); tests/test_config_defaults.py holds hoir_b200.config to it."""
import json
import os

import yaml

REF = "/root/reference/config"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_configs.json")


def load(name):
    with open(os.path.join(REF, name)) as f:
        cfg = yaml.safe_load(f)
    for d in cfg.pop("defaults", []):
        for group, choice in d.items():
            with open(os.path.join(REF, group, f"{choice}.yaml")) as f:
                cfg[group] = yaml.safe_load(f)
    return cfg


if __name__ == "__main__":
    out = {n: load(f"{n}.yaml") for n in ("images_config", "neighborhood_config", "random_interpolation",
                                          "generative_config")}
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", OUT)
