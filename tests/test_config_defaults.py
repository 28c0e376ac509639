"""hoir_b200.config.*_config() carry the reference's YAML values verbatim (runs without overrides must be
interchangeable).  Fixture: tests/golden/reference_configs.json, parsed from /root/reference/config by
tests/golden/make_reference_configs.py."""
import json
import os

import pytest

import hoir_b200 as H
from hoir_b200.config import cfg_to_dict

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "reference_configs.json")))
PAIRS = [("images_config", H.images_config), ("neighborhood_config", H.neighborhood_config),
         ("random_interpolation", H.random_interpolation_config), ("generative_config", H.generative_config)]


def _flat(d, prefix=""):
    out = {}
    for k, v in d.items():
        if isinstance(v, dict):
            out.update(_flat(v, f"{prefix}{k}."))
        else:
            out[f"{prefix}{k}"] = v
    return out


@pytest.mark.parametrize("name,fn", PAIRS)
def test_config_defaults_equal_the_reference_yaml(name, fn):
    want = _flat(GOLD[name])
    got = _flat(cfg_to_dict(fn()))
    for key, val in want.items():
        if isinstance(val, str) and val.startswith("${"):      # OmegaConf interpolation (segments: ${image_size})
            ref_key = val[2:-1]
            assert got[key] == got[ref_key], key
            continue
        assert key in got, f"{name}: missing {key}"
        if isinstance(val, float) or isinstance(got[key], float):
            assert float(got[key]) == float(val), (name, key, got[key], val)
        else:
            assert got[key] == val, (name, key, got[key], val)
    assert set(got) == set(want), (name, sorted(set(got) ^ set(want)))


def test_fixture_is_current_when_the_reference_is_present():
    if not os.path.isdir("/root/reference/config"):
        pytest.skip("reference tree not on this machine")
    import importlib.util
    spec = importlib.util.spec_from_file_location("mk", os.path.join(HERE, "golden", "make_reference_configs.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    for name, _ in PAIRS:
        assert mk.load(f"{name}.yaml") == GOLD[name]
