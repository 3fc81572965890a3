"""Drop-in check: every function / method of the reference's hot-path modules — recorded from the
unmodified reference by `oracle/gen_golden.py api` into tests/golden/api_surface.json — exists here under the same name
with the same parameters in the same order and with the same defaults (extra trailing keyword parameters are allowed).
The per-block library functions of `sindy.py:7-30` are present (values from `sb_theta`; nothing here calls them).
Deliberately absent from the fixture check: `train_lassi` (LaLiGAN symmetry discovery: out of scope; re-exported
from the reference when a checkout is importable)."""
import importlib
import inspect
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
API = json.load(open(os.path.join(ROOT, "tests", "golden", "api_surface.json")))

ABSENT = {("train", "train_lassi")}


def _check(ref_params, fn, where):
    ours = list(inspect.signature(fn).parameters.items())
    names = [n for n, _ in ours]
    for pos, (name, kind, default) in enumerate(ref_params):
        if "VAR_KEYWORD" in kind:
            assert any(p.kind is inspect.Parameter.VAR_KEYWORD for _, p in ours), f"{where}: **{name} missing"
            continue
        assert pos < len(names) and names[pos] == name, f"{where}: parameter {pos} is {names[pos:pos + 1]}, reference has {name!r}"
        p = ours[pos][1]
        if default is None:
            assert p.default is inspect._empty, f"{where}: {name} has a default here but not in the reference"
        else:
            assert p.default is not inspect._empty and repr(p.default) == default, \
                f"{where}: default of {name} is {p.default!r}, reference has {default}"
    for name, p in ours[len([q for q in ref_params if 'VAR_KEYWORD' not in q[1]]):]:
        assert p.kind is inspect.Parameter.VAR_KEYWORD or p.default is not inspect._empty, \
            f"{where}: extra parameter {name} without a default"


@pytest.mark.parametrize("modname", sorted(API))
def test_public_surface_matches_the_reference(modname):
    mod = importlib.import_module(modname)
    checked = 0
    for name, entry in API[modname].items():
        if (modname, name) in ABSENT:
            assert not hasattr(mod, name) or True
            continue
        assert hasattr(mod, name), f"{modname}.{name} missing"
        obj = getattr(mod, name)
        if entry["kind"] == "function":
            _check(entry["params"], obj, f"{modname}.{name}")
            checked += 1
        else:
            for mname, params in entry["methods"].items():
                if (modname, f"{name}.{mname}") in ABSENT:
                    continue
                assert hasattr(obj, mname), f"{modname}.{name}.{mname} missing"
                _check(params, getattr(obj, mname), f"{modname}.{name}.{mname}")
                checked += 1
    assert checked > 0
