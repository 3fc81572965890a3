"""Overlay of the reference's `data_utils` package.

This directory shadows the reference's package of the same name when it comes first on `sys.path`. It replaces only
`ode.py` (batched RK4 -> CUDA rollout) and `smoothing.py` (GP smoother -> device Cholesky); the reference's generators
`damped_oscillator`, `growth`, `lotka`, `selkov` (IC samplers, right-hand sides, CLI — `dataset.py:12` imports
`data_utils.lotka`) must keep resolving, so the reference's own `data_utils` directory — found on `sys.path` or through
$SINDY_B200_REFERENCE — is appended to this package's `__path__`: submodules present here win, everything else falls
through to the reference. Their `from data_utils.ode import *` then binds to this repo's `solve_ode_batch` / `gen_data`.
"""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))


def _reference_data_utils():
    cands = [os.environ.get("SINDY_B200_REFERENCE")] + list(sys.path)
    for base in cands:
        if not base:
            continue
        d = os.path.join(os.path.abspath(base), "data_utils")
        if d != _here and os.path.isfile(os.path.join(d, "lotka.py")):
            return d
    return None


_ref = _reference_data_utils()
if _ref is not None and _ref not in __path__:
    __path__.append(_ref)
