"""sindy_b200 — host-side binding of libsindy_b200.so (hand-written CUDA for sm_100a).

`native` is the ctypes layer over the C ABI (include/sindy_b200.h), `ops` the differentiable operators the
drop-in modules (`sindy.py`, `model_utils.py`, `data_utils/ode.py` one directory up) are written with, `dist`
the sample-sharded multi-GPU step.
"""
from . import native, ops  # noqa: F401
from .native import Library  # noqa: F401

__version__ = "0.1.0"
