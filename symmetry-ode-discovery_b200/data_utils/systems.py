"""The four ODE systems of the reference as members of the SINDy library (f(x) = Θ(x)·Ξᵀ).

Coefficients follow the reference's right-hand sides (`data_utils/damped_oscillator.py:20-24`,
`growth.py:18-22`, `lotka.py:33-41` canonical form, `selkov.py:18-22`) and coincide with the truth tables of
`evaluation/eval_eq.py:88-105`.
"""
from __future__ import annotations

import numpy as np

from sindy_b200.native import Library
from .ode import LibraryODE


def dosc(a=0.1):
    lib = Library(2, 2)
    Xi = np.zeros((2, 6))
    Xi[0, 1], Xi[0, 2] = -a, -1.0
    Xi[1, 1], Xi[1, 2] = 1.0, -a
    return LibraryODE(lib, Xi, "dosc")


def growth():
    lib = Library(2, 2)
    Xi = np.zeros((2, 6))
    Xi[0, 1], Xi[0, 5] = -0.3, 0.1
    Xi[1, 2] = 1.0
    return LibraryODE(lib, Xi, "growth")


def lotka_volterra(a=2 / 3, b=4 / 3, c=1.0, d=1.0):
    """Canonical (log) coordinates: dx0 = a − b·exp(x1), dx1 = c·exp(x0) − d."""
    lib = Library(2, 2, include_exp=True)
    Xi = np.zeros((2, 8))
    Xi[0, 0], Xi[0, 7] = a, -b
    Xi[1, 0], Xi[1, 6] = -d, c
    return LibraryODE(lib, Xi, "lv")


def selkov(a=0.75, b=0.1, c=0.1):
    """dx0 = a − b·x0 − x0·x1², dx1 = c·x0 − x1 + x0·x1² (cubic library, column 8 = x0·x1·x1)."""
    lib = Library(2, 3)
    Xi = np.zeros((2, 10))
    Xi[0, 0], Xi[0, 1], Xi[0, 8] = a, -b, -1.0
    Xi[1, 1], Xi[1, 2], Xi[1, 8] = c, -1.0, 1.0
    return LibraryODE(lib, Xi, "selkov")


SYSTEMS = {"dosc": dosc, "growth": growth, "lv": lotka_volterra, "selkov": selkov}
