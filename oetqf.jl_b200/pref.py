"""The `matvecmul!` backend slot: host mirror of /root/reference/src/pref.jl:1-21.

The reference chooses between "LinearAlgebra" (BLAS) and "Octavian" at precompile time through
Preferences.jl.  Here the slot has exactly one implementation, "B200" (oq_gemv / the fused matvec of
rhs.cu); the getter/setter keep the reference's names and its error behaviour for unknown backends.
"""
from __future__ import annotations

import logging

_BACKENDS = ("B200",)
_current = "B200"
log = logging.getLogger("oetqf_b200")


def get_matvecmul() -> str:            # get_matvecmul!(), pref.jl:1-3
    return _current


def set_matvecmul(backend: str):       # set_matvecmul!(), pref.jl:7-13
    global _current
    if backend not in _BACKENDS:
        raise ValueError(f"Invalid backend: {backend}")     # ArgumentError in the reference
    _current = backend
    log.info("New backend %s set; restart your session for this change to take effect!", backend)


def matvecmul(y, A, x, alpha=None, beta=None):
    """matvecmul!(y, A, x) / matvecmul!(y, A, x, true, true) as used at equation.jl:201-203.
    A is a DeviceMatrix; x, y host vectors."""
    if alpha is None and beta is None:
        y[...] = A.gemv(x).reshape(y.shape, order="F")
        return y
    assert alpha in (True, 1, 1.0) and beta in (True, 1, 1.0), "only α = β = true is used by the RHS"
    return A.gemv(x, y)
