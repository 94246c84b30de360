"""Parameter containers: host mirror of /root/reference/src/BEM/property.jl:10-48 (fields and asserts).

BSON persistence (property.jl:78-107) is a host-side convenience outside the hot path and is not
reproduced; the structs only carry the arrays the RHS kernels read.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List

import numpy as np

from . import _lib


class StateEvolutionLaw:
    pass


class DieterichStateLaw(StateEvolutionLaw):      # property.jl:4-5
    pass


@dataclass
class RateStateQuasiDynamicProperty:
    """property.jl:10-25.  a, b, L, σ are [nx, nξ] arrays; η, vpl, f₀, v₀ scalars."""
    a: np.ndarray
    b: np.ndarray
    L: np.ndarray
    sigma: np.ndarray
    eta: float
    vpl: float
    f0: float = 0.6
    v0: float = 1e-6
    _keep: list = field(default_factory=list, repr=False, compare=False)

    def __post_init__(self):
        assert np.shape(self.a) == np.shape(self.b) == np.shape(self.L) == np.shape(self.sigma)
        assert self.f0 > 0
        assert self.v0 > 0
        assert self.eta > 0
        assert self.vpl > 0

    def c_struct(self) -> _lib.OqFaultProperty:
        arrs = [_lib.f64(x) for x in (self.a, self.b, self.L, self.sigma)]
        self._keep = arrs
        return _lib.OqFaultProperty(*[_lib.dptr(x) for x in arrs], float(self.eta), float(self.vpl),
                                    float(self.f0), float(self.v0))


@dataclass
class DilatancyProperty:
    """property.jl:27-32."""
    tp: np.ndarray
    eps: np.ndarray
    beta: np.ndarray
    p0: np.ndarray
    _keep: list = field(default_factory=list, repr=False, compare=False)

    def c_struct(self) -> _lib.OqDilatancyProperty:
        arrs = [_lib.f64(x) for x in (self.tp, self.eps, self.beta, self.p0)]
        self._keep = arrs
        return _lib.OqDilatancyProperty(*[_lib.dptr(x) for x in arrs])


class ViscosityProperty:
    pass


@dataclass
class PowerLawViscosityProperty(ViscosityProperty):
    """property.jl:34-41.  n holds `power - 1`."""
    gamma: np.ndarray
    n: np.ndarray
    deps0: np.ndarray
    _keep: list = field(default_factory=list, repr=False, compare=False)

    def __post_init__(self):
        assert len(self.deps0) == 6
        assert len(self.gamma) == len(self.n)

    def laws(self):
        return [self]

    def c_struct(self) -> _lib.OqMantleProperty:
        return _mantle_struct(self, self.laws(), self.deps0)


@dataclass
class CompositePowerLawViscosityProperty(ViscosityProperty):
    """property.jl:43-48: the strain rate is the sum over `piter` (equation.jl:286-292)."""
    piter: List[PowerLawViscosityProperty]
    deps0: np.ndarray
    _keep: list = field(default_factory=list, repr=False, compare=False)

    def __post_init__(self):
        assert len(self.deps0) == 6

    def laws(self):
        return list(self.piter)

    def c_struct(self) -> _lib.OqMantleProperty:
        return _mantle_struct(self, self.laws(), self.deps0)


def _mantle_struct(owner, laws, deps0):
    g = _lib.f64(np.stack([np.asarray(p.gamma, dtype=np.float64).reshape(-1) for p in laws]), order="C")
    n = _lib.f64(np.stack([np.asarray(p.n, dtype=np.float64).reshape(-1) for p in laws]), order="C")
    d = _lib.f64(np.asarray(deps0, dtype=np.float64))
    owner._keep = [g, n, d]
    return _lib.OqMantleProperty(len(laws), _lib.dptr(g), _lib.dptr(n), _lib.dptr(d))
