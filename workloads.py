"""Synthetic workloads of BASELINE.json's configs, as plain NumPy inputs.

Shared by tests/, bench.py and __graft_entry__.smoke(); depends on neither the product package nor the
oracle.  Geometry and parameters follow /root/reference/examples/otf-with-mantle.jl (file:line cited
inline); only array shapes change between configs.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

LAM = MU = 3e10                                   # examples/otf-with-mantle.jl:36
CS = 3044.14                                      # :75
VPL = 140e-3 / 365 / 86400                        # :76
V0, F0 = 1e-6, 0.6                                # :77-78
ETA = MU / (2 * CS)                               # :80
AVW, ABVW, DC, SIGMAX = 0.015, 0.0047, 8e-3, 5e7  # :82-85
DEPS0 = np.array([0.0, -1e-12, 0.0, 0.0, 0.0, 0.0])   # :117
YEAR = 365 * 86400.0


@dataclass
class FaultSpec:
    """Arguments of gen_mesh(Val(:RectOkada), x, ξ, Δx, Δξ, dip)."""
    x: float
    xi: float
    dx: float
    dxi: float
    dip: float = 90.0

    @property
    def nx(self):
        return int(round(self.x / self.dx))

    @property
    def nxi(self):
        return int(round(self.xi / self.dxi))


@dataclass
class BoxSpec:
    """Arguments of the structured hex8 box (gen_gmsh_mesh's llx,lly,llz,dx,dy,dz,nx,ny,nz,rfzh)."""
    llx: float
    lly: float
    llz: float
    dx: float
    dy: float
    dz: float
    nx: int
    ny: int
    nz: int
    rfzh: tuple = ()

    def args(self):
        return (self.llx, self.lly, self.llz, self.dx, self.dy, self.dz, self.nx, self.ny, self.nz,
                np.array(self.rfzh) if len(self.rfzh) else None)

    @property
    def n(self):
        return self.nx * self.ny * self.nz


# BASELINE.json configs ------------------------------------------------------------------------------
C1_FAULT = FaultSpec(16e3, 8e3, 500.0, 500.0)                        # 32 x 16, fault only (test scale)
C2_FAULT = FaultSpec(80e3, 8e3, 10e3, 2e3)                           # examples/otf-with-mantle.jl:18
C2_BOX = BoxSpec(-40e3, -2.5e3, -8e3, 80e3, 5e3, -22e3, 4, 3, 3,     # :25-28
                 tuple(np.cumprod(np.ones(3) * 1.5)))
C3_FAULT = FaultSpec(64e3, 16e3, 250.0, 250.0)                       # 256 x 64 = 16 384 cells


def box_for(nx, ny, nz, fault: FaultSpec = C2_FAULT) -> BoxSpec:
    """A mantle box under `fault` with nx*ny*nz cells (layer heights growing 1.5x as in the example)."""
    half = fault.x / 2
    return BoxSpec(-half, -2.5e3 * max(1, ny // 3), -fault.xi, fault.x, 5e3 * max(1, ny // 3), -22e3,
                   nx, ny, nz, tuple(np.cumprod(np.ones(nz) * 1.2)))


def fault_properties(x, z, nx, nxi):
    """a, b, L, σ of examples/otf-with-mantle.jl:86-94 on centroids x[nx], z[nxi] -> arrays [nx, nxi].
    Two velocity-weakening patches: |x| in [L/16, 5L/16] scaled to the fault length, depth 1-6 km
    (scaled to the fault width)."""
    a = np.full((nx, nxi), AVW)
    b = np.full((nx, nxi), AVW - ABVW)
    L = np.full((nx, nxi), DC)
    sig = np.minimum(SIGMAX, 1.5e6 + 18.0e3 * (-z))                  # :89
    sigma = np.repeat(sig[None, :], nx, axis=0)
    xl = (x.max() - x.min() + (x[1] - x[0] if nx > 1 else 0.0))
    s = xl / 80e3
    left = (-25e3 * s <= x) & (x <= -5e3 * s)
    right = (5e3 * s <= x) & (x <= 25e3 * s)
    w = (-z).max() + (z[0] - z[1] if nxi > 1 else 0.0) / 2 if nxi > 1 else 8e3
    sv = w / 8e3
    vert = (-6e3 * sv <= z) & (z <= -1e3 * sv)
    b[np.ix_(left ^ right, vert)] = AVW + ABVW                       # :94
    return a, b, L, sigma


def mantle_properties(cz):
    """Wet-dislocation power law of examples/otf-with-mantle.jl:100-120 -> (γ[ne], n-1 [ne], dϵ0[6])."""
    A, Q, V, r, n = 3e1, 480e3, 11e-6, 1.2, 3.5
    coh, R, crust, kappa = 1000.0, 8.314, 7e3, 8e-7
    z = -np.asarray(cz)
    T = 1673.0 * np.array([math.erf(v / math.sqrt(4 * kappa * 1e6 * YEAR)) for v in z])
    P = 2800 * 9.8 * crust + 3300 * 9.8 * (z - crust)
    gamma = A / (1e6) ** n * coh ** r * np.exp(-(Q + P * V) / R / T)
    return gamma, np.full(len(z), n - 1.0), DEPS0.copy()


def initial_state(nx, nxi, L, cz=None, gamma=None, npow=None, rng=None):
    """Initial conditions of examples/otf-with-mantle.jl:132-149: v = vpl, θ = L/v divided by 1.1 (left
    half) / 2.5 (right half), δ = 0; mantle ϵ = 0, σ lithostatic with σxy balancing dϵ0 (closed form of
    the example's 1-D optimisation).  rng: optional extra ±10 % perturbation of θ."""
    v = np.full((nx, nxi), VPL)
    theta = L / v
    theta[: nx >> 1, :] /= 1.1
    theta[nx >> 1:, :] /= 2.5
    if rng is not None:
        theta = theta * (1 + 0.1 * rng.uniform(-1, 1, size=theta.shape))
    delta = np.zeros((nx, nxi))
    if cz is None:
        return v, theta, delta
    ne = len(cz)
    crust = 7e3
    P = 2800 * 9.8 * crust + 3300 * 9.8 * (-np.asarray(cz) - crust)     # :140
    sigma = np.repeat(P[:, None], 6, axis=1)
    sigma[:, 2] = 0.0
    sigma[:, 4] = 0.0
    # γ (√2 x)^n x = |dϵ0_xy|  ->  x = (|dϵ0| / (γ 2^{n/2}))^{1/(n+1)}          :145-147
    sxy = -(np.abs(DEPS0[1]) / (gamma * 2 ** (npow / 2))) ** (1 / (npow + 1))
    sigma[:, 1] = sxy
    eps = np.zeros((ne, 6))
    return v, theta, eps, sigma, delta
