"""oracle/okada_numeric.py -- numerical oracle for the rectangular dislocation (`dc3d`).  TEST INFRASTRUCTURE ONLY.

`dc3d` lives in the un-vendored GeoGreensFunctions.jl; oracle/okada.c restates Okada's (1992) closed form.  This
file evaluates the same field from its DEFINITION, independently of Okada's tables: Volterra's formula over the
fault rectangle with the half-space (Mindlin) Green's tensor G of oracle/hex8_numeric.py,

    u_m(x) = integral over the fault of  b_i n_j c_ijkl  dG_mk(x, xi)/dxi_l  dS(xi),
    c_ijkl = lam d_ij d_kl + mu (d_ik d_jl + d_il d_jk)        (isotropic),

b = dislocation vector, n = fault normal pointing to the hanging wall.  The source derivative is a complex-step
derivative of G (exact to round-off), the surface integral a tensor Gauss-Legendre rule (the integrand is smooth
for receivers off the fault plane).  Geometry as the reference calls dc3d (src/BEM/GF.jl:49-53): fault plane through
(0, 0, -depth), strike along +x, dip measured from the horizontal; a fault point is
(s, w cos(dip), -depth + w sin(dip)) with s in [al1, al2] along strike and w in [aw1, aw2] up-dip; slip components
(d1, d2, d3) = (strike, up-dip, opening) of the hanging wall relative to the foot wall.
"""
from __future__ import annotations

import numpy as np

from .hex8_numeric import mindlin


def dc3d_displacement(alpha, x, y, z, depth, dip, al1, al2, aw1, aw2, d1, d2, d3, nquad=64):
    """(ux, uy, uz) at (x, y, z <= 0) of the rectangular dislocation, by quadrature.  alpha = (lam+mu)/(lam+2mu)."""
    mu = 1.0
    lam = mu * (2 * alpha - 1) / (1 - alpha)         # alpha = (lam + mu)/(lam + 2 mu)
    sd, cd = np.sin(np.radians(dip)), np.cos(np.radians(dip))
    strike = np.array([1.0, 0.0, 0.0])
    updip = np.array([0.0, cd, sd])
    normal = np.array([0.0, -sd, cd])
    b = d1 * strike + d2 * updip + d3 * normal
    # b_i n_j c_ijkl = lam (b.n) d_kl + mu (b_k n_l + b_l n_k)
    M = lam * float(b @ normal) * np.eye(3) + mu * (np.outer(b, normal) + np.outer(normal, b))
    gp, gw = np.polynomial.legendre.leggauss(nquad)
    s = (al1 + al2) / 2 + gp * (al2 - al1) / 2
    w = (aw1 + aw2) / 2 + gp * (aw2 - aw1) / 2
    S, Wd = np.meshgrid(s, w, indexing="ij")
    Wt = np.outer(gw * (al2 - al1) / 2, gw * (aw2 - aw1) / 2)
    xi = np.stack([S, Wd * cd, -depth + Wd * sd], axis=-1)
    xr = np.array([x, y, z], dtype=float)
    h = 1e-30
    u = np.zeros(3)
    for l in range(3):
        xic = xi.astype(complex)
        xic[..., l] += 1j * h
        dG = np.imag(mindlin(xr.astype(complex), xic, lam, mu)) / h        # dG[..., m, k] / dxi_l
        u += np.einsum("ab,abmk,k->m", Wt, dG, M[:, l])
    return u


def dc3d_gradient(alpha, x, y, z, depth, dip, al1, al2, aw1, aw2, d1, d2, d3, h=None, nquad=64):
    """The nine displacement gradients [d/dx (ux,uy,uz), d/dy (...), d/dz (...)] (dc3d's entries 4..12) of the
    quadrature field above, by SIXTH-ORDER central differences in the receiver coordinate (the field itself is exact
    to round-off, ~1e-15, so a step of a few 1e-3 of the distance to the fault leaves both the truncation error
    h^6 f^(7)/140 and the round-off error eps/h below 1e-11 of the gradient).  Receivers must lie at z <= -3h."""
    if h is None:
        h = 5e-3
    w = np.array([-1.0, 9.0, -45.0, 0.0, 45.0, -9.0, 1.0]) / 60.0
    g = np.zeros(9)
    for ax in range(3):
        acc = np.zeros(3)
        for k, wk in zip(range(-3, 4), w):
            if wk == 0.0:
                continue
            p = [x, y, z]
            p[ax] += k * h
            acc += wk * dc3d_displacement(alpha, p[0], p[1], p[2], depth, dip, al1, al2, aw1, aw2, d1, d2, d3, nquad)
        g[3 * ax: 3 * ax + 3] = acc / h
    return g
