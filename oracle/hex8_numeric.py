"""oracle/hex8_numeric.py -- numerical oracle for the hex8 strain-volume stress kernel.  TEST INFRASTRUCTURE ONLY.

`stress_vol_hex8!` lives in the un-vendored GeoGreensFunctions.jl (Barbot et al. 2017); its source is not
available here, so the kernel is pinned by its DEFINITION, evaluated numerically:

    u_i(x)   = closed-surface integral over the cuboid of  G_ij(x, ξ) m_jk n_k dS(ξ),
    m        = λ tr(ε*) I + 2 μ ε*        (uniform eigenstrain ε* inside the cuboid)
    σ        = λ tr(e) I + 2 μ e,  e = sym(grad u) - ε* [x inside the cuboid]

with G the half-space (Mindlin) Green's tensor in the form of Okada (1992) eqs. 1-2, z up, free surface z = 0.
grad u uses the complex-step derivative of G (G is analytic away from the source), so the only error is the
Gauss-Legendre quadrature of smooth integrands (converges to ~1e-14 for receivers at cell centroids).
The closed form in oracle/hex8.c and the CUDA kernel are both checked against this file.
"""
from __future__ import annotations

import numpy as np


def mindlin(x, xi, lam, mu):
    """G[i, j]: displacement component i at x due to a unit point force in direction j at ξ (ξ3 < 0, x3 <= 0).
    Works on complex x (complex-step differentiation).  x: [3], xi: [..., 3] -> [..., 3, 3]."""
    alpha = (lam + mu) / (lam + 2 * mu)
    xi = np.asarray(xi)
    r = np.stack([x[0] - xi[..., 0], x[1] - xi[..., 1], x[2] - xi[..., 2]], axis=-1)
    Rv = np.stack([x[0] - xi[..., 0], x[1] - xi[..., 1], -x[2] - xi[..., 2]], axis=-1)
    eye = np.eye(3)

    def ua(v):
        n = np.sqrt(np.sum(v * v, axis=-1))[..., None, None]
        vv = v[..., :, None] * v[..., None, :]
        return ((2 - alpha) * eye / n + alpha * vv / n ** 3) / (8 * np.pi * mu)

    Rn = np.sqrt(np.sum(Rv * Rv, axis=-1))
    R3 = Rv[..., 2]
    n1 = Rn[..., None, None]
    RR = Rv[..., :, None] * Rv[..., None, :]
    d3 = np.zeros(3)
    d3[2] = 1.0
    # uB
    t = eye / n1 + RR / n1 ** 3
    w = (Rn + R3)[..., None, None]
    Ri = Rv[..., :, None]
    Rj = Rv[..., None, :]
    dj3 = d3[None, :]
    di3 = d3[:, None]
    extra = (eye / w + (Ri * dj3 - Rj * di3 * (1 - dj3)) / (n1 * w)
             - RR * (1 - di3) * (1 - dj3) / (n1 * w ** 2))
    ub = (t + (1 - alpha) / alpha * extra) / (4 * np.pi * mu)
    # uC
    xi3 = xi[..., 2][..., None, None]
    uc = (1 - 2 * di3) * ((2 - alpha) * (Ri * dj3 - Rj * di3) / n1 ** 3
                          + alpha * xi3 * (eye / n1 ** 3 - 3 * RR / n1 ** 5)) / (4 * np.pi * mu)
    return ua(r) - ua(Rv) + ub + x[2] * uc


def _moment(eps6, lam, mu):
    e = np.array([[eps6[0], eps6[1], eps6[2]], [eps6[1], eps6[3], eps6[4]], [eps6[2], eps6[4], eps6[5]]], dtype=float)
    return lam * np.trace(e) * np.eye(3) + 2 * mu * e, e


def stress_vol_hex8(x, y, z, qx, qy, qz, dx, dy, dz, eps6, mu, nu, nquad=48):
    """σ (xx,xy,xz,yy,yz,zz) at (x,y,z) for the cuboid x∈[qx-dx/2,qx+dx/2], y∈[qy,qy+dy], z∈[qz-dz,qz]
    (the convention of src/BEM/GF.jl:215-221 with θ = 0) carrying eigenstrain eps6 = (xx,xy,xz,yy,yz,zz)."""
    lam = 2 * mu * nu / (1 - 2 * nu)
    m, e0 = _moment(eps6, lam, mu)
    lo = np.array([qx - dx / 2, qy, qz - dz])
    hi = np.array([qx + dx / 2, qy + dy, qz])
    gp, gw = np.polynomial.legendre.leggauss(nquad)
    grad = np.zeros((3, 3))                      # grad[i, l] = d u_i / d x_l
    h = 1e-30
    xr = np.array([x, y, z], dtype=float)
    for k in range(3):                            # face normal direction
        a, b = [d for d in range(3) if d != k]
        pa = (lo[a] + hi[a]) / 2 + gp * (hi[a] - lo[a]) / 2
        pb = (lo[b] + hi[b]) / 2 + gp * (hi[b] - lo[b]) / 2
        wa = gw * (hi[a] - lo[a]) / 2
        wb = gw * (hi[b] - lo[b]) / 2
        A, B = np.meshgrid(pa, pb, indexing="ij")
        Wt = wa[:, None] * wb[None, :]
        for side, sgn in ((lo[k], -1.0), (hi[k], 1.0)):
            pts = np.zeros(A.shape + (3,))
            pts[..., a] = A
            pts[..., b] = B
            pts[..., k] = side
            tvec = sgn * m[:, k]                  # traction m_jk n_k
            for l in range(3):
                xc = xr.astype(complex)
                xc[l] += 1j * h
                G = mindlin(xc, pts, lam, mu)     # [..., i, j]
                dG = G.imag / h
                grad[:, l] += np.einsum("ab,abij,j->i", Wt, dG, tvec)
    e = (grad + grad.T) / 2
    inside = np.all(xr > lo) and np.all(xr < hi)
    if inside:
        e = e - e0
    s = lam * np.trace(e) * np.eye(3) + 2 * mu * e
    return np.array([s[0, 0], s[0, 1], s[0, 2], s[1, 1], s[1, 2], s[2, 2]])


def _self_check():
    lam, mu = 1.3, 0.9
    nu = lam / 2 / (lam + mu)
    eps = [0.3, 0.5, -0.2, -0.4, 0.7, 0.9]
    # box x∈[-1,1], y∈[0,2], z∈[-3,-1]
    for name, p in (("self", (0, 1, -2)), ("neighbour", (2, 1, -2)), ("far", (9, -7, -0.5))):
        s = stress_vol_hex8(*p, 0.0, 0.0, -1.0, 2.0, 2.0, 2.0, eps, mu, nu, nquad=48)
        print(name, " ".join(f"{v: .10e}" for v in s))
    # traction-free surface and equilibrium of the Green's tensor itself
    xi = np.array([0.3, -0.2, -1.7])
    x0 = np.array([1.1, 0.7, 0.0])
    h = 1e-30
    dG = np.zeros((3, 3, 3))
    for l in range(3):
        xc = x0.astype(complex)
        xc[l] += 1j * h
        dG[:, :, l] = mindlin(xc, xi, lam, mu).imag / h
    for j in range(3):
        g = dG[:, j, :]
        e = (g + g.T) / 2
        s = lam * np.trace(e) * np.eye(3) + 2 * mu * e
        print("surface traction for force", j, s[:, 2])


if __name__ == "__main__":
    _self_check()
