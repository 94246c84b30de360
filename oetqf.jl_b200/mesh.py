"""Mesh containers handed to the Green's-function builders (host side, input generators only).

Mirrors /root/reference/src/BEM/mesh.jl: `RectOkadaMesh` (:5-21) with `gen_mesh(Val(:RectOkada), …)`
(:39-56) and the `BEMHex8Mesh` SoA (:58-72).  Gmsh is not part of the hot path: instead of the
Gmsh-backed generator (:95-186) a structured box builder produces the same SoA fields.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from . import _lib


def sincosd(deg: float):
    """sind/cosd with exact values at multiples of 90° (Julia's sincosd)."""
    r = math.fmod(deg, 360.0)
    if r < 0:
        r += 360.0
    exact = {0.0: (0.0, 1.0), 90.0: (1.0, 0.0), 180.0: (0.0, -1.0), 270.0: (-1.0, 0.0)}
    if r in exact:
        return exact[r]
    # long-double evaluation then rounding to double, like the device library's host code
    a = np.longdouble(deg) * (np.longdouble("3.14159265358979323846264338327950288") / np.longdouble(180.0))
    return float(np.sin(a)), float(np.cos(a))


@dataclass
class RectOkadaMesh:
    """src/BEM/mesh.jl:5-21 (field names transliterated: ξ -> xi, Δ -> d)."""
    x: np.ndarray
    dx: float
    nx: int
    ax: np.ndarray          # [nx, 2] strike-cell edges
    xi: np.ndarray
    dxi: float
    nxi: int
    axi: np.ndarray         # [nxi, 2] down-dip cell edges
    y: np.ndarray
    z: np.ndarray
    dep: float
    dip: float
    _keep: list = field(default_factory=list, repr=False)

    def __post_init__(self):
        assert len(self.x) == len(self.ax) == self.nx
        assert len(self.xi) == len(self.axi) == self.nxi == len(self.y) == len(self.z)

    def c_struct(self) -> _lib.OqFaultMesh:
        arrs = [_lib.f64(a) for a in (self.x, self.ax[:, 0], self.ax[:, 1], self.xi, self.axi[:, 0],
                                      self.axi[:, 1], self.y, self.z)]
        self._keep = arrs
        p = [_lib.dptr(a) for a in arrs]
        return _lib.OqFaultMesh(self.nx, self.nxi, p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7],
                                float(self.dx), float(self.dxi), float(self.dep), float(self.dip))


@dataclass
class BEMHex8Mesh:
    """src/BEM/mesh.jl:58-72 (Δx -> dx …; θ kept for interface parity, unused as in the reference)."""
    cx: np.ndarray
    cy: np.ndarray
    cz: np.ndarray
    qx: np.ndarray
    qy: np.ndarray
    qz: np.ndarray
    dx: np.ndarray
    dy: np.ndarray
    dz: np.ndarray
    theta: float = 0.0
    _keep: list = field(default_factory=list, repr=False)

    def __post_init__(self):
        n = len(self.cx)
        assert all(len(a) == n for a in (self.cy, self.cz, self.qx, self.qy, self.qz, self.dx, self.dy, self.dz))

    def __len__(self):
        return len(self.cx)

    def c_struct(self) -> _lib.OqHex8Mesh:
        arrs = [_lib.f64(a) for a in (self.cx, self.cy, self.cz, self.qx, self.qy, self.qz,
                                      self.dx, self.dy, self.dz)]
        self._keep = arrs
        return _lib.OqHex8Mesh(len(self.cx), *[_lib.dptr(a) for a in arrs])


def _range_len(start: float, stop: float, step: float) -> int:
    return int(math.floor((stop - start) / step + 1e-9)) + 1


def gen_mesh(kind: str, *args, **kw):
    """gen_mesh(Val(:RectOkada), x, ξ, Δx, Δξ, dip) -> RectOkadaMesh   (src/BEM/mesh.jl:39-43)
    gen_mesh(Val(:BEMHex8Mesh), llx, lly, llz, dx, dy, dz, nx, ny, nz; rfzh) -> BEMHex8Mesh
    (structured stand-in for gen_gmsh_mesh + the .msh reader, mesh.jl:95-186)."""
    if kind == "RectOkada":
        return _rect_okada(*args, **kw)
    if kind == "BEMHex8Mesh":
        return gen_box_hex8(*args, **kw)
    raise ValueError(f"unknown mesh kind {kind!r}")


def _rect_okada(x: float, xi: float, dx: float, dxi: float, dip: float) -> RectOkadaMesh:
    # _equidist_mesh_downdip, mesh.jl:45-50
    nxi = _range_len(0.0, -xi + dxi, -dxi)
    xic = np.arange(nxi) * (-dxi) - dxi / 2
    axi = np.stack([xic - dxi / 2, xic + dxi / 2], axis=1)
    sd, cd = sincosd(dip)
    # _equidist_mesh_strike, mesh.jl:52-56
    nx = _range_len(-x / 2 + dx / 2, x / 2 - dx / 2, dx)
    xc = (-x / 2 + dx / 2) + np.arange(nx) * dx
    ax = np.stack([xc - dx / 2, xc + dx / 2], axis=1)
    return RectOkadaMesh(xc, dx, nx, ax, xic, dxi, nxi, axi, xic * cd, xic * sd, 0.0, dip)


def gen_box_hex8(llx, lly, llz, dx, dy, dz, nx, ny, nz, rfzh=None) -> BEMHex8Mesh:
    """Axis-aligned box of nx*ny*nz cuboids below the top-surface corner (llx,lly,llz); dz < 0 extends
    downward; layer heights follow normalize(cumsum(rfzh), Inf) (mesh.jl:124-128).  q-point convention of
    mesh.jl:181-183: qx = cx, qy = cy - Δy/2, qz = cz + Δz/2."""
    rfzh = np.ones(nz) if rfzh is None else np.asarray(rfzh, dtype=float)
    assert len(rfzh) == nz
    frac = np.cumsum(rfzh)
    frac = frac / np.max(np.abs(frac))
    ze = llz + np.concatenate([[0.0], frac]) * dz
    xe = llx + np.arange(nx + 1) * (dx / nx)
    ye = lly + np.arange(ny + 1) * (dy / ny)
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    i, j, k = i.ravel(), j.ravel(), k.ravel()          # x fastest, then y, then z layers from the top
    cx = (xe[i] + xe[i + 1]) / 2
    cy = (ye[j] + ye[j + 1]) / 2
    cz = (ze[k] + ze[k + 1]) / 2
    ex, ey, ez = np.abs(xe[i + 1] - xe[i]), np.abs(ye[j + 1] - ye[j]), np.abs(ze[k + 1] - ze[k])
    return BEMHex8Mesh(cx, cy, cz, cx.copy(), cy - ey / 2, cz + ez / 2, ex, ey, ez, 0.0)
