"""oracle/ref.py -- Python face of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  It restates, on the CPU, the reference's algorithms for the two hot paths:

* mesh generators        src/BEM/mesh.jl:39-56 (fault), :58-72,:124-128,:181-183 (hex8 SoA of a box)
* Green's builders       src/BEM/GF.jl:31-296 (C/OpenMP in greens.c, okada.c, hex8.c)
* FFT form of gf11       src/BEM/GF.jl:60-68 and src/BEM/equation.jl:44-61 (numpy pocketfft)
* RHS                    src/BEM/equation.jl:156-292 (C/OpenMP in rhs.c + numpy glue)

Parity with GeoGreensFunctions.jl (un-vendored, absent) is UNPINNED; see oracle/okada.c header.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_dp = C.POINTER(C.c_double)


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboetqf_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("okada.c", "greens.c", "hex8.c", "rhs.c")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboetqf_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.oq_ref_num_threads.restype = C.c_int
    return _LIB


_LIB_LD = None
_ldp = C.POINTER(C.c_longdouble)


def lib_ld():
    """The arbiter build: okada.c / hex8.c in 80-bit extended precision (oracle/Makefile, liboetqf_oracle_ld.so)."""
    global _LIB_LD
    if _LIB_LD is None:
        so = os.path.join(_HERE, "liboetqf_oracle_ld.so")
        srcs = [os.path.join(_HERE, f) for f in ("okada.c", "hex8.c", "hex8_gen.inc", "Makefile")]
        if not os.path.exists(so) or any(os.path.getmtime(s_) > os.path.getmtime(so) for s_ in srcs):
            subprocess.check_call(["make", "-C", _HERE, "-B", "liboetqf_oracle_ld.so"], stdout=subprocess.DEVNULL)
        _LIB_LD = C.CDLL(so)
    return _LIB_LD


def dc3d_ld(alpha, x, y, z, depth, dip, al1, al2, aw1, aw2, d1, d2, d3):
    """dc3d evaluated in extended precision (np.longdouble[12])"""
    assert np.finfo(np.longdouble).nmant >= 63, "needs an 80-bit long double"
    u = np.zeros(12, dtype=np.longdouble)
    d = C.c_longdouble
    lib_ld().oq_refl_dc3d(d(alpha), d(x), d(y), d(z), d(depth), d(dip), d(al1), d(al2), d(aw1), d(aw2),
                          d(d1), d(d2), d(d3), u.ctypes.data_as(_ldp))
    return u


def stress_vol_hex8_ld(x, y, z, qx, qy, qz, dx, dy, dz, eps, mu, nu):
    """stress_vol_hex8 evaluated in extended precision (np.longdouble[6])"""
    assert np.finfo(np.longdouble).nmant >= 63, "needs an 80-bit long double"
    sig = np.zeros(6, dtype=np.longdouble)
    e = np.ascontiguousarray(np.asarray(eps, dtype=np.longdouble))
    d = C.c_longdouble
    lib_ld().oq_refl_stress_vol_hex8(d(x), d(y), d(z), d(qx), d(qy), d(qz), d(dx), d(dy), d(dz),
                                     e.ctypes.data_as(_ldp), d(mu), d(nu), sig.ctypes.data_as(_ldp))
    return sig


def _p(a):
    assert a.dtype == np.float64 and a.flags["F_CONTIGUOUS"] or a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def _f(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def num_threads() -> int:
    return int(lib().oq_ref_num_threads())


def use_all_cores() -> int:
    """Let OpenMP use every core this process may run on (launchers such as torchrun pin OMP_NUM_THREADS=1)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().oq_ref_set_num_threads(int(n))
    return num_threads()


# ----------------------------------------------------------------------------- meshes
@dataclass
class FaultMesh:
    """Restates RectOkadaMesh (src/BEM/mesh.jl:5-21)."""
    x: np.ndarray
    dx: float
    nx: int
    ax: np.ndarray      # [nx, 2]
    xi: np.ndarray
    dxi: float
    nxi: int
    axi: np.ndarray     # [nxi, 2]
    y: np.ndarray
    z: np.ndarray
    dep: float
    dip: float


def sincosd(deg: float):
    s = C.c_double()
    c = C.c_double()
    lib().oq_ref_sincosd(C.c_double(deg), C.byref(s), C.byref(c))
    return s.value, c.value


def _julia_range_len(start, stop, step):
    # length of range(start, stop=stop, step=step) as Julia computes it for floats (to the nearest
    # representable count; inputs here are exact multiples)
    return int(np.floor((stop - start) / step + 1e-9)) + 1


def fault_mesh(x: float, xi: float, dx: float, dxi: float, dip: float) -> FaultMesh:
    """src/BEM/mesh.jl:39-56."""
    nxi = _julia_range_len(0.0, -xi + dxi, -dxi)
    xic = (0.0 + np.arange(nxi) * (-dxi)) - dxi / 2
    axi = np.stack([xic - dxi / 2, xic + dxi / 2], axis=1)
    sd, cd = sincosd(dip)
    y, z = xic * cd, xic * sd
    nx = _julia_range_len(-x / 2 + dx / 2, x / 2 - dx / 2, dx)
    xc = (-x / 2 + dx / 2) + np.arange(nx) * dx
    ax = np.stack([xc - dx / 2, xc + dx / 2], axis=1)
    return FaultMesh(xc, dx, nx, ax, xic, dxi, nxi, axi, y, z, 0.0, dip)


@dataclass
class Hex8Mesh:
    """Restates BEMHex8Mesh (src/BEM/mesh.jl:58-72)."""
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

    @property
    def n(self):
        return len(self.cx)


def hex8_box(llx, lly, llz, dx, dy, dz, nx, ny, nz, rfzh=None) -> Hex8Mesh:
    """Structured restatement of gen_gmsh_mesh + gen_mesh(Val(:BEMHex8Mesh)) for an unrefined box
    (src/BEM/mesh.jl:95-133,149-186): top-surface corner (llx,lly,llz), extents dx,dy and dz<0 downward,
    layer heights normalize(cumsum(rfzh), Inf) (mesh.jl:124-128).  Element order: x fastest, then y,
    then z layers from the top (Gmsh's own numbering is not reproducible without Gmsh; any
    permutation of cells permutes rows/columns of the Green's matrices consistently)."""
    rfzh = np.ones(nz) if rfzh is None else np.asarray(rfzh, dtype=float)
    frac = np.cumsum(rfzh)
    frac = frac / np.max(np.abs(frac))
    zedges = llz + np.concatenate([[0.0], frac]) * dz
    xedges = llx + np.arange(nx + 1) * (dx / nx)
    yedges = lly + np.arange(ny + 1) * (dy / ny)
    cx, cy, cz, ex, ey, ez = [], [], [], [], [], []
    for k in range(nz):
        for j in range(ny):
            for i in range(nx):
                cx.append((xedges[i] + xedges[i + 1]) / 2)
                cy.append((yedges[j] + yedges[j + 1]) / 2)
                cz.append((zedges[k] + zedges[k + 1]) / 2)
                ex.append(abs(xedges[i + 1] - xedges[i]))
                ey.append(abs(yedges[j + 1] - yedges[j]))
                ez.append(abs(zedges[k + 1] - zedges[k]))
    cx, cy, cz, ex, ey, ez = map(np.array, (cx, cy, cz, ex, ey, ez))
    return Hex8Mesh(cx, cy, cz, cx.copy(), cy - ey / 2, cz + ez / 2, ex, ey, ez, 0.0)


def gauss_quadrature(n: int):
    """Tensor Gauss-Legendre rule on [-1,1]^3 with weights normalised to 1 (GF.jl:318-323);
    n=1 is Gmsh's "Gauss1" (one point at the origin)."""
    p, w = np.polynomial.legendre.leggauss(n)
    pts, wts = [], []
    for k in range(n):
        for j in range(n):
            for i in range(n):
                pts += [p[i], p[j], p[k]]
                wts.append(w[i] * w[j] * w[k])
    wts = np.array(wts)
    return np.array(pts), wts / wts.sum()


# ----------------------------------------------------------------------------- dc3d
def dc3d(alpha, x, y, z, depth, dip, al1, al2, aw1, aw2, d1, d2, d3):
    u = np.zeros(12)
    d = C.c_double
    lib().oq_ref_dc3d(d(alpha), d(x), d(y), d(z), d(depth), d(dip), d(al1), d(al2), d(aw1), d(aw2),
                      d(d1), d(d2), d(d3), _p(u))
    return u


def stress_vol_hex8(x, y, z, qx, qy, qz, dx, dy, dz, eps, mu, nu):
    sig = np.zeros(6)
    e = _f(eps)
    d = C.c_double
    lib().oq_ref_stress_vol_hex8(d(x), d(y), d(z), d(qx), d(qy), d(qz), d(dx), d(dy), d(dz), _p(e),
                                 d(mu), d(nu), _p(sig))
    return sig


# ----------------------------------------------------------------------------- Green's builders
def gf_fault_fault(mf: FaultMesh, lam, mu, ftype=0, nrept=2, buffer_ratio=0.0, fourier=False):
    """src/BEM/GF.jl:31-71; returns st[nx,nxi,nxi] (Fortran order) or its strike-wise rFFT."""
    st = np.zeros((mf.nx, mf.nxi, mf.nxi), order="F")
    d = C.c_double
    x, y, z = _f(mf.x), _f(mf.y), _f(mf.z)
    a0, a1 = _f(mf.axi[:, 0]), _f(mf.axi[:, 1])
    lib().oq_ref_gf_fault_fault(mf.nx, mf.nxi, _p(x), d(mf.ax[0, 0]), d(mf.ax[0, 1]), _p(y), _p(z),
                                _p(a0), _p(a1), d(mf.dx), d(mf.dep), d(mf.dip), d(lam), d(mu),
                                int(ftype), int(nrept), d(buffer_ratio), _p(st))
    return gf_fourier(st) if fourier else st


def gf_fourier(st):
    """GF.jl:60-68: rfft along strike of the even extension [st; reverse(st[2:end])] (length 2nx-1)."""
    ext = np.concatenate([st, st[:0:-1]], axis=0)
    return np.asfortranarray(np.fft.rfft(ext, axis=0))


def gf_fault_mantle(mf: FaultMesh, ma: Hex8Mesh, lam, mu, ftype=0, quad=None, nrept=2, buffer_ratio=0.0):
    """src/BEM/GF.jl:123-174; returns [6*ne, nx*nxi] Fortran order."""
    lc, w = gauss_quadrature(1) if quad is None else quad
    lc, w = _f(lc), _f(w)
    assert lc.size == 3 * w.size, "Wrong format of quadrature!"
    st = np.zeros((6 * ma.n, mf.nx * mf.nxi), order="F")
    d = C.c_double
    arrs = [_f(a) for a in (mf.ax[:, 0], mf.ax[:, 1], mf.axi[:, 0], mf.axi[:, 1],
                            ma.cx, ma.cy, ma.cz, ma.dx, ma.dy, ma.dz)]
    lib().oq_ref_gf_fault_mantle(mf.nx, mf.nxi, _p(arrs[0]), _p(arrs[1]), _p(arrs[2]), _p(arrs[3]),
                                 d(mf.dx), d(mf.dep), d(mf.dip), ma.n,
                                 *[_p(a) for a in arrs[4:]], w.size, _p(lc), _p(w),
                                 d(lam), d(mu), int(ftype), int(nrept), d(buffer_ratio), _p(st))
    return st


def gf_mantle_fault(ma: Hex8Mesh, mf: FaultMesh, lam, mu, ftype=0):
    """src/BEM/GF.jl:194-227; returns [nx*nxi, 6*ne] Fortran order."""
    st = np.zeros((mf.nx * mf.nxi, 6 * ma.n), order="F")
    d = C.c_double
    arrs = [_f(a) for a in (ma.qx, ma.qy, ma.qz, ma.dx, ma.dy, ma.dz, mf.x, mf.y, mf.z)]
    lib().oq_ref_gf_mantle_fault(ma.n, *[_p(a) for a in arrs[:6]], mf.nx, mf.nxi,
                                 *[_p(a) for a in arrs[6:]], d(mf.dip), d(lam), d(mu), int(ftype), _p(st))
    return st


def gf_mantle_mantle(ma: Hex8Mesh, lam, mu, quad=None):
    """src/BEM/GF.jl:250-296 without the eigvals diagnostic; returns [6*ne, 6*ne] Fortran order."""
    lc, w = gauss_quadrature(1) if quad is None else quad
    lc, w = _f(lc), _f(w)
    assert lc.size == 3 * w.size, "Wrong format of quadrature!"
    st = np.zeros((6 * ma.n, 6 * ma.n), order="F")
    d = C.c_double
    arrs = [_f(a) for a in (ma.cx, ma.cy, ma.cz, ma.qx, ma.qy, ma.qz, ma.dx, ma.dy, ma.dz)]
    lib().oq_ref_gf_mantle_mantle(ma.n, *[_p(a) for a in arrs], w.size, _p(lc), _p(w),
                                  d(lam), d(mu), _p(st))
    return st


def dense_from_toeplitz(st):
    """test/BEM/tests.jl:46-49: G[(i,j),(k,l)] = st[|i-k|, j, l], rows/cols in vec order i + j*nx."""
    nx, nxi, _ = st.shape
    idx = np.abs(np.arange(nx)[:, None] - np.arange(nx)[None, :])          # [i,k]
    g4 = st[idx]                                                            # [i,k,j,l]
    g4 = np.transpose(g4, (0, 2, 1, 3))                                     # [i,j,k,l]
    return np.asfortranarray(g4.reshape(nx * nxi, nx * nxi, order="F"))


# ----------------------------------------------------------------------------- RHS
def dtau_dt_fft(gf_dft, relv):
    """src/BEM/equation.jl:44-61 with numpy's pocketfft: zero-pad to 2nx-1, rfft, contract over the
    source down-dip index, irfft, keep the first nx rows."""
    nx, nxi, _ = gf_dft.shape
    pad = np.zeros((2 * nx - 1, nxi))
    pad[:nx] = relv
    rd = np.asfortranarray(np.fft.rfft(pad, axis=0))             # [nx, nxi(l)]
    td = np.zeros((nx, nxi), dtype=np.complex128, order="F")
    g = np.asfortranarray(gf_dft, dtype=np.complex128)
    lib().oq_ref_fft_contract(nx, nxi, g.ctypes.data_as(_dp), rd.ctypes.data_as(_dp), td.ctypes.data_as(_dp))
    return np.fft.irfft(td, n=2 * nx - 1, axis=0)[:nx]


def dtau_dt_toeplitz(st, relv):
    nx, nxi, _ = st.shape
    out = np.zeros((nx, nxi), order="F")
    stf, rf = np.asfortranarray(st), np.asfortranarray(relv)
    lib().oq_ref_dtau_dt_toeplitz(nx, nxi, _p(stf), _p(rf), _p(out))
    return out


def gemv(A, x, y=None):
    A = np.asfortranarray(A)
    x = _f(x)
    acc = 0 if y is None else 1
    y = np.zeros(A.shape[0]) if y is None else y
    lib().oq_ref_gemv(A.shape[0], A.shape[1], _p(A), _p(x), _p(y), acc)
    return y


@dataclass
class FaultProp:
    """RateStateQuasiDynamicProperty (src/BEM/property.jl:10-25)."""
    a: np.ndarray
    b: np.ndarray
    L: np.ndarray
    sigma: np.ndarray
    eta: float
    vpl: float
    f0: float = 0.6
    v0: float = 1e-6


@dataclass
class MantleProp:
    """PowerLaw / CompositePowerLaw viscosity (src/BEM/property.jl:34-48): gamma, n are [nlaws, ne];
    n holds power-1."""
    gamma: np.ndarray
    n: np.ndarray
    deps0: np.ndarray


@dataclass
class DilatancyProp:
    """DilatancyProperty (src/BEM/property.jl:27-32)."""
    tp: np.ndarray
    eps: np.ndarray
    beta: np.ndarray
    p0: np.ndarray


def update_strain_rate(pa: MantleProp, sigma):
    ne = sigma.shape[0]
    g = _f(np.atleast_2d(pa.gamma))
    n = _f(np.atleast_2d(pa.n))
    s = np.asfortranarray(sigma, dtype=np.float64)
    out = np.zeros((ne, 6), order="F")
    lib().oq_ref_update_strain_rate(ne, g.shape[0], _p(g), _p(n), _p(s), _p(out))
    return out


def update_fault(pf: FaultProp, dtau, v, theta):
    n = v.size
    arr = [np.asfortranarray(a, dtype=np.float64) for a in (pf.a, pf.b, pf.L, pf.sigma, dtau, v, theta)]
    dv, dth, ddl = (np.zeros(v.shape, order="F") for _ in range(3))
    d = C.c_double
    lib().oq_ref_update_fault(n, *[_p(a) for a in arr[:4]], d(pf.eta), d(pf.f0), d(pf.v0),
                              *[_p(a) for a in arr[4:]], _p(dv), _p(dth), _p(ddl))
    return dv, dth, ddl


def update_fault_dilatancy(pf: FaultProp, dl: DilatancyProp, dtau, v, theta, pr):
    n = v.size
    arr = [np.asfortranarray(a, dtype=np.float64) for a in
           (pf.a, pf.b, pf.L, pf.sigma, dl.tp, dl.eps, dl.beta, dl.p0, dtau, v, theta, pr)]
    dv, dth, ddl, dpr = (np.zeros(v.shape, order="F") for _ in range(4))
    d = C.c_double
    lib().oq_ref_update_fault_dilatancy(n, *[_p(a) for a in arr[:4]], d(pf.f0), d(pf.v0),
                                        *[_p(a) for a in arr[4:]], _p(dv), _p(dth), _p(ddl), _p(dpr))
    return dv, dth, ddl, dpr


def rhs_fault(pf: FaultProp, gf, v, theta, form="fft"):
    """ode() fault-only variant, src/BEM/equation.jl:156-166.  gf: complex DFT kernel (form="fft"),
    real Toeplitz kernel (form="toeplitz") or dense matrix (form="dense")."""
    relv = v - pf.vpl
    if form == "fft":
        dtau = dtau_dt_fft(gf, relv)
    elif form == "toeplitz":
        dtau = dtau_dt_toeplitz(gf, relv)
    else:
        dtau = gemv(gf, relv.reshape(-1, order="F")).reshape(v.shape, order="F")
    return update_fault(pf, dtau, v, theta)


class FaultRhsFFT:
    """The fault-only ode() of src/BEM/equation.jl:156-166 as the reference runs it, with the reference's
    allocation discipline (gen_alloc, equation.jl:18-32: every scratch array is created once) and its
    threading: planned, multi-threaded FFTs (FFTW in the reference; pocketfft with `workers` here), the
    per-receiver contraction and update_fault! in C/OpenMP.  Used as the timed CPU arm of bench.py."""

    def __init__(self, pf: FaultProp, gf_dft, workers=None):
        import scipy.fft as sfft
        self._fft = sfft
        self.pf = pf
        nx, nxi, _ = gf_dft.shape
        self.nx, self.nxi = nx, nxi
        self.workers = int(workers or num_threads())
        self.gf = np.asfortranarray(gf_dft, dtype=np.complex128)
        self.relv = np.zeros((2 * nx - 1, nxi), order="F")                      # zero-padded, :22
        self.td = np.zeros((nx, nxi), dtype=np.complex128, order="F")
        self.prop = [np.asfortranarray(a, dtype=np.float64) for a in (pf.a, pf.b, pf.L, pf.sigma)]
        self.dv, self.dth, self.ddl = (np.zeros((nx, nxi), order="F") for _ in range(3))
        self.dtau = np.zeros((nx, nxi), order="F")

    def __call__(self, v, theta):
        nx, nxi, pf = self.nx, self.nxi, self.pf
        np.subtract(v, pf.vpl, out=self.relv[:nx])                                # :35-42
        rd = self._fft.rfft(self.relv, axis=0, workers=self.workers)              # :46
        rd = rd if rd.flags["F_CONTIGUOUS"] else np.asfortranarray(rd)
        lib().oq_ref_fft_contract(nx, nxi, self.gf.ctypes.data_as(_dp), rd.ctypes.data_as(_dp),
                                  self.td.ctypes.data_as(_dp))                    # :48-54
        full = self._fft.irfft(self.td, n=2 * nx - 1, axis=0, workers=self.workers)   # :56
        self.dtau[...] = full[:nx]                                                # :57-59
        d = C.c_double
        lib().oq_ref_update_fault(nx * nxi, *[_p(a) for a in self.prop], d(pf.eta), d(pf.f0), d(pf.v0),
                                  _p(self.dtau), _p(v), _p(theta), _p(self.dv), _p(self.dth), _p(self.ddl))
        return self.dv, self.dth, self.ddl


def blas_gemv(A, x):
    """y = A x through OpenBLAS dgemv -- the reference's default matvecmul! backend (src/pref.jl:15-16);
    A column-major as Julia holds it."""
    from scipy.linalg import blas
    return blas.dgemv(1.0, A, x)


def rhs_fault_dilatancy(pf, dl, gf, v, theta, pr, form="fft"):
    """ode() dilatancy variant, src/BEM/equation.jl:168-183."""
    relv = v - pf.vpl
    dtau = dtau_dt_fft(gf, relv) if form == "fft" else dtau_dt_toeplitz(gf, relv)
    return update_fault_dilatancy(pf, dl, dtau, v, theta, pr)


def rhs_viscoelastic(pf: FaultProp, pa: MantleProp, gf11, gf12, gf21, gf22, v, theta, sigma, form="fft"):
    """ode() viscoelastic variant, src/BEM/equation.jl:185-205.  Returns (dv, dθ, dϵ, dσ, dδ)."""
    relv = v - pf.vpl
    deps = update_strain_rate(pa, sigma)
    rel = deps - np.asarray(pa.deps0)[None, :]
    if form == "fft":
        dtau = dtau_dt_fft(gf11, relv)
    elif form == "toeplitz":
        dtau = dtau_dt_toeplitz(gf11, relv)
    else:
        dtau = gemv(gf11, relv.reshape(-1, order="F")).reshape(v.shape, order="F")
    relf = rel.reshape(-1, order="F")
    dtau_v = np.ascontiguousarray(dtau.reshape(-1, order="F"))
    gemv(gf21, relf, dtau_v)
    dsig = gemv(gf12, relv.reshape(-1, order="F"))
    gemv(gf22, relf, dsig)
    dv, dth, ddl = update_fault(pf, dtau_v.reshape(v.shape, order="F"), v, theta)
    return dv, dth, deps, dsig.reshape(sigma.shape, order="F"), ddl
