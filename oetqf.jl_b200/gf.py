"""Green's-function builders: the host-side mirror of /root/reference/src/BEM/GF.jl.

`stress_greens_function` keeps the reference's four call shapes (GF.jl:31, :123, :194, :250) and
returns NumPy arrays with the reference's column-major layout; every entry is computed by the
sm_100a kernels behind the C ABI.  `device_*` builders return `DeviceMatrix` handles whose row
shard stays in HBM for the RHS.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .mesh import BEMHex8Mesh, RectOkadaMesh


class FaultType:
    code = -1


class StrikeSlip(FaultType):       # GF.jl:4
    code = 0


class DipSlip(FaultType):          # GF.jl:5
    code = 1


def _ftype_code(ftype) -> int:
    if isinstance(ftype, type) and issubclass(ftype, FaultType):
        return ftype.code
    if isinstance(ftype, FaultType):
        return ftype.code
    if ftype in (0, 1):
        return int(ftype)
    raise TypeError("ftype must be StrikeSlip() or DipSlip()")


def gauss_legendre_hex(n: int):
    """Tensor-product Gauss-Legendre rule with n points per axis on [-1,1]^3, weights normalised to 1
    (GF.jl:318-323); point order: first coordinate fastest."""
    p, w = np.polynomial.legendre.leggauss(n)
    k, j, i = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    coords = np.stack([p[i.ravel()], p[j.ravel()], p[k.ravel()]], axis=1).reshape(-1)
    weights = (w[i] * w[j] * w[k]).ravel()
    return coords, weights / weights.sum()


def gmsh_hex_points_per_axis(order: int) -> int:
    """Points per axis of the hexahedron rule Gmsh >= 4.9 returns for getIntegrationPoints(5, "Gauss<order>")
    (the call behind GF.jl:318-323): <order> is the polynomial ORDER integrated exactly, not the point count --
    "Gauss1": 1 point; "Gauss2" and "Gauss3": the 2x2x2 product rule (examples/otf-with-mantle.jl:64-66);
    from order 4 on (order + 3) // 2 points per axis ("Gauss4": 27 points).  Gmsh itself is not available here: the
    mapping restates Gmsh's documented behaviour; pass an explicit (coords, weights) tuple to be independent of it."""
    if order < 2:
        return 1
    if order < 4:
        return 2
    return (order + 3) // 2


def get_quadrature(qtype):
    """GF.jl:318-328: a "GaussN" name or a (localCoords[3nq], weights[nq]) tuple."""
    if isinstance(qtype, str):
        if not (qtype.startswith("Gauss") and qtype[5:].isdigit()):
            raise ValueError(f"unsupported quadrature {qtype!r}")
        return gauss_legendre_hex(gmsh_hex_points_per_axis(int(qtype[5:])))
    coords, weights = qtype
    coords = np.asarray(coords, dtype=np.float64).reshape(-1)
    weights = np.asarray(weights, dtype=np.float64).reshape(-1)
    assert coords.size == 3 * weights.size, "Wrong format of quadrature!"
    return coords, weights


class _Quad:
    def __init__(self, qtype):
        c, w = get_quadrature(qtype)
        self.c, self.w = _lib.f64(c), _lib.f64(w)
        self.struct = _lib.OqQuadrature(self.w.size, _lib.dptr(self.c), _lib.dptr(self.w))

    def ref(self):
        return C.byref(self.struct)


last_kernel_ms = {"value": None}     # device time of the most recent assembly kernel (CUDA events)


def stress_greens_function(*args, ftype=StrikeSlip(), fourier=True, nrept=2, buffer_ratio=0.0,
                           qtype="Gauss1", checkeigvals=False, fftw_flags=None):
    """The four methods of the reference, dispatched on the mesh arguments:

    stress_greens_function(mf, λ, μ; ftype, fourier, nrept, buffer_ratio)        -> [nx,nξ,nξ]   GF.jl:31
    stress_greens_function(mf, ma, λ, μ; ftype, qtype, nrept, buffer_ratio)     -> [6ne, nf]    GF.jl:123
    stress_greens_function(ma, mf, λ, μ; ftype)                                 -> [nf, 6ne]    GF.jl:194
    stress_greens_function(ma, λ, μ; qtype, checkeigvals)                       -> [6ne, 6ne]   GF.jl:250

    `checkeigvals` defaults to False here: the reference's O(n^3) `eigvals` print (GF.jl:291-294) is a
    diagnostic outside the hot path; pass True to print the same line (NumPy LAPACK on the host).
    """
    lib = _lib.load()
    ms = C.c_double(0.0)
    if buffer_ratio < 0:
        raise AssertionError("Argument `buffer_ratio` must be ≥ 0.")
    if len(args) == 3 and isinstance(args[0], RectOkadaMesh):
        mf, lam, mu = args
        s = mf.c_struct()
        n = mf.nx * mf.nxi * mf.nxi
        out = np.zeros(2 * n if fourier else n)
        _lib.check(lib.oq_gf_fault_fault(C.byref(s), C.c_double(lam), C.c_double(mu), _ftype_code(ftype),
                                         int(bool(fourier)), int(nrept), C.c_double(buffer_ratio),
                                         _lib.dptr(out), C.byref(ms)))
        last_kernel_ms["value"] = ms.value
        if fourier:
            out = out.view(np.complex128)
        return out.reshape((mf.nx, mf.nxi, mf.nxi), order="F")
    if len(args) == 4 and isinstance(args[0], RectOkadaMesh) and isinstance(args[1], BEMHex8Mesh):
        mf, ma, lam, mu = args
        sf, sa, q = mf.c_struct(), ma.c_struct(), _Quad(qtype)
        out = np.zeros((6 * len(ma), mf.nx * mf.nxi), order="F")
        _lib.check(lib.oq_gf_fault_mantle(C.byref(sf), C.byref(sa), q.ref(), C.c_double(lam), C.c_double(mu),
                                          _ftype_code(ftype), int(nrept), C.c_double(buffer_ratio),
                                          _lib.dptr(out), C.byref(ms)))
        last_kernel_ms["value"] = ms.value
        return out
    if len(args) == 4 and isinstance(args[0], BEMHex8Mesh) and isinstance(args[1], RectOkadaMesh):
        ma, mf, lam, mu = args
        sf, sa = mf.c_struct(), ma.c_struct()
        out = np.zeros((mf.nx * mf.nxi, 6 * len(ma)), order="F")
        _lib.check(lib.oq_gf_mantle_fault(C.byref(sa), C.byref(sf), C.c_double(lam), C.c_double(mu),
                                          _ftype_code(ftype), _lib.dptr(out), C.byref(ms)))
        last_kernel_ms["value"] = ms.value
        return out
    if len(args) == 3 and isinstance(args[0], BEMHex8Mesh):
        ma, lam, mu = args
        sa, q = ma.c_struct(), _Quad(qtype)
        out = np.zeros((6 * len(ma), 6 * len(ma)), order="F")
        _lib.check(lib.oq_gf_mantle_mantle(C.byref(sa), q.ref(), C.c_double(lam), C.c_double(mu),
                                           _lib.dptr(out), C.byref(ms)))
        last_kernel_ms["value"] = ms.value
        if checkeigvals:
            print("Maximum real part of eigval is: %.4f" % np.max(np.linalg.eigvals(out).real))
        return out
    raise TypeError("no method matching stress_greens_function for these argument types")


# ------------------------------------------------------------------------------------------------
class DeviceMatrix:
    """Row shard of a Green's matrix resident in HBM (OqMatrix handle)."""

    def __init__(self, handle):
        self._h = handle
        lr, c, gr = C.c_int(), C.c_int(), C.c_int()
        _lib.check(_lib.load().oq_matrix_shape(handle, C.byref(lr), C.byref(c), C.byref(gr)))
        self.local_rows, self.cols, self.global_rows = lr.value, c.value, gr.value

    @property
    def handle(self):
        if self._h is None:
            raise _lib.OqError("matrix already destroyed")
        return self._h

    def to_host(self) -> np.ndarray:
        out = np.zeros((self.local_rows, self.cols), order="F")
        _lib.check(_lib.load().oq_matrix_to_host(self.handle, _lib.dptr(out)))
        return out

    def rows_to_host(self, begin: int, end: int) -> np.ndarray:
        """Local rows [begin, end) of the shard as a C-ordered (rows x cols) array (oq_matrix_rows_to_host)."""
        out = np.zeros((end - begin, self.cols))
        _lib.check(_lib.load().oq_matrix_rows_to_host(self.handle, int(begin), int(end), _lib.dptr(out)))
        return out

    def kernel_ms(self) -> float:
        ms = C.c_double()
        _lib.check(_lib.load().oq_matrix_kernel_ms(self.handle, C.byref(ms)))
        return ms.value

    def assembly_info(self) -> dict:
        """How the shard was assembled (oq_matrix_assembly_info): path pair / tile / classes, pairs, closed-form
        evaluations actually made, device times of the class table and of the dense expansion."""
        info = _lib.OqAssemblyInfo()
        _lib.check(_lib.load().oq_matrix_assembly_info(self.handle, C.byref(info)))
        return {"path": {-1: None, 0: "pair", 1: "tile", 2: "classes"}[info.path], "pairs": info.pairs,
                "unique_pairs": info.unique_pairs, "table_ms": info.table_ms, "expand_ms": info.expand_ms,
                "kernel_ms": info.kernel_ms}

    def form(self) -> dict:
        """Storage form of the operand (oq_matrix_form): "dense" shard or "classes" (table of distinct kernels + class
        maps, csrc/classmat.cuh) and the bytes of HBM it holds."""
        f, b = C.c_int(), C.c_double()
        _lib.check(_lib.load().oq_matrix_form(self.handle, C.byref(f), C.byref(b)))
        return {"form": "classes" if f.value else "dense", "device_bytes": b.value}

    def gemv(self, x, y=None) -> np.ndarray:
        """The matvecmul! slot (src/pref.jl:15-21): y = A x, or y += A x when y is given."""
        x = _lib.f64(np.asarray(x).reshape(-1, order="F"))
        assert x.size == self.cols, "dimension mismatch"
        acc = 0 if y is None else 1
        if y is None:
            y = np.zeros(self.local_rows)
        assert y.dtype == np.float64 and y.size == self.local_rows
        yv = y.reshape(-1, order="F")
        assert np.shares_memory(yv, y)
        _lib.check(_lib.load().oq_gemv(self.handle, _lib.dptr(x), _lib.dptr(yv), acc))
        return y

    def free(self):
        if self._h is not None:
            _lib.load().oq_matrix_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _new_handle():
    return C.c_void_p()


def device_fault_fault(mf, lam, mu, ftype=StrikeSlip(), nrept=2, buffer_ratio=0.0, rows=None) -> DeviceMatrix:
    r0, r1 = rows if rows is not None else (0, mf.nx * mf.nxi)
    s, h = mf.c_struct(), _new_handle()
    _lib.check(_lib.load().oq_matrix_fault_fault(C.byref(s), C.c_double(lam), C.c_double(mu), _ftype_code(ftype),
                                                 int(nrept), C.c_double(buffer_ratio), int(r0), int(r1), C.byref(h)))
    return DeviceMatrix(h)


def device_fault_mantle(mf, ma, lam, mu, ftype=StrikeSlip(), qtype="Gauss1", nrept=2, buffer_ratio=0.0,
                        elems=None, form="dense") -> DeviceMatrix:
    """gf12 on the device.  form="classes": keep the operand as its table of distinct kernels (no dense storage; the RHS
    multiplies from the table, csrc/classmat.cuh); raises OqError when the meshes have no translation classes."""
    e0, e1 = elems if elems is not None else (0, len(ma))
    sf, sa, q, h = mf.c_struct(), ma.c_struct(), _Quad(qtype), _new_handle()
    lib = _lib.load()
    fn = {"dense": lib.oq_matrix_fault_mantle, "classes": lib.oq_matrix_fault_mantle_classes}[form]
    _lib.check(fn(C.byref(sf), C.byref(sa), q.ref(), C.c_double(lam),
                                                  C.c_double(mu), _ftype_code(ftype), int(nrept),
                                                  C.c_double(buffer_ratio), int(e0), int(e1), C.byref(h)))
    return DeviceMatrix(h)


def device_mantle_fault(ma, mf, lam, mu, ftype=StrikeSlip(), rows=None, form="dense") -> DeviceMatrix:
    r0, r1 = rows if rows is not None else (0, mf.nx * mf.nxi)
    sf, sa, h = mf.c_struct(), ma.c_struct(), _new_handle()
    lib = _lib.load()
    fn = {"dense": lib.oq_matrix_mantle_fault, "classes": lib.oq_matrix_mantle_fault_classes}[form]
    _lib.check(fn(C.byref(sa), C.byref(sf), C.c_double(lam), C.c_double(mu),
                                                  _ftype_code(ftype), int(r0), int(r1), C.byref(h)))
    return DeviceMatrix(h)


def device_mantle_mantle(ma, lam, mu, qtype="Gauss1", elems=None, form="dense") -> DeviceMatrix:
    e0, e1 = elems if elems is not None else (0, len(ma))
    sa, q, h = ma.c_struct(), _Quad(qtype), _new_handle()
    lib = _lib.load()
    fn = {"dense": lib.oq_matrix_mantle_mantle, "classes": lib.oq_matrix_mantle_mantle_classes}[form]
    _lib.check(fn(C.byref(sa), q.ref(), C.c_double(lam), C.c_double(mu),
                                                   int(e0), int(e1), C.byref(h)))
    return DeviceMatrix(h)


def device_from_host(a, row_kind="fault", rows=None) -> DeviceMatrix:
    """Upload a reference-layout (column-major) Green's matrix, e.g. one read back from the HDF5 cache of
    examples/otf-with-mantle.jl:39-56."""
    a = _lib.f64(a)
    m, n = a.shape
    kind = 1 if row_kind == "mantle" else 0
    units = m // 6 if kind else m
    r0, r1 = rows if rows is not None else (0, units)
    h = _new_handle()
    _lib.check(_lib.load().oq_matrix_from_host(_lib.dptr(a), m, n, kind, int(r0), int(r1), C.byref(h)))
    return DeviceMatrix(h)


def dc3d_gradient(x, y, z, alpha, dep, dip, al1, al2, aw1, aw2, ftype=StrikeSlip()):
    """Batched gradient rows (entries 4..12) of `dc3d` as called at GF.jl:49-54."""
    x, y, z = (_lib.f64(np.atleast_1d(v)) for v in (x, y, z))
    out = np.zeros((x.size, 9))
    d = C.c_double
    _lib.check(_lib.load().oq_dc3d_gradient(x.size, _lib.dptr(x), _lib.dptr(y), _lib.dptr(z), d(alpha), d(dep),
                                            d(dip), d(al1), d(al2), d(aw1), d(aw2), _ftype_code(ftype),
                                            _lib.dptr(out)))
    return out


def stress_vol_hex8(x, y, z, qx, qy, qz, dx, dy, dz, eps, mu, nu):
    """Batched `stress_vol_hex8!` as called at GF.jl:215-221 (θ = 0)."""
    x, y, z = (_lib.f64(np.atleast_1d(v)) for v in (x, y, z))
    e = _lib.f64(eps)
    out = np.zeros((x.size, 6))
    d = C.c_double
    _lib.check(_lib.load().oq_stress_vol_hex8(x.size, _lib.dptr(x), _lib.dptr(y), _lib.dptr(z), d(qx), d(qy), d(qz),
                                              d(dx), d(dy), d(dz), _lib.dptr(e), d(mu), d(nu), _lib.dptr(out)))
    return out


def max_real_eigval(dm: "DeviceMatrix", k: int = 1, tol: float = 1e-6, maxiter: int = 300) -> float:
    """Largest real part of the spectrum of a square device matrix by implicitly restarted Arnoldi
    (scipy.sparse.linalg.eigs, which='LR') with the matvecs on the GPU (oq_gemv).

    Replaces the reference's O(n^3) `maximum(real, eigvals(st))` print (GF.jl:291-294), which is unusable
    beyond ~1e4 rows; needs the full matrix on this rank (local_rows == cols)."""
    from scipy.sparse.linalg import LinearOperator, eigs
    assert dm.local_rows == dm.cols == dm.global_rows, "needs the whole square matrix on this rank"
    op = LinearOperator((dm.cols, dm.cols), matvec=lambda x: dm.gemv(np.ascontiguousarray(x, dtype=np.float64)),
                        dtype=np.float64)
    vals = eigs(op, k=k, which="LR", tol=tol, maxiter=maxiter, return_eigenvectors=False)
    return float(np.max(vals.real))


def class_form_plan(ma, elems=None) -> dict:
    """Host-only plan of device_mantle_mantle(..., form="classes") for the receivers `elems` (oq_class_form_plan)."""
    e0, e1 = elems if elems is not None else (0, len(ma))
    out = np.zeros(8, dtype=np.int64)
    cma = ma.c_struct()
    _lib.check(_lib.load().oq_class_form_plan(C.byref(cma), int(e0), int(e1), out.ctypes.data_as(C.POINTER(C.c_longlong))))
    keys = ("x_classes", "yz_classes", "worthwhile", "diagonal", "runs", "x_positions", "max_source_group", "receiver_yz_classes")
    return {k: (bool(v) if k in ("worthwhile", "diagonal") else int(v)) for k, v in zip(keys, out)}


def class_window_check(ma, mf, which, begin=0, end=None) -> dict:
    """Host-only self check of the sliding-window plan of gf21 (which=1) / gf12 (which=2) (oq_class_window_check)."""
    end = (mf.nx * mf.nxi if which == 1 else len(ma)) if end is None else end
    out = np.zeros(6, dtype=np.int64)
    cma, cmf = ma.c_struct(), mf.c_struct()
    _lib.check(_lib.load().oq_class_window_check(C.byref(cma), C.byref(cmf), int(which), int(begin), int(end),
                                                 out.ctypes.data_as(C.POINTER(C.c_longlong))))
    return dict(zip(("found", "residues", "runs", "checked", "mismatches", "unreached"), (int(v) for v in out)))


def hex8_pair_classes(ma, mf=None, begin=0, end=None, recv=(), src=()):
    """Host-only view of the class decomposition behind device_mantle_mantle (mf=None) / device_mantle_fault
    (oq_hex8_pair_classes): (counts, rep_recv_x, rep_src_x, rep_recv_yz, rep_src_yz) for the sample pairs."""
    end = (len(ma) if mf is None else mf.nx * mf.nxi) if end is None else end
    recv = np.ascontiguousarray(recv, dtype=np.int32)
    src = np.ascontiguousarray(src, dtype=np.int32)
    n = recv.size
    reps = [np.zeros(max(n, 1), dtype=np.int32) for _ in range(4)]
    counts = np.zeros(3, dtype=np.int64)
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))   # noqa: E731
    cma = ma.c_struct()
    cmf = mf.c_struct() if mf is not None else None
    _lib.check(_lib.load().oq_hex8_pair_classes(C.byref(cma), C.byref(cmf) if cmf is not None else None, int(begin), int(end),
                                                int(n), ip(recv), ip(src), *[ip(r) for r in reps],
                                                counts.ctypes.data_as(C.POINTER(C.c_longlong))))
    return (counts,) + tuple(r[:n] for r in reps)
