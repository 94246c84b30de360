"""`assemble` and the in-place RHS: host mirror of /root/reference/src/BEM/equation.jl:63-205.

`assemble` keeps the reference's three call shapes and returns an `ODEProblem` whose `f(du, u, p, t)`
is the in-place RHS OrdinaryDiffEq would call ("compat mode": host arrays in/out through the C ABI).
`solve` runs the device-resident integrator ("resident mode").
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence

import numpy as np

from . import _lib
from .gf import DeviceMatrix, device_from_host
from .property import (CompositePowerLawViscosityProperty, DieterichStateLaw, DilatancyProperty,
                       PowerLawViscosityProperty, RateStateQuasiDynamicProperty, ViscosityProperty)


class ArrayPartition:
    """RecursiveArrayTools.ArrayPartition stand-in: `.x` is the tuple of component arrays."""

    def __init__(self, *arrays):
        self.x = tuple(np.asarray(a, dtype=np.float64, order="F") if not (
            isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags["F_CONTIGUOUS"]) else a for a in arrays)

    def similar(self):
        return ArrayPartition(*[np.zeros(a.shape, order="F") for a in self.x])

    def copy(self):
        return ArrayPartition(*[np.array(a, order="F") for a in self.x])


def _toeplitz_from_gf(gf: np.ndarray) -> np.ndarray:
    """Accept either form the reference's builder returns (GF.jl:60-70): the real kernel st[nx,nξ,nξ], or
    its strike-wise rFFT, from which the real kernel is recovered (inverse of GF.jl:64-67)."""
    gf = np.asarray(gf)
    if np.iscomplexobj(gf):
        nx = gf.shape[0]
        gf = np.fft.irfft(gf, n=2 * nx - 1, axis=0)[:nx]
    return _lib.f64(gf)


class DeviceProblem:
    """OqProblem handle + the matrices it borrows."""

    def __init__(self, handle, keep, shapes):
        self._h = handle
        self._keep = keep
        self._ptr_cache = {}
        self.shapes = shapes            # global shapes of the state partitions (reference layout)
        n, lens = C.c_int(), (C.c_int * 5)()
        _lib.check(_lib.load().oq_problem_layout(handle, C.byref(n), lens))
        self.nparts = n.value
        self.local_lengths = [lens[i] for i in range(self.nparts)]

    @property
    def handle(self):
        if self._h is None:
            raise _lib.OqError("problem already destroyed")
        return self._h

    def _ptrs(self, arrays: Sequence[np.ndarray], writable: bool):
        # the integrator calls f(du, u, p, t) with the same buffers over and over: cache the pointer tables
        key = tuple((id(a), a.ctypes.data, a.size) if isinstance(a, np.ndarray) else (id(a), 0, 0) for a in arrays) + (writable,)
        hit = self._ptr_cache.get(key)
        if hit is not None:
            return hit
        res = self._ptrs_build(arrays, writable)
        if all(k is a for k, a in zip(res[1], arrays)):      # only cache when no temporary copy was made
            if len(self._ptr_cache) > 64:
                self._ptr_cache.clear()
            self._ptr_cache[key] = res
        return res

    def _ptrs_build(self, arrays: Sequence[np.ndarray], writable: bool):
        assert len(arrays) == self.nparts, f"expected {self.nparts} state partitions"
        out = (_lib.c_double_p * self.nparts)()
        keep = []
        for i, a in enumerate(arrays):
            if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags["F_CONTIGUOUS"]):
                assert not writable, "output partitions must be float64 column-major arrays"
                a = _lib.f64(a)
            assert a.size == self.local_lengths[i], f"partition {i}: {a.size} != {self.local_lengths[i]}"
            keep.append(a)
            out[i] = _lib.dptr(a)
        return out, keep

    def rhs(self, du: Sequence[np.ndarray], u: Sequence[np.ndarray], t: float = 0.0):
        pu, k1 = self._ptrs(u, False)
        pdu, k2 = self._ptrs(du, True)
        _lib.check(_lib.load().oq_rhs(self.handle, C.c_double(t), pu, pdu))

    def set_state(self, u: Sequence[np.ndarray]):
        pu, _ = self._ptrs(u, False)
        _lib.check(_lib.load().oq_state_set(self.handle, pu))

    def get_state(self, u: Sequence[np.ndarray]):
        pu, _ = self._ptrs(u, True)
        _lib.check(_lib.load().oq_state_get(self.handle, pu))

    def get_du(self, du: Sequence[np.ndarray]):
        pu, _ = self._ptrs(du, True)
        _lib.check(_lib.load().oq_state_get_du(self.handle, pu))

    def rhs_resident(self, nevals: int) -> float:
        """nevals device-resident RHS evaluations; returns the CUDA-event time in ms."""
        ms = C.c_double()
        _lib.check(_lib.load().oq_rhs_resident(self.handle, int(nevals), C.byref(ms)))
        return ms.value

    def profile_enable(self, on: bool = True):
        _lib.check(_lib.load().oq_profile_enable(self.handle, int(on)))

    def profile_read(self):
        """(summed matvec kernel time in ms, number of launches) since the last read"""
        ms, n = C.c_double(), C.c_int64()
        _lib.check(_lib.load().oq_profile_read(self.handle, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def rhs_bytes(self) -> float:
        b = C.c_double()
        _lib.check(_lib.load().oq_rhs_bytes(self.handle, C.byref(b)))
        return b.value

    def comm_export(self, rank: int, world: int) -> bytes:
        buf = (C.c_uint8 * _lib.HANDLE_BYTES)()
        _lib.check(_lib.load().oq_comm_export(self.handle, int(rank), int(world), buf))
        return bytes(buf)

    def comm_connect(self, handles: Sequence[bytes]):
        blob = b"".join(handles)
        arr = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        _lib.check(_lib.load().oq_comm_connect(self.handle, arr))

    def free(self):
        if self._h is not None:
            _lib.load().oq_problem_destroy(self._h)
            self._h = None
            self._keep = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


@dataclass
class ODEProblem:
    """SciML's ODEProblem{true}(f, u0, tspan, p) as `assemble` returns it (equation.jl:88,116,153)."""
    f: Callable
    u0: ArrayPartition
    tspan: tuple
    p: DeviceProblem


def ode(du: ArrayPartition, u: ArrayPartition, p: DeviceProblem, t: float):
    """The in-place RHS (equation.jl:156-205): dispatch on the parameter shape happened in `assemble`."""
    p.rhs(du.x, u.x, t)


def _as_device(m, row_kind, rows, keep):
    if isinstance(m, DeviceMatrix):
        return m
    d = device_from_host(m, row_kind=row_kind, rows=rows)
    keep.append(d)
    return d


def assemble(*args, se=DieterichStateLaw(), gf11_form: str = "dense", fault_rows=None, mantle_elems=None,
             **kwargs) -> ODEProblem:
    """assemble(gf, p, u0, tspan)                                   fault only        equation.jl:81-89
    assemble(gf, p, dila, u0, tspan)                              with dilatancy    equation.jl:108-117
    assemble(gf₁₁, gf₁₂, gf₂₁, gf₂₂, pf, pa, u0, tspan)            viscoelastic      equation.jl:141-154

    Green's arguments may be the host arrays the reference's builders return (either form of gf₁₁) or
    `DeviceMatrix` shards already resident in HBM.  `gf11_form`: "dense" (row-sharded dense matvec) or
    "fft" (the reference's translation-invariant form, evaluated on the device from the Toeplitz kernel).
    `fault_rows` / `mantle_elems` select this rank's shard when host arrays are given (default: all).
    """
    assert isinstance(se, DieterichStateLaw), "only the aging law exists in the reference (equation.jl:279)"
    lib = _lib.load()
    keep: list = []
    h = C.c_void_p()
    if len(args) == 4 or (len(args) == 5 and isinstance(args[2], DilatancyProperty)):
        if len(args) == 4:
            gf, pf, u0, tspan = args
            dila = None
        else:
            gf, pf, dila, u0, tspan = args
        assert isinstance(pf, RateStateQuasiDynamicProperty)
        nx, nxi = u0.x[0].shape
        pfs = pf.c_struct()
        dls = dila.c_struct() if dila is not None else None
        g11, st = None, None
        if isinstance(gf, DeviceMatrix):
            g11, form = gf, 0
        elif gf11_form == "fft":
            st, form = _toeplitz_from_gf(gf), 1
        else:
            stt = _toeplitz_from_gf(gf)
            r0, r1 = fault_rows if fault_rows is not None else (0, nx * nxi)
            hh = C.c_void_p()
            _lib.check(lib.oq_matrix_from_toeplitz(_lib.dptr(stt), nx, nxi, int(r0), int(r1), C.byref(hh)))
            g11, form = DeviceMatrix(hh), 0
            keep.append(g11)
        _lib.check(lib.oq_problem_create_fault(nx, nxi, form, g11.handle if g11 else None,
                                               _lib.dptr(st) if st is not None else None, C.byref(pfs),
                                               C.byref(dls) if dls is not None else None, C.byref(h)))
        keep += [g11, st, pf, dila]
        shapes = [a.shape for a in u0.x]
        return ODEProblem(ode, u0, tuple(tspan), DeviceProblem(h, keep, shapes))
    if len(args) == 8:
        gf11, gf12, gf21, gf22, pf, pa, u0, tspan = args
        assert isinstance(pf, RateStateQuasiDynamicProperty) and isinstance(pa, ViscosityProperty)
        nx, nxi = u0.x[0].shape
        ne = u0.x[2].shape[0]
        g12 = _as_device(gf12, "mantle", mantle_elems, keep)
        g21 = _as_device(gf21, "fault", fault_rows, keep)
        g22 = _as_device(gf22, "mantle", mantle_elems, keep)
        g11, st = None, None
        if isinstance(gf11, DeviceMatrix):
            g11, form = gf11, 0
        elif gf11_form == "fft":
            st, form = _toeplitz_from_gf(gf11), 1
        else:
            stt = _toeplitz_from_gf(gf11)
            r0, r1 = fault_rows if fault_rows is not None else (0, nx * nxi)
            hh = C.c_void_p()
            _lib.check(lib.oq_matrix_from_toeplitz(_lib.dptr(stt), nx, nxi, int(r0), int(r1), C.byref(hh)))
            g11, form = DeviceMatrix(hh), 0
        pfs, pas = pf.c_struct(), pa.c_struct()
        _lib.check(lib.oq_problem_create_viscoelastic(nx, nxi, ne, form, g11.handle if g11 else None,
                                                      _lib.dptr(st) if st is not None else None, g12.handle,
                                                      g21.handle, g22.handle, C.byref(pfs), C.byref(pas),
                                                      C.byref(h)))
        keep += [g11, g12, g21, g22, st, pf, pa]
        shapes = [a.shape for a in u0.x]
        return ODEProblem(ode, u0, tuple(tspan), DeviceProblem(h, keep, shapes))
    raise TypeError("no method matching assemble for these arguments")


# ------------------------------------------------------------------------------------------------ solve
class Tsit5:
    """Tsitouras 5(4) Runge-Kutta pair (the algorithm of the reference's tests, test/tests.jl:11)."""
    code = 0


class VCABM5:
    """Variable-coefficient Adams-Bashforth-Moulton PECE of order 5, started with four Tsit5 steps (the
    algorithm of the reference's example, examples/otf-with-mantle.jl:160-162): 2 RHS evaluations per step."""
    code = 1


def _alg_code(alg) -> int:
    if alg is Tsit5 or isinstance(alg, Tsit5):
        return Tsit5.code
    if alg is VCABM5 or isinstance(alg, VCABM5):
        return VCABM5.code
    raise TypeError(f"unsupported algorithm {alg!r}: Tsit5() and VCABM5() are implemented")


@dataclass
class ODESolution:
    t: List[float] = field(default_factory=list)
    u: List[ArrayPartition] = field(default_factory=list)
    du: List[ArrayPartition] = field(default_factory=list)
    retcode: str = "Default"
    stats: dict = field(default_factory=dict)


def solve(prob: ODEProblem, alg=Tsit5(), *, reltol=1e-3, abstol=1e-6, dt=0.0, dtmax=0.0, maxiters=int(1e5),
          stride: int = 1, save_everystep: bool = True, callback: Optional[Callable] = None,
          adaptive: bool = True, local_u0: Optional[Sequence[np.ndarray]] = None,
          async_snapshots: Optional[bool] = None) -> ODESolution:
    """Device-resident counterpart of OrdinaryDiffEq's `solve(prob, alg; reltol, abstol, dt, dtmax,
    maxiters)` for alg = Tsit5() or VCABM5() (defaults reltol=1e-3, abstol=1e-6 as in OrdinaryDiffEq).  `callback(u, t, step)` plays the
    role of wsolve's FunctionCallingCallback (src/io.jl:128-130): it fires at t0 and after every
    `stride`-th accepted step.  `local_u0`: this rank's slices of the state for multi-GPU runs.
    `async_snapshots`: deliver snapshots through the device-side ring of oq_solve (the integration does not wait for
    the callback; a stop request takes effect within one batch of steps).  Default: on when nothing can ask for a
    stop (no user callback), off otherwise."""
    code = _alg_code(alg)
    p = prob.p
    parts0 = list(local_u0) if local_u0 is not None else list(prob.u0.x)
    p.set_state(parts0)
    shapes = [np.shape(a) for a in parts0]
    sol = ODESolution()
    if async_snapshots is None:
        async_snapshots = callback is None
    opts = _lib.OqSolveOptions(float(reltol), float(abstol), float(dt), float(dtmax), float(prob.tspan[1]),
                               int(maxiters), code, 0 if adaptive else 1, 1 if async_snapshots else 0, 0)
    stats = _lib.OqSolveStats()

    def _snap(user, t, step, pu, pdu):
        try:
            u = ArrayPartition(*[np.ctypeslib.as_array(pu[i], shape=(int(np.prod(shapes[i])),)).copy()
                                 .reshape(shapes[i], order="F") for i in range(p.nparts)])
            if save_everystep or callback is None:
                sol.t.append(t)
                sol.u.append(u)
            if callback is not None:
                du = ArrayPartition(*[np.ctypeslib.as_array(pdu[i], shape=(int(np.prod(shapes[i])),)).copy()
                                      .reshape(shapes[i], order="F") for i in range(p.nparts)])
                return 1 if callback(u, t, step, du) else 0
            return 0
        except Exception as exc:      # never unwind through the C frame
            sol.retcode = f"CallbackError: {exc!r}"
            return 1

    need_cb = save_everystep or callback is not None
    cfn = _lib.SNAPSHOT_FN(_snap) if need_cb else C.cast(None, _lib.SNAPSHOT_FN)
    _lib.check(_lib.load().oq_solve(p.handle, C.c_double(prob.tspan[0]), C.byref(opts), int(stride), cfn, None,
                                    C.byref(stats)))
    if not need_cb or not save_everystep:
        u = [np.zeros(s, order="F") for s in shapes]
        p.get_state(u)
        sol.t.append(stats.t)
        sol.u.append(ArrayPartition(*u))
    if sol.retcode == "Default":
        sol.retcode = {0: "Success", 1: "MaxIters", 2: "Unstable", 3: "PeerTimeout", 4: "Terminated"}.get(stats.retcode, "Failure")
    sol.stats = dict(naccept=stats.naccept, nreject=stats.nreject, nf=stats.nrhs, t=stats.t,
                     dt_last=stats.dt_last, dt_next=stats.dt_next)
    return sol
