"""ctypes binding of liboetqf_b200.so (the C ABI declared in include/oetqf_b200.h).

There is no CPU fallback: if the shared library is missing, or no B200 is visible, every compute
call raises.  PyTorch is not needed here; the library owns its device memory.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OETQF_B200_LIB", os.path.join(_HERE, "liboetqf_b200.so"))

c_double_p = C.POINTER(C.c_double)
HANDLE_BYTES = 128


class OqError(RuntimeError):
    pass


class OqFaultMesh(C.Structure):
    _fields_ = [("nx", C.c_int32), ("nxi", C.c_int32),
                ("x", c_double_p), ("ax0", c_double_p), ("ax1", c_double_p),
                ("xi", c_double_p), ("axi0", c_double_p), ("axi1", c_double_p),
                ("y", c_double_p), ("z", c_double_p),
                ("dx", C.c_double), ("dxi", C.c_double), ("dep", C.c_double), ("dip", C.c_double)]


class OqHex8Mesh(C.Structure):
    _fields_ = [("n", C.c_int32)] + [(k, c_double_p) for k in
                                     ("cx", "cy", "cz", "qx", "qy", "qz", "dx", "dy", "dz")]


class OqQuadrature(C.Structure):
    _fields_ = [("nq", C.c_int32), ("coords", c_double_p), ("weights", c_double_p)]


class OqFaultProperty(C.Structure):
    _fields_ = [("a", c_double_p), ("b", c_double_p), ("L", c_double_p), ("sigma", c_double_p),
                ("eta", C.c_double), ("vpl", C.c_double), ("f0", C.c_double), ("v0", C.c_double)]


class OqMantleProperty(C.Structure):
    _fields_ = [("nlaws", C.c_int32), ("gamma", c_double_p), ("n", c_double_p), ("deps0", c_double_p)]


class OqDilatancyProperty(C.Structure):
    _fields_ = [("tp", c_double_p), ("eps", c_double_p), ("beta", c_double_p), ("p0", c_double_p)]


class OqSolveOptions(C.Structure):
    _fields_ = [("reltol", C.c_double), ("abstol", C.c_double), ("dt0", C.c_double), ("dtmax", C.c_double),
                ("tstop", C.c_double), ("maxiters", C.c_int64), ("algorithm", C.c_int32),
                ("fixed_dt", C.c_int32), ("async_snapshots", C.c_int32), ("reserved", C.c_int32)]


class OqSolveStats(C.Structure):
    _fields_ = [("t", C.c_double), ("dt_last", C.c_double), ("dt_next", C.c_double),
                ("naccept", C.c_int64), ("nreject", C.c_int64), ("nrhs", C.c_int64), ("retcode", C.c_int32)]


class OqAssemblyInfo(C.Structure):
    _fields_ = [("path", C.c_int), ("pairs", C.c_int64), ("unique_pairs", C.c_int64),
                ("table_ms", C.c_double), ("expand_ms", C.c_double), ("kernel_ms", C.c_double)]


SNAPSHOT_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_double, C.c_int64,
                          C.POINTER(c_double_p), C.POINTER(c_double_p))

# every symbol include/oetqf_b200.h declares (tests check that the library exports all of them)
EXPORTS = [
    "oq_abi_version", "oq_last_error", "oq_init", "oq_device_count", "oq_kernel_launch_count",
    "oq_measure_fp64_peak", "oq_measure_hbm_copy", "oq_host_register", "oq_host_unregister",
    "oq_gf_fault_fault", "oq_gf_fault_mantle", "oq_gf_mantle_fault", "oq_gf_mantle_mantle",
    "oq_dc3d_gradient", "oq_stress_vol_hex8",
    "oq_matrix_fault_fault", "oq_matrix_from_toeplitz", "oq_matrix_fault_mantle", "oq_matrix_mantle_fault", "oq_matrix_mantle_mantle",
    "oq_matrix_fault_mantle_classes", "oq_matrix_mantle_fault_classes", "oq_matrix_mantle_mantle_classes", "oq_matrix_form", "oq_class_form_plan", "oq_class_window_check",
    "oq_matrix_from_host", "oq_matrix_to_host", "oq_matrix_rows_to_host", "oq_matrix_shape", "oq_matrix_kernel_ms", "oq_matrix_assembly_info", "oq_hex8_pair_classes", "oq_matrix_destroy", "oq_gemv",
    "oq_problem_create_fault", "oq_problem_create_viscoelastic", "oq_problem_destroy", "oq_problem_layout",
    "oq_profile_enable", "oq_profile_read", "oq_rhs_bytes",
    "oq_rhs", "oq_state_set", "oq_state_get", "oq_state_get_du", "oq_rhs_resident", "oq_solve",
    "oq_comm_export", "oq_comm_connect",
]

_lib = None


def load():
    """Load the shared library (no device is touched until the first compute call)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OqError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.oq_last_error.restype = C.c_char_p
    lib.oq_kernel_launch_count.restype = C.c_int64
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        raise OqError(load().oq_last_error().decode("utf-8", "replace"))


def dptr(a: np.ndarray):
    return a.ctypes.data_as(c_double_p)


def f64(a, order="F") -> np.ndarray:
    """float64 array with the reference's (column-major) memory layout."""
    return np.require(np.asarray(a, dtype=np.float64), requirements=["F" if order == "F" else "C", "A"])


def init(device: int = 0):
    check(load().oq_init(int(device)))


def kernel_launch_count() -> int:
    return int(load().oq_kernel_launch_count())


def measure_fp64_peak() -> float:
    v = C.c_double()
    check(load().oq_measure_fp64_peak(C.byref(v)))
    return v.value


def measure_hbm_copy(nbytes: int = 1 << 30) -> float:
    v = C.c_double()
    check(load().oq_measure_hbm_copy(C.c_size_t(nbytes), C.byref(v)))
    return v.value


def host_register(a: np.ndarray):
    """Page-lock and map a host array so that `ode` / oq_rhs reads and writes it without staging copies."""
    check(load().oq_host_register(C.c_void_p(a.ctypes.data), C.c_size_t(a.nbytes)))


def host_unregister(a: np.ndarray):
    check(load().oq_host_unregister(C.c_void_p(a.ctypes.data)))
