"""oetqf_b200 -- B200-native implementation of Oetqf.jl's two data-parallel hot paths.

Host-side mirror of the reference's public entry points for those paths (same names and argument
meaning); every number is produced by the sm_100a kernels in csrc/ through the C ABI declared in
include/oetqf_b200.h.  There is no CPU fallback.
"""
from . import _lib, dist
from ._lib import (OqError, host_register, host_unregister, init, kernel_launch_count, measure_fp64_peak,
                   measure_hbm_copy)
from .equation import ArrayPartition, DeviceProblem, ODEProblem, ODESolution, Tsit5, VCABM5, assemble, ode, solve
from . import io
from .io import wsolve
from .gf import (max_real_eigval, DeviceMatrix, DipSlip, StrikeSlip, dc3d_gradient, device_fault_fault, device_fault_mantle,
                 device_from_host, device_mantle_fault, device_mantle_mantle, gauss_legendre_hex,
                 get_quadrature, hex8_pair_classes, stress_greens_function, stress_vol_hex8)
from .mesh import BEMHex8Mesh, RectOkadaMesh, gen_box_hex8, gen_mesh
from .pref import get_matvecmul, matvecmul, set_matvecmul
from .property import (CompositePowerLawViscosityProperty, DieterichStateLaw, DilatancyProperty,
                       PowerLawViscosityProperty, RateStateQuasiDynamicProperty)

__all__ = [n for n in dir() if not n.startswith("_")]
