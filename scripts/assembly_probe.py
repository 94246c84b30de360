"""Assembly kernels at a moderate size (for ncu captures and timing): K1..K4 once each."""
import sys
sys.path.insert(0, ".")
import numpy as np
import oetqf_b200 as oq
import workloads as W
from oetqf_b200 import gf as gfmod

oq.init(0)
fs = W.C3_FAULT
mf = oq.gen_mesh("RectOkada", fs.x, fs.xi, fs.dx, fs.dxi, fs.dip)
for ft in (oq.StrikeSlip(), oq.DipSlip(), oq.StrikeSlip(), oq.DipSlip()):
    st = oq.stress_greens_function(mf, W.LAM, W.MU, buffer_ratio=1.0, fourier=False, ftype=ft)
    print("K1", type(ft).__name__, gfmod.last_kernel_ms["value"], "ms", st.size / gfmod.last_kernel_ms["value"] * 1e3, "entries/s")
fsm = W.FaultSpec(64e3, 16e3, 1000.0, 1000.0)
mfm = oq.gen_mesh("RectOkada", fsm.x, fsm.xi, fsm.dx, fsm.dxi, fsm.dip)
ma = oq.gen_mesh("BEMHex8Mesh", *W.box_for(32, 8, 8, fsm).args())
import ctypes as C
def kms(m):
    ms = C.c_double(); oq._lib.check(oq._lib.load().oq_matrix_kernel_ms(m.handle, C.byref(ms))); return ms.value
for name, b in (("K2", lambda: oq.device_fault_mantle(mfm, ma, W.LAM, W.MU, buffer_ratio=1.0)),
                ("K3", lambda: oq.device_mantle_fault(ma, mfm, W.LAM, W.MU)),
                ("K4", lambda: oq.device_mantle_mantle(ma, W.LAM, W.MU))):
    m = b(); ms = kms(m)
    print(name, (m.local_rows, m.cols), ms, "ms", m.local_rows * m.cols / ms * 1e3, "entries/s"); m.free()
