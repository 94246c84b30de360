"""Assembly kernels at a moderate size (for ncu captures and timing): K1..K4 once each; the hex8 builders on their
default path (class tables + expansion) and with every pair through the tiled / per-pair kernels."""
import os
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
modes = sys.argv[1:] or ["", "tile"]
for mode in modes:
    os.environ["OQ_HEX8"] = mode
    os.environ["OQ_FAULT_MANTLE"] = "pair" if mode in ("tile", "pair") else ""
    for name, b in (("K2", lambda: oq.device_fault_mantle(mfm, ma, W.LAM, W.MU, buffer_ratio=1.0)),
                    ("K3", lambda: oq.device_mantle_fault(ma, mfm, W.LAM, W.MU)),
                    ("K4", lambda: oq.device_mantle_mantle(ma, W.LAM, W.MU)),
                    ("K4g2", lambda: oq.device_mantle_mantle(ma, W.LAM, W.MU, qtype="Gauss2"))):
        for rep in range(2):
            m = b(); ms = m.kernel_ms(); info = m.assembly_info()
            print(name, f"mode={mode!r}", (m.local_rows, m.cols), "%.4f ms" % ms, "%.3e entries/s" % (m.local_rows * m.cols / ms * 1e3),
                  info["path"], info["unique_pairs"], "table %.4f expand %.4f ms" % (info["table_ms"], info["expand_ms"]),
                  ("expand %.0f GB/s" % (m.local_rows * m.cols * 8 / info["expand_ms"] / 1e6)) if info["expand_ms"] else "")
            m.free()
