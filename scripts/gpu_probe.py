"""Ad-hoc GPU probe (not a bench): peaks, assembly and RHS timings for a few sizes."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import oetqf_b200 as oq  # noqa: E402
import workloads as W  # noqa: E402
from oetqf_b200 import gf as gfmod  # noqa: E402

oq.init(0)
out = {}
out["fp64_peak_tflops"] = oq.measure_fp64_peak() / 1e12
out["hbm_copy_gbs"] = oq.measure_hbm_copy(1 << 30) / 1e9
print(out, flush=True)

for name, fs in (("C1", W.C1_FAULT), ("C3", W.C3_FAULT)):
    mf = oq.gen_mesh("RectOkada", fs.x, fs.xi, fs.dx, fs.dxi, fs.dip)
    for rep in range(3):
        st = oq.stress_greens_function(mf, W.LAM, W.MU, buffer_ratio=1.0, fourier=False)
        ms = gfmod.last_kernel_ms["value"]
    n = st.size
    out[f"{name}_k1_ms"] = ms
    out[f"{name}_k1_entries_per_s"] = n / (ms * 1e-3)
    print(name, "K1", ms, "ms", n / ms * 1e3, "entries/s", flush=True)

# fault->mantle at a moderate size
fs = W.FaultSpec(64e3, 16e3, 1000.0, 1000.0)
mf = oq.gen_mesh("RectOkada", fs.x, fs.xi, fs.dx, fs.dxi, fs.dip)
bs = W.box_for(32, 8, 8, fs)
ma = oq.gen_mesh("BEMHex8Mesh", *bs.args())
for rep in range(2):
    m = oq.device_fault_mantle(mf, ma, W.LAM, W.MU, buffer_ratio=1.0)
    m.free()
g12 = oq.stress_greens_function(mf, ma, W.LAM, W.MU, buffer_ratio=1.0)
ms = gfmod.last_kernel_ms["value"]
print("K2", g12.shape, ms, "ms", g12.size / ms * 1e3, "entries/s", g12.size / 6 / ms * 1e3, "pairs/s", flush=True)
out["k2_ms"] = ms
out["k2_pairs_per_s"] = g12.size / 6 / ms * 1e3

# RHS fault-only dense at C3
fs = W.C3_FAULT
mf = oq.gen_mesh("RectOkada", fs.x, fs.xi, fs.dx, fs.dxi, fs.dip)
a, b, L, sig = W.fault_properties(mf.x, mf.z, mf.nx, mf.nxi)
v, th, dl = W.initial_state(mf.nx, mf.nxi, L, rng=np.random.default_rng(42))
pf = oq.RateStateQuasiDynamicProperty(a, b, L, sig, W.ETA, W.VPL, W.F0, W.V0)
st = oq.stress_greens_function(mf, W.LAM, W.MU, buffer_ratio=1.0, fourier=False)
u0 = oq.ArrayPartition(v, th, dl)
for form in ("dense", "fft"):
    t0 = time.time()
    prob = oq.assemble(st, pf, u0, (0.0, 1.0), gf11_form=form)
    prob.p.set_state(u0.x)
    prob.p.rhs_resident(3)
    ms = prob.p.rhs_resident(20) / 20
    nbytes = 8.0 * (mf.nx * mf.nxi) ** 2
    print("RHS", form, ms, "ms/eval", 1e3 / ms, "evals/s", (nbytes / (ms * 1e-3) / 1e9) if form == "dense" else "",
          "GB/s; setup", time.time() - t0, "s", flush=True)
    out[f"rhs_{form}_ms"] = ms
    prob.p.free()
with open("gpurun_out/probe.json", "w") as fh:
    json.dump(out, fh, indent=1)
