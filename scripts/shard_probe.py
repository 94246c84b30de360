"""Single-GPU probe of the matvec at multi-GPU shard sizes (no communication): rows [0, n) of the 256x64 problem."""
import sys
sys.path.insert(0, ".")
import numpy as np
import oetqf_b200 as oq
import workloads as W

oq.init(0)
fs = W.C3_FAULT
mf = oq.gen_mesh("RectOkada", fs.x, fs.xi, fs.dx, fs.dxi, fs.dip)
a, b, L, sig = W.fault_properties(mf.x, mf.z, mf.nx, mf.nxi)
v, th, dl = W.initial_state(mf.nx, mf.nxi, L, rng=np.random.default_rng(42))
pf = oq.RateStateQuasiDynamicProperty(a, b, L, sig, W.ETA, W.VPL, W.F0, W.V0)
u0 = oq.ArrayPartition(v, th, dl)
for nrows in (16384, 8192, 4096, 2048):
    g11 = oq.device_fault_fault(mf, W.LAM, W.MU, buffer_ratio=1.0, rows=(0, nrows))
    prob = oq.assemble(g11, pf, u0, (0.0, 1.0))
    p = prob.p
    p.set_state(oq.dist.local_state(u0.x, (0, nrows)))
    p.rhs_resident(10)
    ms = p.rhs_resident(400) / 400
    p.profile_enable(True)
    p.rhs_resident(200)
    mv, n = p.profile_read()
    p.profile_enable(False)
    ideal = 8.0 * nrows * 16384 / 6462.1e9 * 1e6
    print(f"rows {nrows:6d}: rhs {ms*1e3:7.1f} us  matvec {mv/n*1e3:7.1f} us  ideal@6462GB/s {ideal:6.1f} us  "
          f"-> {8.0*nrows*16384/(mv/n*1e-3)/1e9:7.0f} GB/s")
    p.free(); g11.free()
