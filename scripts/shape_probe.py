"""Single-GPU probe: the streaming matvec at equal bytes but different row lengths (rows x cols of a fault-only
dense problem): does throughput depend on the number of chunks per row block?"""
import sys
sys.path.insert(0, ".")
import numpy as np
import oetqf_b200 as oq
import workloads as W

oq.init(0)
for nx, nxi, nrows in ((256, 64, 16384), (512, 64, 8192), (1024, 64, 4096), (2048, 64, 2048), (128, 64, 8192), (64, 64, 4096)):
    fs = W.FaultSpec(nx * 250.0, nxi * 250.0, 250.0, 250.0)
    mf = oq.gen_mesh("RectOkada", fs.x, fs.xi, fs.dx, fs.dxi, fs.dip)
    a, b, L, sig = W.fault_properties(mf.x, mf.z, mf.nx, mf.nxi)
    v, th, dl = W.initial_state(mf.nx, mf.nxi, L, rng=np.random.default_rng(42))
    pf = oq.RateStateQuasiDynamicProperty(a, b, L, sig, W.ETA, W.VPL, W.F0, W.V0)
    u0 = oq.ArrayPartition(v, th, dl)
    nf = nx * nxi
    nrows = min(nrows, nf)
    g11 = oq.device_fault_fault(mf, W.LAM, W.MU, buffer_ratio=1.0, rows=(0, nrows))
    prob = oq.assemble(g11, pf, u0, (0.0, 1.0))
    p = prob.p
    p.set_state(oq.dist.local_state(u0.x, (0, nrows)))
    p.rhs_resident(10)
    ms = p.rhs_resident(200) / 200
    gb = 8.0 * nrows * nf / 1e9
    print(f"{nrows:6d} rows x {nf:6d} cols ({gb:5.2f} GB, {nf // 1024:3d} chunks/row block): rhs {ms*1e3:7.1f} us -> {gb / ms * 1e3 / 1e3:6.3f} TB/s", flush=True)
    p.free(); g11.free()
