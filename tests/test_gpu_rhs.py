"""GPU parity: the ODE right-hand side through the C ABI vs the CPU oracle (1e-10 relative per component,
BASELINE.json north_star) and the reference's own FFT == dense test (test/BEM/tests.jl:39-61)."""
import numpy as np
import pytest

import workloads as W
from helpers import meshes, scaled_err
from oracle import ref

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _fault_setup(oq, spec, seed=42):
    mf_o, mf_p = meshes(oq, spec)
    a, b, L, sig = W.fault_properties(mf_o.x, mf_o.z, mf_o.nx, mf_o.nxi)
    rng = np.random.default_rng(seed)
    v, th, dl = W.initial_state(mf_o.nx, mf_o.nxi, L, rng=rng)
    v = v * (1 + 0.3 * rng.uniform(-1, 1, v.shape))
    pf_o = ref.FaultProp(a, b, L, sig, W.ETA, W.VPL, W.F0, W.V0)
    pf_p = oq.RateStateQuasiDynamicProperty(a, b, L, sig, W.ETA, W.VPL, W.F0, W.V0)
    return mf_o, mf_p, pf_o, pf_p, v, th, dl


def _close(got, want, tol=TOL):
    """per-component relative error, with the field's max as the floor for components crossing zero"""
    got, want = np.asarray(got), np.asarray(want)
    den = np.maximum(np.abs(want), 1e-6 * np.max(np.abs(want)))
    return float(np.max(np.abs(got - want) / den)) < tol


@pytest.mark.parametrize("form", ["dense", "fft"])
@pytest.mark.parametrize("spec", [W.C1_FAULT, W.FaultSpec(100.0, 100.0, 10.0, 10.0, 41.0)])
def test_rhs_fault_only(gpu, spec, form):
    oq = gpu
    mf_o, mf_p, pf_o, pf_p, v, th, dl = _fault_setup(oq, spec)
    st = ref.gf_fault_fault(mf_o, W.LAM, W.MU, buffer_ratio=1.0)
    want = ref.rhs_fault(pf_o, st, v, th, form="toeplitz")
    want_fft = ref.rhs_fault(pf_o, ref.gf_fourier(st), v, th, form="fft")    # the reference's own algorithm
    gf = oq.stress_greens_function(mf_p, W.LAM, W.MU, buffer_ratio=1.0)     # Fourier form, as the reference passes it
    u0 = oq.ArrayPartition(v, th, dl)
    prob = oq.assemble(gf, pf_p, u0, (0.0, 1.0), gf11_form=form)
    du = u0.similar()
    prob.f(du, u0, prob.p, 1.0)
    for g, w, wf in zip(du.x, want, want_fft):
        assert _close(g, w)
        assert _close(g, wf, 1e-8)          # FFT round-off of the CPU path itself (rtol sqrt(eps) in the reference test)


def test_fft_conv_equals_dense_reference_test(gpu):
    """test/BEM/tests.jl:39-61 on the device: Toeplitz-form dτ/dt == dense contraction, both slip types"""
    oq = gpu
    spec = W.FaultSpec(100.0, 100.0, 10.0, 10.0, 41.0)
    mf_o, mf_p, pf_o, pf_p, v, th, dl = _fault_setup(oq, spec, seed=1)
    for ft in (oq.StrikeSlip(), oq.DipSlip()):
        gf = oq.stress_greens_function(mf_p, W.LAM, W.MU, ftype=ft)
        u0 = oq.ArrayPartition(v, th, dl)
        outs = []
        for form in ("dense", "fft"):
            prob = oq.assemble(gf, pf_p, u0, (0.0, 1.0), gf11_form=form)
            du = u0.similar()
            prob.f(du, u0, prob.p, 0.0)
            outs.append(du.x[0].copy())
        assert _close(outs[0], outs[1], 1e-12)


def test_rhs_dilatancy(gpu):
    """equation.jl:168-183, 248-276 with positive random properties (test/BEM/tests.jl:73-82 shapes)"""
    oq = gpu
    spec = W.FaultSpec(10.0, 10.0, 2.0, 2.0, 90.0)
    mf_o, mf_p = meshes(oq, spec)
    rng = np.random.default_rng(4)
    shp = (mf_o.nx, mf_o.nxi)
    a, b, L, sig = (rng.uniform(0.5, 1.5, shp) for _ in range(4))
    tp, ed, be, p0 = (rng.uniform(0.5, 1.5, shp) for _ in range(4))
    v, th, dl, pr = (rng.uniform(0.5, 1.5, shp) for _ in range(4))
    pr *= 0.1
    st = ref.gf_fault_fault(mf_o, 1.0, 1.0, buffer_ratio=1.0)
    pf_o = ref.FaultProp(a, b, L, sig, 0.7, 0.3, 0.6, 0.9)
    want = ref.rhs_fault_dilatancy(pf_o, ref.DilatancyProp(tp, ed, be, p0), st, v, th, pr, form="toeplitz")
    pf_p = oq.RateStateQuasiDynamicProperty(a, b, L, sig, 0.7, 0.3, 0.6, 0.9)
    dila = oq.DilatancyProperty(tp, ed, be, p0)
    gf = oq.stress_greens_function(mf_p, 1.0, 1.0, buffer_ratio=1.0)
    u0 = oq.ArrayPartition(v, th, dl, pr)
    for form in ("dense", "fft"):
        prob = oq.assemble(gf, pf_p, dila, u0, (0.0, 1.0), gf11_form=form)
        du = u0.similar()
        prob.f(du, u0, prob.p, 1.0)
        dv, dth, ddl, dpr = want
        for g, w in zip(du.x, (dv, dth, ddl, dpr)):
            assert _close(g, w)


@pytest.mark.parametrize("nlaws", [1, 2])
def test_rhs_viscoelastic_machinery(gpu, nlaws):
    """equation.jl:185-205 with the Okada-built gf11/gf12 and random gf21/gf22 (isolates the RHS machinery
    from the hex8 closed form); power-law and composite viscosity"""
    oq = gpu
    mf_o, mf_p, ma_o, ma_p = meshes(oq, W.C2_FAULT, W.C2_BOX)
    a, b, L, sig = W.fault_properties(mf_o.x, mf_o.z, mf_o.nx, mf_o.nxi)
    g, n, d0 = W.mantle_properties(ma_o.cz)
    rng = np.random.default_rng(8)
    v, th, eps, sg, dl = W.initial_state(mf_o.nx, mf_o.nxi, L, ma_o.cz, g, n, rng=rng)
    sg = sg * (1 + 0.2 * rng.uniform(-1, 1, sg.shape))
    nf, ne = 32, 36
    st = ref.gf_fault_fault(mf_o, W.LAM, W.MU, buffer_ratio=1.0)
    gf12 = ref.gf_fault_mantle(mf_o, ma_o, W.LAM, W.MU, buffer_ratio=1.0)
    gf21 = np.asfortranarray(rng.standard_normal((nf, 6 * ne)) * 1e9)
    gf22 = np.asfortranarray(rng.standard_normal((6 * ne, 6 * ne)) * 1e9)
    gam = np.stack([g * (1 + 0.1 * l) for l in range(nlaws)])
    npw = np.stack([n - 0.5 * l for l in range(nlaws)])
    pa_o = ref.MantleProp(gam, npw, d0)
    pf_o = ref.FaultProp(a, b, L, sig, W.ETA, W.VPL, W.F0, W.V0)
    want = ref.rhs_viscoelastic(pf_o, pa_o, st, gf12, gf21, gf22, v, th, sg, form="toeplitz")
    pf_p = oq.RateStateQuasiDynamicProperty(a, b, L, sig, W.ETA, W.VPL, W.F0, W.V0)
    laws = [oq.PowerLawViscosityProperty(gam[l], npw[l], d0) for l in range(nlaws)]
    pa_p = laws[0] if nlaws == 1 else oq.CompositePowerLawViscosityProperty(laws, d0)
    u0 = oq.ArrayPartition(v, th, eps, sg, dl)
    gf11 = oq.stress_greens_function(mf_p, W.LAM, W.MU, buffer_ratio=1.0)
    gf12p = oq.stress_greens_function(mf_p, ma_p, W.LAM, W.MU, buffer_ratio=1.0)
    for form in ("dense", "fft"):
        prob = oq.assemble(gf11, gf12p, gf21, gf22, pf_p, pa_p, u0, (0.0, 1.0), gf11_form=form)
        du = u0.similar()
        prob.f(du, u0, prob.p, 0.0)
        for gt, w in zip(du.x, want):
            assert gt.shape == w.shape
            assert _close(gt, w), form
    # resident mode gives the same derivative
    prob.p.set_state(u0.x)
    prob.p.rhs_resident(1)
    du2 = u0.similar()
    prob.p.get_du(du2.x)
    for a_, b_ in zip(du.x, du2.x):
        assert np.array_equal(a_, b_)


def test_rhs_is_deterministic_and_rearms(gpu):
    """the split-row partial sums are combined in a fixed order: bitwise repeatable across evaluations"""
    oq = gpu
    mf_o, mf_p, pf_o, pf_p, v, th, dl = _fault_setup(oq, W.C1_FAULT)
    gf = oq.stress_greens_function(mf_p, W.LAM, W.MU, buffer_ratio=1.0)
    u0 = oq.ArrayPartition(v, th, dl)
    prob = oq.assemble(gf, pf_p, u0, (0.0, 1.0))
    outs = []
    for _ in range(3):
        du = u0.similar()
        prob.f(du, u0, prob.p, 0.0)
        outs.append(du.x[0].copy())
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[1], outs[2])


def test_rhs_large_linearity(gpu):
    """BASELINE configs[2] size (256x64, 2.1 GB dense matrix): size-independent properties instead of the
    oracle -- dense form == Toeplitz form, and linearity of dτ/dt in (v - vpl)."""
    oq = gpu
    mf_o, mf_p, pf_o, pf_p, v, th, dl = _fault_setup(oq, W.C3_FAULT)
    gf = oq.stress_greens_function(mf_p, W.LAM, W.MU, buffer_ratio=1.0, fourier=False)
    u0 = oq.ArrayPartition(v, th, dl)
    res = {}
    for form in ("dense", "fft"):
        prob = oq.assemble(gf, pf_p, u0, (0.0, 1.0), gf11_form=form)
        du = u0.similar()
        prob.f(du, u0, prob.p, 0.0)
        res[form] = [x.copy() for x in du.x]
        prob.p.free()
    for a_, b_ in zip(res["dense"], res["fft"]):
        assert _close(a_, b_, 1.5e-8)      # FFT round-off: the reference's own test uses rtol = sqrt(eps) (test/BEM/tests.jl:58)
    # spot-check 64 rows against the oracle's Toeplitz contraction
    st_o = np.asfortranarray(gf)
    relv = v - W.VPL
    rows = np.random.default_rng(0).integers(0, v.size, 64)
    dtau = np.array([sum(st_o[np.abs(i - np.arange(mf_o.nx)), j, l] @ relv[:, l] for l in range(mf_o.nxi))
                     for i, j in zip(rows % mf_o.nx, rows // mf_o.nx)])
    # (the pointwise law is covered at small sizes; here the matvec itself is checked)
    prob = oq.assemble(gf, pf_p, u0, (0.0, 1.0), gf11_form="dense")
    G = prob.p._keep[0]
    y = G.gemv(relv.reshape(-1, order="F"))
    assert _close(y[rows], dtau, 1e-11)


def test_fft_form_non_power_of_two_and_shard(gpu):
    """FFT form with nx = 250 (transform length 512 > 2nx-1 = 499) and on a row shard: identical dτ/dt rows"""
    oq = gpu
    spec = W.FaultSpec(250 * 250.0, 16 * 250.0, 250.0, 250.0)
    mf_o, mf_p, pf_o, pf_p, v, th, dl = _fault_setup(oq, spec, seed=9)
    st = ref.gf_fault_fault(mf_o, W.LAM, W.MU, buffer_ratio=1.0)
    want = ref.rhs_fault(pf_o, st, v, th, form="toeplitz")
    u0 = oq.ArrayPartition(v, th, dl)
    prob = oq.assemble(st, pf_p, u0, (0.0, 1.0), gf11_form="fft")
    du = u0.similar()
    prob.f(du, u0, prob.p, 0.0)
    for g, w in zip(du.x, want):
        assert _close(g, w, 1.5e-8)          # FFT round-off, the reference's own sqrt(eps) criterion
    # dense shard of the same problem: rows [1000, 3000)
    r0, r1 = 1000, 3000
    g11 = oq.device_fault_fault(mf_p, W.LAM, W.MU, buffer_ratio=1.0, rows=(r0, r1))
    probs = oq.assemble(g11, pf_p, u0, (0.0, 1.0))
    loc = oq.dist.local_state(u0.x, (r0, r1))
    dul = [np.zeros_like(a) for a in loc]
    probs.p.rhs(dul, loc, 0.0)          # world = 1: the forcing vector only holds this shard's rows
    # (a single-rank shard sees zero forcing from the rows it does not own, so only the pointwise parts compare)
    assert np.allclose(dul[1], want[1].reshape(-1, order="F")[r0:r1], rtol=1e-12)
    assert np.array_equal(dul[2], v.reshape(-1, order="F")[r0:r1])


def _pinned_like(parts):
    """page-locked host copies of the partitions (what a caller with pinned buffers passes to `ode`)"""
    import torch
    keep, out = [], []
    for a in parts:
        t = torch.empty(a.size, dtype=torch.float64).pin_memory()
        v = t.numpy().reshape(a.shape, order="F")
        v[...] = a
        keep.append(t)
        out.append(v)
    return keep, out


@pytest.mark.parametrize("form", ["dense", "fft"])
def test_rhs_page_locked_buffers_take_the_zero_copy_path(gpu, form):
    """oq_rhs with page-locked host arrays (kernels read u / write du over PCIe themselves) == the staged path
    used for pageable arrays, bit for bit; fault-only and the 5-partition viscoelastic state"""
    oq = gpu
    mf_o, mf_p, pf_o, pf_p, v, th, dl = _fault_setup(oq, W.C1_FAULT, seed=5)
    gf = oq.stress_greens_function(mf_p, W.LAM, W.MU, buffer_ratio=1.0)
    u0 = oq.ArrayPartition(v, th, dl)
    prob = oq.assemble(gf, pf_p, u0, (0.0, 1.0), gf11_form=form)
    du = u0.similar()
    prob.f(du, u0, prob.p, 0.0)                                   # pageable numpy arrays: staged copies
    ku, pu = _pinned_like(u0.x)
    kd, pd = _pinned_like([np.full_like(a, np.nan) for a in u0.x])
    for _ in range(3):                                            # repeated calls reuse the same mapped buffers
        prob.p.rhs(pd, pu, 0.0)
    for g, w in zip(pd, du.x):
        assert np.array_equal(g, w)
    # ordinary numpy arrays registered through the ABI (what a Julia caller does with its own arrays)
    ru = [np.array(a, order="F") for a in u0.x]
    rd = [np.full_like(a, np.nan) for a in ru]
    for a in ru + rd:
        oq.host_register(a)
    try:
        prob.p.rhs(rd, ru, 0.0)
    finally:
        for a in ru + rd:
            oq.host_unregister(a)
    for g, w in zip(rd, du.x):
        assert np.array_equal(g, w)
    # viscoelastic
    mf_o, mf_p, ma_o, ma_p = meshes(oq, W.C2_FAULT, W.C2_BOX)
    a, b, L, sig = W.fault_properties(mf_o.x, mf_o.z, mf_o.nx, mf_o.nxi)
    g, n, d0 = W.mantle_properties(ma_o.cz)
    rng = np.random.default_rng(9)
    v, th, eps, sg, dl = W.initial_state(mf_o.nx, mf_o.nxi, L, ma_o.cz, g, n, rng=rng)
    pf_p = oq.RateStateQuasiDynamicProperty(a, b, L, sig, W.ETA, W.VPL, W.F0, W.V0)
    pa_p = oq.PowerLawViscosityProperty(g, n, d0)
    gf11 = oq.stress_greens_function(mf_p, W.LAM, W.MU, buffer_ratio=1.0)
    gf12 = oq.stress_greens_function(mf_p, ma_p, W.LAM, W.MU, buffer_ratio=1.0)
    gf21 = oq.stress_greens_function(ma_p, mf_p, W.LAM, W.MU)
    gf22 = oq.stress_greens_function(ma_p, W.LAM, W.MU)
    u0 = oq.ArrayPartition(v, th, eps, sg, dl)
    prob = oq.assemble(gf11, gf12, gf21, gf22, pf_p, pa_p, u0, (0.0, 1.0), gf11_form=form)
    du = u0.similar()
    prob.f(du, u0, prob.p, 0.0)
    ku, pu = _pinned_like(u0.x)
    kd, pd = _pinned_like([np.full_like(a, np.nan) for a in u0.x])
    prob.p.rhs(pd, pu, 0.0)
    for g_, w in zip(pd, du.x):
        assert np.array_equal(g_, w)


@pytest.mark.parametrize("env", [{"OQ_MATVEC": "ldg"}, {"OQ_TOEPLITZ": "direct"}, {"OQ_RHS_ZEROCOPY": "0"},
                                 {"OQ_MATVEC_KEEP_MB": "0", "OQ_MATVEC_PINGPONG": "0"}])
def test_validation_twins_stay_correct(gpu, env):
    """the alternative kernels kept behind environment switches (first LDG matvec, direct Toeplitz contraction,
    staged copies for page-locked buffers, matvec without L2 hints / alternating traversal) must pass the same RHS
    parity tests; the switches are read once per
    process, hence the subprocess"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_rhs.py"), "-m", "gpu", "-q",
                          "-x", "-k", "test_rhs_fault_only or test_rhs_viscoelastic_machinery or zero_copy"],
                         env={**os.environ, **env}, capture_output=True, text=True, timeout=600, cwd=root)
    assert res.returncode == 0, res.stdout[-2000:]
