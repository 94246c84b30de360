"""GPU parity: the device-resident Tsit5 integrator vs the CPU oracle integrator driving the CPU oracle RHS.
BASELINE.json north_star: slip-rate and state time series within 1e-6 relative over a fixed window before
the first instability."""
import numpy as np
import pytest

import workloads as W
from helpers import meshes
from oracle import integrator, ref

pytestmark = pytest.mark.gpu


def _pack(parts):
    return np.concatenate([np.asarray(p).reshape(-1, order="F") for p in parts])


def _unpack(u, shapes):
    out, off = [], 0
    for s in shapes:
        n = int(np.prod(s))
        out.append(u[off:off + n].reshape(s, order="F"))
        off += n
    return out


def test_decay_problem_matches_analytic_and_oracle(gpu):
    """the integrator alone, on the linear test problem of the reference's HDF5 test (test/tests.jl:2-7):
    a fault-only problem with zero Green's function has dθ/dt = 1 - vθ/L, dδ/dt = v, dv/dt = -(...)"""
    oq = gpu
    nx, nxi = 4, 3
    rng = np.random.default_rng(0)
    a, b, L, sig = (rng.uniform(0.5, 1.5, (nx, nxi)) for _ in range(4))
    v, th, dl = (rng.uniform(0.5, 1.5, (nx, nxi)) for _ in range(3))
    st = np.zeros((nx, nxi, nxi), order="F")
    pf_p = oq.RateStateQuasiDynamicProperty(a, b, L, sig, 0.7, 0.3, 0.6, 0.9)
    pf_o = ref.FaultProp(a, b, L, sig, 0.7, 0.3, 0.6, 0.9)
    u0 = oq.ArrayPartition(v, th, dl)
    shapes = [x.shape for x in u0.x]

    def f(u):
        vv, tt, _ = _unpack(u, shapes)
        return _pack(ref.rhs_fault(pf_o, st, vv, tt, form="toeplitz"))

    for form in ("dense", "fft"):
        prob = oq.assemble(st, pf_p, u0, (0.0, 2.0), gf11_form=form)
        sol = oq.solve(prob, oq.Tsit5(), reltol=1e-8, abstol=1e-10, dt=1e-3)
        ts, us, stats = integrator.tsit5(f, _pack(u0.x), 0.0, 2.0, reltol=1e-8, abstol=1e-10, dt0=1e-3)
        assert sol.retcode == "Success"
        assert sol.stats["naccept"] == stats["naccept"] and sol.stats["nreject"] == stats["nreject"]
        np.testing.assert_allclose(sol.t, ts, rtol=5e-6)   # the error estimate is a cancelling sum (sum of b~ is 0): FMA-vs-no-FMA round-off moves dt by ~1e-7
        for k in (len(ts) // 2, len(ts) - 1):
            np.testing.assert_allclose(_pack(sol.u[k].x), us[k], rtol=1e-6)
        assert sol.t[-1] == 2.0


def test_fault_cycle_window_matches_oracle(gpu):
    """BASELINE configs[0]-like fault (32x16), aging law, adaptive Tsit5 over a fixed window: v and θ series
    within 1e-6 relative of the CPU oracle at every accepted step"""
    oq = gpu
    spec = W.C1_FAULT
    mf_o, mf_p = meshes(oq, spec)
    a, b, L, sig = W.fault_properties(mf_o.x, mf_o.z, mf_o.nx, mf_o.nxi)
    v, th, dl = W.initial_state(mf_o.nx, mf_o.nxi, L)
    st = ref.gf_fault_fault(mf_o, W.LAM, W.MU, buffer_ratio=1.0)
    pf_o = ref.FaultProp(a, b, L, sig, W.ETA, W.VPL, W.F0, W.V0)
    pf_p = oq.RateStateQuasiDynamicProperty(a, b, L, sig, W.ETA, W.VPL, W.F0, W.V0)
    u0 = oq.ArrayPartition(v, th, dl)
    shapes = [x.shape for x in u0.x]
    tstop = 0.02 * W.YEAR

    def f(u):
        vv, tt, _ = _unpack(u, shapes)
        return _pack(ref.rhs_fault(pf_o, st, vv, tt, form="toeplitz"))

    ts, us, stats = integrator.tsit5(f, _pack(u0.x), 0.0, tstop, reltol=1e-8, abstol=1e-10, dt0=1e-6,
                                     dtmax=0.2 * W.YEAR)
    gf = oq.stress_greens_function(mf_p, W.LAM, W.MU, buffer_ratio=1.0, fourier=False)
    prob = oq.assemble(gf, pf_p, u0, (0.0, tstop), gf11_form="dense")
    sol = oq.solve(prob, oq.Tsit5(), reltol=1e-8, abstol=1e-10, dt=1e-6, dtmax=0.2 * W.YEAR)
    assert sol.retcode == "Success" and len(sol.t) == len(ts)
    np.testing.assert_allclose(sol.t, ts, rtol=5e-6)
    n = v.size
    worst_v = worst_th = 0.0
    for k in range(len(ts)):
        uk = _pack(sol.u[k].x)
        worst_v = max(worst_v, np.max(np.abs(uk[:n] - us[k][:n]) / np.abs(us[k][:n])))
        worst_th = max(worst_th, np.max(np.abs(uk[n:2 * n] - us[k][n:2 * n]) / np.abs(us[k][n:2 * n])))
    assert worst_v < 1e-6 and worst_th < 1e-6, (worst_v, worst_th)


def _check_vcabm5_on_device_grid(sol, f, u0vec, reltol, abstol, dtmax, sel, rtol_state=1e-6):
    """VCABM5 parity, step by step along the grid the device chose (see oracle/integrator.py: vcabm5_on_grid for
    why the two runs are not left to pick their own grids): (1) the oracle, stepping over the same grid with the
    divided-difference form of the formulas, reproduces the device state at EVERY accepted step; (2) every step
    the device accepted has an oracle error estimate <= 1; (3) wherever that estimate is above the round-off
    floor, the device's next step is what the PI controller prescribes."""
    ts = np.array(sol.t)
    us, ee = integrator.vcabm5_on_grid(f, u0vec, ts, reltol, abstol)
    worst = 0.0
    for k in range(len(ts)):
        uk = _pack(sol.u[k].x)
        worst = max(worst, np.max(np.abs(uk[sel] - us[k][sel]) / np.abs(us[k][sel])))
    assert worst < rtol_state, worst
    assert np.all(ee <= 1.0 + 1e-3), ee.max()
    qold, checked = 1e-4, 0
    for k in range(len(ee) - 1):
        dt, nxt = ts[k + 1] - ts[k], ts[k + 2] - ts[k + 1]
        want = min(integrator.pi_next_dt(ee[k], qold, dt), dtmax)
        qold = max(ee[k], 1e-4)
        if ee[k] > 1e-2 and k + 2 < len(ts) - 1:          # above the noise floor, not the tstop-clipped last step
            if sol.stats["nreject"] == 0:
                assert abs(nxt - want) <= 5e-3 * want, (k, nxt, want, ee[k])
            else:
                assert nxt <= want * (1 + 5e-3), (k, nxt, want, ee[k])   # a rejected attempt may have shrunk it
            checked += 1
    return worst, checked


def test_vcabm5_decay_problem_matches_oracle(gpu):
    """the multistep integrator (examples/otf-with-mantle.jl:160-162 asks for VCABM5) on the linear problem of
    test/tests.jl:2-7: device Lagrange form vs the oracle's divided-difference form of the same Adams pair"""
    oq = gpu
    nx, nxi = 4, 3
    rng = np.random.default_rng(3)
    a, b, L, sig = (rng.uniform(0.5, 1.5, (nx, nxi)) for _ in range(4))
    v, th, dl = (rng.uniform(0.5, 1.5, (nx, nxi)) for _ in range(3))
    st = np.zeros((nx, nxi, nxi), order="F")
    pf_p = oq.RateStateQuasiDynamicProperty(a, b, L, sig, 0.7, 0.3, 0.6, 0.9)
    pf_o = ref.FaultProp(a, b, L, sig, 0.7, 0.3, 0.6, 0.9)
    u0 = oq.ArrayPartition(v, th, dl)
    shapes = [x.shape for x in u0.x]

    def f(u):
        vv, tt, _ = _unpack(u, shapes)
        return _pack(ref.rhs_fault(pf_o, st, vv, tt, form="toeplitz"))

    _, _, free = integrator.vcabm5(f, _pack(u0.x), 0.0, 2.0, reltol=1e-8, abstol=1e-10, dt0=1e-3)
    for form in ("dense", "fft"):
        prob = oq.assemble(st, pf_p, u0, (0.0, 2.0), gf11_form=form)
        sol = oq.solve(prob, oq.VCABM5(), reltol=1e-8, abstol=1e-10, dt=1e-3)
        assert sol.retcode == "Success" and sol.t[-1] == 2.0
        n_steps = sol.stats["naccept"] + sol.stats["nreject"]
        assert sol.stats["nf"] == 1 + 6 * 4 + 2 * (n_steps - 4)   # 6 per starting step, 2 per Adams step, 1 initial
        assert abs(sol.stats["naccept"] - free["naccept"]) <= 1   # free-running oracle: same grid up to noise
        worst, checked = _check_vcabm5_on_device_grid(sol, f, _pack(u0.x), 1e-8, 1e-10, 2.0, slice(None), 1e-8)
        assert checked > 30, checked
        # fixed steps (no controller): four Tsit5 steps, then the constant-step Adams pair
        fx = oq.solve(prob, oq.VCABM5(), dt=0.01, adaptive=False)
        assert fx.retcode == "Success" and len(fx.t) == 201 and fx.stats["nf"] == 1 + 24 + 2 * 196
        usf, _ = integrator.vcabm5_on_grid(f, _pack(u0.x), np.array(fx.t))
        np.testing.assert_allclose(_pack(fx.u[-1].x), usf[-1], rtol=1e-10)


def test_vcabm5_fault_cycle_window_matches_oracle_and_tsit5(gpu):
    """configs[0]-like fault: VCABM5 v and θ series within 1e-6 of the CPU oracle at every accepted step, and the
    end state agrees with the Tsit5 run of the same window to the integration tolerance"""
    oq = gpu
    spec = W.C1_FAULT
    mf_o, mf_p = meshes(oq, spec)
    a, b, L, sig = W.fault_properties(mf_o.x, mf_o.z, mf_o.nx, mf_o.nxi)
    v, th, dl = W.initial_state(mf_o.nx, mf_o.nxi, L)
    st = ref.gf_fault_fault(mf_o, W.LAM, W.MU, buffer_ratio=1.0)
    pf_o = ref.FaultProp(a, b, L, sig, W.ETA, W.VPL, W.F0, W.V0)
    pf_p = oq.RateStateQuasiDynamicProperty(a, b, L, sig, W.ETA, W.VPL, W.F0, W.V0)
    u0 = oq.ArrayPartition(v, th, dl)
    shapes = [x.shape for x in u0.x]
    tstop = 0.02 * W.YEAR

    def f(u):
        vv, tt, _ = _unpack(u, shapes)
        return _pack(ref.rhs_fault(pf_o, st, vv, tt, form="toeplitz"))

    kw = dict(reltol=1e-8, abstol=1e-10, dtmax=0.2 * W.YEAR)
    gf = oq.stress_greens_function(mf_p, W.LAM, W.MU, buffer_ratio=1.0, fourier=False)
    prob = oq.assemble(gf, pf_p, u0, (0.0, tstop), gf11_form="fft")
    sol = oq.solve(prob, oq.VCABM5(), dt=1e-6, **kw)
    assert sol.retcode == "Success" and sol.t[-1] == tstop
    n = v.size
    _check_vcabm5_on_device_grid(sol, f, _pack(u0.x), 1e-8, 1e-10, 0.2 * W.YEAR, slice(0, 2 * n), 1e-6)
    rk = oq.solve(prob, oq.Tsit5(), dt=1e-6, **kw)
    assert sol.stats["nf"] < rk.stats["nf"]                 # the point of the multistep method
    np.testing.assert_allclose(_pack(sol.u[-1].x)[:2 * n], _pack(rk.u[-1].x)[:2 * n], rtol=1e-5)


def test_vcabm5_viscoelastic_example_matches_oracle(gpu):
    """configs[1] (the reference's example problem, fault + 36 hex8 cells) stepped with VCABM5 as the example
    does (reltol 1e-6, abstol 1e-8, dt 1e-8, dtmax 0.2 yr: otf-with-mantle.jl:160-162): v and θ within 1e-6 of the
    CPU oracle at every accepted step of a one-year window"""
    oq = gpu
    mf_o, mf_p, ma_o, ma_p = meshes(oq, W.C2_FAULT, W.C2_BOX)
    a, b, L, sig = W.fault_properties(mf_o.x, mf_o.z, mf_o.nx, mf_o.nxi)
    g, n, d0 = W.mantle_properties(ma_o.cz)
    v, th, eps, sg, dl = W.initial_state(mf_o.nx, mf_o.nxi, L, ma_o.cz, g, n)
    st = ref.gf_fault_fault(mf_o, W.LAM, W.MU, buffer_ratio=1.0)
    g12 = ref.gf_fault_mantle(mf_o, ma_o, W.LAM, W.MU, buffer_ratio=1.0)
    g21 = ref.gf_mantle_fault(ma_o, mf_o, W.LAM, W.MU)
    g22 = ref.gf_mantle_mantle(ma_o, W.LAM, W.MU)
    pf_o = ref.FaultProp(a, b, L, sig, W.ETA, W.VPL, W.F0, W.V0)
    pa_o = ref.MantleProp(g[None, :], n[None, :], d0)
    pf_p = oq.RateStateQuasiDynamicProperty(a, b, L, sig, W.ETA, W.VPL, W.F0, W.V0)
    pa_p = oq.PowerLawViscosityProperty(g, n, d0)
    u0 = oq.ArrayPartition(v, th, eps, sg, dl)
    shapes = [x.shape for x in u0.x]
    tstop = 1.0 * W.YEAR

    def f(u):
        vv, tt, _, ss, _ = _unpack(u, shapes)
        return _pack(ref.rhs_viscoelastic(pf_o, pa_o, st, g12, g21, g22, vv, tt, ss, form="toeplitz"))

    prob = oq.assemble(st, g12, g21, g22, pf_p, pa_p, u0, (0.0, tstop))
    sol = oq.solve(prob, oq.VCABM5(), reltol=1e-6, abstol=1e-8, dt=1e-8, dtmax=0.2 * W.YEAR)
    assert sol.retcode == "Success" and sol.t[-1] == tstop
    nf = v.size
    _check_vcabm5_on_device_grid(sol, f, _pack(u0.x), 1e-6, 1e-8, 0.2 * W.YEAR, slice(0, 2 * nf), 1e-6)
    # the full 5-partition end state (ϵ components that stay ~0 by symmetry carry only round-off: absolute floor)
    us, _ = integrator.vcabm5_on_grid(f, _pack(u0.x), np.array(sol.t), 1e-6, 1e-8)
    uk, uo = _pack(sol.u[-1].x), us[-1]
    np.testing.assert_allclose(uk, uo, rtol=1e-6, atol=1e-9 * np.max(np.abs(uo)))


def test_stride_and_callback(gpu):
    """wsolve's saving cadence (src/io.jl:51-58, test/tests.jl:34-41): every stride-th accepted step + t0"""
    oq = gpu
    nx, nxi = 4, 3
    rng = np.random.default_rng(1)
    a, b, L, sig = (rng.uniform(0.5, 1.5, (nx, nxi)) for _ in range(4))
    v, th, dl = (rng.uniform(0.5, 1.5, (nx, nxi)) for _ in range(3))
    st = np.zeros((nx, nxi, nxi), order="F")
    pf_p = oq.RateStateQuasiDynamicProperty(a, b, L, sig, 0.7, 0.3, 0.6, 0.9)
    u0 = oq.ArrayPartition(v, th, dl)
    prob = oq.assemble(st, pf_p, u0, (0.0, 2.0))
    full = oq.solve(prob, oq.Tsit5(), reltol=1e-8, abstol=1e-10, dt=1e-3)
    seen = []
    strided = oq.solve(prob, oq.Tsit5(), reltol=1e-8, abstol=1e-10, dt=1e-3, stride=11,
                       callback=lambda u, t, step, du: seen.append((t, step)) and False)
    nt = len(full.t)
    want = [full.t[i] for i in range(0, nt, 11)]
    if (nt - 1) % 11 != 0:
        want.append(full.t[-1])            # the final state is always delivered
    np.testing.assert_allclose([s[0] for s in seen], want, rtol=0)
    assert np.array_equal(strided.u[0].x[0], full.u[0].x[0])


def test_wsolve_store_append_stride(gpu, tmp_path):
    """test/tests.jl:1-42 restated: wsolve output == solve output, appended storage, strided storage"""
    oq = gpu
    nx, nxi = 4, 3
    rng = np.random.default_rng(2)
    a, b, L, sig = (rng.uniform(0.5, 1.5, (nx, nxi)) for _ in range(4))
    v, th, dl = (rng.uniform(0.5, 1.5, (nx, nxi)) for _ in range(3))
    st = np.zeros((nx, nxi, nxi), order="F")
    pf = oq.RateStateQuasiDynamicProperty(a, b, L, sig, 0.7, 0.3, 0.6, 0.9)
    u0 = oq.ArrayPartition(v, th, dl)
    kw = dict(reltol=1e-8, abstol=1e-10, dt=1e-3)
    prob = oq.assemble(st, pf, u0, (0.0, 2.0))
    sol = oq.solve(prob, oq.Tsit5(), **kw)
    f = str(tmp_path / "out")
    names = ["u1", "u2", "u3"]
    oq.wsolve(prob, oq.Tsit5(), f, 7, oq.io.VThetaDelta, names, "t", **kw)
    assert np.array_equal(oq.io.read(f, "t"), np.array(sol.t))
    for m, n in enumerate(names):
        x = np.stack([u.x[m] for u in sol.u], axis=-1)
        assert np.array_equal(oq.io.read(f, n), x)
    # refusing to overwrite without force (io.jl:119-123)
    assert oq.wsolve(prob, oq.Tsit5(), f, 7, oq.io.VThetaDelta, names, "t", **kw) is None
    # appended storage
    u1 = oq.ArrayPartition(*[x.copy() for x in sol.u[-1].x])
    prob2 = oq.assemble(st, pf, u1, (2.0, 3.0))
    sol2 = oq.solve(prob2, oq.Tsit5(), **kw)
    oq.wsolve(prob2, oq.Tsit5(), f, 7, oq.io.VThetaDelta, names, "t", append=True, **kw)
    assert np.array_equal(oq.io.read(f, "t"), np.concatenate([sol.t, sol2.t]))
    assert oq.io.read(f, "u2").shape == (nx, nxi, len(sol.t) + len(sol2.t))
    # strided storage
    stride = 11
    oq.wsolve(prob, oq.Tsit5(), f, 50, oq.io.VThetaDelta, names, "t", stride=stride, force=True, **kw)
    tt = oq.io.read(f, "t")
    want = [sol.t[i] for i in range(0, len(sol.t), stride)]     # every stride-th callback only (io.jl:51-58): the
    assert np.array_equal(tt, np.array(want))                   # end state is saved only if it is stride-aligned
    # the store is append-only: one chunk file per flush and dataset, nothing rewritten (io.jl:22-82)
    import json
    import os
    meta = json.load(open(os.path.join(f, "meta.json")))
    assert meta["nt"] == len(want) and meta["chunks"] == list(range(0, len(want), 50))
    oq.wsolve(prob, oq.Tsit5(), f, 7, oq.io.VThetaDelta, names, "t", force=True, **kw)
    meta = json.load(open(os.path.join(f, "meta.json")))
    assert meta["chunks"] == list(range(0, len(sol.t), 7))
    assert sorted(x for x in os.listdir(f) if x.startswith("u1.")) == [f"u1.{c:09d}.npy" for c in meta["chunks"]]
    assert np.array_equal(oq.io.read(f, "t"), np.array(sol.t))


def test_max_real_eigval(gpu):
    """replacement of the reference's eigvals print (GF.jl:291-294): Arnoldi with GPU matvecs == LAPACK"""
    oq = gpu
    _, _, _, ma_p = meshes(oq, W.C2_FAULT, W.C2_BOX)
    full = oq.stress_greens_function(ma_p, W.LAM, W.MU)
    want = float(np.max(np.linalg.eigvals(full).real))
    d22 = oq.device_mantle_mantle(ma_p, W.LAM, W.MU)
    got = oq.max_real_eigval(d22)
    assert abs(got - want) <= 1e-5 * max(abs(want), np.max(np.abs(full)) * 1e-6)


def test_example_script_runs(gpu, tmp_path):
    """docs/make.jl:5-10 runs the example as the reference's integration test; same here (short window)"""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("otf_example", os.path.join(root, "examples", "otf_with_mantle.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sol, tt, vv = mod.main(years=0.02, out=str(tmp_path / "otf"), quiet=True)
    # (stride = 100 as in examples/otf-with-mantle.jl:161: this short window only saves t0 -- the end state is stored
    # only when it falls on the stride, io.jl:51-58)
    assert sol.retcode == "Success" and tt[0] == 0.0 and abs(sol.stats["t"] - 0.02 * W.YEAR) < 1e-6
    assert len(tt) == 1 + sol.stats["naccept"] // 100
    assert vv.shape[:2] == (8, 4) and np.all(np.isfinite(vv)) and np.all(vv > 0)
