"""GPU parity of the CLASS FORM of the mantle operands (csrc/classmat.cuh): a Green's matrix kept as the table of its
distinct kernels must (a) hold exactly the entries of the dense builder, (b) multiply like it (the matvecmul! slot,
pref.jl:15-21) and (c) give the oracle's RHS (equation.jl:185-205, 1e-10 per component) in every mix of dense and
class-form operands, on row shards, in resident mode and through the integrator."""
import numpy as np
import pytest

import workloads as W
from helpers import meshes
from oracle import ref

pytestmark = pytest.mark.gpu

FS = W.FaultSpec(32e3, 8e3, 1e3, 1e3)                       # 32 x 8 fault cells of 1 km
BS = W.BoxSpec(-16e3, -6e3, -8e3, 32e3, 12e3, -20e3, 8, 3, 4, tuple(np.cumprod(np.ones(4) * 1.3)))   # 4 km cells in x


def _close(got, want, tol=1e-10):
    got, want = np.asarray(got), np.asarray(want)
    den = np.maximum(np.abs(want), 1e-6 * np.max(np.abs(want)) + 1e-300)
    return float(np.max(np.abs(got - want) / den)) < tol


def _builders(oq, mf, ma, form, elems=None, rows=None):
    g12 = oq.device_fault_mantle(mf, ma, W.LAM, W.MU, buffer_ratio=1.0, elems=elems, form=form[0])
    g21 = oq.device_mantle_fault(ma, mf, W.LAM, W.MU, rows=rows, form=form[1])
    g22 = oq.device_mantle_mantle(ma, W.LAM, W.MU, elems=elems, form=form[2])
    return g12, g21, g22


def test_class_form_holds_the_dense_entries_and_multiplies_like_them(gpu, monkeypatch):
    oq = gpu
    _, mf, _, ma = meshes(oq, FS, BS)
    dense = _builders(oq, mf, ma, ("dense",) * 3)
    cls = _builders(oq, mf, ma, ("classes",) * 3)
    rng = np.random.default_rng(5)
    for d, c in zip(dense, cls):
        assert d.form()["form"] == "dense" and c.form()["form"] == "classes"
        assert c.form()["device_bytes"] > 0          # (on this 96-cell mesh the padded tables are no smaller than the dense shard)
        assert (c.local_rows, c.cols, c.global_rows) == (d.local_rows, d.cols, d.global_rows)
        # (a) the entries: bit for bit (same table, same representatives)
        assert np.array_equal(c.to_host(), d.to_host())
        assert np.array_equal(c.rows_to_host(3, 11), d.rows_to_host(3, 11))
        # (b) the product, plain and accumulating
        x = rng.standard_normal(d.cols)
        yd, yc = d.gemv(x), c.gemv(x)
        assert np.max(np.abs(yd - yc)) <= 1e-13 * np.max(np.abs(yd))
        y0 = rng.standard_normal(d.local_rows) * np.max(np.abs(yd))
        yd2, yc2 = d.gemv(x, y0.copy()), c.gemv(x, y0.copy())
        assert np.max(np.abs(yd2 - yc2)) <= 1e-13 * np.max(np.abs(yd2))
        assert np.array_equal(c.gemv(x), yc)                                   # deterministic
        # the general kernel (validation twin of the diagonal fast path the 6x6 operand takes on this grid)
        monkeypatch.setenv("OQ_CLASSMV", "generic")
        yg = c.gemv(x)
        monkeypatch.delenv("OQ_CLASSMV")
        assert np.max(np.abs(yd - yg)) <= 1e-13 * np.max(np.abs(yd))


def test_class_form_row_shards_equal_the_full_operand(gpu):
    oq = gpu
    _, mf, _, ma = meshes(oq, FS, BS)
    ne, nf = len(ma), mf.nx * mf.nxi
    full = _builders(oq, mf, ma, ("classes",) * 3)
    rng = np.random.default_rng(6)
    xs = [rng.standard_normal(m.cols) for m in full]
    ys = [m.gemv(x) for m, x in zip(full, xs)]
    e0, e1, r0, r1 = 17, 61, 40, 172
    part = _builders(oq, mf, ma, ("classes",) * 3, elems=(e0, e1), rows=(r0, r1))
    nel = e1 - e0
    for idx in (0, 2):                                                           # mantle rows: k*nel + e_local
        got = part[idx].gemv(xs[idx]).reshape(6, nel)
        assert np.array_equal(got, ys[idx].reshape(6, ne)[:, e0:e1])
    assert np.array_equal(part[1].gemv(xs[1]), ys[1][r0:r1])
    assert nf == full[1].local_rows


def _problem_inputs(mf_o, ma_o, seed=11):
    a, b, L, sig = W.fault_properties(mf_o.x, mf_o.z, mf_o.nx, mf_o.nxi)
    g, n, d0 = W.mantle_properties(ma_o.cz)
    rng = np.random.default_rng(seed)
    v, th, eps, sg, dl = W.initial_state(mf_o.nx, mf_o.nxi, L, ma_o.cz, g, n, rng=rng)
    v = v * (1 + 0.3 * rng.uniform(-1, 1, v.shape))
    sg = sg * (1 + 0.2 * rng.uniform(-1, 1, sg.shape))
    return (a, b, L, sig), (g, n, d0), (v, th, eps, sg, dl)


@pytest.mark.parametrize("forms", [("classes", "classes", "classes"), ("dense", "dense", "classes"),
                                   ("classes", "dense", "dense"), ("dense", "classes", "dense"),
                                   ("classes", "classes", "dense")])
@pytest.mark.parametrize("gf11_form", ["dense", "fft"])
def test_rhs_with_class_form_operands_matches_the_oracle(gpu, forms, gf11_form):
    oq = gpu
    mf_o, mf, ma_o, ma = meshes(oq, FS, BS)
    (a, b, L, sig), (g, n, d0), (v, th, eps, sg, dl) = _problem_inputs(mf_o, ma_o)
    o11 = ref.gf_fault_fault(mf_o, W.LAM, W.MU, buffer_ratio=1.0)
    o12 = ref.gf_fault_mantle(mf_o, ma_o, W.LAM, W.MU, buffer_ratio=1.0)
    o21 = ref.gf_mantle_fault(ma_o, mf_o, W.LAM, W.MU)
    o22 = ref.gf_mantle_mantle(ma_o, W.LAM, W.MU)
    want = ref.rhs_viscoelastic(ref.FaultProp(a, b, L, sig, W.ETA, W.VPL, W.F0, W.V0), ref.MantleProp(g, n, d0),
                                o11, o12, o21, o22, v, th, sg, form="toeplitz")
    pf = oq.RateStateQuasiDynamicProperty(a, b, L, sig, W.ETA, W.VPL, W.F0, W.V0)
    pa = oq.PowerLawViscosityProperty(g, n, d0)
    u0 = oq.ArrayPartition(v, th, eps, sg, dl)
    gf11 = oq.stress_greens_function(mf, W.LAM, W.MU, buffer_ratio=1.0)
    g12, g21, g22 = _builders(oq, mf, ma, forms)
    prob = oq.assemble(gf11, g12, g21, g22, pf, pa, u0, (0.0, 1.0), gf11_form=gf11_form)
    du = u0.similar()
    prob.f(du, u0, prob.p, 0.0)
    for gt, w in zip(du.x, want):
        assert gt.shape == w.shape
        assert _close(gt, w, 1e-9), (forms, gf11_form)       # hex8 operands: the oracle's own libm log/atan differ in the last bit
    # against the all-dense product (same entries, other summation order of 600-term sums that cancel): 1e-10 per component
    d12, d21, d22 = _builders(oq, mf, ma, ("dense",) * 3)
    probd = oq.assemble(gf11, d12, d21, d22, pf, pa, u0, (0.0, 1.0), gf11_form=gf11_form)
    dud = u0.similar()
    probd.f(dud, u0, probd.p, 0.0)
    for gt, w in zip(du.x, dud.x):
        assert _close(gt, w, 1e-10), (forms, gf11_form)
    # resident mode (graph replay) delivers the same numbers
    prob.p.set_state(u0.x)
    prob.p.rhs_resident(6)
    du2 = u0.similar()
    prob.p.get_du(du2.x)
    for a_, b_ in zip(du.x, du2.x):
        assert np.array_equal(a_, b_)


def test_solve_with_class_form_operands_follows_the_dense_solution(gpu):
    oq = gpu
    mf_o, mf, ma_o, ma = meshes(oq, W.C2_FAULT, W.box_for(8, 3, 3))
    (a, b, L, sig), (g, n, d0), _ = _problem_inputs(mf_o, ma_o)
    v, th, eps, sg, dl = W.initial_state(mf_o.nx, mf_o.nxi, L, ma_o.cz, g, n)
    pf = oq.RateStateQuasiDynamicProperty(a, b, L, sig, W.ETA, W.VPL, W.F0, W.V0)
    pa = oq.PowerLawViscosityProperty(g, n, d0)
    u0 = oq.ArrayPartition(v, th, eps, sg, dl)
    gf11 = oq.stress_greens_function(mf, W.LAM, W.MU, buffer_ratio=1.0)
    sols = []
    for form in ("dense", "classes"):
        g12, g21, g22 = _builders(oq, mf, ma, (form,) * 3)
        prob = oq.assemble(gf11, g12, g21, g22, pf, pa, u0, (0.0, 1e-2 * W.YEAR))
        for alg in (oq.Tsit5(), oq.VCABM5()):
            sols.append(oq.solve(prob, alg, reltol=1e-6, abstol=1e-8, dt=1e-8, dtmax=0.2 * W.YEAR, maxiters=400,
                                 save_everystep=False))
    for dense, cls in ((sols[0], sols[2]), (sols[1], sols[3])):
        assert dense.retcode == cls.retcode == "Success"
        # the two forms round their sums differently (1e-12 per component): the controller may place a step elsewhere,
        # the solutions must agree far inside the integration tolerance
        same_grid = (dense.stats["naccept"], dense.stats["nreject"]) == (cls.stats["naccept"], cls.stats["nreject"])
        assert abs(dense.stats["naccept"] - cls.stats["naccept"]) <= 2, (dense.stats, cls.stats)
        for x, y in zip(dense.u[-1].x, cls.u[-1].x):          # (strain components that vanish by symmetry hold round-off only)
            # same step grid: the round-off of 35 steps x 6 stages; another grid: two solutions within reltol = 1e-6 of the true one
            err = np.max(np.abs(y - x)) / np.max(np.abs(x))
            assert err <= (1e-7 if same_grid else 1e-5), (err, same_grid, dense.stats, cls.stats)


def test_mesh_without_translation_classes_keeps_the_dense_form(gpu):
    oq = gpu
    _, mf, _, ma = meshes(oq, FS, BS)
    rng = np.random.default_rng(2)
    import copy
    mb = copy.copy(ma)
    for name in ("cx", "cy", "cz", "qx", "qy", "qz", "dx", "dy", "dz"):      # jitter every cell: no two pairs are alike
        arr = np.array(getattr(ma, name), dtype=np.float64)
        setattr(mb, name, arr * (1 + 1e-3 * rng.uniform(-1, 1, arr.shape)))
    with pytest.raises(oq._lib.OqError, match="translation classes"):
        oq.device_mantle_mantle(mb, W.LAM, W.MU, form="classes")
    assert oq.device_mantle_mantle(mb, W.LAM, W.MU).form()["form"] == "dense"
