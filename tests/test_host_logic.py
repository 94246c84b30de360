"""CPU-only tests of the host-side mirror (mesh generators, quadrature, property checks, backend slot),
following the reference's own unit tests where they exist."""
import numpy as np
import pytest

import workloads as W
from oracle import ref


def test_rect_okada_mesh(oq):
    """test/BEM/tests.jl:5-13"""
    m = oq.gen_mesh("RectOkada", 100.0, 50.0, 2.0, 2.0, 33.0)
    assert m.nx == 50 and m.nxi == 25
    assert m.axi[-1][0] == m.xi[-1] - m.dxi / 2
    sd, cd = ref.sincosd(33.0)
    assert m.y[-1] == m.xi[-1] * cd and m.z[-1] == m.xi[-1] * sd
    assert m.x[-1] - m.x[0] == m.dx * (m.nx - 1)
    o = ref.fault_mesh(100.0, 50.0, 2.0, 2.0, 33.0)
    for a, b in [(m.x, o.x), (m.y, o.y), (m.z, o.z), (m.ax, o.ax), (m.axi, o.axi)]:
        assert np.array_equal(a, b)


def test_hex8_box_conventions(oq):
    """test/BEM/tests.jl:15-36 (q-point convention and bounding box)"""
    rng = np.random.default_rng(5)
    llx, lly, llz = rng.random(3)
    dx, dy, dz = rng.random(3) * 5
    nx, ny, nz = (int(v) for v in rng.integers(2, 10, 3))
    me = oq.gen_mesh("BEMHex8Mesh", llx, lly, llz, dx, dy, -dz, nx, ny, nz)
    for arr, n in ((me.cx, nx), (me.cy, ny), (me.cz, nz)):
        assert len(np.unique(np.round(arr, 6))) == n
    np.testing.assert_allclose(me.cx, me.qx)
    np.testing.assert_allclose(me.cy - me.dy / 2, me.qy)
    np.testing.assert_allclose(me.cz + me.dz / 2, me.qz)
    np.testing.assert_allclose(np.max(me.cx + me.dx / 2), llx + dx)
    np.testing.assert_allclose(np.min(me.cx - me.dx / 2), llx)
    np.testing.assert_allclose(np.max(me.cy + me.dy / 2), lly + dy)
    np.testing.assert_allclose(np.min(me.cy - me.dy / 2), lly)
    np.testing.assert_allclose(np.max(me.cz + me.dz / 2), llz)
    np.testing.assert_allclose(np.min(me.cz - me.dz / 2), llz - dz)
    o = ref.hex8_box(llx, lly, llz, dx, dy, -dz, nx, ny, nz)
    for k in ("cx", "cy", "cz", "qx", "qy", "qz", "dx", "dy", "dz"):
        np.testing.assert_array_equal(getattr(me, k), getattr(o, k))


def test_example_box_layers(oq):
    """examples/otf-with-mantle.jl:25-28: layer fractions normalize(cumsum(cumprod(1.5*ones(3))), Inf)"""
    me = oq.gen_mesh("BEMHex8Mesh", *W.C2_BOX.args())
    assert len(me) == 36
    tops = np.unique(np.round(me.cz + me.dz / 2, 6))[::-1]
    np.testing.assert_allclose(tops, -8e3 - 22e3 * np.array([0.0, 1.5 / 7.125, 3.75 / 7.125]))


def test_quadrature(oq):
    c, w = oq.get_quadrature("Gauss1")
    assert np.array_equal(c, [0, 0, 0]) and np.array_equal(w, [1.0])
    c, w = oq.get_quadrature("Gauss2")
    assert w.size == 8 and abs(w.sum() - 1) < 1e-15 and np.allclose(np.abs(c), 1 / np.sqrt(3))
    with pytest.raises(AssertionError, match="Wrong format of quadrature!"):     # GF.jl:326
        oq.get_quadrature((np.zeros(5), np.ones(2)))
    # "Gauss<N>": N is the polynomial order in Gmsh (getIntegrationPoints(5, ...), GF.jl:318-323), not the point count
    for name, npts in (("Gauss0", 1), ("Gauss1", 1), ("Gauss2", 8), ("Gauss3", 8), ("Gauss4", 27), ("Gauss5", 64)):
        c, w = oq.get_quadrature(name)
        assert w.size == npts and c.size == 3 * npts and abs(w.sum() - 1) < 1e-15, name
    co, wo = ref.gauss_quadrature(3)
    c, w = oq.get_quadrature("Gauss4")                                          # the 3x3x3 product rule
    np.testing.assert_array_equal(c, co)
    np.testing.assert_array_equal(w, wo)


def test_property_asserts(oq):
    """property.jl:20-24,39-40,47"""
    z = np.ones((3, 2))
    oq.RateStateQuasiDynamicProperty(z, z, z, z, 1.0, 1.0)
    with pytest.raises(AssertionError):
        oq.RateStateQuasiDynamicProperty(z, z, z, np.ones((2, 2)), 1.0, 1.0)
    with pytest.raises(AssertionError):
        oq.RateStateQuasiDynamicProperty(z, z, z, z, -1.0, 1.0)
    with pytest.raises(AssertionError):
        oq.PowerLawViscosityProperty(np.ones(3), np.ones(3), np.ones(5))
    with pytest.raises(AssertionError):
        oq.PowerLawViscosityProperty(np.ones(3), np.ones(2), np.ones(6))


def test_gemv_backend_setting(oq):
    """test/tests.jl:60-64 restated for the single B200 backend"""
    assert oq.get_matvecmul() == "B200"
    oq.set_matvecmul("B200")
    with pytest.raises(ValueError):
        oq.set_matvecmul("DummyMatVec")


def test_buffer_ratio_assert(oq):
    mf = oq.gen_mesh("RectOkada", 10.0, 10.0, 2.0, 2.0, 90.0)
    with pytest.raises(AssertionError, match="buffer_ratio"):                   # GF.jl:36
        oq.stress_greens_function(mf, 1.0, 1.0, buffer_ratio=-1.0)


def test_workload_shapes():
    assert (W.C1_FAULT.nx, W.C1_FAULT.nxi) == (32, 16)
    assert (W.C2_FAULT.nx, W.C2_FAULT.nxi, W.C2_BOX.n) == (8, 4, 36)
    assert (W.C3_FAULT.nx, W.C3_FAULT.nxi) == (256, 64)
    mf = ref.fault_mesh(W.C2_FAULT.x, W.C2_FAULT.xi, W.C2_FAULT.dx, W.C2_FAULT.dxi, 90.0)
    a, b, L, s = W.fault_properties(mf.x, mf.z, mf.nx, mf.nxi)
    assert (b > a).sum() > 0 and (b < a).sum() > 0          # both VW and VS cells exist
    ma = ref.hex8_box(*W.C2_BOX.args())
    g, n, d0 = W.mantle_properties(ma.cz)
    v, th, eps, sig, dl = W.initial_state(mf.nx, mf.nxi, L, ma.cz, g, n)
    rel = g * (np.sqrt(2) * np.abs(sig[:, 1])) ** n * sig[:, 1]
    np.testing.assert_allclose(rel, d0[1], rtol=1e-10)       # examples/otf-with-mantle.jl:147-148


def test_algorithm_selection(oq):
    """solve/wsolve accept the two algorithms the device integrator provides (Tsit5: test/tests.jl:11; VCABM5:
    examples/otf-with-mantle.jl:160), as instances or classes, and refuse anything else before touching the GPU"""
    import pytest
    from oetqf_b200.equation import _alg_code
    assert _alg_code(oq.Tsit5()) == 0 and _alg_code(oq.Tsit5) == 0
    assert _alg_code(oq.VCABM5()) == 1 and _alg_code(oq.VCABM5) == 1
    with pytest.raises(TypeError):
        _alg_code("RK4")
    with pytest.raises(TypeError):
        oq.wsolve(None, object(), "/nonexistent", 1, None, [], "t")


def test_solve_option_struct_matches_header():
    """OqSolveOptions / OqSolveStats field order and sizes as declared in include/oetqf_b200.h"""
    import ctypes as C
    from oetqf_b200 import _lib
    assert [f[0] for f in _lib.OqSolveOptions._fields_] == ["reltol", "abstol", "dt0", "dtmax", "tstop", "maxiters",
                                                           "algorithm", "fixed_dt", "async_snapshots", "reserved"]
    assert C.sizeof(_lib.OqSolveOptions) == 5 * 8 + 8 + 4 * 4
    assert [f[0] for f in _lib.OqSolveStats._fields_] == ["t", "dt_last", "dt_next", "naccept", "nreject", "nrhs",
                                                         "retcode"]


# ---- class decomposition of the hex8 builders (csrc/greens_classes.cuh), host logic, no device -------------------
def _check_classes(oq, ma, mf, begin, end, rng, n=4000):
    nrecv_total = len(ma) if mf is None else mf.nx * mf.nxi
    recv = rng.integers(begin, end, n).astype(np.int32)
    src = rng.integers(0, len(ma), n).astype(np.int32)
    counts, rjx, rix, rjyz, riyz = oq.hex8_pair_classes(ma, mf, begin, end, recv, src)
    assert counts[2] == (end - begin) * len(ma) and end <= nrecv_total
    if counts[0] == 0:
        return counts
    ext = max(np.ptp(ma.qx) + ma.dx.max(), np.ptp(ma.qy) + ma.dy.max(), 1.0)
    tol = 4e-12 * ext
    if mf is None:      # receivers = cells: centre and half sizes enter (quadrature points are centre + qc * half size)
        x_r, y_r, z_r = ma.cx, ma.cy, ma.cz
        np.testing.assert_allclose(ma.dx[recv], ma.dx[rjx], rtol=2e-12)
        np.testing.assert_allclose(ma.dy[recv], ma.dy[rjyz], rtol=2e-12)
        np.testing.assert_allclose(ma.dz[recv], ma.dz[rjyz], rtol=2e-12)
    else:               # receivers = fault cells f = i + j nx
        x_r = np.tile(mf.x, mf.nxi)
        y_r, z_r = np.repeat(mf.y, mf.nx), np.repeat(mf.z, mf.nx)
    # the pair and its stand-in agree in everything the closed form reads
    np.testing.assert_allclose(x_r[recv] - ma.qx[src], x_r[rjx] - ma.qx[rix], atol=tol, rtol=0)
    np.testing.assert_allclose(ma.dx[src], ma.dx[rix], rtol=2e-12)
    np.testing.assert_allclose(y_r[recv] - ma.qy[src], y_r[rjyz] - ma.qy[riyz], atol=tol, rtol=0)
    np.testing.assert_allclose(z_r[recv], z_r[rjyz], atol=tol, rtol=0)
    np.testing.assert_allclose(ma.qz[src], ma.qz[riyz], atol=tol, rtol=0)
    np.testing.assert_allclose(ma.dy[src], ma.dy[riyz], rtol=2e-12)
    np.testing.assert_allclose(ma.dz[src], ma.dz[riyz], rtol=2e-12)
    return counts


def test_hex8_pair_classes_structured_box(oq):
    """transfinite box (mesh.jl:95-130), uniform in x and y, graded in z: (2nx-1)(2ny-1) nz^2 classes for nx ny nz
    squared pairs; row shards (the multi-GPU builders) decompose on their own receivers"""
    rng = np.random.default_rng(11)
    ma = oq.gen_mesh("BEMHex8Mesh", *W.box_for(12, 6, 5).args())
    counts = _check_classes(oq, ma, None, 0, len(ma), rng)
    assert counts[0] == 2 * 12 - 1 and counts[1] == (2 * 6 - 1) * 5 * 5
    counts = _check_classes(oq, ma, None, 100, 217, rng)
    assert 0 < counts[0] * counts[1] < counts[2]


def test_hex8_pair_classes_fault_receivers(oq):
    """mantle -> fault: vertical and dipping faults (y and z of a fault row are tied together)"""
    rng = np.random.default_rng(12)
    ma = oq.gen_mesh("BEMHex8Mesh", *W.box_for(8, 3, 4).args())
    for dip in (90.0, 60.0):
        mf = oq.gen_mesh("RectOkada", W.C2_FAULT.x, W.C2_FAULT.xi, W.C2_FAULT.dx, W.C2_FAULT.dxi, dip)
        counts = _check_classes(oq, ma, mf, 0, mf.nx * mf.nxi, rng)
        assert counts[0] > 0 and counts[0] * counts[1] <= counts[2]
        _check_classes(oq, ma, mf, 5, mf.nx * mf.nxi - 3, rng)


def test_hex8_pair_classes_irregular_mesh(oq):
    """cells of random position and size share nothing: every pair is its own class (the builders then keep the
    tiled kernels), and the stand-in of a pair is the pair itself"""
    rng = np.random.default_rng(13)
    n = 40
    c = rng.random((3, n)) * 1e4
    d = 100.0 + rng.random((3, n)) * 500.0
    ma = oq.BEMHex8Mesh(c[0], c[1], -2e4 - c[2], c[0].copy(), c[1] - d[1] / 2, -2e4 - c[2] + d[2] / 2, d[0], d[1], d[2])
    counts = _check_classes(oq, ma, None, 0, n, rng)
    assert counts[0] * counts[1] >= counts[2]


def test_class_form_plan(oq):
    """the host-side plan of the class-form mantle->mantle operand (csrc/classmat.cuh): an equidistant x grid takes
    the diagonal kernel (x class = function of the position difference: 2 nx - 1 classes), one run per (y,z) row of
    receivers, also on row shards; a graded x grid keeps its classes but not the diagonal structure; an irregular
    mesh has nothing worth a class form"""
    from oetqf_b200 import gf
    nx, ny, nz = 12, 5, 4
    ma = oq.gen_mesh("BEMHex8Mesh", *W.box_for(nx, ny, nz).args())
    plan = gf.class_form_plan(ma)
    assert plan["worthwhile"] and plan["diagonal"]
    assert plan["x_classes"] == 2 * nx - 1 and plan["x_positions"] == nx and plan["max_source_group"] == nx
    assert plan["yz_classes"] == (2 * ny - 1) * nz * nz and plan["runs"] == ny * nz == plan["receiver_yz_classes"]
    # a shard that starts and ends inside a row of cells: partial rows are runs of their own
    shard = gf.class_form_plan(ma, (nx * 3 + 5, nx * 9 + 2))
    assert shard["diagonal"] and shard["worthwhile"] and shard["runs"] == 7 == shard["receiver_yz_classes"]
    assert shard["x_classes"] == plan["x_classes"]
    # cells growing along x: offsets no longer repeat along x (no diagonal structure), y still does
    xe = np.concatenate([[0.0], np.cumsum(1000.0 * 1.2 ** np.arange(nx))])
    j, i = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    i, j = i.ravel(), j.ravel()
    cx, dx = (xe[i] + xe[i + 1]) / 2, xe[i + 1] - xe[i]
    cy, dy = -2e3 + 1e3 * j + 500.0, np.full(i.size, 1e3)
    cz, dz = np.full(i.size, -9e3), np.full(i.size, 2e3)
    graded = oq.BEMHex8Mesh(cx, cy, cz, cx.copy(), cy - dy / 2, cz + dz / 2, dx, dy, dz)
    gp = gf.class_form_plan(graded)
    assert not gp["diagonal"] and gp["x_classes"] == nx * nx and gp["yz_classes"] == 2 * ny - 1
    # irregular cells
    rng = np.random.default_rng(3)
    n = 30
    c = rng.random((3, n)) * 1e4
    d = 100.0 + rng.random((3, n)) * 500.0
    irr = oq.BEMHex8Mesh(c[0], c[1], -2e4 - c[2], c[0].copy(), c[1] - d[1] / 2, -2e4 - c[2] + d[2] / 2, d[0], d[1], d[2])
    ip = gf.class_form_plan(irr)
    assert not ip["worthwhile"] and not ip["diagonal"]


def test_class_window_plans_of_the_fault_mantle_operands(oq):
    """gf21 / gf12 in class form take the sliding-window kernel per residue of the finer grid (fault cells 4 or 5 to a
    mantle cell): every (receiver, source) pair reached through the plan's maps gets the class the general maps give it,
    every receiver is reached, on full operands and on row shards"""
    from oetqf_b200 import gf
    cases = [(W.FaultSpec(32e3, 8e3, 1e3, 1e3), W.BoxSpec(-16e3, -6e3, -8e3, 32e3, 12e3, -20e3, 8, 3, 4, tuple(np.cumprod(np.ones(4) * 1.3))), 4),
             (W.FaultSpec(50 * 250.0, 8 * 250.0, 250.0, 250.0), W.BoxSpec(-6250.0, -10e3, -2e3, 12500.0, 20e3, -40e3, 10, 5, 4, ()), 5)]
    for fs, bs, q in cases:
        mf = oq.gen_mesh("RectOkada", fs.x, fs.xi, fs.dx, fs.dxi, fs.dip)
        ma = oq.gen_mesh("BEMHex8Mesh", *bs.args())
        nf, ne = mf.nx * mf.nxi, len(ma)
        for which, n in ((1, nf), (2, ne)):
            for b, e in ((0, n), (n // 3 + 1, 2 * n // 3 + 2)):
                res = gf.class_window_check(ma, mf, which, b, e)
                assert res["found"] == 1 and res["residues"] == q, (which, b, e, res)
                assert res["mismatches"] == 0 and res["unreached"] == 0 and res["checked"] >= (e - b) * (ne if which == 1 else nf)
    # a mantle grid that is not commensurate with the fault's: no plan (the operand keeps the general kernel)
    fs = W.FaultSpec(32e3, 8e3, 1e3, 1e3)
    mf = oq.gen_mesh("RectOkada", fs.x, fs.xi, fs.dx, fs.dxi, fs.dip)
    ma = oq.gen_mesh("BEMHex8Mesh", -16e3, -6e3, -8e3, 32e3, 12e3, -20e3, 7, 3, 2, None)
    assert gf.class_window_check(ma, mf, 1)["found"] == 0
