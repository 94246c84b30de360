"""Pins the CPU oracle's Okada restatement (no GPU): published check values, physics invariants and the
known-answer vectors of SURVEY.md Appendix E.  The reference's own tests hold no numeric golden values
for dc3d (parity with GeoGreensFunctions.jl is unpinned); these are the strongest pins available."""
import json
import os

import numpy as np
import pytest

from oracle import ref

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_okada_1985_checklist():
    # Okada (1985) table 2, case 2: x=2, y=3, d=4, dip=70, L=3, W=2 (surface, alpha=2/3)
    u = ref.dc3d(2 / 3, 2, 3, 0, 4, 70, 0, 3, 0, 2, 1, 0, 0)
    np.testing.assert_allclose(u[:3], [-8.689e-3, -4.298e-3, -2.747e-3], rtol=6e-4)
    np.testing.assert_allclose(u[3:6], [-1.220e-3, -8.191e-3, -5.175e-3], rtol=6e-4)
    np.testing.assert_allclose(u[6:9], [2.470e-4, -5.814e-4, 2.945e-4], rtol=6e-4)
    u = ref.dc3d(2 / 3, 2, 3, 0, 4, 70, 0, 3, 0, 2, 0, 1, 0)
    np.testing.assert_allclose(u[:3], [-4.682e-3, -3.527e-2, -3.564e-2], rtol=6e-4)
    u = ref.dc3d(2 / 3, 2, 3, 0, 4, 70, 0, 3, 0, 2, 0, 0, 1)
    np.testing.assert_allclose(u[:3], [-2.660e-4, 1.056e-2, 3.214e-3], rtol=6e-4)


@pytest.mark.parametrize("dip", [0.0, 10.0, 33.0, 41.0, 70.0, 90.0])
def test_gradients_match_finite_differences(dip):
    rng = np.random.default_rng(int(dip) + 1)
    h = 1e-5
    worst = 0.0
    for _ in range(40):
        x, y = rng.uniform(-8, 8, 2)
        z = -rng.uniform(0.3, 9)
        d = rng.uniform(0.0, 1.0, 3)
        args = (5.0, dip, -1.5, 2.5, -3.0, -0.5, *d)
        u = ref.dc3d(0.6, x, y, z, *args)
        scale = np.max(np.abs(u[3:])) + 1e-30
        for ax, (dx, dy, dz) in enumerate([(h, 0, 0), (0, h, 0), (0, 0, h)]):
            up = ref.dc3d(0.6, x + dx, y + dy, z + dz, *args)
            um = ref.dc3d(0.6, x - dx, y - dy, z - dz, *args)
            fd = (up[:3] - um[:3]) / (2 * h)
            worst = max(worst, np.max(np.abs(fd - u[3 + 3 * ax: 6 + 3 * ax])) / scale)
    assert worst < 2e-5      # finite-difference limited


@pytest.mark.parametrize("dip", [0.0, 25.0, 90.0])
def test_free_surface_is_traction_free(dip):
    rng = np.random.default_rng(7)
    lam = mu = 1.0
    alpha = (lam + mu) / (lam + 2 * mu)
    for _ in range(30):
        x, y = rng.uniform(-10, 10, 2)
        for d in ((1, 0, 0), (0, 1, 0), (0, 0, 1)):
            u = ref.dc3d(alpha, x, y, 0.0, 4.0, dip, -2.0, 1.0, -2.5, -0.5, *d)
            ekk = u[3] + u[7] + u[11]
            sxz, syz = mu * (u[5] + u[9]), mu * (u[8] + u[10])
            szz = lam * ekk + 2 * mu * u[11]
            assert max(abs(sxz), abs(syz), abs(szz)) < 1e-12 * max(np.max(np.abs(u[3:])), 1e-30) * 100


def test_singular_and_above_surface_return_zero():
    assert np.all(ref.dc3d(0.6, 0.3, 0.0, -3.0, 4.0, 90.0, -1, 1, -1, 1, 1, 0, 0) == 0)   # on the top edge
    assert np.any(ref.dc3d(0.6, 0.0, 0.0, -4.0, 4.0, 90.0, -1, 1, -1, 1, 1, 0, 0) != 0)   # fault interior is regular
    assert np.all(ref.dc3d(0.6, 1.0, 0.0, -3.0, 4.0, 90.0, -1, 1, -1, 1, 1, 0, 0) == 0)   # on a corner
    assert np.all(ref.dc3d(0.6, 0.3, 2.0, 0.5, 4.0, 90.0, -1, 1, -1, 1, 1, 0, 0) == 0)    # z > 0


def test_appendix_e_known_answers():
    mf = ref.fault_mesh(100.0, 100.0, 10.0, 10.0, 41.0)              # test/BEM/tests.jl:43
    st = ref.gf_fault_fault(mf, 3e10, 3e10, ftype=0, nrept=2, buffer_ratio=0)
    np.testing.assert_allclose([st[0, 0, 0], st[1, 0, 0], st[0, 1, 0], st[9, 9, 0], st[4, 2, 6]],
                               [-2.498417697363009e+09, 5.866850913807347e+08, 3.082244751811971e+08,
                                4.400446523811102e+05, 2.891353409008250e+06], rtol=1e-12)
    st = ref.gf_fault_fault(mf, 3e10, 3e10, ftype=1, nrept=2, buffer_ratio=0)
    np.testing.assert_allclose([st[0, 0, 0], st[1, 0, 0], st[0, 1, 0], st[9, 9, 0], st[4, 2, 6]],
                               [-1.855046154287142e+09, 2.300380786902002e+08, 5.803730065996473e+08,
                                1.318904121574405e+06, 1.791409522708632e+06], rtol=1e-12)
    mf = ref.fault_mesh(80e3, 8e3, 10e3, 2e3, 90.0)                  # examples/otf-with-mantle.jl:18
    st = ref.gf_fault_fault(mf, 3e10, 3e10, ftype=0, nrept=2, buffer_ratio=1)
    np.testing.assert_allclose([st[0, 0, 0], st[1, 0, 0], st[7, 3, 0], st[0, 3, 3], st[2, 1, 2]],
                               [-6.946206284624549e+06, 2.606430253554716e+05, 6.278958444302949e+02,
                                -9.837388526290858e+06, 1.898762473777621e+04], rtol=1e-12)


def test_fft_equals_dense_toeplitz():
    """The reference's own numerical test (test/BEM/tests.jl:39-61) on the oracle: FFT conv == dense."""
    mf = ref.fault_mesh(100.0, 100.0, 10.0, 10.0, 41.0)
    rng = np.random.default_rng(3)
    for ft in (0, 1):
        st = ref.gf_fault_fault(mf, 3e10, 3e10, ftype=ft)
        relv = rng.random((mf.nx, mf.nxi)) - 0.1
        a = ref.dtau_dt_fft(ref.gf_fourier(st), relv)
        b = ref.dtau_dt_toeplitz(st, relv)
        c = ref.gemv(ref.dense_from_toeplitz(st), relv.reshape(-1, order="F")).reshape(relv.shape, order="F")
        np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-6 * np.max(np.abs(b)) * 1e-6)
        np.testing.assert_allclose(c, b, rtol=1e-12, atol=1e-12 * np.max(np.abs(b)))


def test_viscosity_law_exact():
    """test/BEM/tests.jl:142-154: dϵ_dt == A σ τ^n and the composite is the sum (bitwise)."""
    rng = np.random.default_rng(11)
    A, n, A2, n2 = rng.random(4)
    sig = rng.random((1, 6))
    skk = (sig[0, 0] + sig[0, 3] + sig[0, 5]) / 3
    dev = np.array([sig[0, 0] - skk, sig[0, 1], sig[0, 2], sig[0, 3] - skk, sig[0, 4], sig[0, 5] - skk])
    tau = np.sqrt(dev[0] ** 2 + dev[3] ** 2 + dev[5] ** 2 + 2 * (dev[1] ** 2 + dev[2] ** 2 + dev[4] ** 2))
    one = ref.update_strain_rate(ref.MantleProp(np.array([[A]]), np.array([[n]]), np.zeros(6)), sig)
    assert np.array_equal(one[0], A * dev * tau ** n)
    two = ref.update_strain_rate(ref.MantleProp(np.array([[A], [A2]]), np.array([[n], [n2]]), np.zeros(6)), sig)
    assert np.array_equal(two[0], A * dev * tau ** n + A2 * dev * tau ** n2)


def test_golden_fixtures_match_oracle():
    """tests/golden/*.json were produced by tests/golden/make_golden.py from this oracle; a change in the
    oracle's arithmetic shows up here before it can silently move the GPU parity target."""
    with open(os.path.join(GOLD, "okada_kernels.json")) as fh:
        g = json.load(fh)
    for case in g["cases"]:
        fs = case["fault"]
        mf = ref.fault_mesh(*fs)
        st = ref.gf_fault_fault(mf, g["lam"], g["mu"], ftype=case["ftype"], nrept=case["nrept"],
                                buffer_ratio=case["buffer_ratio"])
        idx = np.array(case["index"])
        got = st[idx[:, 0], idx[:, 1], idx[:, 2]]
        np.testing.assert_allclose(got, case["values"], rtol=1e-13)


@pytest.mark.parametrize("dip", [0.0, 17.0, 41.0, 70.0, 90.0])
def test_dc3d_equals_volterra_quadrature_of_the_mindlin_tensor(dip):
    """independent of Okada's tables: the displacement field of the rectangular dislocation from its definition
    (oracle/okada_numeric.py: Volterra's formula, complex-step source derivative of the half-space Green's tensor,
    64x64 Gauss-Legendre) == oracle/okada.c, for all three slip types, interior and surface receivers; the nine
    GRADIENTS -- the only rows the product evaluates -- by sixth-order central differences of the quadrature field,
    also to 1e-10"""
    from oracle.okada_numeric import dc3d_displacement, dc3d_gradient
    rng = np.random.default_rng(100 + int(dip))
    alpha = 0.6
    sd, cd = np.sin(np.radians(dip)), np.cos(np.radians(dip))
    worst_u = worst_g = 0.0
    n = 0
    while n < 10:
        x, y = rng.uniform(-6, 6, 2)
        z = -rng.uniform(0.0, 8.0) if n else 0.0                    # the first receiver sits on the free surface
        geom = (5.0, dip, -1.5, 2.5, -3.0, -0.5)
        # distance from the fault plane (through (0,0,-5), normal (0,-sd,cd)): keep the integrand smooth
        if abs(-(y * sd) + (z + 5.0) * cd) < 0.7:
            continue
        n += 1
        d = rng.uniform(-1.0, 1.0, 3)
        want = ref.dc3d(alpha, x, y, z, *geom, *d)
        got = dc3d_displacement(alpha, x, y, z, *geom, *d)
        worst_u = max(worst_u, np.max(np.abs(got - want[:3])) / np.max(np.abs(want[:3])))
        if z < -0.5 and n % 3 == 0:
            gq = dc3d_gradient(alpha, x, y, z, *geom, *d)
            worst_g = max(worst_g, np.max(np.abs(gq - want[3:])) / np.max(np.abs(want[3:])))
    assert worst_u < 1e-10, worst_u
    assert worst_g < 1e-10, worst_g
