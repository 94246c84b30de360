"""GPU parity: Green's-function assembly through the C ABI vs the CPU oracle on the same inputs.

Okada paths (dc3d gradient rows, fault->fault, fault->mantle): the default kernels keep the published operation
order of DC3D without FMA contraction (csrc/okada_strict.cuh), so they must be BIT-IDENTICAL to the oracle --
asserted with ==.  The closed form is conditioned to ~1e-2 of an entry 1000 cell sizes away, so nothing weaker than
identical rounding meets BASELINE.json's 1e-10 per entry at full size.  With OQ_OKADA=fast (restructured,
FMA-contracted kernels; run as a twin by test_fast_okada_twin) the same tests use 1e-10 relative to the scale of
the receiver's / source's row."""
import json
import os

import numpy as np
import pytest

import workloads as W
from helpers import meshes, rel_err, scaled_err
from oracle import ref

pytestmark = pytest.mark.gpu
TOL = 1e-10
EXACT = os.environ.get("OQ_OKADA") != "fast"
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("dip,ftype", [(90.0, 0), (41.0, 0), (41.0, 1), (10.0, 1), (70.0, 0), (0.0, 1)])
def test_dc3d_gradient_pointwise(gpu, dip, ftype):
    oq = gpu
    rng = np.random.default_rng(17)
    n = 4000
    x, y = rng.uniform(-12, 12, n), rng.uniform(-12, 12, n)
    z = -rng.uniform(0.0, 10.0, n)
    z[:50] = 0.0                                   # free surface
    got = oq.dc3d_gradient(x, y, z, 0.6, 4.0, dip, -1.5, 2.5, -3.0, -0.5, ftype=ftype)
    d = (1.0, 0.0, 0.0) if ftype == 0 else (0.0, 1.0, 0.0)
    want = np.array([ref.dc3d(0.6, x[i], y[i], z[i], 4.0, dip, -1.5, 2.5, -3.0, -0.5, *d)[3:] for i in range(n)])
    if EXACT:
        assert np.array_equal(got, want)            # same operations in the same order: same bits
    assert scaled_err(got, want, axis=1) < TOL      # relative to the largest gradient entry of that receiver


def test_dc3d_singular_edges_and_kxi_ket(gpu):
    """receivers on fault edges (zeros), on the extension lines of edges (KXI/KET branches of the closed
    form) and above the surface: the product must take the same branches as the oracle"""
    oq = gpu
    pts = np.array([[0.3, 0.0, -3.0], [1.0, 0.0, -3.0], [0.3, 2.0, 0.5],        # edge, corner, above surface
                    [-5.0, 0.0, -3.0], [3.0, 0.0, -5.0], [-1.0, 0.0, -9.0],      # extension lines (q = 0)
                    [0.0, 1e-7, -3.5], [1.0 + 1e-7, 0.5, -4.0]])                 # inside the EPS snap distance
    for ftype in (0, 1):
        got = oq.dc3d_gradient(pts[:, 0], pts[:, 1], pts[:, 2], 0.6, 4.0, 90.0, -1, 1, -1, 1, ftype=ftype)
        d = (1.0, 0.0, 0.0) if ftype == 0 else (0.0, 1.0, 0.0)
        want = np.array([ref.dc3d(0.6, *p, 4.0, 90.0, -1, 1, -1, 1, *d)[3:] for p in pts])
        assert np.all(got[:3] == 0.0) and np.all(want[:3] == 0.0)
        if EXACT:
            assert np.array_equal(got, want)
        assert scaled_err(got, want) < TOL


@pytest.mark.parametrize("spec,ftype,nrept,br", [
    (W.FaultSpec(100.0, 100.0, 10.0, 10.0, 41.0), 0, 2, 0.0),      # test/BEM/tests.jl:43
    (W.FaultSpec(100.0, 100.0, 10.0, 10.0, 41.0), 1, 2, 0.0),
    (W.C2_FAULT, 0, 2, 1.0),                                        # examples/otf-with-mantle.jl:18,43
    (W.C1_FAULT, 0, 2, 1.0),                                        # BASELINE configs[0]
    (W.FaultSpec(16e3, 8e3, 500.0, 500.0, 60.0), 1, 1, 0.5),
    (W.FaultSpec(10.0, 10.0, 2.0, 2.0, 90.0), 0, 0, 1.0),           # test/BEM/tests.jl:64-65, no images
])
def test_fault_fault_kernel(gpu, spec, ftype, nrept, br):
    oq = gpu
    mf_o, mf_p = meshes(oq, spec)
    want = ref.gf_fault_fault(mf_o, W.LAM, W.MU, ftype=ftype, nrept=nrept, buffer_ratio=br)
    ft = oq.StrikeSlip() if ftype == 0 else oq.DipSlip()
    got = oq.stress_greens_function(mf_p, W.LAM, W.MU, ftype=ft, fourier=False, nrept=nrept, buffer_ratio=br)
    assert got.shape == want.shape
    if EXACT:
        assert np.array_equal(got, want)                            # bit-identical Toeplitz kernel
    assert rel_err(got, want) < TOL                                 # plain per-entry relative error


def test_fault_fault_golden(gpu):
    oq = gpu
    with open(os.path.join(GOLD, "okada_kernels.json")) as fh:
        g = json.load(fh)
    for case in g["cases"]:
        mf = oq.gen_mesh("RectOkada", *case["fault"])
        ft = oq.StrikeSlip() if case["ftype"] == 0 else oq.DipSlip()
        st = oq.stress_greens_function(mf, g["lam"], g["mu"], ftype=ft, fourier=False, nrept=case["nrept"],
                                       buffer_ratio=case["buffer_ratio"])
        idx = np.array(case["index"])
        assert rel_err(st[idx[:, 0], idx[:, 1], idx[:, 2]], case["values"]) < TOL


def test_fault_fault_fourier_form(gpu):
    """GF.jl:60-68: the default return is the strike-wise rFFT of the even extension"""
    oq = gpu
    mf_o, mf_p = meshes(oq, W.FaultSpec(100.0, 100.0, 10.0, 10.0, 41.0))
    want = ref.gf_fault_fault(mf_o, W.LAM, W.MU, fourier=True)
    got = oq.stress_greens_function(mf_p, W.LAM, W.MU)             # fourier=True is the reference default
    assert got.dtype == np.complex128 and got.shape == want.shape
    assert scaled_err(got, want) < TOL


def test_dense_expansion_and_shards(gpu):
    """test/BEM/tests.jl:46-49: G[(i,j),(k,l)] = st[|i-k|,j,l]; row shards tile the full matrix"""
    oq = gpu
    mf_o, mf_p = meshes(oq, W.FaultSpec(100.0, 100.0, 10.0, 10.0, 41.0))
    want = ref.dense_from_toeplitz(ref.gf_fault_fault(mf_o, W.LAM, W.MU))
    full = oq.device_fault_fault(mf_p, W.LAM, W.MU).to_host()
    if EXACT:
        assert np.array_equal(full, want)
    assert rel_err(full, want) < TOL
    nf = mf_p.nx * mf_p.nxi
    parts = [oq.device_fault_fault(mf_p, W.LAM, W.MU, rows=(a, b)).to_host()
             for a, b in ((0, 37), (37, 37), (37, nf))]
    assert np.array_equal(np.concatenate(parts, axis=0), full)


@pytest.mark.parametrize("quad", ["Gauss1", "Gauss2"])
@pytest.mark.parametrize("ftype", [0, 1])
def test_fault_mantle(gpu, quad, ftype):
    oq = gpu
    mf_o, mf_p, ma_o, ma_p = meshes(oq, W.C2_FAULT, W.C2_BOX)      # KAT-3 geometry incl. the KET line
    q = ref.gauss_quadrature(int(quad[-1]))
    want = ref.gf_fault_mantle(mf_o, ma_o, W.LAM, W.MU, ftype=ftype, quad=q, nrept=2, buffer_ratio=1.0)
    ft = oq.StrikeSlip() if ftype == 0 else oq.DipSlip()
    got = oq.stress_greens_function(mf_p, ma_p, W.LAM, W.MU, ftype=ft, qtype=quad, nrept=2, buffer_ratio=1.0)
    assert got.shape == (6 * 36, 32)
    if EXACT:
        assert np.array_equal(got, want)                            # image sum, quadrature sum and stress epilogue in the reference's order
    # per column (one source): relative to the largest stress that source produces anywhere
    assert scaled_err(got, want, axis=0) < TOL
    # user-supplied quadrature tuple (GF.jl:325-328)
    got2 = oq.stress_greens_function(mf_p, ma_p, W.LAM, W.MU, ftype=ft, qtype=q, nrept=2, buffer_ratio=1.0)
    assert np.array_equal(got, got2)


def test_fault_mantle_kat3(gpu):
    """SURVEY.md Appendix E KAT-3 (receiver on the singular line below a patch edge)"""
    oq = gpu
    _, mf_p, _, ma_p = meshes(oq, W.C2_FAULT, W.C2_BOX)
    g = oq.stress_greens_function(mf_p, ma_p, W.LAM, W.MU, qtype="Gauss1", nrept=2, buffer_ratio=1.0)
    ne = 36
    e = int(np.argmin(np.abs(ma_p.cx + 30e3) + np.abs(ma_p.cy) + np.abs(ma_p.cz + 10315.789473684)))
    col = 0 + 3 * mf_p.nx                                          # strike 1, dip 4 (1-based)
    got = g[e + np.arange(6) * ne, col]
    np.testing.assert_allclose(got, [0, -4.522721347572e+05, 0, 0, 1.453610392276e+05, 0], rtol=1e-9, atol=1e-3)


def test_fault_mantle_element_shards(gpu):
    oq = gpu
    _, mf_p, _, ma_p = meshes(oq, W.C2_FAULT, W.C2_BOX)
    full = oq.device_fault_mantle(mf_p, ma_p, W.LAM, W.MU, buffer_ratio=1.0).to_host()
    ne = 36
    got = np.zeros_like(full)
    for e0, e1 in ((0, 10), (10, 36)):
        part = oq.device_fault_mantle(mf_p, ma_p, W.LAM, W.MU, buffer_ratio=1.0, elems=(e0, e1)).to_host()
        nel = e1 - e0
        for k in range(6):
            got[k * ne + e0: k * ne + e1] = part[k * nel: (k + 1) * nel]
    assert np.array_equal(got, full)


def test_matrix_roundtrip_and_gemv(gpu):
    """the matvecmul! slot (pref.jl:15-21; equation.jl:201-203): 3-argument and α=β=true forms"""
    oq = gpu
    rng = np.random.default_rng(9)
    for m, n in ((50, 144), (144, 50), (1030, 2100), (7, 5000)):
        A = np.asfortranarray(rng.standard_normal((m, n)))
        x, y0 = rng.standard_normal(n), rng.standard_normal(m)
        d = oq.device_from_host(A)
        assert np.array_equal(d.to_host(), A)
        want = ref.gemv(A, x)
        got = d.gemv(x)
        scale = np.abs(A) @ np.abs(x)
        assert np.max(np.abs(got - want) / scale) < 1e-14
        y = y0.copy()
        oq.matvecmul(y, d, x, True, True)
        assert np.max(np.abs(y - (y0 + want)) / (scale + np.abs(y0))) < 1e-14
        r0, r1 = m // 3, m - 1
        part = oq.device_from_host(A, rows=(r0, r1))
        # a shard splits its row blocks across CTAs differently from the full matrix: same sums, other order
        assert np.max(np.abs(part.gemv(x) - got[r0:r1]) / scale[r0:r1]) < 1e-14


@pytest.mark.parametrize("ftype", [0, 1])
def test_fault_mantle_dipping_gauss3(gpu, ftype):
    """test/BEM/tests.jl:85-95 geometry with a 60-degree fault, 3x3x3 product rule as an explicit tuple
    (GF.jl:325-328), no periodic images"""
    oq = gpu
    fs = W.FaultSpec(100.0, 100.0, 10.0, 20.0, 60.0)
    bs = W.BoxSpec(-100.0, -50.0, -120.0, 200.0, 100.0, -30.0, 2, 3, 4)
    mf_o, mf_p, ma_o, ma_p = meshes(oq, fs, bs)
    q = ref.gauss_quadrature(3)
    want = ref.gf_fault_mantle(mf_o, ma_o, 1.0, 1.0, ftype=ftype, quad=q, nrept=0, buffer_ratio=0.0)
    ft = oq.StrikeSlip() if ftype == 0 else oq.DipSlip()
    got = oq.stress_greens_function(mf_p, ma_p, 1.0, 1.0, ftype=ft, qtype=q, nrept=0, buffer_ratio=0.0)
    assert got.shape == (144, 50)
    if EXACT:
        assert np.array_equal(got, want)
    assert scaled_err(got, want, axis=0) < TOL


def test_fast_okada_twin(gpu):
    """the restructured FMA-contracted Okada kernels (OQ_OKADA=fast) pass the same tests at 1e-10 of the row scale;
    the switch is read once per process, hence the subprocess"""
    import subprocess
    import sys
    if not EXACT:
        pytest.skip("already the twin")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-m", "gpu", "-q", "-x",
                          "-k", "dc3d or fault_fault or fault_mantle or dense_expansion"],
                         env={**os.environ, "OQ_OKADA": "fast"}, capture_output=True, text=True, timeout=900, cwd=root)
    assert res.returncode == 0, res.stdout[-2000:]


def test_fault_mantle_class_tables_are_bit_identical(gpu, monkeypatch):
    """K2'' (csrc/greens_classes.cuh): dc3d sees the strike coordinates only through x - al1, x - al2, so pairs whose
    differences are bitwise equal share their six entries: one evaluation per class, copied into the dense shard.
    Every entry is bit-identical to the per-pair kernel's (and hence to the oracle), Gauss1 and Gauss2 receivers,
    vertical and dipping faults, periodic images, element shards."""
    oq = gpu
    cases = [(W.FaultSpec(40e3, 8e3, 2e3, 2e3, 90.0), "Gauss1", 2, 1.0), (W.FaultSpec(40e3, 8e3, 2e3, 2e3, 90.0), "Gauss2", 2, 1.0),
             (W.FaultSpec(40e3, 8e3, 2.5e3, 2e3, 60.0), "Gauss2", 0, 0.0)]      # last: incommensurate grids
    for fs, quad, nrept, br in cases:
        mf_o, mf_p, ma_o, ma_p = meshes(oq, fs, W.box_for(10, 4, 4, fs))
        out = {}
        for mode in ("classes", "pair", ""):
            monkeypatch.setenv("OQ_FAULT_MANTLE", mode)
            m = oq.device_fault_mantle(mf_p, ma_p, W.LAM, W.MU, qtype=quad, nrept=nrept, buffer_ratio=br)
            info = m.assembly_info()
            out[mode] = m.to_host()
            m.free()
            if mode:
                assert info["path"] == mode, (mode, info)
            elif quad == "Gauss1":
                # receivers on cell centres of a uniform box: every pair is a translate of a few (Gauss points sit at
                # irrational offsets, whose differences round differently from cell to cell: fewer bitwise-equal pairs,
                # and the default keeps the per-pair kernel when fewer than 4 pairs share a class)
                assert info["path"] == "classes" and 4 * info["unique_pairs"] <= info["pairs"], info
        assert np.array_equal(out["classes"], out["pair"]) and np.array_equal(out[""], out["pair"])
        if EXACT:
            q = ref.gauss_quadrature(int(quad[-1]))
            want = ref.gf_fault_mantle(mf_o, ma_o, W.LAM, W.MU, ftype=0, quad=q, nrept=nrept, buffer_ratio=br)
            assert np.array_equal(out[""], want)
        monkeypatch.setenv("OQ_FAULT_MANTLE", "")
        ne = len(ma_p)
        for e0, e1 in ((0, 33), (33, ne)):
            part = oq.device_fault_mantle(mf_p, ma_p, W.LAM, W.MU, qtype=quad, nrept=nrept, buffer_ratio=br, elems=(e0, e1)).to_host()
            for k in range(6):
                assert np.array_equal(part[k * (e1 - e0): (k + 1) * (e1 - e0)], out[""][k * ne + e0: k * ne + e1])
