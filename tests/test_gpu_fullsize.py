"""GPU parity at the sizes BASELINE.json names (missing #2/#6):

* configs[2] (256x64): EVERY entry of the Toeplitz kernel, EVERY entry of the 16 384 x 16 384 dense matrix and all
  16 384 rows of dv, dθ, dδ against the CPU oracle (the pattern of /root/reference/test/BEM/tests.jl:39-61);
* configs[3] scale (20 000 fault cells + 19 320 hex8 cells, the largest coupled case that fits one B200): >= 1e4
  randomly sampled entries of each of the four Green's matrices, pulled out of ROW SHARDS built at the full column
  count, against pointwise evaluations of the oracle's dc3d / stress_vol_hex8 with the loops of GF.jl:123-290.
"""
import numpy as np
import pytest

import bench
import workloads as W
from oracle import ref

pytestmark = pytest.mark.gpu
TOL = 1e-10


def test_c3_every_entry_and_every_row(gpu):
    oq = gpu
    fs = W.C3_FAULT
    nf = fs.nx * fs.nxi
    mf, prob, u0, g11 = bench.build_fault_problem(oq, fs, (0, nf))
    p = prob.p
    loc = oq.dist.local_state(u0.x, (0, nf))
    p.set_state(loc)
    p.rhs_resident(1)
    du = [np.zeros(nf) for _ in loc]
    p.get_du(du)
    res = bench.fault_parity(oq, fs, p, g11, (0, nf), du)
    assert res["kernel_entries"] == 256 * 64 * 64 and res["matrix_entries"] == nf * nf and res["rows"] == nf
    assert res["kernel_max_rel_err"] <= TOL, res
    assert res["matrix_rows_max_rel_err"] <= TOL, res
    assert res["rhs_max_rel_err"] <= TOL, res
    # a two-way row split of the same problem gives the same rows (shard builder at full size)
    half = oq.device_fault_fault(mf, W.LAM, W.MU, buffer_ratio=1.0, rows=(nf // 2, nf))
    assert np.array_equal(half.rows_to_host(100, 164), g11.rows_to_host(nf // 2 + 100, nf // 2 + 164))


# ---- pointwise restatements of the builders' loops on the oracle's point kernels -------------------------
def _okada_stress_at(mfo, q1, q2, rx, ry, rz, lam, mu, nrept, lrept):
    """GF.jl:154-169 for one (receiver point, source cell): image sum of dc3d + the 6-stress epilogue"""
    alpha = (lam + mu) / (lam + 2 * mu)
    u = np.zeros(12)
    for r in range(-nrept, nrept + 1):
        u += ref.dc3d(alpha, rx, ry, rz, mfo.dep, mfo.dip, mfo.ax[q1, 0] + r * lrept, mfo.ax[q1, 1] + r * lrept,
                      mfo.axi[q2, 0], mfo.axi[q2, 1], 1.0, 0.0, 0.0)
    exx, eyy, ezz = u[3], u[7], u[11]
    ekk = exx + eyy + ezz
    return np.array([lam * ekk + 2 * mu * exx, mu * (u[4] + u[6]), mu * (u[5] + u[9]),
                     lam * ekk + 2 * mu * eyy, mu * (u[8] + u[10]), lam * ekk + 2 * mu * ezz])


def _hex8_unit(mao, i, x, y, z, p, mu, nu):
    eps = np.zeros(6)
    eps[p] = 1.0
    return ref.stress_vol_hex8(x, y, z, mao.qx[i], mao.qy[i], mao.qz[i], mao.dx[i], mao.dy[i], mao.dz[i], eps, mu, nu)


def _scaled(got, want, scale):
    return float(np.max(np.abs(np.asarray(got) - np.asarray(want)) / scale))


def test_coupled_scale_sampled_entries(gpu):
    oq = gpu
    fs = W.FaultSpec(250 * 250.0, 80 * 250.0, 250.0, 250.0)               # 250 x 80 = 20 000 cells
    bs = W.box_for(46, 20, 21, fs)                                         # 19 320 hex8 cells
    mfo = ref.fault_mesh(fs.x, fs.xi, fs.dx, fs.dxi, fs.dip)
    mao = ref.hex8_box(*bs.args())
    mf = oq.gen_mesh("RectOkada", fs.x, fs.xi, fs.dx, fs.dxi, fs.dip)
    ma = oq.gen_mesh("BEMHex8Mesh", *bs.args())
    nf, ne = mfo.nx * mfo.nxi, mao.n
    assert nf == 20000 and ne == 19320
    lam, mu = W.LAM, W.MU
    nu = lam / 2 / (lam + mu)
    rng = np.random.default_rng(2024)
    nshard, rows_per = 6, 4

    # gf12 fault -> mantle: shards of receiver elements; 6 components x sampled source cells
    checked = 0
    worst = 0.0
    lrept = 2.0 * mfo.dx * mfo.nx
    for e0 in rng.integers(0, ne - rows_per, nshard):
        M = oq.device_fault_mantle(mf, ma, lam, mu, buffer_ratio=1.0, elems=(int(e0), int(e0) + rows_per))
        got = M.rows_to_host(0, M.local_rows)                              # local row k*rows_per + el
        assert got.shape == (6 * rows_per, nf)
        cols = rng.integers(0, nf, 80)
        scale = max(np.max(np.abs(got)), 1e-300)
        for el in range(rows_per):
            e = int(e0) + el
            for j in cols:
                want = _okada_stress_at(mfo, int(j % mfo.nx), int(j // mfo.nx), mao.cx[e], mao.cy[e], mao.cz[e],
                                        lam, mu, 2, lrept)
                g = got[np.arange(6) * rows_per + el, j]
                den = np.maximum(np.abs(want), 1e-3 * scale)                # scale-aware floor (tests/helpers.py)
                worst = max(worst, float(np.max(np.abs(g - want) / den)))
                checked += 6
        M.free()
    assert checked >= 10000 and worst <= TOL, (checked, worst)

    # gf21 mantle -> fault: shards of fault rows; 6 unit strains x sampled source elements
    checked, worst = 0, 0.0
    sd, cd = ref.sincosd(mfo.dip)
    for r0 in rng.integers(0, nf - rows_per, nshard):
        M = oq.device_mantle_fault(ma, mf, lam, mu, rows=(int(r0), int(r0) + rows_per))
        got = M.rows_to_host(0, rows_per)
        assert got.shape == (rows_per, 6 * ne)
        scale = max(np.max(np.abs(got)), 1e-300)
        srcs = rng.integers(0, ne, 80)
        for fl in range(rows_per):
            f = int(r0) + fl
            x, y, z = mfo.x[f % mfo.nx], mfo.y[f // mfo.nx], mfo.z[f // mfo.nx]
            for i in srcs:
                for pc in range(6):
                    S = _hex8_unit(mao, int(i), x, y, z, pc, mu, nu)
                    want = -S[1] * sd + S[2] * cd                           # GF.jl:89-92, strike-slip
                    worst = max(worst, abs(got[fl, pc * ne + i] - want) / max(abs(want), 1e-3 * scale))
                    checked += 1
        M.free()
    assert checked >= 10000 and worst <= TOL, (checked, worst)

    # gf22 mantle -> mantle: shards of receiver elements; 36 entries per sampled pair
    checked, worst = 0, 0.0
    for e0 in rng.integers(0, ne - rows_per, nshard):
        M = oq.device_mantle_mantle(ma, lam, mu, elems=(int(e0), int(e0) + rows_per))
        got = M.rows_to_host(0, M.local_rows)
        assert got.shape == (6 * rows_per, 6 * ne)
        scale = max(np.max(np.abs(got)), 1e-300)
        srcs = rng.integers(0, ne, 14)
        for el in range(rows_per):
            j = int(e0) + el
            for i in srcs:
                for pc in range(6):
                    S = _hex8_unit(mao, int(i), mao.cx[j], mao.cy[j], mao.cz[j], pc, mu, nu)
                    g = got[np.arange(6) * rows_per + el, pc * ne + i]
                    den = np.maximum(np.abs(S), 1e-3 * scale)
                    worst = max(worst, float(np.max(np.abs(g - S) / den)))
                    checked += 6
        M.free()
    assert checked >= 10000 and worst <= TOL, (checked, worst)

    # gf11 at 250 x 80 (non-power-of-two strike count): every Toeplitz entry, and dense rows out of a shard
    st = ref.gf_fault_fault(mfo, lam, mu, buffer_ratio=1.0)
    st_gpu = oq.stress_greens_function(mf, lam, mu, buffer_ratio=1.0, fourier=False)
    assert float(np.max(np.abs(st_gpu - st) / np.abs(st))) <= TOL
    r0 = 12345
    M = oq.device_fault_fault(mf, lam, mu, buffer_ratio=1.0, rows=(r0, r0 + 8))
    want = bench.toeplitz_rows(st, r0, r0 + 8)
    assert float(np.max(np.abs(M.rows_to_host(0, 8) - want) / np.abs(want))) <= TOL
    M.free()
