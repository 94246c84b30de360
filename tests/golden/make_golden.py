"""Regenerates tests/golden/*.json from the CPU oracle (run from the repo root: python tests/golden/make_golden.py).

The reference is Julia and cannot run here, and it ships no golden vectors for this path; these fixtures
freeze the oracle's outputs (themselves pinned by Okada-1985 check values and physics invariants in
tests/test_oracle_okada.py) so that the GPU parity target cannot drift silently."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import ref  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def okada():
    rng = np.random.default_rng(2024)
    cases = []
    for fault, ftype, nrept, br in [((100.0, 100.0, 10.0, 10.0, 41.0), 0, 2, 0.0),
                                    ((100.0, 100.0, 10.0, 10.0, 41.0), 1, 2, 0.0),
                                    ((80e3, 8e3, 10e3, 2e3, 90.0), 0, 2, 1.0),
                                    ((16e3, 8e3, 500.0, 500.0, 90.0), 0, 2, 1.0),
                                    ((16e3, 8e3, 500.0, 500.0, 60.0), 1, 1, 0.5)]:
        mf = ref.fault_mesh(*fault)
        st = ref.gf_fault_fault(mf, 3e10, 3e10, ftype=ftype, nrept=nrept, buffer_ratio=br)
        idx = np.stack([rng.integers(0, mf.nx, 24), rng.integers(0, mf.nxi, 24), rng.integers(0, mf.nxi, 24)], 1)
        cases.append(dict(fault=list(fault), ftype=ftype, nrept=nrept, buffer_ratio=br, index=idx.tolist(),
                          values=[float(st[i, j, l]) for i, j, l in idx]))
    with open(os.path.join(HERE, "okada_kernels.json"), "w") as fh:
        json.dump(dict(lam=3e10, mu=3e10, cases=cases), fh, indent=1)


def hex8():
    """Values from the QUADRATURE oracle (oracle/hex8_numeric.py, 64-point Gauss-Legendre per face axis):
    independent of the closed form they pin."""
    from oracle import hex8_numeric as hn
    rng = np.random.default_rng(77)
    cases = []
    # App. B geometry of SURVEY.md plus random cuboids; receivers at centroid-like distances
    geoms = [(0.0, 0.0, -1.0, 2.0, 2.0, 2.0)] + [
        (rng.uniform(-2, 2), rng.uniform(-2, 2), -rng.uniform(0.2, 3), *rng.uniform(0.5, 3, 3)) for _ in range(5)]
    for g in geoms:
        qx, qy, qz, dx, dy, dz = g
        c = np.array([qx, qy + dy / 2, qz - dz / 2])
        pts = [c,                                         # self (inside)
               c + np.array([dx, 0, 0]), c + np.array([0, -dy, 0]), c + np.array([0, 0, -dz]),   # neighbours
               c + np.array([3 * dx, -2 * dy, 0.0]),
               np.array([c[0] + 2.5 * dx, c[1] + 1.5 * dy, 0.0]),                           # on the free surface
               np.array([c[0] - 4.0, c[1] + 5.0, max(c[2], -0.1 - dz) * 0.3])]
        lam, mu = rng.uniform(0.5, 2), rng.uniform(0.5, 2)
        nu = lam / 2 / (lam + mu)
        eps = rng.uniform(-1, 1, 6)
        for p in pts:
            if p[2] > 0:
                continue
            s = hn.stress_vol_hex8(*p, qx, qy, qz, dx, dy, dz, eps, mu, nu, nquad=64)
            cases.append(dict(point=[float(v) for v in p], geom=[float(v) for v in g], mu=mu, nu=nu,
                              eps=[float(v) for v in eps], sigma=[float(v) for v in s]))
    with open(os.path.join(HERE, "hex8_quadrature.json"), "w") as fh:
        json.dump(dict(cases=cases), fh, indent=1)


def _converged_quadrature(args):
    from oracle import hex8_numeric as hn
    p, g, pc, mu, nu = args
    eps = np.zeros(6)
    eps[pc] = 1.0
    prev, nq = None, 64
    while True:
        cur = hn.stress_vol_hex8(*p, *g, eps, mu, nu, nquad=nq)
        if prev is not None:
            delta = float(np.max(np.abs(cur - prev)) / np.max(np.abs(cur)))
            if delta < 5e-12 or nq >= 512:
                return cur, nq, delta
        prev, nq = cur, nq * 2


def hex8_patterns():
    """>= 200 further cases at the reference's CALL PATTERNS (GF.jl:215-221, :277-283), again from the quadrature
    oracle: mantle->fault (receiver = fault-cell centroid, sources = cells of the example's hex8 box,
    examples/otf-with-mantle.jl:18,25-29) and mantle->mantle (receiver = cell centroid or a Gauss2 point of it;
    itself, stacked / side neighbours, distant cells), unit eigenstrains as the builders pass them, plus scaled copies
    of the geometry out to r/a = 50 and receivers on the free surface."""
    import workloads as W
    from oracle import hex8_numeric as hn
    rng = np.random.default_rng(4242)
    fs, bs = W.C2_FAULT, W.C2_BOX
    mf = ref.fault_mesh(fs.x, fs.xi, fs.dx, fs.dxi, fs.dip)
    ma = ref.hex8_box(*bs.args())
    lam, mu = W.LAM, W.MU
    nu = lam / 2 / (lam + mu)
    cases = []

    todo = []

    def add(kind, p, i, pc):
        todo.append((kind, tuple(float(v) for v in p), int(i), int(pc)))

    # mantle -> fault: 60 (fault cell, hex8 cell, unit strain) triples of the example
    for _ in range(60):
        f, i, pc = int(rng.integers(0, mf.nx * mf.nxi)), int(rng.integers(0, ma.n)), int(rng.integers(0, 6))
        add("mantle_fault", (mf.x[f % mf.nx], mf.y[f // mf.nx], mf.z[f // mf.nx]), i, pc)
    # mantle -> mantle at centroids: self, neighbours in x / y / z, random pairs
    nx, ny = bs.nx, bs.ny
    for j in rng.integers(0, ma.n, 12):
        j = int(j)
        nb = [j, (j + 1) % ma.n, (j + nx) % ma.n, (j + nx * ny) % ma.n, int(rng.integers(0, ma.n))]
        for i in nb:
            add("mantle_mantle", (ma.cx[j], ma.cy[j], ma.cz[j]), i, int(rng.integers(0, 6)))
    # Gauss2 points of the receiver cell (GF.jl:270-276)
    gp = 1 / np.sqrt(3.0)
    for _ in range(40):
        j, i = int(rng.integers(0, ma.n)), int(rng.integers(0, ma.n))
        sgn = rng.choice([-1.0, 1.0], 3)
        p = (ma.cx[j] + sgn[0] * gp * ma.dx[j] / 2, ma.cy[j] + sgn[1] * gp * ma.dy[j] / 2, ma.cz[j] + sgn[2] * gp * ma.dz[j] / 2)
        add("mantle_mantle_gauss2", p, i, int(rng.integers(0, 6)))
    # far receivers (5 to 50 times the SMALLEST cell dimension: the corner sums of the closed form lose ~(r/a)^3 ulps
    # to cancellation, see tests/test_oracle_hex8.py::test_far_field_conditioning) and receivers on the free surface
    for _ in range(30):
        i = int(rng.integers(0, ma.n))
        a = min(ma.dx[i], ma.dy[i], ma.dz[i])
        r = a * rng.uniform(5, 50)
        th, ph = rng.uniform(0, 2 * np.pi), rng.uniform(0.05, 0.45) * np.pi
        c = np.array([ma.cx[i], ma.cy[i], ma.cz[i]])
        p = c + r * np.array([np.cos(th) * np.sin(ph), np.sin(th) * np.sin(ph), -np.cos(ph)])
        add("far", tuple(p), i, int(rng.integers(0, 6)))
    for _ in range(20):
        i = int(rng.integers(0, ma.n))
        add("surface", (ma.cx[i] + rng.uniform(-3, 3) * ma.dx[i], ma.cy[i] + rng.uniform(-3, 3) * ma.dy[i], 0.0), i,
            int(rng.integers(0, 6)))
    # The cells are strongly anisotropic (20 km x 1.7 km x 4.6-11.6 km): a receiver next to a large face needs a very
    # fine rule (a neighbouring centroid: 64 points per face axis leave 5e-3, 256 leave 1e-9, 512 reach 1e-15).  Each
    # case is therefore evaluated with 64, 128, 256, 512 points until two successive rules agree to 5e-12 -- a
    # criterion internal to the quadrature oracle, independent of the closed form it pins.
    from multiprocessing import Pool
    geoms = [(ma.qx[i], ma.qy[i], ma.qz[i], ma.dx[i], ma.dy[i], ma.dz[i]) for (_, _, i, _) in todo]
    with Pool(min(8, os.cpu_count() or 1)) as pool:
        res = pool.map(_converged_quadrature, [(p, g, pc, mu, nu) for (_, p, _, pc), g in zip(todo, geoms)])
    for (kind, p, i, pc), g, (sg, nq, delta) in zip(todo, geoms, res):
        eps = [0.0] * 6
        eps[pc] = 1.0
        cases.append(dict(kind=kind, point=list(p), geom=[float(v) for v in g], mu=mu, nu=nu, eps=eps,
                          sigma=[float(v) for v in sg], nquad=nq, self_convergence=delta))
    with open(os.path.join(HERE, "hex8_patterns.json"), "w") as fh:
        json.dump(dict(lam=lam, mu=mu, cases=cases), fh, indent=0)
    print(len(cases), "hex8 pattern cases; rules used:", sorted(set(c["nquad"] for c in cases)),
          "worst self-convergence", max(c["self_convergence"] for c in cases))


if __name__ == "__main__":
    okada()
    hex8()
    hex8_patterns()
    print("wrote", os.listdir(HERE))
