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


if __name__ == "__main__":
    okada()
    hex8()
    print("wrote", os.listdir(HERE))
