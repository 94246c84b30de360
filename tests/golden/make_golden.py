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


if __name__ == "__main__":
    okada()
    print("wrote", os.listdir(HERE))
