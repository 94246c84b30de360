"""`wsolve`: solve while streaming snapshots to disk -- host mirror of /root/reference/src/io.jl:118-133.

The reference buffers `nstep` snapshots in RAM and appends them to extendable, chunked HDF5 datasets of shape
(shape..., nt) (io.jl:22-82).  HDF5 is not available in this environment, so the store is a directory of `.npy`
files with the same logical layout (one array per name, time as the LAST axis, Fortran order like Julia/HDF5.jl)
plus `meta.json`; `read(store, name)` returns what `h5read(file, name)` would.  Semantics kept: `nstep`
buffering, `stride` down-sampling (every stride-th callback, io.jl:51-58), `append`, `force`, a final flush when
the solve ends early (io.jl:131), and the refusal to overwrite without `force` (io.jl:119-123).
"""
from __future__ import annotations

import json
import logging
import os
import shutil
from typing import Callable, List, Sequence

import numpy as np

from .equation import ODEProblem, _alg_code, solve

log = logging.getLogger("oetqf_b200")


class _Store:
    def __init__(self, path: str, names: Sequence[str], tname: str, append: bool):
        assert tname not in names, f"Duplicate name of {tname} in {list(names)}."
        self.path, self.names, self.tname = path, list(names), tname
        self.t: List[float] = []
        self.u = {n: [] for n in self.names}
        self.total = 0
        if append:
            with open(os.path.join(path, "meta.json")) as fh:
                meta = json.load(fh)
            assert meta["names"] == self.names and meta["tname"] == tname
            self.total = meta["nt"]
        else:
            os.makedirs(path, exist_ok=True)
            self._write_meta()

    def _write_meta(self):
        with open(os.path.join(self.path, "meta.json"), "w") as fh:
            json.dump({"names": self.names, "tname": self.tname, "nt": self.total}, fh)

    def push(self, t, arrays):
        self.t.append(float(t))
        for n, a in zip(self.names, arrays):
            self.u[n].append(np.array(a, order="F"))

    def flush(self):
        if not self.t:
            return
        for name, new in [(self.tname, np.array(self.t))] + [(n, np.stack(self.u[n], axis=-1)) for n in self.names]:
            f = os.path.join(self.path, name + ".npy")
            if self.total > 0:
                new = np.concatenate([np.load(f), new], axis=-1)
            np.save(f, np.asfortranarray(new))
        self.total += len(self.t)
        self.t = []
        self.u = {n: [] for n in self.names}
        self._write_meta()


def read(path: str, name: str) -> np.ndarray:
    """h5read(file, name) of the reference's output file, for this store."""
    return np.load(os.path.join(path, name + ".npy"))


def VThetaDelta(u, t, du):               # 𝐕𝚯𝚫, io.jl:84
    return (u.x[0], u.x[1], u.x[2])


def VThetaEpsRateDelta(u, t, du):        # 𝐕𝚯𝚬′𝚫, io.jl:85-86: the strain RATE comes from the derivative
    return (u.x[0], u.x[1], du.x[2], u.x[4])


def wsolve(prob: ODEProblem, alg, file: str, nstep: int, getu: Callable, ustrs: Sequence[str], tstr: str, *,
           stride: int = 1, append: bool = False, force: bool = False, **kwargs):
    """wsolve(prob, alg, file, nstep, getu, ustrs, tstr; stride, append, force, kwargs...)  (io.jl:118-133).
    `getu(u, t, du)` returns the tuple of arrays to save (du = derivative at t, the role of
    `integrator(t, Val{1})` in the reference's handlers)."""
    _alg_code(alg)                       # raises for algorithms the device integrator does not provide
    if os.path.exists(file) and not force and not append:
        log.info("Overwrite existing file %s must set `force = true`.", file)
        log.info("Aborting computation.")
        return None
    if os.path.exists(file) and not append:
        shutil.rmtree(file)
    store = _Store(file, ustrs, tstr, append)
    assert len(getu(prob.u0, prob.tspan[0], prob.u0)) == len(ustrs), \
        "Unmatched length between solution components and names."

    def cb(u, t, step, du):
        store.push(t, getu(u, t, du))
        if len(store.t) >= nstep:
            store.flush()
        return False

    sol = solve(prob, alg, stride=stride, save_everystep=False, callback=cb, **kwargs)
    store.flush()                        # in case `solve` terminates earlier (io.jl:131)
    return sol
