"""`wsolve`: solve while streaming snapshots to disk -- host mirror of /root/reference/src/io.jl:118-133.

The reference buffers `nstep` snapshots in RAM and appends them to extendable, chunked HDF5 datasets of shape
(shape..., nt) (io.jl:22-82).  `_Store` writes exactly that layout with h5py when it is importable and the file
name ends in .h5/.hdf5; HDF5 is not available in this environment, so the tested path is the append-only
directory of `.npy` chunk files with the same logical layout (one dataset per name, time as the LAST axis,
Fortran order like Julia/HDF5.jl) plus an atomically replaced `meta.json`; `read(store, name)` returns what
`h5read(file, name)` would.  Semantics kept: `nstep` buffering, `stride` down-sampling (every stride-th callback,
io.jl:51-58 -- a completion snapshot that is not stride-aligned is dropped, as in the reference), `append`,
`force`, a final flush when the solve ends early (io.jl:131), and the refusal to overwrite without `force`
(io.jl:119-123).  Snapshots reach the host through the device-side ring of oq_solve (csrc/solve.cu): the
integration does not wait for the callback or the disk.
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


def _have_h5py():
    try:
        import h5py  # noqa: F401
        return True
    except Exception:
        return False


class _Store:
    """Append-only snapshot store with the reference's logical layout: one dataset per name, shape (shape..., nt),
    time LAST, grown by one chunk of up to `nstep` snapshots per flush (io.jl:22-82).

    * `file` ending in .h5 / .hdf5 and h5py importable: a real HDF5 file with the reference's extendable, chunked
      datasets (maxshape None on the time axis, chunk = (shape..., nstep)), readable by the reference's h5read.
    * otherwise (this environment has no HDF5): a directory; every flush writes ONE new .npy chunk file per dataset
      (`<name>.<first index>.npy`) and then replaces `meta.json` atomically (tmp file + rename).  Nothing already on
      disk is read or rewritten, so total I/O and peak memory are O(output), and a crash mid-flush leaves the
      previous meta.json -- hence a consistent store -- behind.
    """

    def __init__(self, path: str, names: Sequence[str], tname: str, append: bool, nstep: int = 1):
        assert tname not in names, f"Duplicate name of {tname} in {list(names)}."
        self.path, self.names, self.tname, self.nstep = path, list(names), tname, max(1, int(nstep))
        self.t: List[float] = []
        self.u = {n: [] for n in self.names}
        self.total = 0
        self.chunks: List[int] = []            # first snapshot index of every chunk on disk
        self.h5 = path.endswith((".h5", ".hdf5")) and _have_h5py()
        if self.h5:
            import h5py
            if append:
                with h5py.File(path, "r") as f:
                    self.total = f[tname].shape[-1]
            else:
                h5py.File(path, "w").close()
            return
        if append:
            with open(os.path.join(path, "meta.json")) as fh:
                meta = json.load(fh)
            assert meta["names"] == self.names and meta["tname"] == tname
            self.total = meta["nt"]
            self.chunks = list(meta.get("chunks", [0] if self.total else []))
        else:
            os.makedirs(path, exist_ok=True)
            self._write_meta()

    def _write_meta(self):
        tmp = os.path.join(self.path, "meta.json.tmp")
        with open(tmp, "w") as fh:
            json.dump({"names": self.names, "tname": self.tname, "nt": self.total, "chunks": self.chunks}, fh)
            fh.flush()
            os.fsync(fh.fileno())
        os.replace(tmp, os.path.join(self.path, "meta.json"))

    def push(self, t, arrays):
        self.t.append(float(t))
        for n, a in zip(self.names, arrays):
            self.u[n].append(np.array(a, order="F"))

    def flush(self):
        if not self.t:
            return
        new = [(self.tname, np.array(self.t))] + [(n, np.stack(self.u[n], axis=-1)) for n in self.names]
        if self.h5:
            import h5py
            with h5py.File(self.path, "r+") as f:
                for name, arr in new:
                    if name not in f:
                        f.create_dataset(name, shape=arr.shape[:-1] + (0,), maxshape=arr.shape[:-1] + (None,),
                                         chunks=arr.shape[:-1] + (self.nstep,), dtype=arr.dtype)
                    d = f[name]
                    d.resize(self.total + arr.shape[-1], axis=d.ndim - 1)
                    d[..., self.total:] = arr
        else:
            for name, arr in new:
                np.save(os.path.join(self.path, f"{name}.{self.total:09d}.npy"), np.asfortranarray(arr))
            self.chunks.append(self.total)
        self.total += len(self.t)
        self.t = []
        self.u = {n: [] for n in self.names}
        if not self.h5:
            self._write_meta()                  # last, atomically: the chunks above become visible together


def read(path: str, name: str) -> np.ndarray:
    """h5read(file, name) of the reference's output file, for this store (chunks are concatenated on read)."""
    if os.path.isfile(path):
        import h5py
        with h5py.File(path, "r") as f:
            return np.asfortranarray(f[name][...])
    with open(os.path.join(path, "meta.json")) as fh:
        meta = json.load(fh)
    if "chunks" not in meta:                     # stores written by the first version: one array per name
        return np.load(os.path.join(path, name + ".npy"))
    parts = [np.load(os.path.join(path, f"{name}.{c:09d}.npy")) for c in meta["chunks"]]
    return np.asfortranarray(np.concatenate(parts, axis=-1)) if parts else np.zeros((0,))


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
        shutil.rmtree(file) if os.path.isdir(file) else os.remove(file)
    store = _Store(file, ustrs, tstr, append, nstep)
    assert len(getu(prob.u0, prob.tspan[0], prob.u0)) == len(ustrs), \
        "Unmatched length between solution components and names."

    def cb(u, t, step, du):
        if step % stride != 0:             # the completion snapshot oq_solve adds; the reference saves only every
            return False                   # stride-th callback (io.jl:51-58)
        store.push(t, getu(u, t, du))
        if len(store.t) >= nstep:
            store.flush()
        return False

    kwargs.setdefault("async_snapshots", True)      # the saving callback never asks for a stop
    sol = solve(prob, alg, stride=stride, save_everystep=False, callback=cb, **kwargs)
    store.flush()                        # in case `solve` terminates earlier (io.jl:131)
    return sol
