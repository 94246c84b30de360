"""BASELINE configs[4]: Green's-matrix assembly sweep, Okada dc3d and hex8 kernels, 1k to 100k elements.

Times the four assembly kernels (CUDA events around the kernel, as recorded by the library) on this rank's
shard.  Matrices that do not fit HBM are assembled as a bounded row shard (entries/s is per-entry work, the
kernels are embarrassingly parallel over rows).  Under torchrun every rank assembles its own shard and the
aggregate rate is reported.

  python scripts/assembly_sweep.py --out gpurun_out/assembly_sweep.json
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oetqf_b200 as oq  # noqa: E402
import workloads as W  # noqa: E402


def kms(m):
    ms = C.c_double()
    oq._lib.check(oq._lib.load().oq_matrix_kernel_ms(m.handle, C.byref(ms)))
    return ms.value


def best(build, reps=3):
    t, shape, info = None, None, None
    for _ in range(reps):
        m = build()
        ms = kms(m)
        shape = (m.local_rows, m.cols)
        if t is None or ms < t:
            t, info = ms, m.assembly_info()
        m.free()
    return t, shape, info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--max-bytes", type=float, default=40e9, help="largest shard to materialise per matrix")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    oq.init(local)
    from oetqf_b200 import gf as gfmod
    rows = []
    # (fault nx, nxi) and mantle (mx, my, mz): total elements from ~1k to ~100k
    # mantle cells are whole multiples of the 250 m fault cells along strike (as in examples/otf-with-mantle.jl:
    # 20 km over 10 km): pairs then fall into translation classes; the last case is the 100k-element configuration on
    # grids that are NOT commensurate (1041.7 m over 250 m), where gf12 / gf21 keep the per-pair / tiled kernels
    cases = [((32, 16), (8, 5, 12)), ((64, 32), (16, 9, 14)), ((128, 64), (32, 9, 28)),
             ((256, 64), (32, 21, 29)), ((250, 80), (50, 33, 48)), ((250, 80), (60, 29, 46))]
    for (nx, nxi), (mx, my, mz) in cases:
        fs = W.FaultSpec(nx * 250.0, nxi * 250.0, 250.0, 250.0)
        bs = W.BoxSpec(-fs.x / 2, -10e3, -fs.xi, fs.x, 20e3, -40e3, mx, my, mz, tuple(np.cumprod(np.ones(mz) * 1.05)))
        mf = oq.gen_mesh("RectOkada", fs.x, fs.xi, fs.dx, fs.dxi, fs.dip)
        ma = oq.gen_mesh("BEMHex8Mesh", *bs.args())
        nf, ne = nx * nxi, len(ma)
        rec = {"fault_cells": nf, "hex8_cells": ne, "elements": nf + ne}
        # K1: Toeplitz-unique entries
        t = None
        for _ in range(3):
            st = oq.stress_greens_function(mf, W.LAM, W.MU, buffer_ratio=1.0, fourier=False)
            ms = gfmod.last_kernel_ms["value"]
            t = ms if t is None else min(t, ms)
        rec["okada_fault_fault"] = {"unique_entries": int(st.size), "ms": t, "entries_per_s": st.size / (t * 1e-3),
                                    "dense_equivalent_entries_per_s": float(nf) ** 2 / (t * 1e-3)}
        # shards bounded by --max-bytes, split over ranks
        e_rank = oq.dist.shard_range(ne, world, rank)
        f_rank = oq.dist.shard_range(nf, world, rank, align=4)

        def cap_elems(cols):
            n = int(args.max_bytes / (8.0 * 6 * cols))
            return (e_rank[0], min(e_rank[1], e_rank[0] + max(1, n)))

        def cap_rows(cols):
            n = int(args.max_bytes / (8.0 * cols))
            return (f_rank[0], min(f_rank[1], f_rank[0] + max(4, n)))

        el12, el22, r21 = cap_elems(nf), cap_elems(6 * ne), cap_rows(6 * ne)
        for name, build in (
                ("okada_fault_mantle", lambda: oq.device_fault_mantle(mf, ma, W.LAM, W.MU, buffer_ratio=1.0, elems=el12)),
                ("hex8_mantle_fault", lambda: oq.device_mantle_fault(ma, mf, W.LAM, W.MU, rows=r21)),
                ("hex8_mantle_mantle", lambda: oq.device_mantle_mantle(ma, W.LAM, W.MU, elems=el22))):
            ms, shape, info = best(build)
            n = shape[0] * shape[1]
            rec[name] = {"shard_shape": list(shape), "ms": ms, "entries_per_s_per_gpu": n / (ms * 1e-3), "path": info["path"],
                         "pairs": info["pairs"], "closed_form_evaluations": info["unique_pairs"], "table_ms": info["table_ms"],
                         "expand_ms": info["expand_ms"],
                         "expand_gbs": n * 8 / (info["expand_ms"] * 1e-3) / 1e9 if info["expand_ms"] else None}
        rows.append(rec)
        if rank == 0:
            print(json.dumps(rec), flush=True)
    if rank == 0 and args.out:
        with open(args.out, "w") as fh:
            json.dump({"n_gpus": world, "note": "entries/s per GPU; shards bounded to %g bytes" % args.max_bytes,
                       "fp64_peak_tflops_measured": oq.measure_fp64_peak() / 1e12, "sweep": rows}, fh, indent=1)


if __name__ == "__main__":
    main()
