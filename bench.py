#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 hot path (contract: see the task statement / DESIGN.md §6).

Workload (config.workload): BASELINE.json configs[2], "fault-only synthetic 256x64 mesh (16k elements), dense
Green's matvec RHS" -- the configuration the metric's 1/2/4/8-GPU scaling is quoted on and the largest RHS
config that fits one GPU.  A step = ONE evaluation of the ODE right-hand side (equation.jl:156-166) on the
resident state: forcing kernel + fused dense matvec with the rate-and-state epilogue.  The fp64 Green's matrix
(2.15 GB) is far larger than the 126 MB L2, so every step streams it from HBM (no L2 flush needed).

  value        RHS evaluations / s, state and matrix resident in HBM (device-timed, max over ranks)
  e2e          the same through the reference-facing call prob.f(du, u, p, t) with HOST buffers (page-locked):
               every step moves u host->device and du device->host inside the timed region (the kernels read /
               write the mapped host arrays over PCIe themselves; pageable arrays would be staged by copies)
  roofline     the fused matvec kernel: algorithmic bytes / its CUDA-event time vs the measured HBM copy peak
  cpu_baseline the oracle port of the reference's own CPU algorithm (FFT form of equation.jl:44-61, OpenMP +
               pocketfft) on this box's host cores, bounded sample
  extra        Green's assembly throughput (entries/s, fp64 pipe fraction) and the example config

N > 1 (torchrun): the matrix is row-sharded (strong scaling); every RHS all-gathers v - vpl through peer stores
over NVLink (comm.cu).  `--impl reference` times the CPU port alone (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import workloads as W  # noqa: E402

METRIC = "rhs_evals_per_s"
UNIT = "evals/s"
WORKLOAD = "BASELINE configs[2]: fault-only 256x64 (16384 cells), dense fp64 Green's matvec RHS, aging law"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    return 6650.0, "B200_PROFILING.md fallback (of fallback)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                 parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def bench_config(fs, world):
    """The `config` object of the JSON line -- identical in both arms (ours and --impl reference) at the same N."""
    import oetqf_b200 as oq
    nf = fs.nx * fs.nxi
    r0, r1 = oq.dist.shard_range(nf, world, 0, align=4)
    return {"workload": WORKLOAD, "nx": fs.nx, "nxi": fs.nxi, "rows_per_rank": r1 - r0,
            "matrix_bytes_per_rank": 8.0 * (r1 - r0) * nf,
            "l2": "inputs larger than L2 (2.15 GB fp64 matrix vs 126 MB): no flush between steps",
            "parallelism": f"row-sharded x{world}, peer-store all-gather" if world > 1 else "single GPU"}


def bench_state(fs, x, z, rng_seed=42):
    """Properties and state of the benchmark problem: the example's fields with theta perturbed by
    +-10 % and v by +-30 % (a state with v == vpl everywhere would make v - vpl, hence the whole matvec, zero)."""
    a, b, L, sig = W.fault_properties(x, z, fs.nx, fs.nxi)
    rng = np.random.default_rng(rng_seed)
    v, th, dl = W.initial_state(fs.nx, fs.nxi, L, rng=rng)
    v = np.asfortranarray(v * (1 + 0.3 * rng.uniform(-1, 1, v.shape)))
    return a, b, L, sig, v, np.asfortranarray(th), dl


def toeplitz_rows(st, f0, f1):
    """Rows [f0, f1) of the dense matrix G[(i,j),(k,l)] = st[|i-k|, j, l] (test/BEM/tests.jl:46-49) as a C-ordered
    [rows, nx*nxi] array (column k + l*nx), built from sliding windows over the even extension."""
    from numpy.lib.stride_tricks import sliding_window_view
    nx, nxi, _ = st.shape
    out = np.empty((f1 - f0, nx * nxi))
    f = f0
    while f < f1:
        j, i0 = divmod(f, nx)
        i1 = min(nx, i0 + (f1 - f))
        ext = np.concatenate([st[::-1, j, :], st[1:, j, :]], axis=0)               # ext[m + nx-1] = st[|m|, j, :]
        win = sliding_window_view(ext, nx, axis=0)                                 # [w, l, k] = ext[w + k, l]
        out[f - f0: f - f0 + (i1 - i0)] = win[::-1][i0:i1].reshape(i1 - i0, nx * nxi)   # w = nx-1-i
        f += i1 - i0
    return out


def comp_rel_err(got, want):
    """per-component relative error; components crossing zero are measured against 1e-6 of the field's maximum"""
    got, want = np.asarray(got), np.asarray(want)
    den = np.maximum(np.abs(want), 1e-6 * np.max(np.abs(want)) + 1e-300)
    return float(np.max(np.abs(got - want) / den))


PARITY_TOL = 1e-10


def fault_parity(oq, fs, p, g11, rows, du_local, threads=None, block=1024):
    """Oracle check of this rank's shard on the benchmark problem: (1) every entry of the Toeplitz kernel the GPU
    assembles against the CPU restatement of GF.jl:31-58; (2) every entry of this rank's dense rows against the
    expansion of the oracle kernel; (3) every RHS component of this rank's rows against the oracle's Toeplitz-form
    RHS (equation.jl:156-166).  Returns a dict of maxima (relative errors)."""
    from oracle import ref
    if threads:
        ref.lib().oq_ref_set_num_threads(int(threads))
    else:
        ref.use_all_cores()
    r0, r1 = rows
    mfo = ref.fault_mesh(fs.x, fs.xi, fs.dx, fs.dxi, fs.dip)
    a, b, L, sig, v, th, _ = bench_state(fs, mfo.x, mfo.z)
    st = ref.gf_fault_fault(mfo, W.LAM, W.MU, buffer_ratio=1.0)
    mf = oq.gen_mesh("RectOkada", fs.x, fs.xi, fs.dx, fs.dxi, fs.dip)
    st_gpu = oq.stress_greens_function(mf, W.LAM, W.MU, buffer_ratio=1.0, fourier=False)
    kern_err = float(np.max(np.abs(st_gpu - st) / np.abs(st)))
    kern_identical = bool(np.array_equal(st_gpu, st))
    rows_err = 0.0
    for b0 in range(0, r1 - r0, block):
        b1 = min(r1 - r0, b0 + block)
        got = g11.rows_to_host(b0, b1)
        want = toeplitz_rows(st, r0 + b0, r0 + b1)
        rows_err = max(rows_err, float(np.max(np.abs(got - want) / np.abs(want))))
    pf = ref.FaultProp(a, b, L, sig, W.ETA, W.VPL, W.F0, W.V0)
    want = ref.rhs_fault(pf, st, v, th, form="toeplitz")
    rhs_err = 0.0
    for g, w in zip(du_local, want):
        w = np.asarray(w).reshape(-1, order="F")
        den = np.maximum(np.abs(w), 1e-6 * np.max(np.abs(w)) + 1e-300)           # scale of the WHOLE field
        rhs_err = max(rhs_err, float(np.max(np.abs(np.asarray(g) - w[r0:r1]) / den[r0:r1]))) if r1 > r0 else rhs_err
    return {"kernel_max_rel_err": kern_err, "matrix_rows_max_rel_err": rows_err, "rhs_max_rel_err": rhs_err,
            "rows": r1 - r0, "matrix_entries": (r1 - r0) * fs.nx * fs.nxi, "kernel_entries": int(st.size),
            "kernel_bit_identical": kern_identical}


def build_fault_problem(oq, fs, rows, rng_seed=42):
    mf = oq.gen_mesh("RectOkada", fs.x, fs.xi, fs.dx, fs.dxi, fs.dip)
    a, b, L, sig, v, th, dl = bench_state(fs, mf.x, mf.z, rng_seed)
    pf = oq.RateStateQuasiDynamicProperty(a, b, L, sig, W.ETA, W.VPL, W.F0, W.V0)
    g11 = oq.device_fault_fault(mf, W.LAM, W.MU, buffer_ratio=1.0, rows=rows)
    u0 = oq.ArrayPartition(v, th, dl)
    prob = oq.assemble(g11, pf, u0, (0.0, 1.0))
    return mf, prob, u0, g11


# ------------------------------------------------------------------------------------------ reference arm
def cpu_reference(fs, budget_s=10.0, steps=None, warmup=3, min_s=2.0):
    """The reference's own CPU algorithm for this path, restated (oracle port): FFT form of the fault-fault
    interaction (equation.jl:44-61) + update_fault! (equation.jl:233-246) with the reference's allocation
    discipline (scratch created once, gen_alloc) on all host threads -- ref.FaultRhsFFT.  Timed for at least
    `min_s` seconds (whole multiples of `steps` evaluations when given)."""
    from oracle import ref
    ref.use_all_cores()
    mf = ref.fault_mesh(fs.x, fs.xi, fs.dx, fs.dxi, fs.dip)
    a, b, L, sig, v, th, _ = bench_state(fs, mf.x, mf.z)
    pf = ref.FaultProp(a, b, L, sig, W.ETA, W.VPL, W.F0, W.V0)
    t0 = time.perf_counter()
    # Green's kernel: a bounded sample of source rows (the full 256x64 kernel is 5.2 M dc3d calls)
    nsrc = min(mf.nxi, 8)
    sub = ref.FaultMesh(mf.x, mf.dx, mf.nx, mf.ax, mf.xi[:nsrc], mf.dxi, nsrc, mf.axi[:nsrc], mf.y[:nsrc],
                        mf.z[:nsrc], mf.dep, mf.dip)
    ta = time.perf_counter()
    st_sub = ref.gf_fault_fault(sub, W.LAM, W.MU, buffer_ratio=1.0)
    asm_s = time.perf_counter() - ta
    asm_entries_per_s = st_sub.size / asm_s
    # RHS timing needs a kernel of the full shape; its values do not affect the time
    rng = np.random.default_rng(0)
    st = np.asfortranarray(rng.standard_normal((mf.nx, mf.nxi, mf.nxi)) * 1e6)
    plan = ref.FaultRhsFFT(pf, ref.gf_fourier(st))
    for _ in range(max(3, warmup)):
        plan(v, th)
    n, t1 = 0, time.perf_counter()
    while True:
        for _ in range(steps or 16):
            plan(v, th)
        n += steps or 16
        el = time.perf_counter() - t1
        if el >= (min_s if steps is not None else budget_s):
            break
    fft_evals = n / el
    # dense form of the same RHS (what the GPU arm streams) through OpenBLAS dgemv, the reference's default
    # matvecmul! backend (pref.jl:15-16), on a row sample far larger than the CPU caches
    nf = mf.nx * mf.nxi
    rows = min(nf, 4096)
    A = np.asfortranarray(rng.standard_normal((rows, nf)))
    x = rng.standard_normal(nf)
    ref.blas_gemv(A, x)
    m, t2 = 0, time.perf_counter()
    while time.perf_counter() - t2 < 2.0:
        ref.blas_gemv(A, x)
        m += 1
    dense_s_per_eval = (time.perf_counter() - t2) / m * (nf / rows)
    return dict(value=fft_evals, unit=UNIT, cores=ref.num_threads(), kind="port",
                sample=f"{n} FFT-form RHS evaluations of the full 256x64 problem in {el:.1f} s (reference algorithm, "
                       f"equation.jl:44-61; scratch preallocated, pocketfft + OpenMP); dense form extrapolated from "
                       f"an OpenBLAS dgemv over {rows} of {nf} rows ({A.nbytes / 1e6:.0f} MB)",
                dense_form_evals_per_s=1.0 / dense_s_per_eval,
                okada_assembly_entries_per_s=asm_entries_per_s, wall_s=time.perf_counter() - t0), el / n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fs = W.C3_FAULT
    base, s_per = cpu_reference(fs, steps=max(1, args.steps), warmup=max(3, args.warmup))
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * s_per, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": bench_config(fs, args.gpus),
            "algorithm": "FFT/Toeplitz form of the fault-fault interaction (equation.jl:44-61), CPU port of the "
                         "reference's own path; host cores only (the GPUs are idle in this arm)",
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ extras
ASSEMBLY_FLOPS = os.path.join(ROOT, "profiles", "r02_assembly_flops.json")


def _fp64_roofline(name, entries, kernel_ms, fp64_peak, flops):
    """fp64 roofline of one assembly kernel: executed fp64 flop per entry (dadd + dmul + 2 dfma, frozen from the
    ncu capture of the same kernel at HEAD, profiles/r02_assembly_flops.json) x entries / CUDA-event time, against
    the DFMA peak measured in this run (oq_measure_fp64_peak)."""
    rec = (flops or {}).get(name)
    if not rec:
        return None
    tf = rec["flop_per_entry"] * entries / (kernel_ms * 1e-3) / 1e12
    return {"bound": "fp64", "achieved": tf, "peak": fp64_peak / 1e12, "unit": "TFLOP/s", "frac": tf / (fp64_peak / 1e12),
            "flop_per_entry": rec["flop_per_entry"], "flop_source": rec.get("source"),
            "peak_source": "DFMA microbenchmark in this run (oq_measure_fp64_peak)"}


def assembly_extras(oq, fp64_peak):
    """Green's assembly throughput: entries/s and the fp64 roofline fraction of K1-K4."""
    from oetqf_b200 import gf as gfmod
    flops = None
    if os.path.exists(ASSEMBLY_FLOPS):
        with open(ASSEMBLY_FLOPS) as fh:
            flops = json.load(fh).get("kernels")
    out = {}
    fs = W.C3_FAULT
    mf = oq.gen_mesh("RectOkada", fs.x, fs.xi, fs.dx, fs.dxi, fs.dip)
    for label, ft, key in (("okada_fault_fault_256x64", oq.StrikeSlip(), "gf_fault_fault_kernel<0>"),
                           ("okada_fault_fault_256x64_dipslip", oq.DipSlip(), "gf_fault_fault_kernel<1>")):
        best = None
        for _ in range(4):
            st = oq.stress_greens_function(mf, W.LAM, W.MU, buffer_ratio=1.0, fourier=False, ftype=ft)
            ms = gfmod.last_kernel_ms["value"]
            best = ms if best is None else min(best, ms)
        out[label] = {"unique_entries": int(st.size), "kernel_ms": best, "entries_per_s": st.size / (best * 1e-3),
                      "dense_equivalent_entries_per_s": (mf.nx * mf.nxi) ** 2 / (best * 1e-3),
                      "roofline": _fp64_roofline(key, st.size, best, fp64_peak, flops)}
    fsm = W.FaultSpec(64e3, 16e3, 1000.0, 1000.0)
    mfm = oq.gen_mesh("RectOkada", fsm.x, fsm.xi, fsm.dx, fsm.dxi, fsm.dip)
    ma = oq.gen_mesh("BEMHex8Mesh", *W.box_for(32, 8, 8, fsm).args())
    hbm_peak, hbm_src = measured_peaks()

    def timed(builder, reps=3):
        best, shape, info = None, None, None
        for _ in range(reps):
            m = builder()
            shape = (m.local_rows, m.cols)
            ms = _matrix_kernel_ms(m)
            if best is None or ms < best:
                best, info = ms, m.assembly_info()
            m.free()
        return best, shape, info

    # Okada fault->mantle.  Default path on commensurate grids: one dc3d evaluation per class of pairs with bitwise
    # equal arguments + dense expansion (HBM-write bound); its fp64-bound twin (every pair, OQ_FAULT_MANTLE=pair) is
    # timed beside it against the DFMA peak.
    saved12 = os.environ.get("OQ_FAULT_MANTLE")
    build12 = lambda: oq.device_fault_mantle(mfm, ma, W.LAM, W.MU, buffer_ratio=1.0)   # noqa: E731
    best, shape, info = timed(build12)
    n = shape[0] * shape[1]
    rec = {"shape": list(shape), "kernel_ms": best, "entries_per_s": n / (best * 1e-3), "path": info["path"],
           "pairs": info["pairs"], "closed_form_evaluations": info["unique_pairs"], "table_ms": info["table_ms"],
           "expand_ms": info["expand_ms"], "roofline": None}
    if info["path"] == "classes" and info["expand_ms"] > 0:
        gbs = n * 8 / (info["expand_ms"] * 1e-3) / 1e9
        rec["roofline"] = {"bound": "hbm", "kernel": "expand_classes_kernel", "achieved": gbs, "peak": hbm_peak,
                           "unit": "GB/s", "frac": gbs / hbm_peak, "algorithmic_bytes": n * 8, "peak_source": hbm_src}
    else:
        rec["roofline"] = _fp64_roofline("gf_fault_mantle_kernel<0>", n, best, fp64_peak, flops)
    os.environ["OQ_FAULT_MANTLE"] = "pair"
    pbest, _, pinfo = timed(build12)
    rec["every_pair_kernel"] = {"kernel_ms": pbest, "entries_per_s": n / (pbest * 1e-3), "path": pinfo["path"],
                                "roofline": _fp64_roofline("gf_fault_mantle_kernel<0>", n, pbest, fp64_peak, flops)}
    if saved12 is None:
        os.environ.pop("OQ_FAULT_MANTLE", None)
    else:
        os.environ["OQ_FAULT_MANTLE"] = saved12
    out["okada_fault_mantle"] = rec
    # hex8 builders.  Default path: one closed-form evaluation per translation class of pairs + dense expansion
    # (csrc/greens_classes.cuh) -- bounded by the HBM writes of the shard; its fp64-bound twins (every pair through
    # the tiled kernels, OQ_HEX8=tile) are timed beside it against the DFMA peak.
    saved = os.environ.get("OQ_HEX8")
    for name, key, builder in (
            ("hex8_mantle_fault", "gf_mantle_fault_tile_kernel<0>", lambda: oq.device_mantle_fault(ma, mfm, W.LAM, W.MU)),
            ("hex8_mantle_mantle", "gf_mantle_mantle_tile_kernel", lambda: oq.device_mantle_mantle(ma, W.LAM, W.MU))):
        os.environ["OQ_HEX8"] = saved or ""
        best, shape, info = timed(builder)
        n = shape[0] * shape[1]
        rec = {"shape": list(shape), "kernel_ms": best, "entries_per_s": n / (best * 1e-3), "path": info["path"],
               "pairs": info["pairs"], "closed_form_evaluations": info["unique_pairs"],
               "table_ms": info["table_ms"], "expand_ms": info["expand_ms"], "roofline": None}
        if info["path"] == "classes" and info["expand_ms"] > 0:
            gbs = n * 8 / (info["expand_ms"] * 1e-3) / 1e9
            rec["roofline"] = {"bound": "hbm", "kernel": "expand_classes_kernel", "achieved": gbs, "peak": hbm_peak,
                               "unit": "GB/s", "frac": gbs / hbm_peak, "algorithmic_bytes": n * 8,
                               "note": "bytes of the dense shard written once; the class table is re-read from L2",
                               "peak_source": hbm_src}
        os.environ["OQ_HEX8"] = "tile"
        tbest, _, tinfo = timed(builder)
        rec["every_pair_tile_kernels"] = {"kernel_ms": tbest, "entries_per_s": n / (tbest * 1e-3), "path": tinfo["path"],
                                          "roofline": _fp64_roofline(key, n, tbest, fp64_peak, flops)}
        out[name] = rec
    if saved is None:
        os.environ.pop("OQ_HEX8", None)
    else:
        os.environ["OQ_HEX8"] = saved
    out["fp64_peak_tflops_measured"] = fp64_peak / 1e12
    return out


def _matrix_kernel_ms(m):
    # device time of the assembly kernel recorded by the library (CUDA events around the launch)
    import ctypes as C
    from oetqf_b200 import _lib
    ms = C.c_double()
    _lib.check(_lib.load().oq_matrix_kernel_ms(m.handle, C.byref(ms)))
    return ms.value


def fft_form_extra(oq):
    """The same 256x64 RHS in the reference's own translation-invariant (FFT) form on the GPU (§8f row 1)."""
    fs = W.C3_FAULT
    mf = oq.gen_mesh("RectOkada", fs.x, fs.xi, fs.dx, fs.dxi, fs.dip)
    a, b, L, sig = W.fault_properties(mf.x, mf.z, mf.nx, mf.nxi)
    v, th, dl = W.initial_state(mf.nx, mf.nxi, L, rng=np.random.default_rng(42))
    pf = oq.RateStateQuasiDynamicProperty(a, b, L, sig, W.ETA, W.VPL, W.F0, W.V0)
    st = oq.stress_greens_function(mf, W.LAM, W.MU, buffer_ratio=1.0, fourier=False)
    u0 = oq.ArrayPartition(v, th, dl)
    prob = oq.assemble(st, pf, u0, (0.0, 1.0), gf11_form="fft")
    prob.p.set_state(u0.x)
    prob.p.rhs_resident(20)
    ms = prob.p.rhs_resident(1000) / 1000
    # the same algorithm the CPU reference arm runs, end to end through prob.f(du, u, p, t) with page-locked buffers
    import torch
    hu = [torch.empty(x.size, dtype=torch.float64).pin_memory() for x in u0.x]
    hdu = [torch.empty(x.size, dtype=torch.float64).pin_memory() for x in u0.x]
    u_np = [h.numpy().reshape(x.shape, order="F") for h, x in zip(hu, u0.x)]
    du_np = [h.numpy().reshape(x.shape, order="F") for h, x in zip(hdu, u0.x)]
    for dst, x in zip(u_np, u0.x):
        dst[...] = x
    for _ in range(20):
        prob.p.rhs(du_np, u_np, 0.0)
    n = 1000
    t0 = time.perf_counter()
    for _ in range(n):
        prob.p.rhs(du_np, u_np, 0.0)
    e2e = n / (time.perf_counter() - t0)
    return {"workload": "256x64 fault-only RHS, FFT/Toeplitz form (equation.jl:44-61) on the GPU",
            "rhs_evals_per_s": 1e3 / ms, "rhs_us": 1e3 * ms, "bytes_per_eval": prob.p.rhs_bytes(),
            "e2e_evals_per_s": e2e, "e2e_note": "host u -> device -> host du every call, page-locked buffers"}


def example_extra(oq):
    """BASELINE configs[1]: the example problem (N_f = 32, N_e = 36; 0.5 MB of matrices: launch-latency bound)."""
    fs, bs = W.C2_FAULT, W.C2_BOX
    mf = oq.gen_mesh("RectOkada", fs.x, fs.xi, fs.dx, fs.dxi, fs.dip)
    ma = oq.gen_mesh("BEMHex8Mesh", *bs.args())
    a, b, L, sig = W.fault_properties(mf.x, mf.z, mf.nx, mf.nxi)
    g, n, d0 = W.mantle_properties(ma.cz)
    v, th, eps, sg, dl = W.initial_state(mf.nx, mf.nxi, L, ma.cz, g, n)
    pf = oq.RateStateQuasiDynamicProperty(a, b, L, sig, W.ETA, W.VPL, W.F0, W.V0)
    pa = oq.PowerLawViscosityProperty(g, n, d0)
    d11 = oq.device_fault_fault(mf, W.LAM, W.MU, buffer_ratio=1.0)
    d12 = oq.device_fault_mantle(mf, ma, W.LAM, W.MU, buffer_ratio=1.0)
    d21 = oq.device_mantle_fault(ma, mf, W.LAM, W.MU)
    d22 = oq.device_mantle_mantle(ma, W.LAM, W.MU)
    u0 = oq.ArrayPartition(v, th, eps, sg, dl)
    prob = oq.assemble(d11, d12, d21, d22, pf, pa, u0, (0.0, 0.1 * W.YEAR))
    prob.p.set_state(u0.x)
    prob.p.rhs_resident(20)
    ms = prob.p.rhs_resident(500) / 500
    t0 = time.perf_counter()
    sol = oq.solve(prob, oq.Tsit5(), reltol=1e-6, abstol=1e-8, dt=1e-8, dtmax=0.2 * W.YEAR, maxiters=300,
                   save_everystep=False)
    wall = time.perf_counter() - t0
    steps = sol.stats["naccept"] + sol.stats["nreject"]
    # the example's own algorithm (otf-with-mantle.jl:160): simulated time per wall second is what the user sees
    year = {}
    for name, alg in (("tsit5", oq.Tsit5()), ("vcabm5", oq.VCABM5())):
        prob1 = oq.assemble(d11, d12, d21, d22, pf, pa, u0, (0.0, 1.0 * W.YEAR))
        t0 = time.perf_counter()
        s1 = oq.solve(prob1, alg, reltol=1e-6, abstol=1e-8, dt=1e-8, dtmax=0.2 * W.YEAR, maxiters=100000,
                      save_everystep=False)
        year[name] = {"wall_s": time.perf_counter() - t0, "steps": s1.stats["naccept"] + s1.stats["nreject"],
                      "rhs_evals": s1.stats["nf"], "retcode": s1.retcode}
    return {"workload": "BASELINE configs[1]: examples/otf-with-mantle.jl (8x4 fault, 4x3x3 hex8 mantle), coupled RHS",
            "rhs_evals_per_s": 1e3 / ms, "rhs_us": 1e3 * ms, "tsit5_steps_per_s": steps / wall,
            "tsit5_rhs_per_s": 6 * steps / wall, "one_simulated_year": year}


def class_form_extra(oq):
    """A coupled problem (4 096 fault cells + 4 608 hex8 cells) with the three mantle operands dense (what the north
    star streams) and in class form (csrc/classmat.cuh: tables of the distinct kernels, no dense storage): bytes held,
    RHS evaluations/s, and the two derivatives against each other."""
    fs = W.FaultSpec(32e3, 8e3, 250.0, 250.0)                                   # 128 x 32
    bs = W.BoxSpec(-fs.x / 2, -10e3, -fs.xi, fs.x, 20e3, -40e3, 32, 9, 16, tuple(np.cumprod(np.ones(16) * 1.1)))
    mf = oq.gen_mesh("RectOkada", fs.x, fs.xi, fs.dx, fs.dxi, fs.dip)
    ma = oq.gen_mesh("BEMHex8Mesh", *bs.args())
    a, b, L, sig = W.fault_properties(mf.x, mf.z, mf.nx, mf.nxi)
    g, n, d0 = W.mantle_properties(ma.cz)
    v, th, eps, sg, dl = W.initial_state(mf.nx, mf.nxi, L, ma.cz, g, n, rng=np.random.default_rng(42))
    v = v * (1 + 0.3 * np.random.default_rng(1).uniform(-1, 1, v.shape))
    pf = oq.RateStateQuasiDynamicProperty(a, b, L, sig, W.ETA, W.VPL, W.F0, W.V0)
    pa = oq.PowerLawViscosityProperty(g, n, d0)
    u0 = oq.ArrayPartition(v, th, eps, sg, dl)
    gf11 = oq.stress_greens_function(mf, W.LAM, W.MU, buffer_ratio=1.0)
    out = {"workload": f"coupled RHS, {mf.nx * mf.nxi} fault cells + {len(ma)} hex8 cells, gf11 in FFT form"}
    dus = {}
    for form in ("dense", "classes"):
        ops = (oq.device_fault_mantle(mf, ma, W.LAM, W.MU, buffer_ratio=1.0, form=form),
               oq.device_mantle_fault(ma, mf, W.LAM, W.MU, form=form),
               oq.device_mantle_mantle(ma, W.LAM, W.MU, form=form))
        prob = oq.assemble(gf11, *ops, pf, pa, u0, (0.0, 1.0), gf11_form="fft")
        prob.p.set_state(u0.x)
        prob.p.rhs_resident(5)
        ms = prob.p.rhs_resident(50) / 50
        du = u0.similar()
        prob.p.get_du(du.x)
        dus[form] = du
        out[form] = {"rhs_evals_per_s": 1e3 / ms, "rhs_ms": ms, "operand_bytes": sum(m.form()["device_bytes"] for m in ops)}
        del prob
        for m in ops:
            m.free()
    out["speedup"] = out["dense"]["rhs_ms"] / out["classes"]["rhs_ms"]
    out["max_rel_diff_vs_dense"] = max(comp_rel_err(x, y) for x, y in zip(dus["classes"].x, dus["dense"].x))
    return out


def _matvec_kernel_name(world):
    """the kernel oq_rhs launches for the dense operands (csrc/rhs.cu: matvec_variant): OQ_MATVEC overrides, else the
    fused panel kernel on one GPU and the forcing + streaming pair on row shards"""
    v = os.environ.get("OQ_MATVEC")
    if v == "ldg":
        return "matvec_fused_kernel"
    if v == "stream" or (v != "panel" and world > 1):
        return "matvec_stream_kernel"
    return "matvec_panel_kernel"


def hbm_only_extra(args):
    """The headline run again in a fresh process with the L2-residency hints and the alternating traversal switched
    off (every byte of the matrix comes from HBM on every evaluation): the plain streaming roofline."""
    import subprocess
    env = {**os.environ, "OQ_MATVEC_KEEP_MB": "0", "OQ_MATVEC_PINGPONG": "0"}
    try:
        res = subprocess.run([sys.executable, os.path.abspath(__file__), "--steps", str(args.steps), "--warmup",
                              str(args.warmup), "--no-extra", "--no-cpu", "--no-parity"], env=env, capture_output=True, text=True,
                             timeout=600)
        d = json.loads(res.stdout.strip().splitlines()[-1])
    except Exception as exc:              # a failed side run must not take the other extras down
        return {"error": repr(exc)}
    return {"value": d["value"], "unit": d["unit"], "roofline_achieved_gbs": d["roofline"]["achieved"],
            "roofline_frac": d["roofline"]["frac"], "kernel_ms": d["roofline"]["kernel_ms"],
            "env": "OQ_MATVEC_KEEP_MB=0 OQ_MATVEC_PINGPONG=0"}


# ------------------------------------------------------------------------------------------ main arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    import oetqf_b200 as oq

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with torchrun (one rank per GPU)")
    torch.cuda.set_device(local)
    oq.init(local)
    if world > 1:
        # NCCL prints its version banner on stdout at communicator creation: send fd 1 to stderr until the
        # first collective has completed, so that stdout carries exactly one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    fs = W.C3_FAULT
    nf = fs.nx * fs.nxi
    # contiguous row shards in rank order, multiples of 4 rows (the matvec's row-block size)
    r0, r1 = oq.dist.shard_range(nf, world, rank, align=4)
    mf, prob, u0, g11 = build_fault_problem(oq, fs, (r0, r1))
    p = prob.p
    if world > 1:
        oq.dist.connect(p)
    loc = oq.dist.local_state(u0.x, (r0, r1))
    p.set_state(loc)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident value ------------------------------------------------------------------
    nwarm = max(3, args.warmup)
    ms_w = p.rhs_resident(nwarm)
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ~0.3 s of further untimed evaluations while the clock sampler starts up: the GPUs stay at their working clocks
    # (an idle pause here lets them drop, and the timed region of K = 20 steps is only 1-6 ms long).  The count is the
    # same on every rank (evaluations are collective).
    t = torch.tensor([ms_w / nwarm], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    nkeep = int(min(5000, max(50, 300.0 / max(1e-3, float(t.item())))))
    p.rhs_resident(nkeep)
    torch.cuda.synchronize()
    launches0 = oq.kernel_launch_count()
    barrier()
    ms = p.rhs_resident(args.steps)          # the timed region: K evaluations, CUDA events on the library's stream
    barrier()
    launches = oq.kernel_launch_count() - launches0
    # second pass of the same K evaluations with CUDA events around every matvec launch (roofline line)
    p.profile_enable(True)
    barrier()
    ms_prof = p.rhs_resident(args.steps)
    barrier()
    mv_ms, mv_n = p.profile_read()
    p.profile_enable(False)
    if rank == 0:
        time.sleep(0.15)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = args.steps / (ms_max * 1e-3)

    # ---- end to end through prob.f(du, u, p, t) with pinned host buffers ---------------------------
    hu = [torch.empty(a.size, dtype=torch.float64).pin_memory() for a in loc]
    hdu = [torch.empty(a.size, dtype=torch.float64).pin_memory() for a in loc]
    for h, a in zip(hu, loc):
        h.numpy()[:] = a
    u_np = [h.numpy() for h in hu]
    du_np = [h.numpy() for h in hdu]
    for _ in range(max(3, args.warmup)):
        p.rhs(du_np, u_np, 0.0)
    ksteps = args.steps
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(ksteps):
        p.rhs(du_np, u_np, 0.0)          # u read from / du written to host memory by the kernels; synchronous
    ev1.record()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0     # host clock: the call is synchronous and includes both copies
    barrier()
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = ksteps / float(t.item())
    bytes_in = sum(a.nbytes for a in u_np)
    bytes_out = sum(a.nbytes for a in du_np)

    # ---- parity of THIS run's shard against the CPU oracle, on every rank, at every N ---------------
    parity = None
    if not args.no_parity:
        p.set_state(loc)
        p.rhs_resident(1)
        du_loc = [np.zeros(a.size) for a in loc]
        p.get_du(du_loc)
        ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        mine = fault_parity(oq, fs, p, g11, (r0, r1), du_loc, threads=max(1, ncpu // world))
        # the host-buffer path must deliver the same numbers as the resident one
        mine["e2e_vs_resident"] = max(comp_rel_err(a, b) for a, b in zip(du_np, du_loc)) if r1 > r0 else 0.0
        errs = torch.tensor([mine["kernel_max_rel_err"], mine["matrix_rows_max_rel_err"], mine["rhs_max_rel_err"],
                             mine["e2e_vs_resident"]], dtype=torch.float64, device="cuda")
        cnt = torch.tensor([mine["rows"], mine["matrix_entries"]], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(errs, op=dist.ReduceOp.MAX)
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        e = [float(x) for x in errs.tolist()]
        parity = {"max_rel_err": max(e[0], e[1], e[2]), "tol": PARITY_TOL, "rows": int(cnt[0].item()),
                  "ranks": world, "kernel_max_rel_err": e[0], "kernel_entries": mine["kernel_entries"],
                  "matrix_rows_max_rel_err": e[1], "matrix_entries": int(cnt[1].item()), "rhs_max_rel_err": e[2],
                  "e2e_vs_resident_max_rel_diff": e[3], "kernel_bit_identical_rank0": mine["kernel_bit_identical"],
                  "oracle": "oracle/ CPU restatement: gf_fault_fault (GF.jl:31-58), dense expansion "
                            "(test/BEM/tests.jl:46-49), Toeplitz-form RHS (equation.jl:156-166); every row of every rank",
                  "pass": bool(max(e) <= PARITY_TOL)}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        if parity is not None and not parity["pass"]:
            sys.exit(3)
        return

    peak_gbs, peak_src = measured_peaks()
    rhs_bytes = p.rhs_bytes()
    mat_bytes = 8.0 * (r1 - r0) * nf
    mv_avg_ms = mv_ms / max(1, mv_n)
    achieved = rhs_bytes / (mv_avg_ms * 1e-3) / 1e9
    traffic = None
    # the ncu capture is of the single-GPU launch (all 16384 rows); no capture exists at the shard sizes of N > 1
    traffic_src = None
    for name in ("r02_matvec_traffic.json", "r01_matvec_traffic.json"):
        prof = os.path.join(ROOT, "profiles", name)
        if world == 1 and os.path.exists(prof):
            with open(prof) as fh:
                traffic = json.load(fh).get("dram_bytes_per_launch")
            traffic_src = f"profiles/{name}: dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this launch"
            break
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "untimed_steps_before_timing": nwarm + nkeep,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": bench_config(fs, world),
        "algorithm": "dense row-sharded fp64 matvec with fused friction epilogue (the form the north star names)",
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                     "frac": achieved / peak_gbs, "traffic": traffic, "traffic_source": traffic_src,
                     "kernel": _matvec_kernel_name(world),
                     "kernel_ms": mv_avg_ms, "algorithmic_bytes_per_launch": rhs_bytes, "peak_source": peak_src,
                     "kernel_share_of_step": mv_ms / ms_prof if ms_prof > 0 else None,
                     "timing": "CUDA events around each matvec launch, second pass of the same K steps"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": bytes_in, "d2h_bytes_per_step": bytes_out},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if parity is not None:
        line["parity"] = parity
    if world == 1 and not args.no_cpu:
        base, _ = cpu_reference(fs, budget_s=10.0)
        line["cpu_baseline"] = base
        if not args.no_extra:
            try:
                fp64_peak = oq.measure_fp64_peak()
                line["extra"] = {"assembly": assembly_extras(oq, fp64_peak), "example": example_extra(oq),
                                 "fft_form": fft_form_extra(oq), "class_form": class_form_extra(oq),
                                 "hbm_only": hbm_only_extra(args),
                                 "hbm_copy_gbs_own_kernel": oq.measure_hbm_copy(1 << 30) / 1e9}
                # the plain streaming fraction (every byte from HBM on every evaluation) belongs next to the
                # headline fraction, which includes what L2 keeps between back-to-back evaluations
                line["roofline"]["frac_hbm_only"] = line["extra"]["hbm_only"].get("roofline_frac")
            except Exception as exc:      # extras must never take the headline line down
                line["extra"] = {"error": repr(exc)}
    print(json.dumps(line))
    sys.stdout.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if parity is not None and not parity["pass"]:
        sys.stderr.write(f"PARITY FAILURE: {parity}\n")
        sys.exit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg (used by the hbm_only sub-run)")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle parity block (used by the hbm_only sub-run)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
