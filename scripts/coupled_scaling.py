"""Coupled fault + mantle problem at BASELINE configs[3] scale (as far as HBM allows): assembly straight into
row shards, RHS throughput, roofline.  Run directly (1 GPU) or under torchrun (N GPUs).

  python scripts/coupled_scaling.py --nx 250 --nxi 80 --mantle 40 20 24     # 20k fault cells + 19.2k hex8 cells
  python scripts/coupled_scaling.py --form classes --gf11 fft --mantle 50 40 40   # configs[3] as written: 20k + 80k cells,
                                                                                  # mantle operands in class form (csrc/classmat.cuh)
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oetqf_b200 as oq  # noqa: E402
import workloads as W  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=250)
    ap.add_argument("--nxi", type=int, default=80)
    # odd ny: the fault plane y = 0 is not a cell face; 50 cells along strike = 5 fault cells each (commensurate grids:
    # the pairs fall into translation classes, csrc/greens_classes.cuh)
    ap.add_argument("--mantle", type=int, nargs=3, default=[50, 17, 23])
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--out", default="")
    ap.add_argument("--form", default="dense", choices=["dense", "classes"],
                    help="storage of gf12 / gf21 / gf22: dense row shards (the north star's form) or class form")
    ap.add_argument("--gf11", default="dense", choices=["dense", "fft"])
    ap.add_argument("--check-sources", type=int, default=3,
                    help="class form: check gf22 x against columns evaluated pointwise (oq_stress_vol_hex8) for this many sources")
    ap.add_argument("--solve-steps", type=int, default=0,
                    help="also take this many Tsit5 steps with the device-resident integrator (oq_solve) and report the time per step")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    oq.init(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    fs = W.FaultSpec(args.nx * 250.0, args.nxi * 250.0, 250.0, 250.0)
    mx, my, mz = args.mantle
    bs = W.BoxSpec(-fs.x / 2, -10e3, -fs.xi, fs.x, 20e3, -40e3, mx, my, mz, tuple(np.cumprod(np.ones(mz) * 1.1)))
    mf = oq.gen_mesh("RectOkada", fs.x, fs.xi, fs.dx, fs.dxi, fs.dip)
    ma = oq.gen_mesh("BEMHex8Mesh", *bs.args())
    nf, ne = mf.nx * mf.nxi, len(ma)
    rows = oq.dist.shard_range(nf, world, rank, align=4)
    elems = oq.dist.shard_range(ne, world, rank)
    t0 = time.perf_counter()
    import ctypes as C

    def kms(m):
        ms = C.c_double()
        oq._lib.check(oq._lib.load().oq_matrix_kernel_ms(m.handle, C.byref(ms)))
        return ms.value

    if args.gf11 == "dense":
        d11 = oq.device_fault_fault(mf, W.LAM, W.MU, buffer_ratio=1.0, rows=rows)
    else:
        d11 = oq.stress_greens_function(mf, W.LAM, W.MU, buffer_ratio=1.0)      # Fourier form, as the reference passes it
    d12 = oq.device_fault_mantle(mf, ma, W.LAM, W.MU, buffer_ratio=1.0, elems=elems, form=args.form)
    d21 = oq.device_mantle_fault(ma, mf, W.LAM, W.MU, rows=rows, form=args.form)
    d22 = oq.device_mantle_mantle(ma, W.LAM, W.MU, elems=elems, form=args.form)
    torch.cuda.synchronize()
    asm_wall = time.perf_counter() - t0
    asm = {"gf11_ms": kms(d11) if args.gf11 == "dense" else oq.gf.last_kernel_ms["value"], "gf12_ms": kms(d12), "gf21_ms": kms(d21), "gf22_ms": kms(d22),
           "gf12_entries_per_s": d12.local_rows * d12.cols / (kms(d12) * 1e-3),
           "gf21_entries_per_s": d21.local_rows * d21.cols / (kms(d21) * 1e-3),
           "gf22_entries_per_s": d22.local_rows * d22.cols / (kms(d22) * 1e-3)}
    for key, m in (("gf12", d12), ("gf21", d21), ("gf22", d22)):
        info = m.assembly_info()
        asm[key + "_path"] = info["path"]
        asm[key + "_closed_form_evaluations"] = info["unique_pairs"]
        asm[key + "_pairs"] = info["pairs"]
        asm[key + "_table_ms"], asm[key + "_expand_ms"] = info["table_ms"], info["expand_ms"]
    a, b, L, sig = W.fault_properties(mf.x, mf.z, mf.nx, mf.nxi)
    g, n, d0 = W.mantle_properties(ma.cz)
    v, th, eps, sg, dl = W.initial_state(mf.nx, mf.nxi, L, ma.cz, g, n, rng=np.random.default_rng(42))
    pf = oq.RateStateQuasiDynamicProperty(a, b, L, sig, W.ETA, W.VPL, W.F0, W.V0)
    pa = oq.PowerLawViscosityProperty(g, n, d0)
    u0 = oq.ArrayPartition(v, th, eps, sg, dl)
    prob = oq.assemble(d11, d12, d21, d22, pf, pa, u0, (0.0, 1.0), gf11_form=args.gf11, fault_rows=rows)
    p = prob.p
    if world > 1:
        oq.dist.connect(p)
    p.set_state(oq.dist.local_state(u0.x, rows, elems, kind="viscoelastic"))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    p.rhs_resident(args.warmup)
    barrier()
    ms = p.rhs_resident(args.steps)
    barrier()
    p.profile_enable(True)
    ms2 = p.rhs_resident(args.steps)
    mv_ms, mv_n = p.profile_read()
    p.profile_enable(False)
    t = torch.tensor([ms, asm["gf22_ms"], asm["gf21_ms"], asm["gf12_ms"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    du = [np.zeros(k) for k in p.local_lengths]
    p.get_du(du)
    finite = all(np.all(np.isfinite(x)) for x in du)
    # class form at full scale: gf22 x for x supported on a few source cells == the same columns evaluated pointwise
    # through the closed form (oq_stress_vol_hex8), on every receiver of this rank's shard
    col_err = None
    if args.form == "classes" and args.check_sources > 0 and elems[1] > elems[0]:
        rng = np.random.default_rng(7)
        src = elems[0] + rng.choice(elems[1] - elems[0], size=args.check_sources, replace=False)   # sources inside the shard:
        # the columns' largest entries (self and neighbour terms) are then among the receivers checked
        x = np.zeros(6 * ne)
        e0, e1 = elems
        want = np.zeros((6, e1 - e0))
        nu = W.LAM / 2 / (W.LAM + W.MU)
        for i in src:
            for pc in range(6):
                c = rng.uniform(0.5, 1.5)
                x[pc * ne + i] = c
                eps6 = np.zeros(6)
                eps6[pc] = 1.0
                col = oq.stress_vol_hex8(ma.cx[e0:e1], ma.cy[e0:e1], ma.cz[e0:e1], ma.qx[i], ma.qy[i], ma.qz[i],
                                         ma.dx[i], ma.dy[i], ma.dz[i], eps6, W.MU, nu)       # [receivers, 6]
                want += c * col.T
        got = d22.gemv(x).reshape(6, e1 - e0)
        col_err = float(np.max(np.abs(got - want)) / np.max(np.abs(want)))
    solve_rec = None
    if args.solve_steps > 0:
        loc = oq.dist.local_state(u0.x, rows, elems, kind="viscoelastic")
        barrier()
        t1 = time.perf_counter()
        sol = oq.solve(prob, oq.Tsit5(), reltol=1e-6, abstol=1e-8, dt=1e-6, dtmax=0.2 * W.YEAR, maxiters=args.solve_steps,
                       save_everystep=False, local_u0=loc)
        barrier()
        wall = time.perf_counter() - t1
        steps = sol.stats["naccept"] + sol.stats["nreject"]
        solve_rec = {"algorithm": "Tsit5 (oq_solve, state resident on the GPUs)", "steps": steps, "accepted": sol.stats["naccept"],
                     "rhs_evaluations": sol.stats["nf"], "simulated_s": sol.stats["t"], "wall_s": wall,
                     "ms_per_step": 1e3 * wall / max(1, steps), "retcode": sol.retcode,
                     "finite": bool(all(np.all(np.isfinite(np.asarray(x))) for x in sol.u[-1].x))}
    forms = {k: m.form() for k, m in (("gf12", d12), ("gf21", d21), ("gf22", d22))}
    fma = sum((m.local_rows * m.cols) for m in (d12, d21, d22)) if args.form == "classes" else 0
    if rank == 0:
        byts = p.rhs_bytes()
        line = {"workload": f"coupled: {nf} fault cells + {ne} hex8 cells ({6 * ne} mantle rows)", "n_gpus": world,
                "rhs_evals_per_s": args.steps / (float(t[0]) * 1e-3), "ms_per_eval": float(t[0]) / args.steps,
                "matrix_bytes_per_rank": byts, "matvec_ms": mv_ms / max(1, mv_n),
                "matvec_gbs": byts / (mv_ms / max(1, mv_n) * 1e-3) / 1e9 if mv_ms > 0 else None, "finite": bool(finite),
                "assembly_ms_max_over_ranks": {"gf22": float(t[1]), "gf21": float(t[2]), "gf12": float(t[3])},
                "assembly_rank0": asm, "assembly_wall_s": asm_wall,
                "form": args.form, "gf11": args.gf11, "operands_rank0": forms,
                "dense_equivalent_bytes_total": 8.0 * (nf * nf + 2 * 6 * ne * nf + (6 * ne) ** 2), "solve": solve_rec}
        if args.form == "classes":
            peak = oq.measure_fp64_peak()
            line["class_form"] = {"fma_per_eval_rank0": fma, "achieved_tflops": 2 * fma / (float(t[0]) / args.steps * 1e-3) / 1e12,
                                  "fp64_peak_tflops": peak / 1e12,
                                  "frac_of_fp64_peak": 2 * fma / (float(t[0]) / args.steps * 1e-3) / peak,
                                  "gf22_columns_vs_pointwise_closed_form_max_rel_err": col_err,
                                  "note": "fp64 FMA from shared-memory tables (csrc/classmat.cuh): window kernel where the grids are commensurate, general kernel (<= 25 % of the DFMA peak) otherwise"}
        print(json.dumps(line))
        if args.out:
            with open(args.out, "w") as fh:
                json.dump(line, fh, indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
