"""The reference's example, /root/reference/examples/otf-with-mantle.jl, on the B200 path.

Same problem statement, geometry, parameters and call sequence (line numbers of the Julia script in comments);
differences: the mantle box comes from the structured builder instead of Gmsh, the integrator is the
device-resident VCABM5 (or Tsit5 with --alg tsit5), and snapshots go to an .npy store instead of HDF5.

    python examples/otf_with_mantle.py [--years 0.1] [--out /tmp/otf_output] [--alg vcabm5|tsit5]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oetqf_b200 as oq  # noqa: E402
import workloads as W  # noqa: E402


def main(years=0.1, out="/tmp/otf_output", quiet=False, alg="vcabm5"):
    oq.init(0)
    # -- meshes (otf-with-mantle.jl:18, :25-29)
    mf = oq.gen_mesh("RectOkada", 80e3, 8e3, 10e3, 2e3, 90.0)
    ma = oq.gen_mesh("BEMHex8Mesh", -40e3, -2.5e3, -8e3, 80e3, 5e3, -22e3, 4, 3, 3,
                     rfzh=np.cumprod(np.ones(3) * 1.5))
    # -- Green's functions (:36-56)
    lam = mu = 3e10
    t0 = time.perf_counter()
    gf11 = oq.stress_greens_function(mf, lam, mu, buffer_ratio=1)                      # fault -> fault
    gf12 = oq.stress_greens_function(mf, ma, lam, mu, buffer_ratio=1, qtype="Gauss1")  # fault -> mantle
    gf21 = oq.stress_greens_function(ma, mf, lam, mu)                                  # mantle -> fault
    gf22 = oq.stress_greens_function(ma, lam, mu, qtype="Gauss1")                      # mantle -> mantle
    t_gf = time.perf_counter() - t0
    # -- parameters (:75-121)
    a, b, L, sigma = W.fault_properties(mf.x, mf.z, mf.nx, mf.nxi)
    pf = oq.RateStateQuasiDynamicProperty(a, b, L, sigma, W.ETA, W.VPL, W.F0, W.V0)
    gamma, nm1, deps0 = W.mantle_properties(ma.cz)
    pa = oq.PowerLawViscosityProperty(gamma, nm1, deps0)          # note: n - 1 is stored, as in the reference
    # -- initial conditions (:132-149)
    v, theta, eps, sig, delta = W.initial_state(mf.nx, mf.nxi, L, ma.cz, gamma, nm1)
    # -- assemble and solve (:153-162)
    uinit = oq.ArrayPartition(v, theta, eps, sig, delta)
    prob = oq.assemble(gf11, gf12, gf21, gf22, pf, pa, uinit, (0.0, years * W.YEAR))
    handler = lambda u, t, du: (u.x[0], u.x[1], du.x[2], u.x[2], u.x[3], u.x[4])      # noqa: E731  (:158)
    t0 = time.perf_counter()
    algorithm = oq.VCABM5() if alg == "vcabm5" else oq.Tsit5()                          # (:160)
    sol = oq.wsolve(prob, algorithm, out, 100, handler, ["v", "θ", "dϵ", "ϵ", "σ", "δ"], "t",
                    reltol=1e-6, abstol=1e-8, dtmax=0.2 * W.YEAR, dt=1e-8, maxiters=int(1e7), stride=100, force=True)
    t_solve = time.perf_counter() - t0
    tt = oq.io.read(out, "t")
    vv = oq.io.read(out, "v")
    if not quiet:
        print(f"Green's functions: {t_gf:.3f} s; solve: {t_solve:.3f} s, {sol.stats['naccept']} accepted / "
              f"{sol.stats['nreject']} rejected steps, {sol.stats['nf']} RHS evaluations ({alg}), "
              f"retcode {sol.retcode}")
        print(f"saved {len(tt)} snapshots to {out}: t[-1] = {tt[-1] / W.YEAR:.4f} yr, "
              f"max slip rate over the run = {vv.max():.3e} m/s (plate rate {W.VPL:.3e})")
    return sol, tt, vv


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--years", type=float, default=0.1)
    ap.add_argument("--out", default="/tmp/otf_output")
    ap.add_argument("--alg", choices=("vcabm5", "tsit5"), default="vcabm5")
    args = ap.parse_args()
    main(args.years, args.out, alg=args.alg)
