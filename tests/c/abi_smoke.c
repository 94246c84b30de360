/*
 * tests/c/abi_smoke.c -- the drop-in boundary exercised WITHOUT Python: a plain C program (gcc) that includes only
 * include/oetqf_b200.h, links liboetqf_b200.so, and walks the path a `ccall` binding walks
 * (/root/reference/src/pref.jl:1-21, src/BEM/equation.jl:141-154, src/io.jl:128-130):
 *   meshes (mesh.jl:39-56 and the structured hex8 box) -> oq_gf_* (host arrays, GF.jl:31,123,194,250)
 *   -> oq_matrix_* (device-resident row shards) -> oq_problem_create_viscoelastic (assemble, equation.jl:141-154)
 *   -> oq_rhs (the (du,u,p,t) call, equation.jl:185-205) -> oq_solve with a snapshot callback -> oq_state_get.
 * Everything it computes is written to a binary file that tests/test_gpu_cabi.py compares with the same calls made
 * through the ctypes binding and with the CPU oracle.  The _Static_asserts pin the struct layouts the Julia shim
 * (oetqf.jl_b200/julia/OetqfB200.jl) mirrors.
 *
 * build: gcc -std=c11 -O1 -I include tests/c/abi_smoke.c -o <exe> -L oetqf.jl_b200 -loetqf_b200 -lm
 */
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "oetqf_b200.h"

_Static_assert(sizeof(OqFaultMesh) == 104 && offsetof(OqFaultMesh, x) == 8 && offsetof(OqFaultMesh, dx) == 72, "OqFaultMesh layout");
_Static_assert(sizeof(OqHex8Mesh) == 80 && offsetof(OqHex8Mesh, cx) == 8, "OqHex8Mesh layout");
_Static_assert(sizeof(OqQuadrature) == 24 && offsetof(OqQuadrature, coords) == 8, "OqQuadrature layout");
_Static_assert(sizeof(OqFaultProperty) == 64 && offsetof(OqFaultProperty, eta) == 32, "OqFaultProperty layout");
_Static_assert(sizeof(OqMantleProperty) == 32 && offsetof(OqMantleProperty, gamma) == 8, "OqMantleProperty layout");
_Static_assert(sizeof(OqDilatancyProperty) == 32, "OqDilatancyProperty layout");
_Static_assert(sizeof(OqSolveOptions) == 64 && offsetof(OqSolveOptions, maxiters) == 40 &&
               offsetof(OqSolveOptions, algorithm) == 48 && offsetof(OqSolveOptions, async_snapshots) == 56, "OqSolveOptions layout");
_Static_assert(sizeof(OqAssemblyInfo) == 48 && offsetof(OqAssemblyInfo, pairs) == 8 && offsetof(OqAssemblyInfo, table_ms) == 24, "OqAssemblyInfo layout");
_Static_assert(sizeof(OqSolveStats) == 56 && offsetof(OqSolveStats, naccept) == 24 && offsetof(OqSolveStats, retcode) == 48, "OqSolveStats layout");

#define CHECK(call)                                                                         \
    do {                                                                                    \
        if ((call) != 0) {                                                                  \
            fprintf(stderr, "%s failed: %s\n", #call, oq_last_error());                     \
            return 1;                                                                       \
        }                                                                                   \
    } while (0)

static int nsnap = 0;
static double last_t = 0;
static int on_snapshot(void *user, double t, int64_t step, const double *const *u, const double *const *du)
{
    (void)user; (void)step; (void)u; (void)du;
    ++nsnap;
    last_t = t;
    return 0;
}

static void put(FILE *f, const char *name, const double *a, size_t n)
{
    char tag[32] = {0};
    strncpy(tag, name, 31);
    unsigned long long nn = n;
    fwrite(tag, 1, 32, f);
    fwrite(&nn, sizeof nn, 1, f);
    fwrite(a, sizeof(double), n, f);
}

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: %s out.bin\n", argv[0]); return 2; }
    if (oq_abi_version() != OQ_ABI_VERSION) { fprintf(stderr, "header / library ABI mismatch\n"); return 1; }
    CHECK(oq_init(0));

    /* ---- examples/otf-with-mantle.jl:18,25-29: 80 km x 8 km fault in 10 km x 2 km cells, 4x3x3 hex8 box ---- */
    enum { NX = 8, NXI = 4, BX = 4, BY = 3, BZ = 3, NE = BX * BY * BZ, NF = NX * NXI };
    const double lam = 3e10, mu = 3e10;
    double x[NX], ax0[NX], ax1[NX], xi[NXI], axi0[NXI], axi1[NXI], y[NXI], z[NXI];
    for (int i = 0; i < NX; ++i) { x[i] = -40e3 + 5e3 + i * 10e3; ax0[i] = x[i] - 5e3; ax1[i] = x[i] + 5e3; }
    for (int j = 0; j < NXI; ++j) {
        xi[j] = (0.0 + j * (-2e3)) - 1e3; axi0[j] = xi[j] - 1e3; axi1[j] = xi[j] + 1e3;
        y[j] = xi[j] * 0.0; z[j] = xi[j] * 1.0;                       /* dip = 90: cosd = 0, sind = 1 */
    }
    OqFaultMesh mf = {NX, NXI, x, ax0, ax1, xi, axi0, axi1, y, z, 10e3, 2e3, 0.0, 90.0};
    double cx[NE], cy[NE], cz[NE], qx[NE], qy[NE], qz[NE], dx[NE], dy[NE], dz[NE];
    {
        const double rf[BZ] = {1.5, 2.25, 3.375};                     /* cumprod(1.5 * ones(3)) */
        double ze[BZ + 1] = {-8e3}, acc = 0, tot = 1.5 + 2.25 + 3.375;
        for (int k = 0; k < BZ; ++k) { acc += rf[k]; ze[k + 1] = -8e3 + acc / tot * (-22e3); }
        int e = 0;
        for (int k = 0; k < BZ; ++k)
            for (int j = 0; j < BY; ++j)
                for (int i = 0; i < BX; ++i, ++e) {
                    const double x0 = -40e3 + i * (80e3 / BX), x1 = -40e3 + (i + 1) * (80e3 / BX);
                    const double y0 = -2.5e3 + j * (5e3 / BY), y1 = -2.5e3 + (j + 1) * (5e3 / BY);
                    cx[e] = (x0 + x1) / 2; cy[e] = (y0 + y1) / 2; cz[e] = (ze[k] + ze[k + 1]) / 2;
                    dx[e] = fabs(x1 - x0); dy[e] = fabs(y1 - y0); dz[e] = fabs(ze[k + 1] - ze[k]);
                    qx[e] = cx[e]; qy[e] = cy[e] - dy[e] / 2; qz[e] = cz[e] + dz[e] / 2;   /* mesh.jl:181-183 */
                }
    }
    OqHex8Mesh ma = {NE, cx, cy, cz, qx, qy, qz, dx, dy, dz};

    FILE *out = fopen(argv[1], "wb");
    if (!out) { perror("fopen"); return 2; }

    /* ---- the four stress_greens_function methods, host arrays in the reference's layouts ---- */
    static double g11[NX * NXI * NXI], g12[6 * NE * NF], g21[NF * 6 * NE], g22[6 * NE * 6 * NE];
    CHECK(oq_gf_fault_fault(&mf, lam, mu, OQ_STRIKE_SLIP, 0, 2, 1.0, g11, NULL));
    CHECK(oq_gf_fault_mantle(&mf, &ma, NULL, lam, mu, OQ_STRIKE_SLIP, 2, 1.0, g12, NULL));
    CHECK(oq_gf_mantle_fault(&ma, &mf, lam, mu, OQ_STRIKE_SLIP, g21, NULL));
    CHECK(oq_gf_mantle_mantle(&ma, NULL, lam, mu, g22, NULL));
    put(out, "g11", g11, NX * NXI * NXI); put(out, "g12", g12, 6 * NE * NF);
    put(out, "g21", g21, NF * 6 * NE);    put(out, "g22", g22, 6 * NE * 6 * NE);

    /* ---- device-resident shards + the matvecmul! slot (pref.jl:15-21) ---- */
    OqMatrix *d11, *d12, *d21, *d22;
    CHECK(oq_matrix_fault_fault(&mf, lam, mu, OQ_STRIKE_SLIP, 2, 1.0, 0, NF, &d11));
    CHECK(oq_matrix_fault_mantle(&mf, &ma, NULL, lam, mu, OQ_STRIKE_SLIP, 2, 1.0, 0, NE, &d12));
    CHECK(oq_matrix_mantle_fault(&ma, &mf, lam, mu, OQ_STRIKE_SLIP, 0, NF, &d21));
    CHECK(oq_matrix_mantle_mantle(&ma, NULL, lam, mu, 0, NE, &d22));
    OqAssemblyInfo info;
    CHECK(oq_matrix_assembly_info(d22, &info));
    if (info.path < 0 || info.unique_pairs <= 0 || info.unique_pairs > info.pairs || info.pairs != (int64_t)NE * NE) {
        fprintf(stderr, "oq_matrix_assembly_info: path %d, %lld of %lld pairs\n", info.path, (long long)info.unique_pairs, (long long)info.pairs);
        return 1;
    }
    double xv[6 * NE], yv[NF];
    for (int i = 0; i < 6 * NE; ++i) xv[i] = sin(0.37 * i) * 1e-14;
    CHECK(oq_gemv(d21, xv, yv, 0));
    put(out, "gemv21", yv, NF);
    /* the same operand kept in class form (no dense storage): same handle type, same product */
    {
        OqMatrix *c22;
        int form = -1;
        double bytes = 0, yd[6 * NE], yc[6 * NE], worst = 0, scale = 0;
        CHECK(oq_matrix_mantle_mantle_classes(&ma, NULL, lam, mu, 0, NE, &c22));
        CHECK(oq_matrix_form(c22, &form, &bytes));
        CHECK(oq_gemv(d22, xv, yd, 0));
        CHECK(oq_gemv(c22, xv, yc, 0));
        for (int i = 0; i < 6 * NE; ++i) { worst = fmax(worst, fabs(yd[i] - yc[i])); scale = fmax(scale, fabs(yd[i])); }
        if (form != 1 || bytes <= 0 || !(worst <= 1e-12 * scale)) {
            fprintf(stderr, "class form: form %d, %g bytes, gemv differs by %g of %g\n", form, bytes, worst, scale);
            return 1;
        }
        put(out, "gemv22_classes", yc, 6 * NE);
        oq_matrix_destroy(c22);
    }

    /* ---- assemble (equation.jl:141-154) with the example's properties ---- */
    double a[NF], b[NF], L[NF], sg[NF], gam[NE], npw[NE], deps0[6] = {0, -1e-12, 0, 0, 0, 0};
    for (int j = 0; j < NXI; ++j)
        for (int i = 0; i < NX; ++i) {
            const int f = i + j * NX;
            a[f] = 0.015; b[f] = 0.015 - 0.0047 + ((i == 2 || i == 5) && j >= 1 ? 0.0094 : 0.0); L[f] = 8e-3;
            sg[f] = fmin(5e7, 1.5e6 + 18e3 * (-z[j]));
        }
    for (int e = 0; e < NE; ++e) { gam[e] = 1e-37 * (1.0 + 0.01 * e); npw[e] = 2.5; }
    const double vpl = 140e-3 / 365 / 86400;
    OqFaultProperty pf = {a, b, L, sg, 3e10 / (2 * 3044.14), vpl, 0.6, 1e-6};
    OqMantleProperty pa = {1, gam, npw, deps0};
    OqProblem *prob;
    CHECK(oq_problem_create_viscoelastic(NX, NXI, NE, OQ_GF11_DENSE, d11, NULL, d12, d21, d22, &pf, &pa, &prob));
    int nparts, lens[5];
    CHECK(oq_problem_layout(prob, &nparts, lens));
    if (nparts != 5 || lens[0] != NF || lens[2] != 6 * NE) { fprintf(stderr, "unexpected layout\n"); return 1; }

    /* ---- the (du, u, p, t) call with ordinary host arrays ---- */
    static double v[NF], th[NF], eps[6 * NE], sig[6 * NE], dl[NF];
    static double dv[NF], dth[NF], deps[6 * NE], dsig[6 * NE], ddl[NF];
    for (int f = 0; f < NF; ++f) { v[f] = vpl * (1.0 + 0.2 * sin(1.3 * f)); th[f] = L[f] / vpl / (f % NX < NX / 2 ? 1.1 : 2.5); dl[f] = 0; }
    for (int e = 0; e < NE; ++e)
        for (int k = 0; k < 6; ++k) {
            eps[e + k * NE] = 0;
            sig[e + k * NE] = (k == 0 || k == 3 || k == 5) ? 2e8 + 1e6 * e : (k == 1 ? -3e6 - 1e4 * e : 1e3 * (e - 7));
        }
    const double *u_parts[5] = {v, th, eps, sig, dl};
    double *du_parts[5] = {dv, dth, deps, dsig, ddl};
    CHECK(oq_rhs(prob, 0.0, u_parts, du_parts));
    put(out, "u_v", v, NF); put(out, "u_th", th, NF); put(out, "u_sig", sig, 6 * NE);
    put(out, "a", a, NF); put(out, "b", b, NF); put(out, "L", L, NF); put(out, "sigma", sg, NF); put(out, "gamma", gam, NE);
    put(out, "dv", dv, NF); put(out, "dth", dth, NF); put(out, "deps", deps, 6 * NE); put(out, "dsig", dsig, 6 * NE);
    put(out, "ddl", ddl, NF);

    /* ---- resident integration with a snapshot callback (io.jl:128-130) ---- */
    CHECK(oq_state_set(prob, u_parts));
    OqSolveOptions opt = {1e-6, 1e-8, 1e-8, 0.2 * 365 * 86400.0, 1e-3 * 365 * 86400.0, 200, OQ_ALG_TSIT5, 0, 0, 0};
    OqSolveStats st;
    CHECK(oq_solve(prob, 0.0, &opt, 1, on_snapshot, NULL, &st));
    double *fin[5] = {dv, dth, deps, dsig, ddl};
    CHECK(oq_state_get(prob, fin));
    put(out, "fin_v", dv, NF); put(out, "fin_th", dth, NF);
    double stats[6] = {st.t, (double)st.naccept, (double)st.nreject, (double)st.nrhs, (double)st.retcode, (double)nsnap};
    put(out, "stats", stats, 6);
    fclose(out);
    if (st.retcode != 0 || nsnap != st.naccept + 1 || last_t != st.t) {
        fprintf(stderr, "solve: retcode %d, %d snapshots for %lld steps, last t %g vs %g\n", st.retcode, nsnap,
                (long long)st.naccept, last_t, st.t);
        return 1;
    }

    /* NULL / malformed arguments fail with a message, never crash */
    if (oq_rhs(prob, 0.0, NULL, du_parts) == 0 || strlen(oq_last_error()) == 0) { fprintf(stderr, "NULL not rejected\n"); return 1; }
    oq_problem_destroy(prob);
    oq_matrix_destroy(d11); oq_matrix_destroy(d12); oq_matrix_destroy(d21); oq_matrix_destroy(d22);
    printf("abi_smoke ok: %lld steps, %lld rhs evaluations, %lld kernel launches\n", (long long)st.naccept,
           (long long)st.nrhs, (long long)oq_kernel_launch_count());
    return 0;
}
