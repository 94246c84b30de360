/*
 * oracle/hex8.c -- CPU restatement of `stress_vol_hex8!` as the reference calls it
 * (src/BEM/GF.jl:215-221 and :277-283; theta = 0).  TEST INFRASTRUCTURE ONLY.
 *
 * The routine belongs to the un-vendored GeoGreensFunctions.jl (Barbot et al. 2017, BSSA 107(2)); its
 * source is not available, so this file implements the DEFINITION of the kernel (stress of a uniform
 * eigenstrain in a cuboid of an elastic half-space) through the closed form derived
 * in oetqf.jl_b200/derive/hex8_derive.py (hex8_gen.inc is its plain-C output).  Parity with GeoGreensFunctions.jl is UNPINNED; the closed form is
 * pinned against an independent quadrature evaluation of the same definition (oracle/hex8_numeric.py,
 * tests/test_oracle_hex8.py) and against physical invariants.
 *
 * Geometry convention (src/BEM/mesh.jl:181-183, GF.jl:218): the cuboid spans
 *     x in [qx - dx/2, qx + dx/2],  y in [qy, qy + dy],  z in [qz - dz, qz]      (z up, z <= 0)
 * eps = (xx, xy, xz, yy, yz, zz) tensor components; output sigma in the same order.
 */
#include <math.h>
#include <string.h>

#include "hex8_gen.inc"
#define HEX8_NEED(mask) 1
#define HEX8_GROUP_BARRIER

static void corner_inputs(double r1, double r2, double r3, double *R, double *w, double *q, double *L, double *A,
                          double *iR, double *iw, double *iq)
{
    const double rs[3] = {r1, r2, r3};
    const double n2 = r1 * r1 + r2 * r2 + r3 * r3;
    const double n = sqrt(n2);
    *R = n;
    *iR = 1.0 / n;
    for (int c = 0; c < 3; ++c) {
        const int a = (c + 1) % 3, b = (c + 2) % 3;
        q[c] = rs[a] * rs[a] + rs[b] * rs[b];
        /* R + R_c without cancellation when R_c < 0 */
        w[c] = rs[c] >= 0.0 ? n + rs[c] : q[c] / (n - rs[c]);
        L[c] = log(w[c]);
        A[c] = atan(rs[a] * rs[b] / (rs[c] * n));
        iw[c] = 1.0 / w[c];
        iq[c] = 1.0 / q[c];
    }
}

/* The closed form is singular on the lines through the cuboid's edges; receivers within `nudge` of such a
 * line are moved off it along one axis (same rule as the product kernel, hex8_dev.cuh). */
static void regularise(double *r1, double *r2, double *r3, double nudge)
{
    const int t1 = fabs(*r1) < nudge, t2 = fabs(*r2) < nudge, t3 = fabs(*r3) < nudge;
    if (t1 && (t2 || t3)) *r1 = nudge;
    else if (t2 && t3) *r2 = nudge;
}

static void basis_real(double R1, double R2, double R3, double R, double w1, double w2, double w3, double q1,
                       double q2, double q3, double iR, double iw1, double iw2, double iw3, double iq1, double iq2,
                       double iq3, double L1, double L2, double L3, double A1, double A2, double A3, double sgn,
                       double *acc)
{
    (void)w1; (void)w2; (void)w3; (void)q1; (void)q2; (void)q3; (void)iw1; (void)iw2; (void)iw3;
    (void)iq1; (void)iq2; (void)iq3; (void)L1; (void)L2; (void)L3; (void)A1; (void)A2; (void)A3; (void)R; (void)iR;
#define ACC(b) acc[b]
    HEX8_BASIS_REAL_BODY
#undef ACC
}

static void basis_image(double R1, double R2, double R3, double R, double w1, double w2, double w3, double q1,
                        double q2, double q3, double iR, double iw1, double iw2, double iw3, double iq1, double iq2,
                        double iq3, double L1, double L2, double L3, double A1, double A2, double A3, double Ba,
                        double Bb, double sgn, double *acc)
{
    (void)w1; (void)w2; (void)w3; (void)q1; (void)q2; (void)q3; (void)iw1; (void)iw2; (void)iw3;
    (void)iq1; (void)iq2; (void)iq3; (void)L1; (void)L2; (void)L3; (void)A1; (void)A2; (void)A3; (void)Ba; (void)Bb;
#define ACC(b) acc[b]
    HEX8_BASIS_IMAGE_BODY
#undef ACC
}

/* Q[(il),(jk)] (times 8*pi*mu) from the 8-corner sums of the basis functions */
static void hex8_strain_kernels(double x, double y, double z, double qx, double qy, double qz,
                                double dx, double dy, double dz, double al, double *Q)
{
    double accr[HEX8_NB_REAL], acci[HEX8_NB_IMAGE];
    memset(accr, 0, sizeof(accr));
    memset(acci, 0, sizeof(acci));
    const double xs[2] = {qx - dx / 2, qx + dx / 2};
    const double ys[2] = {qy, qy + dy};
    const double zs[2] = {qz - dz, qz};
    const double nudge = 1e-6 * fmin(dx, fmin(dy, dz));
    for (int c3 = 0; c3 < 2; ++c3)
        for (int c2 = 0; c2 < 2; ++c2)
            for (int c1 = 0; c1 < 2; ++c1) {
                const double sgn = ((c1 + c2 + c3) & 1) ? 1.0 : -1.0;   /* s1*s2*s3 with s = -1 at the lower limit */
                double r1 = x - xs[c1], r2 = y - ys[c2], r3r = z - zs[c3];
                double R, w[3], q[3], L[3], A[3], iR, iw[3], iq[3];
                /* real source */
                regularise(&r1, &r2, &r3r, nudge);
                corner_inputs(r1, r2, r3r, &R, w, q, L, A, &iR, iw, iq);
                basis_real(r1, r2, r3r, R, w[0], w[1], w[2], q[0], q[1], q[2], iR, iw[0], iw[1], iw[2],
                           iq[0], iq[1], iq[2], L[0], L[1], L[2], A[0], A[1], A[2], sgn, accr);
                /* image source */
                r1 = x - xs[c1]; r2 = y - ys[c2];
                double r3 = -z - zs[c3];
                regularise(&r1, &r2, &r3, nudge);
                corner_inputs(r1, r2, r3, &R, w, q, L, A, &iR, iw, iq);
                basis_image(r1, r2, r3, R, w[0], w[1], w[2], q[0], q[1], q[2], iR, iw[0], iw[1], iw[2],
                            iq[0], iq[1], iq[2], L[0], L[1], L[2], A[0], A[1], A[2], atan(r1 / r2), atan(r2 / r1),
                            sgn, acci);
            }
    const double x3 = z, ial = 1.0 / al;
#define ACCR(b) accr[b]
#define ACCI(b) acci[b]
    HEX8_COMBINE_BODY
#undef ACCR
#undef ACCI
}

void oq_ref_stress_vol_hex8(double x, double y, double z, double qx, double qy, double qz,
                            double dx, double dy, double dz, const double *eps,
                            double mu, double nu, double *sig)
{
    const double lam = 2.0 * mu * nu / (1.0 - 2.0 * nu);
    const double alpha = (lam + mu) / (lam + 2.0 * mu);
    double Q[36];
    hex8_strain_kernels(x, y, z, qx, qy, qz, dx, dy, dz, alpha, Q);
    /* moment density m = lam tr(eps) I + 2 mu eps, pairs (xx,xy,xz,yy,yz,zz) */
    const double tr = eps[0] + eps[3] + eps[5];
    const double m[6] = {lam * tr + 2 * mu * eps[0], 2 * mu * eps[1], 2 * mu * eps[2],
                         lam * tr + 2 * mu * eps[3], 2 * mu * eps[4], lam * tr + 2 * mu * eps[5]};
    const double pref = 1.0 / (8.0 * 3.14159265358979323846 * mu);
    double e[6];
    for (int a = 0; a < 6; ++a) {
        double s = 0.0;
        for (int b = 0; b < 6; ++b) s += m[b] * Q[6 * a + b];
        e[a] = pref * s;
    }
    const int inside = x > qx - dx / 2 && x < qx + dx / 2 && y > qy && y < qy + dy && z > qz - dz && z < qz;
    if (inside)
        for (int a = 0; a < 6; ++a) e[a] -= eps[a];
    const double ekk = e[0] + e[3] + e[5];
    sig[0] = lam * ekk + 2 * mu * e[0];
    sig[1] = 2 * mu * e[1];
    sig[2] = 2 * mu * e[2];
    sig[3] = lam * ekk + 2 * mu * e[3];
    sig[4] = 2 * mu * e[4];
    sig[5] = lam * ekk + 2 * mu * e[5];
}
