/*
 * oracle/okada.c -- CPU restatement of Okada (1992) DC3D.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker, not the product: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * What it restates: the call `dc3d(x, y, z, alpha, dep, dip, al1, al2, aw1, aw2, d1, d2, d3, cache)`
 * made by the reference at src/BEM/GF.jl:49-54 and src/BEM/GF.jl:156-161.  The arithmetic itself
 * lives in the un-vendored dependency GeoGreensFunctions.jl (Project.toml:11, compat "0.2", no
 * Manifest => patch version unpinned), which is absent from /root/reference, so this file restates
 * the PUBLISHED algorithm (Okada 1992, BSSA 82(2), tables 6-9; executable spec in SURVEY.md
 * Appendix A) -- parity with GeoGreensFunctions.jl itself is UNPINNED.  It is pinned instead by
 * Okada (1985) table-2 checklist values, finite-difference consistency, free-surface tractions and
 * the SURVEY Appendix-E known-answer vectors (tests/test_oracle_okada.py).
 *
 * Output ordering (confirmed by src/BEM/GF.jl:77-78,83-85,163-169):
 *   u[0..2] = ux,uy,uz ; u[3..5] = d/dx (ux,uy,uz) ; u[6..8] = d/dy ; u[9..11] = d/dz
 */
#include <math.h>
#include <string.h>

#define OQ_EPS 1e-6
#define OQ_PI2 6.283185307179586476925286766559

/* sind/cosd with exact values at multiples of 90 degrees (Julia's sincosd behaviour). */
void oq_ref_sincosd(double deg, double *s, double *c)
{
    double r = fmod(deg, 360.0);
    if (r < 0) r += 360.0;
    if (r == 0.0)        { *s = 0.0;  *c = 1.0;  return; }
    if (r == 90.0)       { *s = 1.0;  *c = 0.0;  return; }
    if (r == 180.0)      { *s = 0.0;  *c = -1.0; return; }
    if (r == 270.0)      { *s = -1.0; *c = 0.0;  return; }
    long double a = (long double)deg * 3.14159265358979323846264338327950288L / 180.0L;
    *s = (double)sinl(a);
    *c = (double)cosl(a);
}

typedef struct {
    double alp1, alp2, alp3, alp4, alp5;
    double sd, cd, sdsd, cdcd, sdcd;
} med_t;

typedef struct {
    double xi, et, q;
    double xi2, et2, q2, r, r2, r3, r5, y, d, tt;
    double alx, ale, x11, y11, x32, y32;
    double ey, ez, fy, fz, gy, gz, hy, hz;
} geo_t;

static void medium(double alpha, double dip, med_t *m)
{
    m->alp1 = (1.0 - alpha) / 2.0;
    m->alp2 = alpha / 2.0;
    m->alp3 = (1.0 - alpha) / alpha;
    m->alp4 = 1.0 - alpha;
    m->alp5 = alpha;
    double sd, cd;
    oq_ref_sincosd(dip, &sd, &cd);
    if (fabs(cd) < OQ_EPS) {
        cd = 0.0;
        sd = (sd > 0.0) ? 1.0 : -1.0;
    }
    m->sd = sd; m->cd = cd;
    m->sdsd = sd * sd; m->cdcd = cd * cd; m->sdcd = sd * cd;
}

static void corner(double xi, double et, double q, const med_t *m, int kxi, int ket, geo_t *g)
{
    if (fabs(xi) < OQ_EPS) xi = 0.0;
    if (fabs(et) < OQ_EPS) et = 0.0;
    if (fabs(q)  < OQ_EPS) q  = 0.0;
    double sd = m->sd, cd = m->cd;
    g->xi = xi; g->et = et; g->q = q;
    g->xi2 = xi * xi; g->et2 = et * et; g->q2 = q * q;
    g->r2 = g->xi2 + g->et2 + g->q2;
    g->r = sqrt(g->r2);
    g->r3 = g->r * g->r2;
    g->r5 = g->r3 * g->r2;
    g->y = et * cd + q * sd;
    g->d = et * sd - q * cd;
    g->tt = (q == 0.0) ? 0.0 : atan(xi * et / (q * g->r));
    if (kxi) {
        g->alx = -log(g->r - xi); g->x11 = 0.0; g->x32 = 0.0;
    } else {
        double rxi = g->r + xi;
        g->alx = log(rxi);
        g->x11 = 1.0 / (g->r * rxi);
        g->x32 = (g->r + rxi) * g->x11 * g->x11 / g->r;
    }
    if (ket) {
        g->ale = -log(g->r - et); g->y11 = 0.0; g->y32 = 0.0;
    } else {
        double ret = g->r + et;
        g->ale = log(ret);
        g->y11 = 1.0 / (g->r * ret);
        g->y32 = (g->r + ret) * g->y11 * g->y11 / g->r;
    }
    g->ey = sd / g->r - g->y * q / g->r3;
    g->ez = cd / g->r + g->d * q / g->r3;
    g->fy = g->d / g->r3 + g->xi2 * g->y32 * sd;
    g->fz = g->y / g->r3 + g->xi2 * g->y32 * cd;
    g->gy = 2.0 * g->x11 * sd - g->y * q * g->x32;
    g->gz = 2.0 * g->x11 * cd + g->d * q * g->x32;
    g->hy = g->d * q * g->x32 + xi * q * g->y32 * sd;
    g->hz = g->y * q * g->x32 + xi * q * g->y32 * cd;
}

static void accumulate(double *u, double disl, const double *v)
{
    if (disl != 0.0)
        for (int i = 0; i < 12; ++i) u[i] += disl / OQ_PI2 * v[i];
}

/* part A: infinite-medium terms */
static void part_a(const med_t *m, const geo_t *g, double d1, double d2, double d3, double *u)
{
    double xi = g->xi, et = g->et, q = g->q, xi2 = g->xi2, q2 = g->q2, r = g->r, r3 = g->r3;
    double y = g->y, d = g->d, tt = g->tt, alx = g->alx, ale = g->ale;
    double x11 = g->x11, y11 = g->y11, y32 = g->y32;
    double ey = g->ey, ez = g->ez, fy = g->fy, fz = g->fz, gy = g->gy, gz = g->gz, hy = g->hy, hz = g->hz;
    double a1 = m->alp1, a2 = m->alp2, sd = m->sd, cd = m->cd;
    double xy = xi * y11, qx = q * x11, qy = q * y11;
    memset(u, 0, 12 * sizeof(double));
    double S[12] = {
        tt / 2 + a2 * xi * qy, a2 * q / r, a1 * ale - a2 * q * qy,
        -a1 * qy - a2 * xi2 * q * y32, -a2 * xi * q / r3, a1 * xy + a2 * xi * q2 * y32,
        a1 * xy * sd + a2 * xi * fy + d / 2 * x11, a2 * ey, a1 * (cd / r + qy * sd) - a2 * q * fy,
        a1 * xy * cd + a2 * xi * fz + y / 2 * x11, a2 * ez, -a1 * (sd / r - qy * cd) - a2 * q * fz };
    double D[12] = {
        a2 * q / r, tt / 2 + a2 * et * qx, a1 * alx - a2 * q * qx,
        -a2 * xi * q / r3, -qy / 2 - a2 * et * q / r3, a1 / r + a2 * q2 / r3,
        a2 * ey, a1 * d * x11 + xy / 2 * sd + a2 * et * gy, a1 * y * x11 - a2 * q * gy,
        a2 * ez, a1 * y * x11 + xy / 2 * cd + a2 * et * gz, -a1 * d * x11 - a2 * q * gz };
    double T[12] = {
        -a1 * ale - a2 * q * qy, -a1 * alx - a2 * q * qx, tt / 2 - a2 * (et * qx + xi * qy),
        -a1 * xy + a2 * xi * q2 * y32, -a1 / r + a2 * q2 / r3, -a1 * qy - a2 * q * q2 * y32,
        -a1 * (cd / r + qy * sd) - a2 * q * fy, -a1 * y * x11 - a2 * q * gy, a1 * (d * x11 + xy * sd) + a2 * q * hy,
        a1 * (sd / r - qy * cd) - a2 * q * fz, a1 * d * x11 - a2 * q * gz, a1 * (y * x11 + xy * cd) + a2 * q * hz };
    accumulate(u, d1, S); accumulate(u, d2, D); accumulate(u, d3, T);
}

/* part B: free-surface correction terms */
static void part_b(const med_t *m, const geo_t *g, double d1, double d2, double d3, double *u)
{
    double xi = g->xi, et = g->et, q = g->q, xi2 = g->xi2, q2 = g->q2, r = g->r, r3 = g->r3;
    double y = g->y, d = g->d, tt = g->tt, ale = g->ale;
    double x11 = g->x11, y11 = g->y11, y32 = g->y32;
    double ey = g->ey, ez = g->ez, fy = g->fy, fz = g->fz, gy = g->gy, gz = g->gz, hy = g->hy, hz = g->hz;
    double a3 = m->alp3, sd = m->sd, cd = m->cd, sdsd = m->sdsd, cdcd = m->cdcd, sdcd = m->sdcd;
    double rd = r + d, d11 = 1.0 / (r * rd);
    double aj2 = xi * y / rd * d11, aj5 = -(d + y * y / rd) * d11;
    double ai3, ai4, ak1, ak3, aj3, aj6;
    if (cd != 0.0) {
        if (xi == 0.0) {
            ai4 = 0.0;
        } else {
            double x = sqrt(xi2 + q2);
            ai4 = 1.0 / cdcd * (xi / rd * sdcd
                  + 2.0 * atan((et * (x + q * cd) + x * (r + x) * sd) / (xi * (r + x) * cd)));
        }
        ai3 = (y * cd / rd - ale + sd * log(rd)) / cdcd;
        ak1 = xi * (d11 - y11 * sd) / cd;
        ak3 = (q * y11 - y * d11) / cd;
        aj3 = (ak1 - aj2 * sd) / cd;
        aj6 = (ak3 - aj5 * sd) / cd;
    } else {
        double rd2 = rd * rd;
        ai3 = (et / rd + y * q / rd2 - ale) / 2;
        ai4 = xi * y / rd2 / 2;
        ak1 = xi * q / rd * d11;
        ak3 = sd / rd * (xi2 * d11 - 1.0);
        aj3 = -xi / rd2 * (q2 * d11 - 0.5);
        aj6 = -y / rd2 * (xi2 * d11 - 0.5);
    }
    double xy = xi * y11;
    double ai1 = -xi / rd * cd - ai4 * sd;
    double ai2 = log(rd) + ai3 * sd;
    double ak2 = 1.0 / r + ak3 * sd;
    double ak4 = xy * cd - ak1 * sd;
    double aj1 = aj5 * cd - aj6 * sd;
    double aj4 = -xy - aj2 * cd + aj3 * sd;
    double qx = q * x11, qy = q * y11;
    memset(u, 0, 12 * sizeof(double));
    double S[12] = {
        -xi * qy - tt - a3 * ai1 * sd, -q / r + a3 * y / rd * sd, q * qy - a3 * ai2 * sd,
        xi2 * q * y32 - a3 * aj1 * sd, xi * q / r3 - a3 * aj2 * sd, -xi * q2 * y32 - a3 * aj3 * sd,
        -xi * fy - d * x11 + a3 * (xy + aj4) * sd, -ey + a3 * (1.0 / r + aj5) * sd, q * fy - a3 * (qy - aj6) * sd,
        -xi * fz - y * x11 + a3 * ak1 * sd, -ez + a3 * y * d11 * sd, q * fz + a3 * ak2 * sd };
    double D[12] = {
        -q / r + a3 * ai3 * sdcd, -et * qx - tt - a3 * xi / rd * sdcd, q * qx + a3 * ai4 * sdcd,
        xi * q / r3 + a3 * aj4 * sdcd, et * q / r3 + qy + a3 * aj5 * sdcd, -q2 / r3 + a3 * aj6 * sdcd,
        -ey + a3 * aj1 * sdcd, -et * gy - xy * sd + a3 * aj2 * sdcd, q * gy + a3 * aj3 * sdcd,
        -ez - a3 * ak3 * sdcd, -et * gz - xy * cd - a3 * xi * d11 * sdcd, q * gz - a3 * ak4 * sdcd };
    double T[12] = {
        q * qy - a3 * ai3 * sdsd, q * qx + a3 * xi / rd * sdsd, et * qx + xi * qy - tt - a3 * ai4 * sdsd,
        -xi * q2 * y32 - a3 * aj4 * sdsd, -q2 / r3 - a3 * aj5 * sdsd, q * q2 * y32 - a3 * aj6 * sdsd,
        q * fy - a3 * aj1 * sdsd, q * gy - a3 * aj2 * sdsd, -q * hy - a3 * aj3 * sdsd,
        q * fz + a3 * ak3 * sdsd, q * gz + a3 * xi * d11 * sdsd, -q * hz + a3 * ak4 * sdsd };
    accumulate(u, d1, S); accumulate(u, d2, D); accumulate(u, d3, T);
}

/* part C: depth-dependent terms (multiplied by z by the caller) */
static void part_c(const med_t *m, const geo_t *g, double z, double d1, double d2, double d3, double *u)
{
    double xi = g->xi, et = g->et, q = g->q, xi2 = g->xi2, et2 = g->et2, q2 = g->q2;
    double r = g->r, r2 = g->r2, r3 = g->r3, r5 = g->r5, y = g->y, d = g->d;
    double x11 = g->x11, y11 = g->y11, x32 = g->x32, y32 = g->y32;
    double a4 = m->alp4, a5 = m->alp5, sd = m->sd, cd = m->cd, sdsd = m->sdsd, cdcd = m->cdcd, sdcd = m->sdcd;
    double c = d + z;
    double x53 = (8.0 * r2 + 9.0 * r * xi + 3.0 * xi2) * x11 * x11 * x11 / r2;
    double y53 = (8.0 * r2 + 9.0 * r * et + 3.0 * et2) * y11 * y11 * y11 / r2;
    double h = q * cd - z;
    double z32 = sd / r3 - h * y32;
    double z53 = 3.0 * sd / r5 - h * y53;
    double y0 = y11 - xi2 * y32;
    double z0 = z32 - xi2 * z53;
    double ppy = cd / r3 + q * y32 * sd;
    double ppz = sd / r3 - q * y32 * cd;
    double qq = z * y32 + z32 + z0;
    double qqy = 3.0 * c * d / r5 - qq * sd;
    double qqz = 3.0 * c * y / r5 - qq * cd + q * y32;
    double xy = xi * y11, qy = q * y11;
    double qr = 3.0 * q / r5;
    double cdr = (c + d) / r3;
    double yy0 = y / r3 - y0 * cd;
    memset(u, 0, 12 * sizeof(double));
    double S[12] = {
        a4 * xy * cd - a5 * xi * q * z32,
        a4 * (cd / r + 2.0 * qy * sd) - a5 * c * q / r3,
        a4 * qy * cd - a5 * (c * et / r3 - z * y11 + xi2 * z32),
        a4 * y0 * cd - a5 * q * z0,
        -a4 * xi * (cd / r3 + 2.0 * q * y32 * sd) + a5 * c * xi * qr,
        -a4 * xi * q * y32 * cd + a5 * xi * (3.0 * c * et / r5 - qq),
        -a4 * xi * ppy * cd - a5 * xi * qqy,
        a4 * 2.0 * (d / r3 - y0 * sd) * sd - y / r3 * cd - a5 * (cdr * sd - et / r3 - c * y * qr),
        -a4 * q / r3 + yy0 * sd + a5 * (cdr * cd + c * d * qr - (y0 * cd + q * z0) * sd),
        a4 * xi * ppz * cd - a5 * xi * qqz,
        a4 * 2.0 * (y / r3 - y0 * cd) * sd + d / r3 * cd - a5 * (cdr * cd + c * d * qr),
        yy0 * cd - a5 * (cdr * sd - c * y * qr - y0 * sdsd + q * z0 * cd) };
    double D[12] = {
        a4 * cd / r - qy * sd - a5 * c * q / r3,
        a4 * y * x11 - a5 * c * et * q * x32,
        -d * x11 - xy * sd - a5 * c * (x11 - q2 * x32),
        -a4 * xi / r3 * cd + a5 * c * xi * qr + xi * q * y32 * sd,
        -a4 * y / r3 + a5 * c * et * qr,
        d / r3 - y0 * sd + a5 * c / r3 * (1.0 - 3.0 * q2 / r2),
        -a4 * et / r3 + y0 * sdsd - a5 * (cdr * sd - c * y * qr),
        a4 * (x11 - y * y * x32) - a5 * c * ((d + 2.0 * q * cd) * x32 - y * et * q * x53),
        xi * ppy * sd + y * d * x32 + a5 * c * ((y + 2.0 * q * sd) * x32 - y * q2 * x53),
        -q / r3 + y0 * sdcd - a5 * (cdr * cd + c * d * qr),
        a4 * y * d * x32 - a5 * c * ((y - 2.0 * q * sd) * x32 + d * et * q * x53),
        -xi * ppz * sd + x11 - d * d * x32 - a5 * c * ((d - 2.0 * q * cd) * x32 - d * q2 * x53) };
    double T[12] = {
        -a4 * (sd / r + qy * cd) - a5 * (z * y11 - q2 * z32),
        a4 * 2.0 * xy * sd + d * x11 - a5 * c * (x11 - q2 * x32),
        a4 * (y * x11 + xy * cd) + a5 * q * (c * et * x32 + xi * z32),
        a4 * xi / r3 * sd + xi * q * y32 * cd + a5 * xi * (3.0 * c * et / r5 - 2.0 * z32 - z0),
        a4 * 2.0 * y0 * sd - d / r3 + a5 * c / r3 * (1.0 - 3.0 * q2 / r2),
        -a4 * yy0 - a5 * (c * et * qr - q * z0),
        a4 * (q / r3 + y0 * sdcd) + a5 * (z / r3 * cd + c * d * qr - q * z0 * sd),
        -a4 * 2.0 * xi * ppy * sd - y * d * x32 + a5 * c * ((y + 2.0 * q * sd) * x32 - y * q2 * x53),
        -a4 * (xi * ppy * cd - x11 + y * y * x32) + a5 * (c * ((d + 2.0 * q * cd) * x32 - y * et * q * x53) + xi * qqy),
        -et / r3 + y0 * cdcd - a5 * (z / r3 * sd - c * y * qr - y0 * sdsd + q * z0 * cd),
        a4 * 2.0 * xi * ppz * sd - x11 + d * d * x32 - a5 * c * ((d - 2.0 * q * cd) * x32 - d * q2 * x53),
        a4 * (xi * ppz * cd + y * d * x32) + a5 * (c * ((y - 2.0 * q * sd) * x32 + d * et * q * x53) + xi * qqz) };
    accumulate(u, d1, S); accumulate(u, d2, D); accumulate(u, d3, T);
}

typedef struct { double xi[2], et[2], q; int kxi[2], ket[2]; int singular; } setup_t;

static void setup(double x, double y, double dd, const med_t *m,
                  double al1, double al2, double aw1, double aw2, setup_t *s)
{
    double sd = m->sd, cd = m->cd;
    s->xi[0] = x - al1; s->xi[1] = x - al2;
    for (int k = 0; k < 2; ++k) if (fabs(s->xi[k]) < OQ_EPS) s->xi[k] = 0.0;
    double p = y * cd + dd * sd;
    double q = y * sd - dd * cd;
    s->et[0] = p - aw1; s->et[1] = p - aw2;
    if (fabs(q) < OQ_EPS) q = 0.0;
    for (int k = 0; k < 2; ++k) if (fabs(s->et[k]) < OQ_EPS) s->et[k] = 0.0;
    s->q = q;
    s->singular = (q == 0.0) &&
        ((s->xi[0] * s->xi[1] <= 0.0 && s->et[0] * s->et[1] == 0.0) ||
         (s->et[0] * s->et[1] <= 0.0 && s->xi[0] * s->xi[1] == 0.0));
    double r12 = sqrt(s->xi[0] * s->xi[0] + s->et[1] * s->et[1] + q * q);
    double r21 = sqrt(s->xi[1] * s->xi[1] + s->et[0] * s->et[0] + q * q);
    double r22 = sqrt(s->xi[1] * s->xi[1] + s->et[1] * s->et[1] + q * q);
    s->kxi[0] = (s->xi[0] < 0.0 && r21 + s->xi[1] < OQ_EPS);
    s->kxi[1] = (s->xi[0] < 0.0 && r22 + s->xi[1] < OQ_EPS);
    s->ket[0] = (s->et[0] < 0.0 && r12 + s->et[1] < OQ_EPS);
    s->ket[1] = (s->et[0] < 0.0 && r22 + s->et[1] < OQ_EPS);
}

void oq_ref_dc3d(double alpha, double x, double y, double z, double depth, double dip,
                 double al1, double al2, double aw1, double aw2,
                 double d1, double d2, double d3, double *u)
{
    memset(u, 0, 12 * sizeof(double));
    if (z > 0.0) return;
    med_t m; medium(alpha, dip, &m);
    double sd = m.sd, cd = m.cd;
    double acc[12] = {0};
    setup_t s; geo_t g;
    double A[12], B[12], C[12], du[12];

    /* real source */
    setup(x, y, depth + z, &m, al1, al2, aw1, aw2, &s);
    if (s.singular) return;
    for (int k = 0; k < 2; ++k)
        for (int j = 0; j < 2; ++j) {
            corner(s.xi[j], s.et[k], s.q, &m, s.kxi[k], s.ket[j], &g);
            part_a(&m, &g, d1, d2, d3, A);
            for (int i = 0; i < 12; i += 3) {
                du[i]     = -A[i];
                du[i + 1] = -A[i + 1] * cd + A[i + 2] * sd;
                du[i + 2] = -A[i + 1] * sd - A[i + 2] * cd;
                if (i == 9) { du[9] = -du[9]; du[10] = -du[10]; du[11] = -du[11]; }
            }
            double sgn = (j + k == 1) ? -1.0 : 1.0;
            for (int i = 0; i < 12; ++i) acc[i] += sgn * du[i];
        }

    /* image source */
    setup(x, y, depth - z, &m, al1, al2, aw1, aw2, &s);
    if (s.singular) return;
    for (int k = 0; k < 2; ++k)
        for (int j = 0; j < 2; ++j) {
            corner(s.xi[j], s.et[k], s.q, &m, s.kxi[k], s.ket[j], &g);
            part_a(&m, &g, d1, d2, d3, A);
            part_b(&m, &g, d1, d2, d3, B);
            part_c(&m, &g, z, d1, d2, d3, C);
            for (int i = 0; i < 12; i += 3) {
                du[i]     = A[i] + B[i] + z * C[i];
                du[i + 1] = (A[i + 1] + B[i + 1] + z * C[i + 1]) * cd - (A[i + 2] + B[i + 2] + z * C[i + 2]) * sd;
                du[i + 2] = (A[i + 1] + B[i + 1] - z * C[i + 1]) * sd + (A[i + 2] + B[i + 2] - z * C[i + 2]) * cd;
                if (i == 9) {
                    du[9]  += C[0];
                    du[10] += C[1] * cd - C[2] * sd;
                    du[11] -= C[1] * sd + C[2] * cd;
                }
            }
            double sgn = (j + k == 1) ? -1.0 : 1.0;
            for (int i = 0; i < 12; ++i) acc[i] += sgn * du[i];
        }
    memcpy(u, acc, sizeof(acc));
}
