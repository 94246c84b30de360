/*
 * oracle/greens.c -- CPU restatement of the reference's Green's-function builders.
 * TEST INFRASTRUCTURE ONLY (see oracle/okada.c header): the checker and the timed CPU baseline.
 *
 * Loop structure follows the reference: threads over SOURCES, serial loops over receivers,
 * quadrature points and periodic images (src/BEM/GF.jl:42-58, :141-172, :206-225, :262-290).
 * All matrices are column-major exactly as the Julia arrays they restate.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

void oq_ref_sincosd(double deg, double *s, double *c);
void oq_ref_dc3d(double alpha, double x, double y, double z, double depth, double dip,
                 double al1, double al2, double aw1, double aw2,
                 double d1, double d2, double d3, double *u);
void oq_ref_stress_vol_hex8(double x, double y, double z, double qx, double qy, double qz,
                            double dx, double dy, double dz, const double *eps,
                            double mu, double nu, double *sig);

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline asks for all cores explicitly */
void oq_ref_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int oq_ref_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* src/BEM/GF.jl:76-87 : shear traction from the 12-vector of dc3d (1-based u[k] -> u[k-1]) */
static double shear_traction_dc3d(int ftype, const double *u, double lam, double mu, double dip)
{
    double s1, c1;
    if (ftype == 0) {                       /* StrikeSlip */
        oq_ref_sincosd(dip, &s1, &c1);
        double sxy = mu * (u[4] + u[6]);
        double sxz = mu * (u[5] + u[9]);
        return -sxy * s1 + sxz * c1;
    }
    oq_ref_sincosd(2.0 * dip, &s1, &c1);    /* DipSlip */
    double szz = (lam + 2 * mu) * u[11] + lam * u[3] + lam * u[7];
    double syy = (lam + 2 * mu) * u[7] + lam * u[3] + lam * u[11];
    double syz = mu * (u[10] + u[8]);
    return (szz - syy) / 2 * s1 + syz * c1;
}

/* src/BEM/GF.jl:89-96 */
static double shear_traction_vol(int ftype, const double *s, double dip)
{
    double s1, c1;
    if (ftype == 0) {
        oq_ref_sincosd(dip, &s1, &c1);
        return -s[1] * s1 + s[2] * c1;
    }
    oq_ref_sincosd(2.0 * dip, &s1, &c1);
    return (s[5] - s[3]) / 2 * s1 + s[4] * c1;
}

/*
 * Fault -> fault, Toeplitz-compressed kernel st[nx, nxi, nxi] (col-major), src/BEM/GF.jl:31-58.
 * x[nx], ax0/ax1 = edges of strike cell #1 (mesh.ax[1]); xi-arrays of length nxi.
 */
void oq_ref_gf_fault_fault(int nx, int nxi, const double *x, double ax0, double ax1,
                           const double *y, const double *z, const double *axi0, const double *axi1,
                           double dx, double dep, double dip, double lam, double mu,
                           int ftype, int nrept, double buffer_ratio, double *st)
{
    double lrept = (buffer_ratio + 1.0) * (dx * nx);
    double alpha = (lam + mu) / (lam + 2 * mu);
    double ud1 = ftype == 0 ? 1.0 : 0.0, ud2 = ftype == 0 ? 0.0 : 1.0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int l = 0; l < nxi; ++l) {
        double u[12], c[12];
        for (int j = 0; j < nxi; ++j)
            for (int i = 0; i < nx; ++i) {
                memset(u, 0, sizeof(u));
                for (int p = -nrept; p <= nrept; ++p) {
                    double jump = p * lrept;
                    oq_ref_dc3d(alpha, x[i], y[j], z[j], dep, dip, ax0 + jump, ax1 + jump,
                                axi0[l], axi1[l], ud1, ud2, 0.0, c);
                    for (int k = 0; k < 12; ++k) u[k] += c[k];
                }
                st[i + (size_t)nx * (j + (size_t)nxi * l)] = shear_traction_dc3d(ftype, u, lam, mu, dip);
            }
    }
}

/*
 * Fault -> mantle, st[6*ne, nx*nxi] col-major, src/BEM/GF.jl:123-174.
 * ax0/ax1[nx]: strike-cell edges; quadrature: lc[3*nq] local coords in [-1,1], w[nq] weights.
 */
void oq_ref_gf_fault_mantle(int nx, int nxi, const double *ax0, const double *ax1,
                            const double *axi0, const double *axi1,
                            double dx, double dep, double dip,
                            int ne, const double *cx, const double *cy, const double *cz,
                            const double *ex, const double *ey, const double *ez,
                            int nq, const double *lc, const double *w,
                            double lam, double mu, int ftype, int nrept, double buffer_ratio, double *st)
{
    double lrept = (buffer_ratio + 1.0) * (dx * nx);
    double alpha = (lam + mu) / (lam + 2 * mu);
    double ud1 = ftype == 0 ? 1.0 : 0.0, ud2 = ftype == 0 ? 0.0 : 1.0;
    size_t nrow = 6 * (size_t)ne;
    int ndisl = nx * nxi;
    memset(st, 0, sizeof(double) * nrow * ndisl);
#pragma omp parallel for schedule(dynamic, 4)
    for (int j = 0; j < ndisl; ++j) {
        double u[12], c[12];
        int q1 = j % nx, q2 = j / nx;
        double *col = st + nrow * j;
        for (int i = 0; i < ne; ++i)
            for (int k = 0; k < nq; ++k) {
                double rx = cx[i] + lc[3 * k] * ex[i] / 2;
                double ry = cy[i] + lc[3 * k + 1] * ey[i] / 2;
                double rz = cz[i] + lc[3 * k + 2] * ez[i] / 2;
                memset(u, 0, sizeof(u));
                for (int p = -nrept; p <= nrept; ++p) {
                    double jump = p * lrept;
                    oq_ref_dc3d(alpha, rx, ry, rz, dep, dip, ax0[q1] + jump, ax1[q1] + jump,
                                axi0[q2], axi1[q2], ud1, ud2, 0.0, c);
                    for (int t = 0; t < 12; ++t) u[t] += c[t];
                }
                double lekk = lam * (u[3] + u[7] + u[11]);
                col[i]          += w[k] * (lekk + 2 * mu * u[3]);
                col[i + ne]     += w[k] * (mu * (u[4] + u[6]));
                col[i + 2 * ne] += w[k] * (mu * (u[5] + u[9]));
                col[i + 3 * ne] += w[k] * (lekk + 2 * mu * u[7]);
                col[i + 4 * ne] += w[k] * (mu * (u[8] + u[10]));
                col[i + 5 * ne] += w[k] * (lekk + 2 * mu * u[11]);
            }
    }
}

/*
 * Mantle -> fault, st[nx*nxi, 6*ne] col-major, src/BEM/GF.jl:194-227.  The reference evaluates the
 * volume kernel once per unit strain component p (outer loop over p); restated identically.
 */
void oq_ref_gf_mantle_fault(int ne, const double *qx, const double *qy, const double *qz,
                            const double *ex, const double *ey, const double *ez,
                            int nx, int nxi, const double *x, const double *y, const double *z,
                            double dip, double lam, double mu, int ftype, double *st)
{
    double nu = lam / 2 / (lam + mu);
    size_t ndisl = (size_t)nx * nxi;
    for (int p = 0; p < 6; ++p) {
        double eps[6] = {0, 0, 0, 0, 0, 0};
        eps[p] = 1.0;
#pragma omp parallel for schedule(dynamic, 4)
        for (int j = 0; j < ne; ++j) {
            double sig[6];
            size_t jcol = (size_t)p * ne + j;
            for (size_t i = 0; i < ndisl; ++i) {
                int q1 = (int)(i % nx), q2 = (int)(i / nx);
                oq_ref_stress_vol_hex8(x[q1], y[q2], z[q2], qx[j], qy[j], qz[j],
                                       ex[j], ey[j], ez[j], eps, mu, nu, sig);
                st[i + ndisl * jcol] = shear_traction_vol(ftype, sig, dip);
            }
        }
    }
}

/* Mantle -> mantle, st[6*ne, 6*ne] col-major, src/BEM/GF.jl:250-290 (eigvals diagnostic excluded). */
void oq_ref_gf_mantle_mantle(int ne, const double *cx, const double *cy, const double *cz,
                             const double *qx, const double *qy, const double *qz,
                             const double *ex, const double *ey, const double *ez,
                             int nq, const double *lc, const double *w,
                             double lam, double mu, double *st)
{
    double nu = lam / 2 / (lam + mu);
    size_t n6 = 6 * (size_t)ne;
    memset(st, 0, sizeof(double) * n6 * n6);
    for (int p = 0; p < 6; ++p) {
        double eps[6] = {0, 0, 0, 0, 0, 0};
        eps[p] = 1.0;
#pragma omp parallel for schedule(dynamic, 4)
        for (int i = 0; i < ne; ++i) {          /* source */
            double sig[6];
            size_t icol = (size_t)p * ne + i;
            for (int j = 0; j < ne; ++j)        /* receiver */
                for (int k = 0; k < nq; ++k) {
                    double rx = cx[j] + lc[3 * k] * ex[j] / 2;
                    double ry = cy[j] + lc[3 * k + 1] * ey[j] / 2;
                    double rz = cz[j] + lc[3 * k + 2] * ez[j] / 2;
                    oq_ref_stress_vol_hex8(rx, ry, rz, qx[i], qy[i], qz[i],
                                           ex[i], ey[i], ez[i], eps, mu, nu, sig);
                    for (int t = 0; t < 6; ++t)
                        st[(size_t)t * ne + j + n6 * icol] += sig[t] * w[k];
                }
        }
    }
}
