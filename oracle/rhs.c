/*
 * oracle/rhs.c -- CPU restatement of the reference's ODE right-hand side (src/BEM/equation.jl).
 * TEST INFRASTRUCTURE ONLY (see oracle/okada.c header): the checker and the timed CPU baseline.
 *
 * Layouts are the reference's (Julia column-major): fault fields [nx, nxi] -> i + j*nx;
 * mantle fields [ne, 6] -> e + k*ne with k = xx,xy,xz,yy,yz,zz.
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

/* equation.jl:35-42  relvnp = v - vpl */
void oq_ref_relative_velocity(int n, const double *v, double vpl, double *relv)
{
#pragma omp parallel for
    for (int i = 0; i < n; ++i) relv[i] = v[i] - vpl;
}

/*
 * Dense-equivalent of equation.jl:44-61: E[i,j] = sum_{k,l} st[|i-k|, j, l] * relv[k, l]
 * (the contraction the reference's own test builds at test/BEM/tests.jl:46-58).
 */
void oq_ref_dtau_dt_toeplitz(int nx, int nxi, const double *st, const double *relv, double *out)
{
#pragma omp parallel for collapse(2)
    for (int j = 0; j < nxi; ++j)
        for (int i = 0; i < nx; ++i) {
            double acc = 0.0;
            for (int l = 0; l < nxi; ++l) {
                const double *g = st + (size_t)nx * (j + (size_t)nxi * l);
                const double *r = relv + (size_t)nx * l;
                for (int k = 0; k < nx; ++k) acc += g[abs(i - k)] * r[k];
            }
            out[i + (size_t)nx * j] = acc;
        }
}

/* y = A x (accumulate = 0) or y += A x (accumulate = 1); A is m x n column-major (equation.jl:201-203) */
void oq_ref_gemv(int m, int n, const double *A, const double *x, double *y, int accumulate)
{
#pragma omp parallel
    {
#pragma omp for
        for (int i = 0; i < m; ++i) if (!accumulate) y[i] = 0.0;
        /* row blocks per thread, column sweep inside: streams A once, unit stride */
#pragma omp for schedule(static)
        for (int ib = 0; ib < (m + 255) / 256; ++ib) {
            int i0 = ib * 256, i1 = i0 + 256 < m ? i0 + 256 : m;
            for (int j = 0; j < n; ++j) {
                const double *col = A + (size_t)m * j;
                double xj = x[j];
                for (int i = i0; i < i1; ++i) y[i] += col[i] * xj;
            }
        }
    }
}

/*
 * equation.jl:207-222 + :285-292.  nlaws = 1 for PowerLawViscosityProperty; >1 restates
 * CompositePowerLawViscosityProperty (sum over laws in order).  gamma, npow: [nlaws][ne];
 * npow holds "power - 1" as the reference stores it (property.jl:26).
 */
void oq_ref_update_strain_rate(int ne, int nlaws, const double *gamma, const double *npow,
                               const double *sigma, double *deps)
{
#pragma omp parallel for
    for (int i = 0; i < ne; ++i) {
        double skk = (sigma[i] + sigma[i + 3 * (size_t)ne] + sigma[i + 5 * (size_t)ne]) / 3;
        double sxx = sigma[i] - skk;
        double syy = sigma[i + 3 * (size_t)ne] - skk;
        double szz = sigma[i + 5 * (size_t)ne] - skk;
        double sxy = sigma[i + (size_t)ne], sxz = sigma[i + 2 * (size_t)ne], syz = sigma[i + 4 * (size_t)ne];
        double tn = sqrt(sxx * sxx + syy * syy + szz * szz + 2 * (sxy * sxy + sxz * sxz + syz * syz));
        double comp[6] = {sxx, sxy, sxz, syy, syz, szz};
        for (int k = 0; k < 6; ++k) {
            double ans = 0.0;
            for (int l = 0; l < nlaws; ++l)
                ans += gamma[(size_t)l * ne + i] * comp[k] * pow(tn, npow[(size_t)l * ne + i]);
            deps[i + (size_t)k * ne] = ans;
        }
    }
}

/* equation.jl:224-230 */
void oq_ref_relative_strain_rate(int ne, const double *deps, const double *deps0, double *rel)
{
#pragma omp parallel for
    for (int i = 0; i < ne; ++i)
        for (int k = 0; k < 6; ++k) rel[i + (size_t)k * ne] = deps[i + (size_t)k * ne] - deps0[k];
}

/* equation.jl:233-246 with dθ_dt of :279 (aging law) */
void oq_ref_update_fault(int n, const double *a, const double *b, const double *L, const double *sig,
                         double eta, double f0, double v0,
                         const double *dtau, const double *v, const double *theta,
                         double *dv, double *dtheta, double *ddelta)
{
#pragma omp parallel for
    for (int i = 0; i < n; ++i) {
        double psi1 = exp((f0 + b[i] * log(v0 * fmax(0.0, theta[i]) / L[i])) / a[i]) / (2 * v0);
        double psi2 = sig[i] * psi1 / hypot(1.0, v[i] * psi1);
        double dmu_dv = a[i] * psi2;
        double dmu_dth = b[i] / theta[i] * v[i] * psi2;
        dtheta[i] = 1 - v[i] * theta[i] / L[i];
        dv[i] = (dtau[i] - dmu_dth * dtheta[i]) / (dmu_dv + eta);
        ddelta[i] = v[i];
    }
}

/* equation.jl:248-276 with :279 and :282 (dilatancy variant) */
void oq_ref_update_fault_dilatancy(int n, const double *a, const double *b, const double *L, const double *sig,
                                   double f0, double v0,
                                   const double *tp, const double *epsd, const double *beta, const double *p0,
                                   const double *dtau, const double *v, const double *theta, const double *pr,
                                   double *dv, double *dtheta, double *ddelta, double *dpr)
{
#pragma omp parallel for
    for (int i = 0; i < n; ++i) {
        dtheta[i] = 1 - v[i] * theta[i] / L[i];
        dpr[i] = -(pr[i] - p0[i]) / tp[i] + epsd[i] / beta[i] / theta[i] * dtheta[i];
        double af = a[i] / f0, bf = b[i] / f0;
        double vf = fmax(0.0, v[i] / v0);
        double tf = fmax(0.0, theta[i] * v0 / L[i]);
        double vfa1 = pow(vf, af - 1), tfb1 = pow(tf, bf - 1);
        double vfa = pow(vf, af), tfb = pow(tf, bf);
        dv[i] = (dtau[i] + f0 * dpr[i] * vfa * tfb
                 - f0 * (sig[i] - pr[i]) * vfa * tfb1 * bf * v0 / L[i] * dtheta[i])
              / (f0 * (sig[i] - pr[i]) * vfa1 * tfb * af / v0);
        ddelta[i] = v[i];
    }
}

/*
 * equation.jl:46-54: dτ_dft[i,j] = Σ_l gf[i,j,l] * relv_dft[i,l] on interleaved complex arrays,
 * threads over j as the reference's @batch loop.  gf: [nx,nxi,nxi] complex, rd: [nx,nxi] complex.
 */
void oq_ref_fft_contract(int nx, int nxi, const double *gf, const double *rd, double *out)
{
#pragma omp parallel for schedule(static)
    for (int j = 0; j < nxi; ++j) {
        double *o = out + 2 * (size_t)nx * j;
        for (int i = 0; i < 2 * nx; ++i) o[i] = 0.0;
        for (int l = 0; l < nxi; ++l) {
            const double *g = gf + 2 * (size_t)nx * (j + (size_t)nxi * l);
            const double *r = rd + 2 * (size_t)nx * l;
            for (int i = 0; i < nx; ++i) {
                const double gr = g[2 * i], gi = g[2 * i + 1], rr = r[2 * i], ri = r[2 * i + 1];
                o[2 * i] += gr * rr - gi * ri;
                o[2 * i + 1] += gr * ri + gi * rr;
            }
        }
    }
}
