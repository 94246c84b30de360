// greens.cu -- Green's-function assembly kernels (hot path 1) and the OqMatrix handle.
//
// Replaces the four `stress_greens_function` methods of /root/reference/src/BEM/GF.jl (:31,:123,:194,:250).
// One thread per (receiver, source) pair; element geometry is staged in shared memory; the periodic
// image loop (GF.jl:47,154), the quadrature loop (GF.jl:146,270) and the stress/traction projections
// (GF.jl:76-96,163-169) are fused into the kernels.  All device matrices are row-major so that the
// RHS matvec (rhs.cu) streams each row with coalesced 128-bit loads.
#include <algorithm>

#include "common.cuh"
#include "greens_okada.cuh"
#include "hex8_dev.cuh"
#include "greens_classes.cuh"

namespace oq {

// Dense expansion G[(i,j),(k,l)] = st[|i-k|, j, l] (test/BEM/tests.jl:46-49), rows [r0, r1).
__global__ void __launch_bounds__(256)
expand_toeplitz_kernel(const double* __restrict__ st, int nx, int nxi, int r0, int nrows, size_t ld,
                       double* __restrict__ G)
{
    const int nf = nx * nxi;
    for (int row = blockIdx.y; row < nrows; row += gridDim.y) {      // grid.y is capped at 65535
        const int f = r0 + row;
        const int i = f % nx, j = f / nx;
        for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < (int)ld; c += gridDim.x * blockDim.x) {
            double v = 0.0;
            if (c < nf) {
                const int k = c % nx, l = c / nx;
                const int dk = i > k ? i - k : k - i;
                v = st[dk + (size_t)nx * (j + (size_t)nxi * l)];
            }
            G[(size_t)row * ld + c] = v;
        }
    }
}

// Strike-wise DFT of the even extension [st; reverse(st[2:end])] of length N = 2nx-1 (GF.jl:60-68).
// The sequence is real and even, so the transform is the real cosine sum; one thread per output.
__global__ void __launch_bounds__(256)
toeplitz_dft_kernel(const double* __restrict__ st, int nx, int npairs, double* __restrict__ out_complex)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)nx * npairs) return;
    const int k = (int)(t % nx);
    const size_t pair = t / nx;
    const double* s = st + pair * nx;
    const int N = 2 * nx - 1;
    double acc = 0.0;
    for (int m = nx - 1; m >= 1; --m) {
        // cos(2π k m / N) with exact integer argument reduction
        const int km = (int)(((long long)k * m) % N);
        acc += s[m] * cospi(2.0 * (double)km / (double)N);
    }
    out_complex[2 * t] = s[0] + 2.0 * acc;
    out_complex[2 * t + 1] = 0.0;
}

// ---- K3: mantle -> fault (GF.jl:194-227) -----------------------------------------------------------
// One geometry evaluation serves all six unit strains (the reference re-evaluates six times).
// thread t -> (source element e fastest, receiver fault cell); writes G[fl, p*ne + e].
template <int SLIP>
__global__ void __launch_bounds__(kHex8Threads, OQ_HEX8_MINB)
gf_mantle_fault_kernel(Hex8Geom a, FaultGeom f, double mu, double nu, int slip, double s1, double c1,
                       double s2, double c2, int r0, int nrows, size_t ld, double* __restrict__ G)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)a.n * nrows) return;
    const int e = (int)(t % a.n);
    const int fl = (int)(t / a.n);
    const int fc = r0 + fl;
    const int q1 = fc % f.nx, q2 = fc / f.nx;
    extern __shared__ double hex8_acc[];
    double* row = G + (size_t)fl * ld + e;
    const size_t ne = a.n;
    // strike-slip traction needs only the xy,xz strain rows; dip-slip only yy,yz,zz (GF.jl:89-96)
    constexpr int kNeed = SLIP == kStrikeSlip ? 0x06 : 0x38;
    hex8_stress_emit<kNeed>(f.x[q1], f.y[q2], f.z[q2], a.qx[e], a.qy[e], a.qz[e], a.dx[e], a.dy[e], a.dz[e], mu, nu,
                     hex8_acc + threadIdx.x, [&](int pc, const double (&S)[6]) {
                         row[(size_t)pc * ne] = shear_traction_stress(slip, S, s1, c1, s2, c2);
                     });
}

// ---- K4: mantle -> mantle (GF.jl:250-290) ----------------------------------------------------------
// thread t -> (source element i fastest, receiver element j); 36 outputs G[(k*nel + jl), p*ne + i].
__global__ void __launch_bounds__(kHex8Threads, OQ_HEX8_MINB)
gf_mantle_mantle_kernel(Hex8Geom a, double mu, double nu, const double* __restrict__ qc,
                        const double* __restrict__ qw, int nq, int e_begin, int nel, size_t ld,
                        double* __restrict__ G)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)a.n * nel) return;
    const int i = (int)(t % a.n);
    const int jl = (int)(t / a.n);
    const int j = e_begin + jl;
    const double cx = a.cx[j], cy = a.cy[j], cz = a.cz[j];
    const double hx = a.dx[j] / 2, hy = a.dy[j] / 2, hz = a.dz[j] / 2;
    const double qx = a.qx[i], qy = a.qy[i], qz = a.qz[i], ex = a.dx[i], ey = a.dy[i], ez = a.dz[i];
    extern __shared__ double hex8_acc[];
    const size_t ne = a.n;
    for (int w = 0; w < nq; ++w) {
        const double rx = cx + qc[3 * w] * hx;
        const double ry = cy + qc[3 * w + 1] * hy;
        const double rz = cz + qc[3 * w + 2] * hz;
        const double wt = qw[w];
        // the thread owns its 36 entries: the first quadrature point stores, later ones accumulate in place
        hex8_stress_emit(rx, ry, rz, qx, qy, qz, ex, ey, ez, mu, nu, hex8_acc + threadIdx.x,
                         [&](int pc, const double (&S)[6]) {
#pragma unroll
                             for (int k = 0; k < 6; ++k) {
                                 double* dst = G + ((size_t)k * nel + jl) * ld + (size_t)pc * ne + i;
                                 *dst = (w == 0) ? S[k] * wt : *dst + S[k] * wt;
                             }
                         });
    }
}

// ---- K3'/K4': the hex8 kernels on TILES of source cells that share vertices -----------------------------------
// A tile is up to kTileC source cells and the up to kTileV distinct mesh vertices they touch (a 4x4x4 block of a
// conforming mesh: 64 cells, 125 vertices instead of 512 corners).  One CTA takes one tile and a run of receivers;
// per receiver (and quadrature point) thread v evaluates the basis + combination at vertex v ONCE (hex8_vertex_kernels,
// the transcendental-heavy part), the 36 strain kernels of every vertex go to shared memory, then thread c forms
// the signed 8-vertex sum of cell c and applies the stress / traction epilogue.  Same closed form, same entries
// (GF.jl:206-225, :262-290), ~4x fewer basis evaluations.
constexpr int kTileC = 64;
constexpr int kTileV = 128;       // == kHex8Threads: one vertex per thread
static_assert(kTileV == kHex8Threads, "one vertex per thread");

struct Hex8TileView {
    const double *vx, *vy, *vz;          // [ntiles][kTileV]
    const int* cell;                     // [ntiles][kTileC] global source cell, -1: unused slot
    const unsigned char* corner;         // [ntiles][kTileC][8] local vertex of corner c1 + 2 c2 + 4 c3
    const int* counts;                   // [ntiles][2]: cells, vertices
    const double* nudge;                 // [ntiles]
    int ntiles;
};

struct TileSmem {
    double vx[kTileV], vy[kTileV], vz[kTileV];
    int cell[kTileC];
    unsigned char corner[kTileC][8];
    int ncell, nvert;
    double nudge;
};

__device__ __forceinline__ void tile_load(const Hex8TileView& T, int tile, TileSmem& ts)
{
    const int tid = threadIdx.x;
    if (tid < kTileV) {
        ts.vx[tid] = T.vx[(size_t)tile * kTileV + tid];
        ts.vy[tid] = T.vy[(size_t)tile * kTileV + tid];
        ts.vz[tid] = T.vz[(size_t)tile * kTileV + tid];
    }
    if (tid < kTileC) {
        ts.cell[tid] = T.cell[(size_t)tile * kTileC + tid];
#pragma unroll
        for (int k = 0; k < 8; ++k) ts.corner[tid][k] = T.corner[((size_t)tile * kTileC + tid) * 8 + k];
    }
    if (tid == 0) { ts.ncell = T.counts[2 * tile]; ts.nvert = T.counts[2 * tile + 1]; ts.nudge = T.nudge[tile]; }
    __syncthreads();
}

// signed sum over the eight vertices of a cell of the per-vertex strain kernels in shared memory
template <int NEED>
__device__ __forceinline__ void tile_cell_kernels(const double* qv, const unsigned char (&cn)[8], double (&Q)[36])
{
#pragma unroll
    for (int m = 0; m < 36; ++m) {
        if (!((NEED >> (m / 6)) & 1)) { Q[m] = 0.0; continue; }
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const double v = qv[m * kHex8Threads + cn[k]];
            s += (((k & 1) + ((k >> 1) & 1) + (k >> 2)) & 1) ? v : -v;       // s1 s2 s3, s = -1 at the lower limit
        }
        Q[m] = s;
    }
}

template <int SLIP>
__global__ void __launch_bounds__(kHex8Threads, OQ_HEX8_MINB)
gf_mantle_fault_tile_kernel(Hex8TileView T, Hex8Geom a, FaultGeom f, double mu, double nu, int slip, double s1, double c1,
                            double s2, double c2, int r0, int nrows, int rows_per_cta, size_t ld, double* __restrict__ G)
{
    extern __shared__ double hex8_acc[];
    __shared__ TileSmem ts;
    tile_load(T, blockIdx.x, ts);
    constexpr int kNeed = SLIP == kStrikeSlip ? 0x06 : 0x38;     // GF.jl:89-96: xy,xz or yy,yz,zz strain rows
    const int tid = threadIdx.x;
    const double lam = 2.0 * mu * nu / (1.0 - 2.0 * nu);
    const double alpha = (lam + mu) / (lam + 2.0 * mu);
    const size_t ne = a.n;
    const int fl0 = blockIdx.y * rows_per_cta, fl1 = min(nrows, fl0 + rows_per_cta);
    for (int fl = fl0; fl < fl1; ++fl) {
        const int fc = r0 + fl;
        const int q1 = fc % f.nx, q2 = fc / f.nx;
        const double x = f.x[q1], y = f.y[q2], z = f.z[q2];
        if (tid < ts.nvert) {
            double Q[36];
            hex8_vertex_kernels<kNeed>(x, y, z, ts.vx[tid], ts.vy[tid], ts.vz[tid], ts.nudge, alpha, hex8_acc + tid, Q);
#pragma unroll
            for (int m = 0; m < 36; ++m)
                if ((kNeed >> (m / 6)) & 1) hex8_acc[m * kHex8Threads + tid] = Q[m];
        }
        __syncthreads();
        if (tid < ts.ncell) {
            const int e = ts.cell[tid];
            double Q[36];
            tile_cell_kernels<kNeed>(hex8_acc, ts.corner[tid], Q);
            const double qx = a.qx[e], qy = a.qy[e], qz = a.qz[e], dx = a.dx[e], dy = a.dy[e], dz = a.dz[e];
            const bool inside = x > qx - 0.5 * dx && x < qx + 0.5 * dx && y > qy && y < qy + dy && z > qz - dz && z < qz;
            double* row = G + (size_t)fl * ld + e;
            hex8_stress_from_kernels(Q, inside, mu, nu, [&](int pc, const double (&S)[6]) {
                row[(size_t)pc * ne] = shear_traction_stress(slip, S, s1, c1, s2, c2);
            });
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kHex8Threads, OQ_HEX8_MINB)
gf_mantle_mantle_tile_kernel(Hex8TileView T, Hex8Geom a, double mu, double nu, const double* __restrict__ qc,
                             const double* __restrict__ qw, int nq, int e_begin, int nel, int rows_per_cta, size_t ld,
                             double* __restrict__ G)
{
    extern __shared__ double hex8_acc[];
    __shared__ TileSmem ts;
    tile_load(T, blockIdx.x, ts);
    const int tid = threadIdx.x;
    const double lam = 2.0 * mu * nu / (1.0 - 2.0 * nu);
    const double alpha = (lam + mu) / (lam + 2.0 * mu);
    const size_t ne = a.n;
    const int jl0 = blockIdx.y * rows_per_cta, jl1 = min(nel, jl0 + rows_per_cta);
    for (int jl = jl0; jl < jl1; ++jl) {
        const int j = e_begin + jl;
        const double cx = a.cx[j], cy = a.cy[j], cz = a.cz[j];
        const double hx = a.dx[j] / 2, hy = a.dy[j] / 2, hz = a.dz[j] / 2;
        for (int w = 0; w < nq; ++w) {
            const double rx = cx + qc[3 * w] * hx;
            const double ry = cy + qc[3 * w + 1] * hy;
            const double rz = cz + qc[3 * w + 2] * hz;
            const double wt = qw[w];
            if (tid < ts.nvert) {
                double Q[36];
                hex8_vertex_kernels<0x3f>(rx, ry, rz, ts.vx[tid], ts.vy[tid], ts.vz[tid], ts.nudge, alpha, hex8_acc + tid, Q);
#pragma unroll
                for (int m = 0; m < 36; ++m) hex8_acc[m * kHex8Threads + tid] = Q[m];
            }
            __syncthreads();
            if (tid < ts.ncell) {
                const int i = ts.cell[tid];
                double Q[36];
                tile_cell_kernels<0x3f>(hex8_acc, ts.corner[tid], Q);
                const double qx = a.qx[i], qy = a.qy[i], qz = a.qz[i], ex = a.dx[i], ey = a.dy[i], ez = a.dz[i];
                const bool inside = rx > qx - 0.5 * ex && rx < qx + 0.5 * ex && ry > qy && ry < qy + ey && rz > qz - ez && rz < qz;
                // the thread owns its 36 entries: the first quadrature point stores, later ones accumulate in place
                hex8_stress_from_kernels(Q, inside, mu, nu, [&](int pc, const double (&S)[6]) {
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        double* dst = G + ((size_t)k * nel + jl) * ld + (size_t)pc * ne + i;
                        *dst = (w == 0) ? S[k] * wt : *dst + S[k] * wt;
                    }
                });
            }
            __syncthreads();
        }
    }
}

// ---- K3''/K4'': the hex8 kernels on CLASSES of pairs (greens_classes.cuh) --------------------------------------
// One thread per class (u1, u23): the per-pair code of K3/K4 on the coordinates of a representative receiver and
// source of the x group and of the (y,z) group.  T[(k*6 + p)][u23][u1]  (K4),  T[p][u23][u1]  (K3).
template <int SLIP>
__global__ void __launch_bounds__(kHex8Threads, OQ_HEX8_MINB)
gf_mantle_fault_class_kernel(Hex8Geom a, FaultGeom f, const int* __restrict__ rep_r1, const int* __restrict__ rep_s1,
                             const int* __restrict__ rep_r23, const int* __restrict__ rep_s23, int n1, int n23,
                             double mu, double nu, int slip, double s1, double c1, double s2, double c2,
                             double* __restrict__ T)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)n1 * n23) return;
    const int u1 = (int)(t % n1), u23 = (int)(t / n1);
    const int q1 = rep_r1[u1] % f.nx, q2 = rep_r23[u23] / f.nx;      // representatives are global cell indices
    const int i1 = rep_s1[u1], i23 = rep_s23[u23];
    extern __shared__ double hex8_acc[];
    constexpr int kNeed = SLIP == kStrikeSlip ? 0x06 : 0x38;
    double* dst = T + (size_t)u23 * n1 + u1;
    const size_t stride = (size_t)n1 * n23;
    hex8_stress_emit<kNeed>(f.x[q1], f.y[q2], f.z[q2], a.qx[i1], a.qy[i23], a.qz[i23], a.dx[i1], a.dy[i23], a.dz[i23], mu, nu,
                            hex8_acc + threadIdx.x, [&](int pc, const double (&S)[6]) {
                                dst[(size_t)pc * stride] = shear_traction_stress(slip, S, s1, c1, s2, c2);
                            });
}

__global__ void __launch_bounds__(kHex8Threads, OQ_HEX8_MINB)
gf_mantle_mantle_class_kernel(Hex8Geom a, const int* __restrict__ rep_r1, const int* __restrict__ rep_s1,
                              const int* __restrict__ rep_r23, const int* __restrict__ rep_s23, int n1, int n23,
                              double mu, double nu, const double* __restrict__ qc,
                              const double* __restrict__ qw, int nq, double* __restrict__ T)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)n1 * n23) return;
    const int u1 = (int)(t % n1), u23 = (int)(t / n1);
    const int j1 = rep_r1[u1], j23 = rep_r23[u23], i1 = rep_s1[u1], i23 = rep_s23[u23];   // global elements
    const double cx = a.cx[j1], cy = a.cy[j23], cz = a.cz[j23];
    const double hx = a.dx[j1] / 2, hy = a.dy[j23] / 2, hz = a.dz[j23] / 2;
    const double qx = a.qx[i1], qy = a.qy[i23], qz = a.qz[i23], ex = a.dx[i1], ey = a.dy[i23], ez = a.dz[i23];
    extern __shared__ double hex8_acc[];
    double* dst0 = T + (size_t)u23 * n1 + u1;
    const size_t stride = (size_t)n1 * n23;
    for (int w = 0; w < nq; ++w) {
        const double rx = cx + qc[3 * w] * hx;
        const double ry = cy + qc[3 * w + 1] * hy;
        const double rz = cz + qc[3 * w + 2] * hz;
        const double wt = qw[w];
        hex8_stress_emit(rx, ry, rz, qx, qy, qz, ex, ey, ez, mu, nu, hex8_acc + threadIdx.x,
                         [&](int pc, const double (&S)[6]) {
#pragma unroll
                             for (int k = 0; k < 6; ++k) {
                                 double* dst = dst0 + (size_t)(k * 6 + pc) * stride;
                                 *dst = (w == 0) ? S[k] * wt : *dst + S[k] * wt;
                             }
                         });
    }
}

// ---- batched direct evaluations (parity probes of the two closed forms) --------------------------
template <int SLIP>
__global__ void dc3d_gradient_kernel(int n, const double* x, const double* y, const double* z, OkadaMedium m,
                                     double dep, double al1, double al2, double aw1, double aw2, double* out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    double g[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) g[k] = 0.0;
    okada_gradient<SLIP>(m, x[t], y[t], z[t], dep, al1, al2, aw1, aw2, g);
#pragma unroll
    for (int k = 0; k < 9; ++k) out[(size_t)t * 9 + k] = g[k] * kInv2Pi;
}

__global__ void hex8_stress_kernel(int n, const double* x, const double* y, const double* z, double qx, double qy,
                                   double qz, double dx, double dy, double dz, const double* eps, double mu,
                                   double nu, double* out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    extern __shared__ double hex8_acc[];
    double v[6] = {0, 0, 0, 0, 0, 0};
    hex8_stress_emit(x[t], y[t], z[t], qx, qy, qz, dx, dy, dz, mu, nu, hex8_acc + threadIdx.x,
                     [&](int pc, const double (&S)[6]) {
#pragma unroll
                         for (int k = 0; k < 6; ++k) v[k] += eps[pc] * S[k];
                     });
#pragma unroll
    for (int k = 0; k < 6; ++k) out[(size_t)t * 6 + k] = v[k];
}

// row-major [rows x ld] -> column-major [rows x cols]
__global__ void __launch_bounds__(256)
rowmajor_to_colmajor_kernel(const double* __restrict__ in, int rows, int cols, size_t ld, double* __restrict__ out)
{
    __shared__ double tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        const int r = r0 + dy, c = c0 + threadIdx.x;
        tile[dy][threadIdx.x] = (r < rows && c < cols) ? in[(size_t)r * ld + c] : 0.0;
    }
    __syncthreads();
    for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        const int c = c0 + dy, r = r0 + threadIdx.x;
        if (r < rows && c < cols) out[(size_t)c * rows + r] = tile[threadIdx.x][dy];
    }
}

// column-major host matrix rows -> row-major shard (inverse of the above, with a row gather)
__global__ void __launch_bounds__(256)
colmajor_to_rowmajor_kernel(const double* __restrict__ in, int rows, int cols, size_t ld, double* __restrict__ out)
{
    __shared__ double tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        const int c = c0 + dy, r = r0 + threadIdx.x;
        tile[dy][threadIdx.x] = (r < rows && c < cols) ? in[(size_t)c * rows + r] : 0.0;
    }
    __syncthreads();
    for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        const int r = r0 + dy, c = c0 + threadIdx.x;
        if (r < rows && c < (int)ld) out[(size_t)r * ld + c] = (c < cols) ? tile[threadIdx.x][dy] : 0.0;
    }
}

// ---- host-side helpers -------------------------------------------------------------------------
// Okada kernels: the bit-reproducible published operation order by default; OQ_OKADA=fast selects the
// restructured, FMA-contracted form (~2x fewer fp64 instructions, equal to 1e-10 only relative to a row's scale)
static bool okada_strict_enabled()
{
    static const bool strict = [] { const char* e = getenv("OQ_OKADA"); return !(e && strcmp(e, "fast") == 0); }();
    return strict;
}

struct DevFaultMesh {
    DevBuf<double> buf;
    FaultGeom g{};
    int upload(const OqFaultMesh* mf)
    {
        OQ_CHECK(mf && mf->nx > 0 && mf->nxi > 0, "fault mesh is empty");
        OQ_CHECK(mf->x && mf->ax0 && mf->ax1 && mf->y && mf->z && mf->axi0 && mf->axi1, "fault mesh has NULL arrays");
        const size_t nx = mf->nx, nxi = mf->nxi;
        std::vector<double> h(3 * nx + 4 * nxi);
        double* q = h.data();
        auto put = [&](const double* src, size_t n) { memcpy(q, src, n * sizeof(double)); q += n; };
        put(mf->x, nx); put(mf->ax0, nx); put(mf->ax1, nx);
        put(mf->y, nxi); put(mf->z, nxi); put(mf->axi0, nxi); put(mf->axi1, nxi);
        OQ_TRY(buf.upload(h.data(), h.size()));
        g.x = buf.p; g.ax0 = g.x + nx; g.ax1 = g.ax0 + nx;
        g.y = g.ax1 + nx; g.z = g.y + nxi; g.axi0 = g.z + nxi; g.axi1 = g.axi0 + nxi;
        g.nx = mf->nx; g.nxi = mf->nxi; g.dep = mf->dep;
        return 0;
    }
};

struct DevHex8Mesh {
    DevBuf<double> buf;
    Hex8Geom g{};
    int upload(const OqHex8Mesh* ma)
    {
        OQ_CHECK(ma && ma->n > 0, "hex8 mesh is empty");
        OQ_CHECK(ma->cx && ma->cy && ma->cz && ma->qx && ma->qy && ma->qz && ma->dx && ma->dy && ma->dz,
                 "hex8 mesh has NULL arrays");
        const size_t n = ma->n;
        std::vector<double> h(9 * n);
        const double* src[9] = {ma->cx, ma->cy, ma->cz, ma->qx, ma->qy, ma->qz, ma->dx, ma->dy, ma->dz};
        for (int k = 0; k < 9; ++k) memcpy(h.data() + k * n, src[k], n * sizeof(double));
        OQ_TRY(buf.upload(h.data(), h.size()));
        const double* p = buf.p;
        g.cx = p; g.cy = p + n; g.cz = p + 2 * n; g.qx = p + 3 * n; g.qy = p + 4 * n; g.qz = p + 5 * n;
        g.dx = p + 6 * n; g.dy = p + 7 * n; g.dz = p + 8 * n; g.n = ma->n;
        return 0;
    }
};

struct DevQuad {
    DevBuf<double> c, w;
    int nq = 0;
    int upload(const OqQuadrature* q)
    {
        static const double c1[3] = {0, 0, 0}, w1[1] = {1.0};   // "Gauss1", GF.jl:103
        if (!q) { nq = 1; OQ_TRY(c.upload(c1, 3)); return w.upload(w1, 1); }
        // GF.jl:326: @assert length(qtype[1]) == 3 * length(qtype[2]) "Wrong format of quadrature!"
        OQ_CHECK(q->nq > 0 && q->coords && q->weights, "Wrong format of quadrature!");
        nq = q->nq;
        OQ_TRY(c.upload(q->coords, 3 * (size_t)nq));
        return w.upload(q->weights, nq);
    }
};

// Cuts a hex8 source mesh into tiles of cells that share vertices (host side of K3'/K4').  Corner coordinates are
// formed exactly as the pair kernels form them (x0 = qx - dx/2, x0 + dx; qy, qy + dy; qz - dz, qz); coordinates of
// neighbouring cells that agree to 1e-12 of the mesh extent are one vertex plane (conforming meshes built from
// centroids and sizes differ by an ulp or two); anything less regular simply shares fewer vertices.
struct DevHex8Tiles {
    DevBuf<double> vx, vy, vz, nudge;
    DevBuf<int> cell, counts;
    DevBuf<unsigned char> corner;
    Hex8TileView v{};
    double corners_per_vertex = 0.0;     // 8 * cells / vertices over all tiles (sharing factor, for the record)

    static void planes(const std::vector<double>& vals, std::vector<double>& uniq, std::vector<int>& index)
    {
        const size_t n = vals.size();
        std::vector<size_t> ord(n);
        for (size_t i = 0; i < n; ++i) ord[i] = i;
        std::sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return vals[a] < vals[b]; });
        const double span = n ? vals[ord[n - 1]] - vals[ord[0]] : 0.0;
        const double tol = 1e-12 * (span > 0 ? span : 1.0);
        index.assign(n, 0);
        uniq.clear();
        for (size_t k = 0; k < n; ++k) {
            const double v = vals[ord[k]];
            if (uniq.empty() || v - uniq.back() > tol) uniq.push_back(v);
            index[ord[k]] = (int)uniq.size() - 1;
        }
    }

    int build(const OqHex8Mesh* ma)
    {
        const int n = ma->n;
        std::vector<double> xs(2 * (size_t)n), ys(2 * (size_t)n), zs(2 * (size_t)n);
        for (int i = 0; i < n; ++i) {
            const double x0 = ma->qx[i] - 0.5 * ma->dx[i];
            xs[2 * i] = x0; xs[2 * i + 1] = x0 + ma->dx[i];
            ys[2 * i] = ma->qy[i]; ys[2 * i + 1] = ma->qy[i] + ma->dy[i];
            zs[2 * i] = ma->qz[i] - ma->dz[i]; zs[2 * i + 1] = ma->qz[i];
        }
        std::vector<double> ux, uy, uz;
        std::vector<int> ix, iy, iz;
        planes(xs, ux, ix); planes(ys, uy, iy); planes(zs, uz, iz);
        // group cells by 4x4x4 blocks of plane indices
        struct Key { int kx, ky, kz, cell; };
        std::vector<Key> keys(n);
        for (int i = 0; i < n; ++i) keys[i] = {ix[2 * i] / 4, iy[2 * i] / 4, iz[2 * i] / 4, i};
        std::sort(keys.begin(), keys.end(), [](const Key& a, const Key& b) {
            if (a.kz != b.kz) return a.kz < b.kz;
            if (a.ky != b.ky) return a.ky < b.ky;
            if (a.kx != b.kx) return a.kx < b.kx;
            return a.cell < b.cell;
        });
        std::vector<double> hvx, hvy, hvz, hnudge;
        std::vector<int> hcell, hcounts;
        std::vector<unsigned char> hcorner;
        size_t total_vertices = 0;
        auto flush = [&](const std::vector<int>& cells) {
            // one tile; vertices in order of first use
            std::vector<long long> vkey;
            std::vector<int> vids;
            const size_t t = hcounts.size() / 2;
            hvx.resize((t + 1) * kTileV, 0.0); hvy.resize((t + 1) * kTileV, 0.0); hvz.resize((t + 1) * kTileV, 0.0);
            hcell.resize((t + 1) * kTileC, -1); hcorner.resize((t + 1) * kTileC * 8, 0);
            double nudge = 1e300;
            int nv = 0;
            for (size_t c = 0; c < cells.size(); ++c) {
                const int i = cells[c];
                hcell[t * kTileC + c] = i;
                nudge = std::min(nudge, 1e-6 * std::min(ma->dx[i], std::min(ma->dy[i], ma->dz[i])));
                for (int k = 0; k < 8; ++k) {
                    const int jx = ix[2 * i + (k & 1)], jy = iy[2 * i + ((k >> 1) & 1)], jz = iz[2 * i + (k >> 2)];
                    const long long key = ((long long)jz * (long long)uy.size() + jy) * (long long)ux.size() + jx;
                    int found = -1;
                    for (int q = 0; q < nv; ++q) if (vkey[q] == key) { found = q; break; }
                    if (found < 0) {
                        found = nv++;
                        vkey.push_back(key);
                        hvx[t * kTileV + found] = ux[jx]; hvy[t * kTileV + found] = uy[jy]; hvz[t * kTileV + found] = uz[jz];
                    }
                    hcorner[(t * kTileC + c) * 8 + k] = (unsigned char)found;
                }
            }
            hcounts.push_back((int)cells.size()); hcounts.push_back(nv);
            hnudge.push_back(nudge);
            total_vertices += nv;
        };
        auto vkey_of = [&](int i, int k) {
            return ((long long)iz[2 * i + (k >> 2)] * (long long)uy.size() + iy[2 * i + ((k >> 1) & 1)]) * (long long)ux.size() +
                   ix[2 * i + (k & 1)];
        };
        // cells of one 4x4x4 block form a tile; a block of an irregular mesh may hold more cells / vertices than a
        // tile takes: cut it greedily (vertex set kept incrementally)
        std::vector<int> part;
        std::vector<long long> pkeys;
        for (int k = 0; k <= n; ++k) {
            const bool brk = k == n || (k > 0 && (keys[k].kx != keys[k - 1].kx || keys[k].ky != keys[k - 1].ky ||
                                                  keys[k].kz != keys[k - 1].kz));
            if (brk && !part.empty()) { flush(part); part.clear(); pkeys.clear(); }
            if (k == n) break;
            const int c = keys[k].cell;
            int fresh = 0;
            long long add[8];
            for (int q = 0; q < 8; ++q) {
                const long long key = vkey_of(c, q);
                if (std::find(pkeys.begin(), pkeys.end(), key) == pkeys.end() && std::find(add, add + fresh, key) == add + fresh)
                    add[fresh++] = key;
            }
            if ((int)part.size() + 1 > kTileC || (int)pkeys.size() + fresh > kTileV) {
                flush(part); part.clear(); pkeys.clear();
                fresh = 0;
                for (int q = 0; q < 8; ++q) {
                    const long long key = vkey_of(c, q);
                    if (std::find(add, add + fresh, key) == add + fresh) add[fresh++] = key;
                }
            }
            part.push_back(c);
            pkeys.insert(pkeys.end(), add, add + fresh);
        }
        const int ntiles = (int)(hcounts.size() / 2);
        corners_per_vertex = total_vertices ? 8.0 * n / (double)total_vertices : 0.0;
        OQ_TRY(vx.upload(hvx.data(), hvx.size())); OQ_TRY(vy.upload(hvy.data(), hvy.size())); OQ_TRY(vz.upload(hvz.data(), hvz.size()));
        OQ_TRY(nudge.upload(hnudge.data(), hnudge.size()));
        OQ_TRY(cell.upload(hcell.data(), hcell.size())); OQ_TRY(counts.upload(hcounts.data(), hcounts.size()));
        OQ_TRY(corner.upload(hcorner.data(), hcorner.size()));
        v.vx = vx.p; v.vy = vy.p; v.vz = vz.p; v.cell = cell.p; v.corner = corner.p; v.counts = counts.p; v.nudge = nudge.p;
        v.ntiles = ntiles;
        return 0;
    }
};

// hex8 builders: by default the class tables (greens_classes.cuh) when the mesh has at least 4 pairs per class, else
// the tiles with shared vertices; OQ_HEX8 = pair | tile | classes forces one path (validation twins)
enum { kHex8Auto = -1, kHex8Pair = 0, kHex8Tile = 1, kHex8Classes = 2 };
static int hex8_mode()
{
    const char* e = getenv("OQ_HEX8");           // read on every call: tests switch between the twins
    if (!e || !*e) return kHex8Auto;
    if (strcmp(e, "pair") == 0) return kHex8Pair;
    if (strcmp(e, "tile") == 0) return kHex8Tile;
    if (strcmp(e, "classes") == 0) return kHex8Classes;
    return kHex8Auto;
}

// class path worthwhile?  (forced: always when the classes could be built)
static bool hex8_use_classes(int mode, bool built, const Hex8PairClasses& pc, int outputs_per_pair)
{
    if (!built || (mode != kHex8Auto && mode != kHex8Classes)) return false;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return false;
    const double table_bytes = (double)pc.classes * outputs_per_pair * sizeof(double);
    if (table_bytes > 0.5 * (double)free_b) return false;
    return mode == kHex8Classes || pc.worthwhile;
}

static int make_okada_params(const OqFaultMesh* mf, double lam, double mu, int ftype, int nrept,
                             double buffer_ratio, OkadaParams* p)
{
    OQ_CHECK(buffer_ratio >= 0, "Argument `buffer_ratio` must be >= 0.");   // GF.jl:36,131
    OQ_CHECK(nrept >= 0, "nrept must be >= 0");
    OQ_CHECK(ftype == OQ_STRIKE_SLIP || ftype == OQ_DIP_SLIP, "unknown fault type %d", ftype);
    double sd, cd;
    sincosd(mf->dip, &sd, &cd);
    p->m = make_okada_medium((lam + mu) / (lam + 2 * mu), sd, cd);
    p->lam = lam; p->mu = mu;
    p->s1 = sd; p->c1 = cd;
    sincosd(2 * mf->dip, &p->s2, &p->c2);
    p->lrept = (buffer_ratio + 1.0) * (mf->dx * mf->nx);
    p->nrept = nrept;
    return 0;
}

static int alloc_matrix(OqMatrix* M, int row_kind, int row_begin, int row_end, int global_rows, int cols, bool dense = true)
{
    M->row_kind = row_kind; M->row_begin = row_begin; M->row_end = row_end;
    M->global_rows = global_rows; M->cols = cols;
    M->local_rows = (row_kind == OQ_ROWS_MANTLE ? 6 : 1) * (row_end - row_begin);
    M->ld = round_up((size_t)cols, 16);
    // A leading dimension that is a multiple of 8 KB makes every row of a chunk (and every CTA's stream) start
    // on the same HBM channel phase; one extra 128-byte line per row spreads them (OQ_LD_PAD=0 disables).
    static const bool pad = [] { const char* e = getenv("OQ_LD_PAD"); return !(e && e[0] == '0'); }();
    if (pad && M->ld % 1024 == 0) M->ld += 16;
    return dense ? M->d.alloc((size_t)M->local_rows * M->ld) : 0;        // class form: no dense storage
}

// Toeplitz kernel on the device: st[nx*nxi*nxi]
static int assemble_toeplitz(const OqFaultMesh* mf, double lam, double mu, int ftype, int nrept,
                             double buffer_ratio, DevBuf<double>& st, double* kernel_ms)
{
    OkadaParams p;
    DevFaultMesh dm;
    OQ_TRY(make_okada_params(mf, lam, mu, ftype, nrept, buffer_ratio, &p));
    OQ_TRY(dm.upload(mf));
    const size_t total = (size_t)mf->nx * mf->nxi * mf->nxi;
    OQ_TRY(st.alloc(total));
    const int threads = 128;
    const unsigned blocks = (unsigned)((total + threads - 1) / threads);
    const size_t smem = 4 * (size_t)mf->nxi * sizeof(double);
    OQ_CHECK(smem <= 48 * 1024, "nxi = %d too large for the shared-memory stage", mf->nxi);
    EventTimer tm;
    OQ_TRY(tm.start());
    if (okada_strict_enabled()) launch_fault_fault_strict(ftype, blocks, smem, dm.g, p, st.p);
    else if (ftype == OQ_STRIKE_SLIP) gf_fault_fault_kernel<kStrikeSlip, false><<<blocks, threads, smem>>>(dm.g, p, st.p);
    else gf_fault_fault_kernel<kDipSlip, false><<<blocks, threads, smem>>>(dm.g, p, st.p);
    OQ_LAUNCHED();
    OQ_TRY(tm.stop(kernel_ms));
    return 0;
}

// dense copy of a class-form shard (parity checks and host copies only; the RHS never expands)
static int dense_of_class_form(const OqMatrix* M, DevBuf<double>& dense)
{
    OQ_TRY(dense.alloc((size_t)M->local_rows * M->ld));
    OQ_TRY(expand_class_operand(*M->cls, M->ld, dense.p));
    OQ_CUDA(cudaDeviceSynchronize());
    return 0;
}

static int matrix_to_host_colmajor(const OqMatrix* M, double* out)
{
    DevBuf<double> dense;
    if (M->cls) OQ_TRY(dense_of_class_form(M, dense));
    const double* src = M->cls ? dense.p : M->d.p;
    DevBuf<double> cm;
    OQ_TRY(cm.alloc((size_t)M->local_rows * M->cols));
    dim3 grid((M->cols + 31) / 32, (M->local_rows + 31) / 32), block(32, 8);
    rowmajor_to_colmajor_kernel<<<grid, block>>>(src, M->local_rows, M->cols, M->ld, cm.p);
    OQ_LAUNCHED();
    OQ_CUDA(cudaMemcpy(out, cm.p, cm.n * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

}  // namespace oq

using namespace oq;

extern "C" {

int oq_gf_fault_fault(const OqFaultMesh* mf, double lambda, double mu, int ftype, int fourier, int nrept,
                      double buffer_ratio, double* out, double* kernel_ms)
{
    OQ_CHECK(mf && out, "NULL argument");
    OQ_TRY(enter());
    DevBuf<double> st;
    OQ_TRY(assemble_toeplitz(mf, lambda, mu, ftype, nrept, buffer_ratio, st, kernel_ms));
    if (!fourier) {
        OQ_CUDA(cudaMemcpy(out, st.p, st.n * sizeof(double), cudaMemcpyDeviceToHost));
        return 0;
    }
    DevBuf<double> dft;
    OQ_TRY(dft.alloc(2 * st.n));
    const int npairs = mf->nxi * mf->nxi;
    toeplitz_dft_kernel<<<(unsigned)((st.n + 255) / 256), 256>>>(st.p, mf->nx, npairs, dft.p);
    OQ_LAUNCHED();
    OQ_CUDA(cudaMemcpy(out, dft.p, dft.n * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

int oq_matrix_fault_fault(const OqFaultMesh* mf, double lambda, double mu, int ftype, int nrept,
                          double buffer_ratio, int row_begin, int row_end, OqMatrix** out)
{
    OQ_CHECK(mf && out, "NULL argument");
    OQ_TRY(enter());
    const int nf = mf->nx * mf->nxi;
    OQ_CHECK(0 <= row_begin && row_begin <= row_end && row_end <= nf, "row range [%d,%d) outside [0,%d)",
             row_begin, row_end, nf);
    DevBuf<double> st;
    double ms = 0;
    OQ_TRY(assemble_toeplitz(mf, lambda, mu, ftype, nrept, buffer_ratio, st, &ms));
    OqMatrix* M = new OqMatrix();
    if (alloc_matrix(M, OQ_ROWS_FAULT, row_begin, row_end, nf, nf)) { delete M; return 1; }
    M->kernel_ms = ms;
    if (M->local_rows > 0) {
        dim3 grid((unsigned)((M->ld + 255) / 256), M->local_rows > 65535 ? 65535 : M->local_rows);
        if (grid.x > 64) grid.x = 64;
        expand_toeplitz_kernel<<<grid, 256>>>(st.p, mf->nx, mf->nxi, row_begin, M->local_rows, M->ld, M->d.p);
        g_launches.fetch_add(1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { delete M; return fail("expand_toeplitz_kernel: %s", cudaGetErrorString(e)); }
    }
    *out = M;
    return 0;
}

int oq_matrix_from_toeplitz(const double* st_host, int nx, int nxi, int row_begin, int row_end, OqMatrix** out)
{
    OQ_CHECK(st_host && out && nx > 0 && nxi > 0, "bad argument");
    OQ_TRY(enter());
    const int nf = nx * nxi;
    OQ_CHECK(0 <= row_begin && row_begin <= row_end && row_end <= nf, "row range [%d,%d) outside [0,%d)",
             row_begin, row_end, nf);
    DevBuf<double> st;
    OQ_TRY(st.upload(st_host, (size_t)nx * nxi * nxi));
    OqMatrix* M = new OqMatrix();
    if (alloc_matrix(M, OQ_ROWS_FAULT, row_begin, row_end, nf, nf)) { delete M; return 1; }
    if (M->local_rows > 0) {
        dim3 grid((unsigned)((M->ld + 255) / 256), M->local_rows > 65535 ? 65535 : M->local_rows);
        if (grid.x > 64) grid.x = 64;
        expand_toeplitz_kernel<<<grid, 256>>>(st.p, nx, nxi, row_begin, M->local_rows, M->ld, M->d.p);
        g_launches.fetch_add(1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { delete M; return fail("expand_toeplitz_kernel: %s", cudaGetErrorString(e)); }
    }
    *out = M;
    return 0;
}

// x positions of receivers and sources on their common grid + the coarse grid (origin, step) of `coarse` for the
// sliding-window class kernel; false: no such grid (the operand keeps the general kernel)
static bool window_grid(const std::vector<double>& xr, const std::vector<double>& xs, int coarse_side /*0 both, 1 sources, 2 receivers*/,
                        std::vector<int>& pr, std::vector<int>& ps, long long* c0, long long* qstep)
{
    std::vector<double> all(xr);
    all.insert(all.end(), xs.begin(), xs.end());
    std::vector<int> pos;
    if (!grid_positions(all, pos)) return false;
    pr.assign(pos.begin(), pos.begin() + xr.size());
    ps.assign(pos.begin() + xr.size(), pos.end());
    return coarse_grid(coarse_side == 2 ? pr : coarse_side == 1 ? ps : pos, *c0, *qstep);
}

static int build_fault_mantle(const OqFaultMesh* mf, const OqHex8Mesh* ma, const OqQuadrature* quad, double lambda,
                              double mu, int ftype, int nrept, double buffer_ratio, int e_begin, int e_end,
                              OqMatrix** out, bool keep = false)
{
    OQ_CHECK(mf && ma && out, "NULL argument");
    OQ_TRY(enter());
    OQ_CHECK(0 <= e_begin && e_begin <= e_end && e_end <= ma->n, "element range [%d,%d) outside [0,%d)", e_begin,
             e_end, ma->n);
    OkadaParams p;
    DevFaultMesh dmf;
    DevHex8Mesh dma;
    DevQuad dq;
    OQ_TRY(make_okada_params(mf, lambda, mu, ftype, nrept, buffer_ratio, &p));
    OQ_TRY(dmf.upload(mf));
    OQ_TRY(dma.upload(ma));
    OQ_TRY(dq.upload(quad));
    const int nf = mf->nx * mf->nxi, nel = e_end - e_begin;
    OqMatrix* M = new OqMatrix();
    if (alloc_matrix(M, OQ_ROWS_MANTLE, e_begin, e_end, 6 * ma->n, nf, !keep)) { delete M; return 1; }
    if (nel > 0) {
        const size_t total = (size_t)nf * nel;
        M->pairs = (long long)total;
        // classes of pairs with bitwise equal dc3d arguments (greens_classes.cuh); OQ_FAULT_MANTLE = pair | classes forces
        const char* env = getenv("OQ_FAULT_MANTLE");
        const int mode = keep ? kHex8Classes : !env || !*env ? kHex8Auto : (strcmp(env, "pair") == 0 ? kHex8Pair : (strcmp(env, "classes") == 0 ? kHex8Classes : kHex8Auto));
        Hex8PairClasses pc;
        bool built = false;
        if (mode != kHex8Pair) {
            static const double c1[3] = {0, 0, 0};
            built = fault_mantle_classes(mf, ma, quad ? quad->coords : c1, dq.nq, nrept, p.lrept, e_begin, e_end, pc);
        }
        const bool classes = hex8_use_classes(mode, built, pc, 6) && (!keep || pc.worthwhile);
        if (keep && !classes) { delete M; return fail("fault -> mantle: the pairs of these meshes do not fall into translation classes (keep the dense form)"); }
        DevPairClasses dpc;
        DevBuf<double> table;
        if (!classes && M->d.zero()) { delete M; return 1; }
        if (classes && (dpc.upload(pc) || table.alloc((size_t)pc.classes * 6))) { delete M; return 1; }
        const bool strict = okada_strict_enabled();
        EventTimer tm;
        int rc = tm.start();
        if (!rc && classes) {
            OkadaClassLaunch cl{dq.c.p, dq.w.p, dq.nq, dpc.rep_r1.p, dpc.rep_s1.p, dpc.rep_r23.p, dpc.rep_s23.p, pc.g1.n, pc.g23.n, table.p};
            const unsigned nb = (unsigned)((pc.classes + 127) / 128);
            if (strict) launch_fault_mantle_class_strict(ftype, dmf.g, dma.g, p, cl);
            else if (ftype == OQ_STRIKE_SLIP)
                gf_fault_mantle_class_kernel<kStrikeSlip, false><<<nb, 128>>>(dmf.g, dma.g, p, cl.qc, cl.qw, cl.nq, cl.rep_r1, cl.rep_s1,
                                                                              cl.rep_r23, cl.rep_s23, cl.n1, cl.n23, cl.T);
            else
                gf_fault_mantle_class_kernel<kDipSlip, false><<<nb, 128>>>(dmf.g, dma.g, p, cl.qc, cl.qw, cl.nq, cl.rep_r1, cl.rep_s1,
                                                                           cl.rep_r23, cl.rep_s23, cl.n1, cl.n23, cl.T);
            g_launches.fetch_add(1);
            rc = tm.stop(&M->table_ms);
            if (!rc && keep) {
                M->cls.reset(new ClassOperand());
                // receivers = mantle cells (coarse grid along x), sources = fault cells (finer): sliding window per residue
                std::vector<double> xr(ma->cx, ma->cx + ma->n), xs(nf);
                for (int j = 0; j < nf; ++j) xs[j] = mf->x[j % mf->nx];
                std::vector<int> pr, ps;
                long long c0 = 0, qstep = 1;
                const bool grid = window_grid(xr, xs, 2, pr, ps, &c0, &qstep);
                rc = make_class_operand(pc, table, 6, 1, nel, nf, *M->cls, grid ? pr.data() + e_begin : nullptr, grid ? ps.data() : nullptr,
                                        2, c0, qstep);
            } else {
            if (!rc) rc = tm.start();
            if (!rc) {
                dim3 grid((unsigned)std::min<size_t>(((size_t)nf + 255) / 256, 64), (unsigned)std::min(nel, 65535));
                expand_classes_kernel<6, 1><<<grid, 256>>>(table.p, dpc.v, nel, nf, M->ld, M->d.p);
                g_launches.fetch_add(1);
                rc = tm.stop(&M->expand_ms);
            }
            }
            M->kernel_ms = M->table_ms + M->expand_ms;
            M->path = kHex8Classes; M->unique_pairs = pc.classes;
        } else if (!rc) {
            const unsigned blocks = (unsigned)((total + 127) / 128);
            if (strict)
                launch_fault_mantle_strict(ftype, blocks, dmf.g, dma.g, p, dq.c.p, dq.w.p, dq.nq, e_begin, nel, M->ld, M->d.p);
            else if (ftype == OQ_STRIKE_SLIP)
                gf_fault_mantle_kernel<kStrikeSlip, false><<<blocks, 128>>>(dmf.g, dma.g, p, dq.c.p, dq.w.p, dq.nq, e_begin,
                                                                             nel, M->ld, M->d.p);
            else
                gf_fault_mantle_kernel<kDipSlip, false><<<blocks, 128>>>(dmf.g, dma.g, p, dq.c.p, dq.w.p, dq.nq, e_begin,
                                                                          nel, M->ld, M->d.p);
            g_launches.fetch_add(1);
            rc = tm.stop(&M->kernel_ms);
            M->path = kHex8Pair; M->unique_pairs = M->pairs;
        }
        if (!rc && cudaGetLastError() != cudaSuccess) rc = fail("fault->mantle kernels failed to launch");
        if (rc) { delete M; return rc; }
    }
    *out = M;
    return 0;
}

int oq_matrix_fault_mantle(const OqFaultMesh* mf, const OqHex8Mesh* ma, const OqQuadrature* quad, double lambda,
                           double mu, int ftype, int nrept, double buffer_ratio, int e_begin, int e_end,
                           OqMatrix** out)
{
    return build_fault_mantle(mf, ma, quad, lambda, mu, ftype, nrept, buffer_ratio, e_begin, e_end, out);
}

int oq_matrix_fault_mantle_classes(const OqFaultMesh* mf, const OqHex8Mesh* ma, const OqQuadrature* quad, double lambda,
                                   double mu, int ftype, int nrept, double buffer_ratio, int e_begin, int e_end,
                                   OqMatrix** out)
{
    return build_fault_mantle(mf, ma, quad, lambda, mu, ftype, nrept, buffer_ratio, e_begin, e_end, out, true);
}

int oq_gf_fault_mantle(const OqFaultMesh* mf, const OqHex8Mesh* ma, const OqQuadrature* quad, double lambda,
                       double mu, int ftype, int nrept, double buffer_ratio, double* outp, double* kernel_ms)
{
    OQ_CHECK(outp && ma, "NULL argument");
    OqMatrix* M = nullptr;
    OQ_TRY(build_fault_mantle(mf, ma, quad, lambda, mu, ftype, nrept, buffer_ratio, 0, ma->n, &M));
    if (kernel_ms) *kernel_ms = M->kernel_ms;
    int rc = matrix_to_host_colmajor(M, outp);
    delete M;
    return rc;
}

static int hex8_smem_optin()
{
    static bool done = false;
    if (!done) {
        OQ_CUDA(cudaFuncSetAttribute(gf_mantle_fault_kernel<kStrikeSlip>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHex8SmemBytes));
        OQ_CUDA(cudaFuncSetAttribute(gf_mantle_fault_kernel<kDipSlip>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHex8SmemBytes));
        OQ_CUDA(cudaFuncSetAttribute(gf_mantle_mantle_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHex8SmemBytes));
        OQ_CUDA(cudaFuncSetAttribute(hex8_stress_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHex8SmemBytes));
        OQ_CUDA(cudaFuncSetAttribute(gf_mantle_fault_tile_kernel<kStrikeSlip>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHex8SmemBytes));
        OQ_CUDA(cudaFuncSetAttribute(gf_mantle_fault_tile_kernel<kDipSlip>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHex8SmemBytes));
        OQ_CUDA(cudaFuncSetAttribute(gf_mantle_mantle_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHex8SmemBytes));
        OQ_CUDA(cudaFuncSetAttribute(gf_mantle_fault_class_kernel<kStrikeSlip>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHex8SmemBytes));
        OQ_CUDA(cudaFuncSetAttribute(gf_mantle_fault_class_kernel<kDipSlip>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHex8SmemBytes));
        OQ_CUDA(cudaFuncSetAttribute(gf_mantle_mantle_class_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHex8SmemBytes));
        done = true;
    }
    return 0;
}

static int build_mantle_fault(const OqHex8Mesh* ma, const OqFaultMesh* mf, double lambda, double mu, int ftype,
                              int row_begin, int row_end, OqMatrix** out, bool keep = false)
{
    OQ_CHECK(mf && ma && out, "NULL argument");
    OQ_TRY(enter());
    const int nf = mf->nx * mf->nxi;
    OQ_CHECK(0 <= row_begin && row_begin <= row_end && row_end <= nf, "row range [%d,%d) outside [0,%d)",
             row_begin, row_end, nf);
    OQ_CHECK(ftype == OQ_STRIKE_SLIP || ftype == OQ_DIP_SLIP, "unknown fault type %d", ftype);
    DevFaultMesh dmf;
    DevHex8Mesh dma;
    OQ_TRY(hex8_smem_optin());
    OQ_TRY(dmf.upload(mf));
    OQ_TRY(dma.upload(ma));
    double s1, c1, s2, c2;
    sincosd(mf->dip, &s1, &c1);
    sincosd(2 * mf->dip, &s2, &c2);
    const double nu = lambda / 2 / (lambda + mu);   // GF.jl:203
    OqMatrix* M = new OqMatrix();
    if (alloc_matrix(M, OQ_ROWS_FAULT, row_begin, row_end, nf, 6 * ma->n, !keep)) { delete M; return 1; }
    if (M->local_rows > 0) {
        const size_t total = (size_t)ma->n * M->local_rows;
        const int mode = keep ? kHex8Classes : hex8_mode();
        M->pairs = (long long)total;
        // classes of (fault cell, hex8 cell) pairs: x group = (x_f - q_x, dx), (y,z) group = (y_f - q_y, dy, z_f, q_z, dz)
        Hex8PairClasses pc;
        bool built = false;
        if (mode == kHex8Auto || mode == kHex8Classes) built = mantle_fault_classes(ma, mf, row_begin, row_end, pc);
        const bool classes = hex8_use_classes(mode, built, pc, 6) && (!keep || pc.worthwhile);
        if (keep && !classes) { delete M; return fail("mantle -> fault: the pairs of these meshes do not fall into translation classes (keep the dense form)"); }
        DevHex8Tiles tiles;
        const bool tiled = !classes && mode != kHex8Pair;
        if (!classes && M->d.zero()) { delete M; return 1; }
        if (tiled && tiles.build(ma)) { delete M; return 1; }
        DevPairClasses dpc;
        DevBuf<double> table;
        if (classes && (dpc.upload(pc) || table.alloc((size_t)pc.classes * 6))) { delete M; return 1; }
        EventTimer tm;
        int rc = tm.start();
        if (!rc && classes) {
            const unsigned nb = (unsigned)((pc.classes + kHex8Threads - 1) / kHex8Threads);
            if (ftype == OQ_STRIKE_SLIP)
                gf_mantle_fault_class_kernel<kStrikeSlip><<<nb, kHex8Threads, kHex8SmemBytes>>>(
                    dma.g, dmf.g, dpc.rep_r1.p, dpc.rep_s1.p, dpc.rep_r23.p, dpc.rep_s23.p, pc.g1.n, pc.g23.n, mu, nu,
                    ftype, s1, c1, s2, c2, table.p);
            else
                gf_mantle_fault_class_kernel<kDipSlip><<<nb, kHex8Threads, kHex8SmemBytes>>>(
                    dma.g, dmf.g, dpc.rep_r1.p, dpc.rep_s1.p, dpc.rep_r23.p, dpc.rep_s23.p, pc.g1.n, pc.g23.n, mu, nu,
                    ftype, s1, c1, s2, c2, table.p);
            g_launches.fetch_add(1);
            rc = tm.stop(&M->table_ms);
            if (!rc && keep) {
                M->cls.reset(new ClassOperand());
                // receivers = fault cells (fine grid along x), sources = mantle cells (coarse): sliding window per residue
                std::vector<double> xr(nf), xs(ma->qx, ma->qx + ma->n);
                for (int j = 0; j < nf; ++j) xr[j] = mf->x[j % mf->nx];
                std::vector<int> pr, ps;
                long long c0 = 0, qstep = 1;
                const bool grid = window_grid(xr, xs, 1, pr, ps, &c0, &qstep);
                rc = make_class_operand(pc, table, 1, 6, M->local_rows, ma->n, *M->cls, grid ? pr.data() + row_begin : nullptr,
                                        grid ? ps.data() : nullptr, 1, c0, qstep);
            } else {
            if (!rc) rc = tm.start();
            if (!rc) {
                dim3 grid((unsigned)std::min<size_t>(((size_t)ma->n + 255) / 256, 64), (unsigned)std::min(M->local_rows, 65535));
                expand_classes_kernel<1, 6><<<grid, 256>>>(table.p, dpc.v, M->local_rows, ma->n, M->ld, M->d.p);
                g_launches.fetch_add(1);
                rc = tm.stop(&M->expand_ms);
            }
            }
            M->kernel_ms = M->table_ms + M->expand_ms;
            M->path = kHex8Classes; M->unique_pairs = pc.classes;
        } else if (!rc && tiled) {
            // receivers per CTA: enough CTAs to fill the GPU several times over, tile data reused across the run
            int rpc = 16;
            while (rpc > 1 && (long long)tiles.v.ntiles * ((M->local_rows + rpc - 1) / rpc) < 148LL * 2 * 8) rpc >>= 1;
            dim3 grid((unsigned)tiles.v.ntiles, (unsigned)((M->local_rows + rpc - 1) / rpc));
            if (ftype == OQ_STRIKE_SLIP)
                gf_mantle_fault_tile_kernel<kStrikeSlip><<<grid, kHex8Threads, kHex8SmemBytes>>>(
                    tiles.v, dma.g, dmf.g, mu, nu, ftype, s1, c1, s2, c2, row_begin, M->local_rows, rpc, M->ld, M->d.p);
            else
                gf_mantle_fault_tile_kernel<kDipSlip><<<grid, kHex8Threads, kHex8SmemBytes>>>(
                    tiles.v, dma.g, dmf.g, mu, nu, ftype, s1, c1, s2, c2, row_begin, M->local_rows, rpc, M->ld, M->d.p);
            g_launches.fetch_add(1);
            rc = tm.stop(&M->kernel_ms);
            M->path = kHex8Tile; M->unique_pairs = M->pairs;
        } else if (!rc) {
            const unsigned nb = (unsigned)((total + kHex8Threads - 1) / kHex8Threads);
            if (ftype == OQ_STRIKE_SLIP)
                gf_mantle_fault_kernel<kStrikeSlip><<<nb, kHex8Threads, kHex8SmemBytes>>>(
                    dma.g, dmf.g, mu, nu, ftype, s1, c1, s2, c2, row_begin, M->local_rows, M->ld, M->d.p);
            else
                gf_mantle_fault_kernel<kDipSlip><<<nb, kHex8Threads, kHex8SmemBytes>>>(
                    dma.g, dmf.g, mu, nu, ftype, s1, c1, s2, c2, row_begin, M->local_rows, M->ld, M->d.p);
            g_launches.fetch_add(1);
            rc = tm.stop(&M->kernel_ms);
            M->path = kHex8Pair; M->unique_pairs = M->pairs;
        }
        if (!rc && cudaGetLastError() != cudaSuccess) rc = fail("hex8 mantle->fault kernels failed to launch");
        if (rc) { delete M; return rc; }
    }
    *out = M;
    return 0;
}

int oq_matrix_mantle_fault(const OqHex8Mesh* ma, const OqFaultMesh* mf, double lambda, double mu, int ftype,
                           int row_begin, int row_end, OqMatrix** out)
{
    return build_mantle_fault(ma, mf, lambda, mu, ftype, row_begin, row_end, out);
}

int oq_matrix_mantle_fault_classes(const OqHex8Mesh* ma, const OqFaultMesh* mf, double lambda, double mu, int ftype,
                                   int row_begin, int row_end, OqMatrix** out)
{
    return build_mantle_fault(ma, mf, lambda, mu, ftype, row_begin, row_end, out, true);
}

int oq_gf_mantle_fault(const OqHex8Mesh* ma, const OqFaultMesh* mf, double lambda, double mu, int ftype,
                       double* outp, double* kernel_ms)
{
    OQ_CHECK(outp && mf, "NULL argument");
    OqMatrix* M = nullptr;
    OQ_TRY(build_mantle_fault(ma, mf, lambda, mu, ftype, 0, mf->nx * mf->nxi, &M));
    if (kernel_ms) *kernel_ms = M->kernel_ms;
    int rc = matrix_to_host_colmajor(M, outp);
    delete M;
    return rc;
}

static int build_mantle_mantle(const OqHex8Mesh* ma, const OqQuadrature* quad, double lambda, double mu,
                               int e_begin, int e_end, OqMatrix** out, bool keep = false)
{
    OQ_CHECK(ma && out, "NULL argument");
    OQ_TRY(enter());
    OQ_CHECK(0 <= e_begin && e_begin <= e_end && e_end <= ma->n, "element range [%d,%d) outside [0,%d)", e_begin,
             e_end, ma->n);
    DevHex8Mesh dma;
    DevQuad dq;
    OQ_TRY(hex8_smem_optin());
    OQ_TRY(dma.upload(ma));
    OQ_TRY(dq.upload(quad));
    const double nu = lambda / 2 / (lambda + mu);   // GF.jl:259
    const int nel = e_end - e_begin;
    OqMatrix* M = new OqMatrix();
    if (alloc_matrix(M, OQ_ROWS_MANTLE, e_begin, e_end, 6 * ma->n, 6 * ma->n, !keep)) { delete M; return 1; }
    if (nel > 0) {
        const size_t total = (size_t)ma->n * nel;
        const int mode = keep ? kHex8Classes : hex8_mode();
        M->pairs = (long long)total;
        // classes of (receiver cell, source cell) pairs: x group = (c_x - q_x, receiver dx, source dx),
        // (y,z) group = (c_y - q_y, both dy, receiver c_z and dz, source q_z and dz)
        Hex8PairClasses pc;
        bool built = false;
        if (mode == kHex8Auto || mode == kHex8Classes) built = mantle_mantle_classes(ma, e_begin, e_end, pc);
        const bool classes = hex8_use_classes(mode, built, pc, 36) && (!keep || pc.worthwhile);
        if (keep && !classes) { delete M; return fail("mantle -> mantle: the cell pairs of this mesh do not fall into translation classes worth a class form (fewer than 4 pairs per class, or a table that does not fit the free HBM): keep the dense form"); }
        DevHex8Tiles tiles;
        const bool tiled = !classes && mode != kHex8Pair;
        if (!classes && M->d.zero()) { delete M; return 1; }
        if (tiled && tiles.build(ma)) { delete M; return 1; }
        DevPairClasses dpc;
        DevBuf<double> table;
        if (classes && (dpc.upload(pc) || table.alloc((size_t)pc.classes * 36))) { delete M; return 1; }
        EventTimer tm;
        int rc = tm.start();
        if (!rc && classes) {
            const unsigned nb = (unsigned)((pc.classes + kHex8Threads - 1) / kHex8Threads);
            gf_mantle_mantle_class_kernel<<<nb, kHex8Threads, kHex8SmemBytes>>>(
                dma.g, dpc.rep_r1.p, dpc.rep_s1.p, dpc.rep_r23.p, dpc.rep_s23.p, pc.g1.n, pc.g23.n, mu, nu, dq.c.p,
                dq.w.p, dq.nq, table.p);
            g_launches.fetch_add(1);
            rc = tm.stop(&M->table_ms);
            if (!rc && keep) {
                M->cls.reset(new ClassOperand());
                // integer x positions of receivers (c_x) and sources (q_x) on their common grid: the sliding-window kernel
                std::vector<double> xr(ma->cx, ma->cx + ma->n), xs(ma->qx, ma->qx + ma->n);
                std::vector<int> pr, ps;
                long long c0 = 0, qstep = 1;
                const bool grid = window_grid(xr, xs, 0, pr, ps, &c0, &qstep);
                // walk order of the source (y,z) classes: by layer (q_z, dz, dy), then along y
                std::vector<int> rep(pc.g23.ns, -1), walk(pc.g23.ns);
                for (int e = 0; e < ma->n; ++e) if (rep[pc.g23.scls[e]] < 0) rep[pc.g23.scls[e]] = e;
                for (int b = 0; b < pc.g23.ns; ++b) walk[b] = b;
                std::stable_sort(walk.begin(), walk.end(), [&](int a, int b) {
                    const int ea = rep[a], eb = rep[b];
                    if (ma->qz[ea] != ma->qz[eb]) return ma->qz[ea] > ma->qz[eb];
                    if (ma->dz[ea] != ma->dz[eb]) return ma->dz[ea] < ma->dz[eb];
                    if (ma->dy[ea] != ma->dy[eb]) return ma->dy[ea] < ma->dy[eb];
                    return ma->qy[ea] < ma->qy[eb];
                });
                rc = make_class_operand(pc, table, 6, 6, nel, ma->n, *M->cls, grid ? pr.data() + e_begin : nullptr, grid ? ps.data() : nullptr,
                                        0, c0, qstep, &walk);
            } else {
            if (!rc) rc = tm.start();
            if (!rc) {
                dim3 grid((unsigned)std::min<size_t>(((size_t)ma->n + 255) / 256, 64), (unsigned)std::min(nel, 65535));
                expand_classes_kernel<6, 6><<<grid, 256>>>(table.p, dpc.v, nel, ma->n, M->ld, M->d.p);
                g_launches.fetch_add(1);
                rc = tm.stop(&M->expand_ms);
            }
            }
            M->kernel_ms = M->table_ms + M->expand_ms;
            M->path = kHex8Classes; M->unique_pairs = pc.classes;
        } else if (!rc && tiled) {
            int rpc = 16;
            while (rpc > 1 && (long long)tiles.v.ntiles * ((nel + rpc - 1) / rpc) < 148LL * 2 * 8) rpc >>= 1;
            dim3 grid((unsigned)tiles.v.ntiles, (unsigned)((nel + rpc - 1) / rpc));
            gf_mantle_mantle_tile_kernel<<<grid, kHex8Threads, kHex8SmemBytes>>>(tiles.v, dma.g, mu, nu, dq.c.p, dq.w.p, dq.nq,
                                                                                 e_begin, nel, rpc, M->ld, M->d.p);
            g_launches.fetch_add(1);
            rc = tm.stop(&M->kernel_ms);
            M->path = kHex8Tile; M->unique_pairs = M->pairs;
        } else if (!rc) {
            gf_mantle_mantle_kernel<<<(unsigned)((total + kHex8Threads - 1) / kHex8Threads), kHex8Threads, kHex8SmemBytes>>>(dma.g, mu, nu, dq.c.p, dq.w.p, dq.nq,
                                                                              e_begin, nel, M->ld, M->d.p);
            g_launches.fetch_add(1);
            rc = tm.stop(&M->kernel_ms);
            M->path = kHex8Pair; M->unique_pairs = M->pairs;
        }
        if (!rc && cudaGetLastError() != cudaSuccess) rc = fail("hex8 mantle->mantle kernels failed to launch");
        if (rc) { delete M; return rc; }
    }
    *out = M;
    return 0;
}

int oq_matrix_mantle_mantle(const OqHex8Mesh* ma, const OqQuadrature* quad, double lambda, double mu, int e_begin,
                            int e_end, OqMatrix** out)
{
    return build_mantle_mantle(ma, quad, lambda, mu, e_begin, e_end, out);
}

int oq_matrix_mantle_mantle_classes(const OqHex8Mesh* ma, const OqQuadrature* quad, double lambda, double mu,
                                    int e_begin, int e_end, OqMatrix** out)
{
    return build_mantle_mantle(ma, quad, lambda, mu, e_begin, e_end, out, true);
}

int oq_gf_mantle_mantle(const OqHex8Mesh* ma, const OqQuadrature* quad, double lambda, double mu, double* outp,
                        double* kernel_ms)
{
    OQ_CHECK(outp && ma, "NULL argument");
    OqMatrix* M = nullptr;
    OQ_TRY(build_mantle_mantle(ma, quad, lambda, mu, 0, ma->n, &M));
    if (kernel_ms) *kernel_ms = M->kernel_ms;
    int rc = matrix_to_host_colmajor(M, outp);
    delete M;
    return rc;
}

int oq_dc3d_gradient(int n, const double* x, const double* y, const double* z, double alpha, double dep, double dip,
                     double al1, double al2, double aw1, double aw2, int ftype, double* out9)
{
    OQ_CHECK(n >= 0 && (n == 0 || (x && y && z && out9)), "NULL argument");
    OQ_CHECK(ftype == OQ_STRIKE_SLIP || ftype == OQ_DIP_SLIP, "unknown fault type %d", ftype);
    if (n == 0) return 0;
    OQ_TRY(enter());
    DevBuf<double> dx, dy, dz, dout;
    OQ_TRY(dx.upload(x, n)); OQ_TRY(dy.upload(y, n)); OQ_TRY(dz.upload(z, n));
    OQ_TRY(dout.alloc((size_t)n * 9));
    double sd, cd;
    sincosd(dip, &sd, &cd);
    const OkadaMedium m = make_okada_medium(alpha, sd, cd);
    const int blocks = (n + 127) / 128;
    if (okada_strict_enabled())
        launch_dc3d_gradient_strict(ftype, n, dx.p, dy.p, dz.p, m, dep, al1, al2, aw1, aw2, dout.p);
    else if (ftype == OQ_STRIKE_SLIP)
        dc3d_gradient_kernel<kStrikeSlip><<<blocks, 128>>>(n, dx.p, dy.p, dz.p, m, dep, al1, al2, aw1, aw2, dout.p);
    else
        dc3d_gradient_kernel<kDipSlip><<<blocks, 128>>>(n, dx.p, dy.p, dz.p, m, dep, al1, al2, aw1, aw2, dout.p);
    OQ_LAUNCHED();
    OQ_CUDA(cudaMemcpy(out9, dout.p, dout.n * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

int oq_stress_vol_hex8(int n, const double* x, const double* y, const double* z, double qx, double qy, double qz,
                       double dx, double dy, double dz, const double* eps6, double mu, double nu, double* out6)
{
    OQ_CHECK(n >= 0 && eps6 && (n == 0 || (x && y && z && out6)), "NULL argument");
    if (n == 0) return 0;
    OQ_TRY(enter());
    OQ_TRY(hex8_smem_optin());
    DevBuf<double> bx, by, bz, be, dout;
    OQ_TRY(bx.upload(x, n)); OQ_TRY(by.upload(y, n)); OQ_TRY(bz.upload(z, n)); OQ_TRY(be.upload(eps6, 6));
    OQ_TRY(dout.alloc((size_t)n * 6));
    hex8_stress_kernel<<<(n + kHex8Threads - 1) / kHex8Threads, kHex8Threads, kHex8SmemBytes>>>(n, bx.p, by.p, bz.p, qx, qy, qz, dx, dy, dz, be.p, mu, nu, dout.p);
    OQ_LAUNCHED();
    OQ_CUDA(cudaMemcpy(out6, dout.p, dout.n * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

int oq_matrix_from_host(const double* a, int m, int n, int row_kind, int row_begin, int row_end, OqMatrix** out)
{
    OQ_CHECK(a && out && m > 0 && n > 0, "bad argument");
    OQ_TRY(enter());
    OQ_CHECK(row_kind == OQ_ROWS_FAULT || row_kind == OQ_ROWS_MANTLE, "bad row_kind");
    const int units = row_kind == OQ_ROWS_MANTLE ? m / 6 : m;
    OQ_CHECK(row_kind != OQ_ROWS_MANTLE || m % 6 == 0, "mantle matrices need 6*ne rows");
    OQ_CHECK(0 <= row_begin && row_begin <= row_end && row_end <= units, "row range [%d,%d) outside [0,%d)",
             row_begin, row_end, units);
    OqMatrix* M = new OqMatrix();
    if (alloc_matrix(M, row_kind, row_begin, row_end, m, n)) { delete M; return 1; }
    const int nel = row_end - row_begin;
    if (nel > 0) {
        // gather the shard's rows on the host into a compact column-major block, then transpose on device
        const int lr = M->local_rows;
        std::vector<double> h((size_t)lr * n);
        for (int c = 0; c < n; ++c)
            for (int r = 0; r < lr; ++r) {
                const int gr = row_kind == OQ_ROWS_MANTLE ? (r / nel) * units + row_begin + (r % nel) : row_begin + r;
                h[(size_t)c * lr + r] = a[(size_t)c * m + gr];
            }
        DevBuf<double> cm;
        int rc = cm.upload(h.data(), h.size());
        if (!rc) {
            dim3 grid((unsigned)((M->ld + 31) / 32), (lr + 31) / 32), block(32, 8);
            colmajor_to_rowmajor_kernel<<<grid, block>>>(cm.p, lr, n, M->ld, M->d.p);
            g_launches.fetch_add(1);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) rc = fail("colmajor_to_rowmajor_kernel: %s", cudaGetErrorString(e));
        }
        if (rc) { delete M; return rc; }
    }
    *out = M;
    return 0;
}

int oq_matrix_to_host(const OqMatrix* a, double* outp)
{
    OQ_CHECK(a && outp, "NULL argument");
    OQ_TRY(enter());
    if (a->local_rows == 0) return 0;
    return matrix_to_host_colmajor(a, outp);
}

int oq_matrix_rows_to_host(const OqMatrix* a, int local_begin, int local_end, double* outp)
{
    OQ_CHECK(a && outp, "NULL argument");
    OQ_CHECK(0 <= local_begin && local_begin <= local_end && local_end <= a->local_rows,
             "local row range [%d,%d) outside [0,%d)", local_begin, local_end, a->local_rows);
    OQ_TRY(enter());
    if (local_end == local_begin) return 0;
    DevBuf<double> dense;
    if (a->cls) OQ_TRY(dense_of_class_form(a, dense));
    const double* src = a->cls ? dense.p : a->d.p;
    OQ_CUDA(cudaMemcpy2D(outp, (size_t)a->cols * sizeof(double), src + (size_t)local_begin * a->ld,
                         a->ld * sizeof(double), (size_t)a->cols * sizeof(double), (size_t)(local_end - local_begin),
                         cudaMemcpyDeviceToHost));
    return 0;
}

int oq_matrix_shape(const OqMatrix* a, int* local_rows, int* cols, int* global_rows)
{
    OQ_CHECK(a, "NULL matrix");
    if (local_rows) *local_rows = a->local_rows;
    if (cols) *cols = a->cols;
    if (global_rows) *global_rows = a->global_rows;
    return 0;
}

int oq_matrix_form(const OqMatrix* a, int* form, double* device_bytes)
{
    OQ_CHECK(a, "NULL matrix");
    if (form) *form = a->cls ? 1 : 0;
    if (device_bytes) {
        if (a->cls) {
            const ClassOperand& c = *a->cls;
            *device_bytes = c.table_bytes + 8.0 * (double)(c.xg.n + c.dxg.n) +
                            4.0 * (double)(c.rc1.n + c.sc1.n + c.D1.n + c.D23.n + c.rc23.n + c.sc23.n + c.rg_items.n + c.sg_ptr.n +
                                           c.sg_order.n + c.xmap.n + c.csg.n + c.dxmap.n + c.dout_map.n + c.dcta_m0.n +
                                           c.cta_row.n + c.cta_begin.n + c.cta_count.n + c.dcta_row.n + c.dcta_begin.n + c.dcta_count.n);
        } else *device_bytes = 8.0 * (double)a->d.n;
    }
    return 0;
}

int oq_matrix_kernel_ms(const OqMatrix* a, double* ms)
{
    OQ_CHECK(a && ms, "NULL argument");
    *ms = a->kernel_ms;
    return 0;
}

int oq_hex8_pair_classes(const OqHex8Mesh* ma, const OqFaultMesh* mf, int begin, int end, int n, const int* recv,
                          const int* src, int* rep_recv_x, int* rep_src_x, int* rep_recv_yz, int* rep_src_yz,
                          long long* counts)
{
    OQ_CHECK(ma && ma->n > 0 && counts, "NULL argument");
    OQ_CHECK(n >= 0 && (n == 0 || (recv && src && rep_recv_x && rep_src_x && rep_recv_yz && rep_src_yz)), "NULL argument");
    const int limit = mf ? mf->nx * mf->nxi : ma->n;
    OQ_CHECK(0 <= begin && begin < end && end <= limit, "receiver range [%d,%d) outside [0,%d)", begin, end, limit);
    Hex8PairClasses pc;
    const bool built = mf ? mantle_fault_classes(ma, mf, begin, end, pc) : mantle_mantle_classes(ma, begin, end, pc);
    counts[0] = built ? pc.g1.n : 0; counts[1] = built ? pc.g23.n : 0;
    counts[2] = (long long)(end - begin) * ma->n;
    if (!built) return 0;
    for (int k = 0; k < n; ++k) {
        OQ_CHECK(begin <= recv[k] && recv[k] < end && 0 <= src[k] && src[k] < ma->n, "sample pair %d out of range", k);
        const int r = recv[k] - begin;
        const int c1 = pc.g1.D[(size_t)pc.g1.rcls[r] * pc.g1.ns + pc.g1.scls[src[k]]];
        const int c23 = pc.g23.D[(size_t)pc.g23.rcls[r] * pc.g23.ns + pc.g23.scls[src[k]]];
        rep_recv_x[k] = pc.g1.rep_r[c1]; rep_src_x[k] = pc.g1.rep_s[c1];
        rep_recv_yz[k] = pc.g23.rep_r[c23]; rep_src_yz[k] = pc.g23.rep_s[c23];
    }
    return 0;
}

int oq_class_form_plan(const OqHex8Mesh* ma, int e_begin, int e_end, long long* out8)
{
    OQ_CHECK(ma && ma->n > 0 && out8, "NULL argument");
    OQ_CHECK(0 <= e_begin && e_begin < e_end && e_end <= ma->n, "element range [%d,%d) outside [0,%d)", e_begin, e_end, ma->n);
    for (int i = 0; i < 8; ++i) out8[i] = 0;
    Hex8PairClasses pc;
    if (!mantle_mantle_classes(ma, e_begin, e_end, pc)) return 0;
    const int nel = e_end - e_begin;
    out8[0] = pc.g1.n; out8[1] = pc.g23.n; out8[2] = pc.worthwhile ? 1 : 0;
    std::vector<double> xr(ma->cx, ma->cx + ma->n), xs(ma->qx, ma->qx + ma->n);
    std::vector<int> pr, ps;
    long long c0 = 0, qstep = 1;
    OffsetPlan pl;
    const bool ok = window_grid(xr, xs, 0, pr, ps, &c0, &qstep) &&
                    plan_offsets(pc, 0, pr.data() + e_begin, ps.data(), nel, ma->n, c0, qstep, 6, kCdBlk, 0, pl);
    out8[3] = ok ? 1 : 0; out8[4] = ok ? (long long)pl.drow.size() : 0; out8[5] = ok ? pl.NS : 0;
    int max_sg = 0;
    std::vector<int> cnt(pc.g23.ns, 0);
    for (int s2 = 0; s2 < ma->n; ++s2) max_sg = std::max(max_sg, ++cnt[pc.g23.scls[s2]]);
    out8[6] = max_sg; out8[7] = pc.g23.nr;
    return 0;
}

// Host-only self check of the sliding-window plan of a fault <-> mantle operand (tests): which = 1 mantle -> fault
// (receivers = fault cells [begin, end)), 2 fault -> mantle (receivers = elements [begin, end)).  out6 = { plan found,
// residues Q, runs, (receiver, source) pairs checked, pairs whose class through the plan differs from the class maps,
// receivers not reachable through out_map }.
int oq_class_window_check(const OqHex8Mesh* ma, const OqFaultMesh* mf, int which, int begin, int end, long long* out6)
{
    OQ_CHECK(ma && mf && out6 && (which == 1 || which == 2), "bad argument");
    for (int i = 0; i < 6; ++i) out6[i] = 0;
    const int nf = mf->nx * mf->nxi, ne = ma->n;
    Hex8PairClasses pc;
    std::vector<double> xr, xs;
    int K = 1;
    if (which == 1) {
        OQ_CHECK(0 <= begin && begin < end && end <= nf, "row range");
        if (!mantle_fault_classes(ma, mf, begin, end, pc)) return 0;
        xr.resize(nf); for (int j = 0; j < nf; ++j) xr[j] = mf->x[j % mf->nx];
        xs.assign(ma->qx, ma->qx + ne);
    } else {
        OQ_CHECK(0 <= begin && begin < end && end <= ne, "element range");
        static const double c1[3] = {0, 0, 0};
        const double lrept = 2.0 * (mf->dx * mf->nx);
        if (!fault_mantle_classes(mf, ma, c1, 1, 2, lrept, begin, end, pc)) return 0;
        xr.assign(ma->cx, ma->cx + ne);
        xs.resize(nf); for (int j = 0; j < nf; ++j) xs[j] = mf->x[j % mf->nx];
        K = 6;
    }
    const int nr = end - begin, ns = which == 1 ? ne : nf;
    std::vector<int> pr, ps;
    long long c0 = 0, qstep = 1;
    if (!window_grid(xr, xs, which, pr, ps, &c0, &qstep)) return 0;
    OffsetPlan pl;
    if (!plan_offsets(pc, which, pr.data() + begin, ps.data(), nr, ns, c0, qstep, K, kCdBlk, 0, pl)) return 0;
    const int npad = pl.NS;
    if (!plan_offsets(pc, which, pr.data() + begin, ps.data(), nr, ns, c0, qstep, K, kCdBlk, npad, pl)) return 0;
    out6[0] = 1; out6[1] = pl.Q; out6[2] = (long long)pl.drow.size();
    const int nd = pl.MR + pl.NS - 1;
    std::vector<char> seen(nr, 0);
    long long checked = 0, bad = 0;
    for (size_t run = 0; run < pl.drow.size(); ++run)
        for (int m = 0; m < pl.dcnt[run]; ++m)
            for (int k = 0; k < 6; ++k) {
                const int o = pl.out_map[(size_t)(pl.dbeg[run] + m) * 6 + k];
                if (o < 0) continue;
                const int r = which == 1 ? o : o % nr;                     // K = 1: y index = r; K = 6: k*nr + r
                if (which == 2 && o / nr != k) { ++bad; continue; }
                seen[r] = 1;
                if (pc.g23.rcls[r] != pl.drow[run]) { ++bad; continue; }
                const int mp = pl.dm0[run] + m;
                // every source through xmap
                for (int g = 0; g < pc.g23.ns; ++g)
                    for (int j = 0; j < npad; ++j)
                        for (int c = 0; c < 6; ++c) {
                            const int xm = pl.xmap[((size_t)g * npad + j) * 6 + c];
                            if (xm < 0) continue;
                            const int s2 = which == 1 ? xm % ns : xm;       // P = 6: c*ns + s; P = 1: s
                            if (which == 1 && (xm / ns != c)) { ++bad; continue; }
                            if (which == 1 && c != 0) continue;            // one check per source
                            if (pc.g23.scls[s2] != g) { ++bad; continue; }
                            const int slot = which == 1 ? k : c;
                            const int want = pc.g1.D[(size_t)pc.g1.rcls[r] * pc.g1.ns + pc.g1.scls[s2]];
                            const int got = pl.dcls[(size_t)slot * nd + (mp - j + pl.NS - 1)];
                            ++checked;
                            if (got != want) ++bad;
                        }
                if (which == 2 && k > 0) break;                             // rows of one receiver share the pairs
            }
    out6[3] = checked; out6[4] = bad;
    for (int r = 0; r < nr; ++r) if (!seen[r]) ++out6[5];
    return 0;
}

int oq_matrix_assembly_info(const OqMatrix* a, OqAssemblyInfo* info)
{
    OQ_CHECK(a && info, "NULL argument");
    info->path = a->path; info->pairs = a->pairs; info->unique_pairs = a->unique_pairs;
    info->table_ms = a->table_ms; info->expand_ms = a->expand_ms; info->kernel_ms = a->kernel_ms;
    return 0;
}

int oq_matrix_destroy(OqMatrix* a)
{
    if (a) { enter(); delete a; }
    return 0;
}

}  // extern "C"
