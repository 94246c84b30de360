// classmat.cuh -- K9: the RHS matvec of a Green's operand kept in CLASS FORM (included by rhs.cu).
//
// The three mantle operands of equation.jl:201-203 (`dτ += gf₂₁·reldϵ`, `dσ = gf₁₂·relv`, `dσ += gf₂₂·reldϵ`) are dense
// in the reference, and the dense form is what the north star streams from HBM.  On the meshes the package builds
// (Gmsh transfinite boxes, mesh.jl:95-130; the equidistant fault, mesh.jl:39-56) their (receiver, source) pairs fall
// into translation classes (greens_classes.cuh): G[(k, r), (p, s)] = T[class(r, s)][k][p] with a few thousand to a
// few million distinct 6x6 (6x1, 1x6) blocks instead of (6 N_e)² entries -- the same invariance the reference itself
// exploits for the fault (Toeplitz kernel, GF.jl:31-71), extended to the mantle.  These kernels multiply straight
// from the table:
//
//     y[k, r] = y_in[k, r] + Σ_s Σ_p T[D23[rc23(r), sc23(s)]][D1[rc1(r), sc1(s)]][k][p] · x[p, s]
//
// so a problem whose dense gf₂₂ would need 1.84 TB (BASELINE configs[3]: 80 000 cells) keeps a few GB of tables, and the
// evaluation is bound by the fp64 pipe / shared-memory bandwidth instead of HBM.  Nothing about the mesh is assumed
// beyond what the class maps say (they are found numerically); a mesh without structure has no class form and keeps
// the dense operands.
//
// Common structure.  One CTA = a run of receivers of one (y,z) class (they share the row D23[rc23, :]) x all sources.
// Sources are walked group by group ((y,z) class of the source) in ascending order of the pair's (y,z) class.  Per
// group the CTA needs the slab T[c23][:][:] (n1 classes x K*P doubles), the group's forcing values and the x classes
// of its sources: a producer warp fetches them with 1-D bulk TMA copies into a two-stage ring in shared memory
// (full/empty mbarriers), so the fetch of group i+1 overlaps the arithmetic on group i.  The forcing values are
// gathered once per evaluation into group order by class_gather_x_kernel (which is also the kernel that waits for the
// peers' slices of the forcing vector).  Sums are formed in a fixed order (sources ascending within a slice, slices
// folded in order; kCmParts CTAs per run each walk a quarter of the groups and class_fold_kernel adds the quarters in order):
// results are bitwise reproducible and independent of the number of ranks.
#pragma once

#include "tma.cuh"

namespace oq {

constexpr int kCmRb = 64;                    // receivers per CTA of the general kernel
constexpr int kCmSlices = 4;                 // source slices
constexpr int kCmConsumers = kCmRb * kCmSlices;      // 256 compute threads
constexpr int kCmThreads = kCmConsumers + 32;        // + the producer warp
constexpr int kCmStages = 2;
constexpr int kCmParts = 4;                  // CTAs per receiver run: each walks a quarter of the source groups (class_fold_kernel adds them up)
constexpr int kCmU = 4;                      // interleaved partial sums per output in the general kernel

struct ClassMvArgs {
    const double* Tm;             // general: [n23][n1][ts]; diagonal: [n23][noff][ts] (offset order, zero padded)
    int ts, n1, ns1, ns23;
    const int *rc1, *D1, *D23;
    const int *rg_items, *sg_ptr, *sg_order, *cta_row, *cta_begin, *cta_count;
    const double* xg;             // forcing values in group order [ns23][xstride][PX] (class_gather_x_kernel)
    const int* csg;               // x class of the sources in group order [ns23][xstride]
    int xstride;                  // slots per group (multiple of 8)
    int nr;                       // local receiver units
    const double* y_in;           // optional: accumulate onto (may alias y_out)
    double* y_out;                // [K][nr]
    const int* done;              // optional device flag: integration complete, skip
    double* part;                 // [kCmParts][K][nr] partial sums of the source-group quarters (blockIdx.y)
    // diagonal kernel
    const int *run_m0, *out_map;  // first coarse position of every run; flat y index of (run entry, row), -1: none
    int noff, L, nout;            // padded offsets per slab, source positions per slice, K * nr
    int d1_smem;                  // general kernel: the receivers' rows of D1 are staged in shared memory
};

// xg[slot][p] = x[p*ns + xmap[slot]] (0 where xmap < 0): the forcing vector in the order the CTAs consume it.  Also the
// one kernel of the class-form path that waits for the peers' slices (consumer_parity).
template <int P>
__global__ void __launch_bounds__(256)
class_gather_x_kernel(const int* __restrict__ xmap, size_t nslots, int ns, const double* x0, size_t x_stride, PeerWait pw,
                      const int* done, double* __restrict__ xg)
{
    constexpr int PX = (P + 1) & ~1;
    if (done && *reinterpret_cast<const volatile int*>(done)) return;
    const double* x = x0 + consumer_parity(pw) * x_stride;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < nslots; t += (size_t)gridDim.x * blockDim.x) {
        const int s = __ldg(xmap + t);
#pragma unroll
        for (int p = 0; p < PX; ++p) xg[t * PX + p] = (s >= 0 && p < P) ? x[(size_t)p * ns + s] : 0.0;
    }
}

// the same with a flat map: xg[t] = x[map[t]] (0 where map < 0) -- the sliding-window kernel's [group][position][column]
__global__ void __launch_bounds__(256)
class_gather_flat_kernel(const int* __restrict__ map, size_t n, const double* x0, size_t x_stride, PeerWait pw, const int* done,
                         double* __restrict__ xg)
{
    if (done && *reinterpret_cast<const volatile int*>(done)) return;
    const double* x = x0 + consumer_parity(pw) * x_stride;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
        const int m = __ldg(map + t);
        xg[t] = m >= 0 ? x[m] : 0.0;
    }
}

__device__ __forceinline__ void cm_bulk_load(void* dst, const void* src, unsigned bytes, uint64_t* bar)
{
    // pieces of <= 16 KB (every piece a multiple of 16 bytes)
    char* d = static_cast<char*>(dst);
    const char* s = static_cast<const char*>(src);
    while (bytes) {
        const unsigned n = bytes > 16384u ? 16384u : bytes;
        tma_load_1d(d, s, n, bar);
        d += n; s += n; bytes -= n;
    }
}

// ---- general kernel: every (receiver, source) pair looks its x class up in D1 ------------------------------
// The rows of D1 of the CTA's receivers sit in shared memory as 16-bit class ids (row stride = 2 mod 4 entries, i.e. an
// odd number of 32-bit words: conflict free) when they fit (d1_smem).
template <int K, int P, bool D1S>
__global__ void __launch_bounds__(kCmThreads)
class_matvec_kernel(const __grid_constant__ ClassMvArgs a)
{
    constexpr int KP = K * P;
    constexpr int PX = (P + 1) & ~1;
    static_assert(KP % 2 == 0, "class blocks are loaded as double2");
    extern __shared__ __align__(16) double cm_smem[];
    __shared__ __align__(8) uint64_t full_bar[kCmStages], empty_bar[kCmStages];
    if (a.done && *reinterpret_cast<const volatile int*>(a.done)) return;
    const unsigned slab_bytes = (unsigned)((size_t)a.n1 * a.ts * sizeof(double));
    const unsigned x_bytes = (unsigned)((size_t)a.xstride * PX * sizeof(double));
    const unsigned c_bytes = (unsigned)((size_t)a.xstride * sizeof(int));
    const size_t stage_doubles = ((size_t)slab_bytes + x_bytes + c_bytes) / sizeof(double);      // all multiples of 16 bytes
    const int row = a.cta_row[blockIdx.x], begin = a.cta_begin[blockIdx.x], count = a.cta_count[blockIdx.x];
    const int* d23row = a.D23 + (size_t)row * a.ns23;
    const int* order = a.sg_order + (size_t)row * a.ns23;
    const int it0 = (int)((long long)a.ns23 * blockIdx.y / kCmParts), it1 = (int)((long long)a.ns23 * (blockIdx.y + 1) / kCmParts);
    const int t = threadIdx.x;
    if (t == 0) {
        for (int s = 0; s < kCmStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], kCmConsumers); }
        fence_mbar_init();
    }
    __syncthreads();
    if (t >= kCmConsumers) {                                    // ---- producer warp: one lane fetches group after group
        if (t == kCmConsumers) {
            for (int it = it0; it < it1; ++it) {
                const int stage = (it - it0) % kCmStages;
                const int sg = order[it];
                const int c23 = d23row[sg];
                mbar_wait(&empty_bar[stage], (((unsigned)((it - it0) / kCmStages)) & 1u) ^ 1u);
                double* st = cm_smem + (size_t)stage * stage_doubles;
                mbar_arrive_expect_tx(&full_bar[stage], slab_bytes + x_bytes + c_bytes);
                cm_bulk_load(st, a.Tm + (size_t)c23 * a.n1 * a.ts, slab_bytes, &full_bar[stage]);
                cm_bulk_load(reinterpret_cast<char*>(st) + slab_bytes, a.xg + (size_t)sg * a.xstride * PX, x_bytes, &full_bar[stage]);
                cm_bulk_load(reinterpret_cast<char*>(st) + slab_bytes + x_bytes, a.csg + (size_t)sg * a.xstride, c_bytes, &full_bar[stage]);
            }
        }
        return;
    }
    // thread (slice, receiver)
    const int i = t % kCmRb, sl = t / kCmRb;
    const bool active = i < count;
    const int r = active ? a.rg_items[begin + i] : 0;
    const int* d1row = a.D1 + (size_t)(active ? a.rc1[r] : 0) * a.ns1;
    // the receivers' rows of D1 in shared memory (behind the stages and the fold area)
    unsigned short* d1s = reinterpret_cast<unsigned short*>(cm_smem + (size_t)kCmStages * stage_doubles + (size_t)kCmConsumers * K);
    const int ns1p = cm_d1_stride(a.ns1);
    if (D1S) {
        for (int q = t; q < count * a.ns1; q += kCmConsumers) {
            const int ii = q / a.ns1, b = q - ii * a.ns1;
            d1s[ii * ns1p + b] = (unsigned short)__ldg(a.D1 + (size_t)a.rc1[a.rg_items[begin + ii]] * a.ns1 + b);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kCmConsumers) : "memory");
    }
    const unsigned short* d1mine = d1s + i * ns1p;
    // kCmU interleaved partial sums per output (sources j, j + 4, j + 8, ... of a slice are in flight together: kCmU
    // times the independent work per thread; the association order stays a function of the source index alone)
    double acc[kCmU][K];
#pragma unroll
    for (int u = 0; u < kCmU; ++u)
#pragma unroll
        for (int k = 0; k < K; ++k) acc[u][k] = 0.0;
    auto lookup = [&](int b) -> int { return D1S ? (int)d1mine[b] : __ldg(d1row + b); };
    auto pair_fma = [&](const double* Ts, const double* xs, int c1, int j, double (&sum)[K]) {
        const double2* t2 = reinterpret_cast<const double2*>(Ts + (size_t)c1 * a.ts);
        double tv[KP];
#pragma unroll
        for (int q = 0; q < KP / 2; ++q) { const double2 v = t2[q]; tv[2 * q] = v.x; tv[2 * q + 1] = v.y; }
        const double* xv = xs + j * PX;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const double xp = xv[p];
#pragma unroll
            for (int k = 0; k < K; ++k) sum[k] = fma(tv[k * P + p], xp, sum[k]);
        }
    };
    for (int it = it0; it < it1; ++it) {
        const int stage = (it - it0) % kCmStages;
        const int sg = order[it];
        const int sn = a.sg_ptr[sg + 1] - a.sg_ptr[sg];
        const double* Ts = cm_smem + (size_t)stage * stage_doubles;
        const double* xs = Ts + (size_t)a.n1 * a.ts;
        const int* cs = reinterpret_cast<const int*>(xs + (size_t)a.xstride * PX);
        mbar_wait(&full_bar[stage], ((unsigned)((it - it0) / kCmStages)) & 1u);
        if (active) {
            int j = sl;
            for (; j + (kCmU - 1) * kCmSlices < sn; j += kCmU * kCmSlices) {
                int c1[kCmU];
#pragma unroll
                for (int u = 0; u < kCmU; ++u) c1[u] = lookup(cs[j + u * kCmSlices]);
#pragma unroll
                for (int u = 0; u < kCmU; ++u) pair_fma(Ts, xs, c1[u], j + u * kCmSlices, acc[u]);
            }
#pragma unroll
            for (int u = 0; u < kCmU - 1; ++u)
                if (j + u * kCmSlices < sn) pair_fma(Ts, xs, lookup(cs[j + u * kCmSlices]), j + u * kCmSlices, acc[u]);
        }
        mbar_arrive(&empty_bar[stage]);
    }
#pragma unroll
    for (int u = 1; u < kCmU; ++u)
#pragma unroll
        for (int k = 0; k < K; ++k) acc[0][k] += acc[u][k];
    // fold the slices in order (own region behind the stages)
    double* red = cm_smem + (size_t)kCmStages * stage_doubles;  // [slices][rb][K]
    if (sl > 0 && active) {
#pragma unroll
        for (int k = 0; k < K; ++k) red[((size_t)sl * kCmRb + i) * K + k] = acc[0][k];
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kCmConsumers) : "memory");
    if (sl == 0 && active) {
        for (int s = 1; s < kCmSlices; ++s) {
#pragma unroll
            for (int k = 0; k < K; ++k) acc[0][k] += red[((size_t)s * kCmRb + i) * K + k];
        }
#pragma unroll
        for (int k = 0; k < K; ++k) a.part[((size_t)blockIdx.y * K + k) * a.nr + r] = acc[0][k];
    }
}

// ---- diagonal fast path (6x6 operands on one equidistant x grid) ---------------------------------------------
// In class_matvec_kernel every FMA reads its own table entry from shared memory (128 B/clk per SM against 64 DFMA/clk:
// at most 25 % of the fp64 pipe).  When the x class of a pair depends on the DIFFERENCE of the integer x positions
// only (the Toeplitz structure of GF.jl:31-71), the pairs (r+1, s+1) and (r, s) share their 6x6 block: a thread that
// owns G consecutive receivers (one output component k) and walks the sources in position order keeps a sliding
// window of G block-rows (6 doubles each) in registers -- one new row (48 B) per G*6 FMAs instead of one per 6.
// The table is stored in OFFSET order (offset = receiver position - source position), zero padded on both sides, so
// the window of a CTA run starting at position p0 is ONE contiguous piece of the slab: padded offsets [p0, p0 + ndp).
//   thread (slice, blk, k): receivers p0 + blk*G + g (g < G), output k, source positions [slice*L, (slice+1)*L)
//   Td[dd] = block of offset (p0 + dd - (npad - 1)), dd = (blk*G + g) - j + npad - 1
// Shared-memory layout of a slab: class dd at dd*ts + 4*(dd >> 3) doubles -- every run of 8 classes starts 32 bytes
// later than a dense layout would put it, which makes the 128-bit loads of a warp (6 lanes of one receiver block read
// 288 contiguous bytes, the next block sits 8 classes = 19 x 128 bytes further) bank-conflict free.  One bulk copy per run.
// BLK receiver blocks of 8 per CTA: 8 (192 compute threads + the producer warp) or, when a shard has too few runs to
// fill the GPU, 4 (96 + 32) -- a receiver's sum does not depend on the run it sits in

__device__ __forceinline__ int cd_slab_off(int dd, int ts) { return dd * ts + 4 * (dd >> 3); }

template <int BLK>
__global__ void __launch_bounds__(BLK * 6 * kCdSlices + 32, BLK == 8 ? 2 : 3)
class_matvec_diag_kernel(const __grid_constant__ ClassMvArgs a)
{
    constexpr int G = kCdG;
    constexpr int kCdConsumers = BLK * 6 * kCdSlices;
    extern __shared__ __align__(16) double cm_smem[];
    __shared__ __align__(8) uint64_t full_bar[kCmStages], empty_bar[kCmStages];
    if (a.done && *reinterpret_cast<const volatile int*>(a.done)) return;
    const int npad = kCdSlices * a.L;                           // padded source positions (= xstride)
    const int ndp = BLK * G + npad;                             // padded diagonals of a run (multiple of 8)
    const int ts = a.ts;
    const unsigned run_bytes = (unsigned)(8 * ts * sizeof(double));
    const unsigned slab_bytes = (unsigned)((size_t)ndp * ts * sizeof(double));
    const unsigned x_bytes = (unsigned)((size_t)npad * 6 * sizeof(double));
    const size_t slab_doubles = (size_t)ndp * ts + (size_t)ndp / 2;
    const size_t stage_doubles = slab_doubles + (size_t)npad * 6;
    const int row = a.cta_row[blockIdx.x], begin = a.cta_begin[blockIdx.x], count = a.cta_count[blockIdx.x];
    const int it0 = (int)((long long)a.ns23 * blockIdx.y / kCmParts), it1 = (int)((long long)a.ns23 * (blockIdx.y + 1) / kCmParts);
    const int t = threadIdx.x;
    if (t == 0) {
        for (int s = 0; s < kCmStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], kCdConsumers); }
        fence_mbar_init();
    }
    __syncthreads();
    if (t >= kCdConsumers) {                                    // ---- producer warp: one lane fetches group after group
        if (t == kCdConsumers) {
            const int p0 = a.run_m0[blockIdx.x];
            const int* d23row = a.D23 + (size_t)row * a.ns23;
            const int* order = a.sg_order + (size_t)row * a.ns23;
            for (int it = it0; it < it1; ++it) {
                const int stage = (it - it0) % kCmStages;
                const int sg = order[it];
                const int c23 = d23row[sg];
                mbar_wait(&empty_bar[stage], (((unsigned)((it - it0) / kCmStages)) & 1u) ^ 1u);
                double* st = cm_smem + (size_t)stage * stage_doubles;
                mbar_arrive_expect_tx(&full_bar[stage], slab_bytes + x_bytes);
                const double* src = a.Tm + ((size_t)c23 * a.noff + p0) * ts;
                for (int q = 0; q < ndp / 8; ++q)
                    tma_load_1d(st + cd_slab_off(8 * q, ts), src + (size_t)8 * q * ts, run_bytes, &full_bar[stage]);
                cm_bulk_load(st + slab_doubles, a.xg + (size_t)sg * npad * 6, x_bytes, &full_bar[stage]);
            }
        }
        return;
    }
    const int sl = t / (BLK * 6), u = t % (BLK * 6), blk = u / 6, k = u % 6;
    const bool active = blk * G < count;
    double acc[G];
#pragma unroll
    for (int g = 0; g < G; ++g) acc[g] = 0.0;
    const int jb = sl * a.L;
    for (int it = it0; it < it1; ++it) {
        const int stage = (it - it0) % kCmStages;
        const double* Td = cm_smem + (size_t)stage * stage_doubles;
        const double* xs = Td + slab_doubles;
        mbar_wait(&full_bar[stage], ((unsigned)((it - it0) / kCmStages)) & 1u);
        if (active) {
            // window: logical g at source position j sits at dd = blk*G + g - j + npad - 1; physical slot (g - jj) mod G
            double W[G][6];
            const double* trow = Td + (size_t)k * 6;
            int dd0 = blk * G - jb + npad - 1;                  // logical 0 at j = jb
#pragma unroll
            for (int g = 1; g < G; ++g) {
                const double2* s2 = reinterpret_cast<const double2*>(trow + cd_slab_off(dd0 + g, ts));
                const double2 v0 = s2[0], v1 = s2[1], v2 = s2[2];
                W[g][0] = v0.x; W[g][1] = v0.y; W[g][2] = v1.x; W[g][3] = v1.y; W[g][4] = v2.x; W[g][5] = v2.y;
            }
            for (int j = jb; j < jb + a.L; j += G) {
#pragma unroll
                for (int jj = 0; jj < G; ++jj) {
                    {
                        const int slot = (G - jj) % G;
                        const double2* s2 = reinterpret_cast<const double2*>(trow + cd_slab_off(dd0 - jj, ts));
                        const double2 v0 = s2[0], v1 = s2[1], v2 = s2[2];
                        W[slot][0] = v0.x; W[slot][1] = v0.y; W[slot][2] = v1.x; W[slot][3] = v1.y; W[slot][4] = v2.x; W[slot][5] = v2.y;
                    }
                    const double2* x2 = reinterpret_cast<const double2*>(xs + (size_t)(j + jj) * 6);
                    const double2 a0 = x2[0], a1 = x2[1], a2 = x2[2];
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        const int slot = (g - jj + G) % G;
                        double v = acc[g];
                        v = fma(W[slot][0], a0.x, v); v = fma(W[slot][1], a0.y, v); v = fma(W[slot][2], a1.x, v);
                        v = fma(W[slot][3], a1.y, v); v = fma(W[slot][4], a2.x, v); v = fma(W[slot][5], a2.y, v);
                        acc[g] = v;
                    }
                }
                dd0 -= G;
            }
        }
        mbar_arrive(&empty_bar[stage]);
    }
    // fold the slices in order (own region behind the stages)
    double* red = cm_smem + (size_t)kCmStages * stage_doubles;  // [slices][blk][G][6]
    if (sl > 0 && active) {
#pragma unroll
        for (int g = 0; g < G; ++g) red[(((size_t)sl * BLK + blk) * G + g) * 6 + k] = acc[g];
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kCdConsumers) : "memory");
    if (sl == 0 && active) {
        for (int s = 1; s < kCdSlices; ++s) {
#pragma unroll
            for (int g = 0; g < G; ++g) acc[g] += red[(((size_t)s * BLK + blk) * G + g) * 6 + k];
        }
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int m = blk * G + g;
            if (m < count) {
                const int o = a.out_map[(size_t)(begin + m) * 6 + k];
                if (o >= 0) a.part[(size_t)blockIdx.y * a.nout + o] = acc[g];
            }
        }
    }
}

// y_out = (y_in) + the quarters in order
__global__ void __launch_bounds__(256)
class_fold_kernel(const double* __restrict__ part, size_t n, const double* y_in, double* y_out, const int* done)
{
    if (done && *reinterpret_cast<const volatile int*>(done)) return;
    const size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n) return;
    double v = part[o];
#pragma unroll
    for (int p = 1; p < kCmParts; ++p) v += part[(size_t)p * n + o];
    y_out[o] = (y_in ? y_in[o] : 0.0) + v;
}

template <class Kern>
static int cm_set_smem(Kern kern, size_t need, size_t& have)
{
    if (need > have) {
        OQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
        have = need;
    }
    return 0;
}

// y_out = (y_in) + A x for a class-form operand; x / x_stride as in MatOperand
static int class_matvec(const OqMatrix* A, const double* x, size_t x_stride, const double* y_in, double* y_out,
                        const PeerWait& pw, const int* done, cudaStream_t st)
{
    const ClassOperand& c = *A->cls;
    if (c.nctas == 0) return 0;
    // OQ_CLASSMV=generic keeps the general kernel (validation twin of the diagonal fast path)
    const char* env = getenv("OQ_CLASSMV");
    const bool diag = c.diag_ok && c.ndctas > 0 && !(env && strcmp(env, "generic") == 0);
    ClassMvArgs a{};
    a.ts = c.ts; a.n1 = c.n1; a.ns1 = c.ns1; a.ns23 = c.ns23;
    a.rc1 = c.rc1.p; a.D1 = c.D1.p; a.D23 = c.D23.p;
    a.rg_items = c.rg_items.p; a.sg_ptr = c.sg_ptr.p; a.sg_order = c.sg_order.p;
    a.nr = c.nr; a.y_in = y_in; a.y_out = y_out; a.done = done; a.part = c.part.p;
    const size_t nout = (size_t)c.K * c.nr;
    const unsigned fblocks = (unsigned)((nout + 255) / 256);
    if (diag) {
        // 1. the forcing vector by (source group, coarse position, column) -- waits for the peers
        const size_t n = (size_t)c.ns23 * kCdSlices * c.dL * 6;
        class_gather_flat_kernel<<<(unsigned)std::min<size_t>((n + 255) / 256, 148 * 8), 256, 0, st>>>(c.dxmap.p, n, x, x_stride, pw, done, c.dxg.p);
        OQ_LAUNCHED();
        // 2. the product
        a.ts = 38; a.xg = c.dxg.p; a.xstride = kCdSlices * c.dL;
        a.Tm = c.Td.p; a.noff = c.noff; a.L = c.dL; a.nout = (int)nout;
        a.cta_row = c.dcta_row.p; a.cta_begin = c.dcta_begin.p; a.cta_count = c.dcta_count.p;
        a.run_m0 = c.dcta_m0.p; a.out_map = c.dout_map.p;
        static size_t dsmem_set[2] = {48 * 1024, 48 * 1024};
        if (c.dblk == 8) {
            OQ_TRY(cm_set_smem(class_matvec_diag_kernel<8>, c.dsmem, dsmem_set[0]));
            class_matvec_diag_kernel<8><<<dim3(c.ndctas, kCmParts), 8 * 6 * kCdSlices + 32, c.dsmem, st>>>(a);
        } else {
            OQ_TRY(cm_set_smem(class_matvec_diag_kernel<4>, c.dsmem, dsmem_set[1]));
            class_matvec_diag_kernel<4><<<dim3(c.ndctas, kCmParts), 4 * 6 * kCdSlices + 32, c.dsmem, st>>>(a);
        }
        OQ_LAUNCHED();
        class_fold_kernel<<<fblocks, 256, 0, st>>>(c.part.p, nout, y_in, y_out, done);
        OQ_LAUNCHED();
        return 0;
    }
    // 1. the forcing vector in group order (waits for the peers)
    {
        const size_t nslots = (size_t)c.ns23 * c.xstride;
        const unsigned gblocks = (unsigned)std::min<size_t>((nslots + 255) / 256, 148 * 8);
        if (c.P == 6) class_gather_x_kernel<6><<<gblocks, 256, 0, st>>>(c.xmap.p, nslots, c.ns, x, x_stride, pw, done, c.xg.p);
        else class_gather_x_kernel<1><<<gblocks, 256, 0, st>>>(c.xmap.p, nslots, c.ns, x, x_stride, pw, done, c.xg.p);
        OQ_LAUNCHED();
        a.xg = c.xg.p; a.xstride = c.xstride;
    }
    // 2. the product
    a.Tm = c.Tm.p; a.csg = c.csg.p; a.d1_smem = c.d1_smem ? 1 : 0;
    a.cta_row = c.cta_row.p; a.cta_begin = c.cta_begin.p; a.cta_count = c.cta_count.p;
    static size_t smem_set[6] = {48 * 1024, 48 * 1024, 48 * 1024, 48 * 1024, 48 * 1024, 48 * 1024};
#define OQ_CM_LAUNCH(KK, PP, DD, IDX)                                                          \
    do {                                                                                        \
        OQ_TRY(cm_set_smem(class_matvec_kernel<KK, PP, DD>, c.smem, smem_set[IDX]));           \
        class_matvec_kernel<KK, PP, DD><<<dim3(c.nctas, kCmParts), kCmThreads, c.smem, st>>>(a); \
    } while (0)
    if (c.K == 6 && c.P == 6) { if (c.d1_smem) OQ_CM_LAUNCH(6, 6, true, 0); else OQ_CM_LAUNCH(6, 6, false, 1); }
    else if (c.K == 6 && c.P == 1) { if (c.d1_smem) OQ_CM_LAUNCH(6, 1, true, 2); else OQ_CM_LAUNCH(6, 1, false, 3); }
    else if (c.K == 1 && c.P == 6) { if (c.d1_smem) OQ_CM_LAUNCH(1, 6, true, 4); else OQ_CM_LAUNCH(1, 6, false, 5); }
    else return fail("class-form operand with %dx%d blocks is not supported", c.K, c.P);
#undef OQ_CM_LAUNCH
    OQ_LAUNCHED();
    class_fold_kernel<<<fblocks, 256, 0, st>>>(c.part.p, nout, y_in, y_out, done);
    OQ_LAUNCHED();
    return 0;
}

}  // namespace oq
