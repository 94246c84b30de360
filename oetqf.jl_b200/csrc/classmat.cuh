// classmat.cuh -- K9: the RHS matvec of a Green's operand kept in CLASS FORM (included by rhs.cu).
//
// The three mantle operands of equation.jl:201-203 (`dτ += gf₂₁·reldϵ`, `dσ = gf₁₂·relv`, `dσ += gf₂₂·reldϵ`) are dense
// in the reference, and the dense form is what the north star streams from HBM.  On the meshes the package builds
// (Gmsh transfinite boxes, mesh.jl:95-130; the equidistant fault, mesh.jl:39-56) their (receiver, source) pairs fall
// into translation classes (greens_classes.cuh): G[(k, r), (p, s)] = T[class(r, s)][k][p] with a few thousand to a
// few million distinct 6x6 (6x1, 1x6) blocks instead of (6 N_e)² entries -- the same invariance the reference itself
// exploits for the fault (Toeplitz kernel, GF.jl:31-71), extended to the mantle.  This kernel multiplies straight
// from the table:
//
//     y[k, r] = y_in[k, r] + Σ_s Σ_p T[D23[rc23(r), sc23(s)]][D1[rc1(r), sc1(s)]][k][p] · x[p, s]
//
// so a problem whose dense gf₂₂ would need 1.84 TB (BASELINE configs[3]: 80 000 cells) keeps a 3.6 GB table, and the
// evaluation is bound by the fp64 pipe / shared-memory bandwidth instead of HBM.  Nothing about the mesh is assumed
// beyond what the class maps say (they are found numerically); a mesh without structure has no class form and keeps
// the dense operands.
//
// One CTA = a run of <= rb receivers of one (y,z) class (they share the row D23[rc23, :]) x all sources.  Sources are
// walked group by group ((y,z) class of the source): the slab T[c23][:][:] of the group (n1 classes x K*P doubles,
// class-major, padded so that 128-bit loads of consecutive classes are bank-conflict free) and the group's forcing
// values are staged in shared memory, then thread (slice, receiver) accumulates its K outputs over every nsl-th
// source of the group.  Sums are formed in a fixed order (sources ascending within a slice, slices folded in order):
// results are bitwise reproducible and independent of the number of ranks (a rank's receivers see all sources).
#pragma once

namespace oq {

constexpr int kCmThreads = 256;

struct ClassMvArgs {
    const double* Tm;
    int ts, n1, ns1, ns23;
    const int *rc1, *sc1, *D1, *D23;
    const int *rg_items, *sg_ptr, *sg_items, *sg_order, *cta_row, *cta_begin, *cta_count;
    int rb, max_sg;
    int nr, ns;                   // local receiver units, source units (= stride between the P planes of x)
    const double* x;              // copy 0 of the forcing vector
    size_t x_stride;              // distance to copy 1 (parity of the evaluation)
    const double* y_in;           // optional: accumulate onto (may alias y_out)
    double* y_out;                // [K][nr]
    PeerWait pw;
    const int* done;              // optional device flag: integration complete, skip
};

template <int K, int P>
__global__ void __launch_bounds__(kCmThreads)
class_matvec_kernel(const __grid_constant__ ClassMvArgs a)
{
    constexpr int KP = K * P;
    constexpr int PX = (P + 1) & ~1;
    static_assert(KP % 2 == 0, "class blocks are loaded as double2");
    extern __shared__ __align__(16) double cm_smem[];
    if (a.done && *reinterpret_cast<const volatile int*>(a.done)) return;
    double* Ts = cm_smem;                                       // [n1][ts]
    double* xs = Ts + (size_t)a.n1 * a.ts;                      // [max_sg][PX]
    int* cs = reinterpret_cast<int*>(xs + (size_t)a.max_sg * PX);   // [max_sg] x class of the group's sources
    const double* x = a.x + consumer_parity(a.pw) * a.x_stride;
    const int row = a.cta_row[blockIdx.x], begin = a.cta_begin[blockIdx.x], count = a.cta_count[blockIdx.x];
    const int rb = a.rb, nsl = kCmThreads / rb;
    const int i = threadIdx.x % rb, sl = threadIdx.x / rb;
    const bool active = i < count;
    const int r = active ? a.rg_items[begin + i] : 0;
    const int* d1row = a.D1 + (size_t)(active ? a.rc1[r] : 0) * a.ns1;
    const int* d23row = a.D23 + (size_t)row * a.ns23;
    double acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.0;
    const int n2 = a.n1 * a.ts / 2;
    const int* order = a.sg_order + (size_t)row * a.ns23;
    for (int it = 0; it < a.ns23; ++it) {
        const int sg = order[it];
        const int s0 = a.sg_ptr[sg], sn = a.sg_ptr[sg + 1] - s0;
        if (sn == 0) continue;
        const int c23 = d23row[sg];
        __syncthreads();                                        // the previous group has been consumed
        {
            const double2* src = reinterpret_cast<const double2*>(a.Tm + (size_t)c23 * a.n1 * a.ts);
            double2* dst = reinterpret_cast<double2*>(Ts);
            for (int q = threadIdx.x; q < n2; q += kCmThreads) dst[q] = __ldg(src + q);
        }
        for (int j = threadIdx.x; j < sn; j += kCmThreads) {
            const int s = a.sg_items[s0 + j];
            cs[j] = a.sc1[s];
#pragma unroll
            for (int p = 0; p < P; ++p) xs[j * PX + p] = x[(size_t)p * a.ns + s];
        }
        __syncthreads();
        if (active) {
            for (int j = sl; j < sn; j += nsl) {
                const int c1 = __ldg(d1row + cs[j]);
                const double2* t2 = reinterpret_cast<const double2*>(Ts + (size_t)c1 * a.ts);
                double tv[KP];
#pragma unroll
                for (int q = 0; q < KP / 2; ++q) { const double2 v = t2[q]; tv[2 * q] = v.x; tv[2 * q + 1] = v.y; }
                const double* xv = xs + j * PX;
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const double xp = xv[p];
#pragma unroll
                    for (int k = 0; k < K; ++k) acc[k] = fma(tv[k * P + p], xp, acc[k]);
                }
            }
        }
    }
    // fold the slices in order
    __syncthreads();
    double* red = cm_smem;                                      // [nsl][rb][K]
    if (sl > 0 && active) {
#pragma unroll
        for (int k = 0; k < K; ++k) red[((size_t)sl * rb + i) * K + k] = acc[k];
    }
    __syncthreads();
    if (sl == 0 && active) {
        for (int s = 1; s < nsl; ++s) {
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] += red[((size_t)s * rb + i) * K + k];
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const size_t o = (size_t)k * a.nr + r;
            a.y_out[o] = (a.y_in ? a.y_in[o] : 0.0) + acc[k];
        }
    }
}


// ---- diagonal fast path (6x6 operands on one equidistant x grid) ---------------------------------------------
// In class_matvec_kernel every FMA reads its own table entry from shared memory (128 B/clk per SM against 64 DFMA/clk:
// at most 25 % of the fp64 pipe).  When the x class of a pair depends on the DIFFERENCE of the integer x positions
// only (the Toeplitz structure of GF.jl:31-71), the pairs (r+1, s+1) and (r, s) share their 6x6 block: a thread that
// owns G consecutive receivers (one output component k) and walks the sources in position order keeps a sliding
// window of G blocks-rows (6 doubles each) in registers -- one new row (48 B) per G*6 FMAs instead of one per 6.
//   thread (slice, blk, k): receivers p0 + blk*G + g (g < G), output k, source positions [slice*L, (slice+1)*L)
//   Td[dd] = row-block of the class of offset (p0 + dd - (npad - 1)), dd = (blk*G + g) - j + npad - 1   (zero outside)
constexpr int kCdThreads = kCdBlk * 6 * kCdSlices;      // 192

struct ClassDiagArgs {
    ClassMvArgs b;
    const int *diag, *rpos, *sg_bypos, *rg_items_pos;
    int npos, L;
};

__global__ void __launch_bounds__(kCdThreads)
class_matvec_diag_kernel(const __grid_constant__ ClassDiagArgs A)
{
    constexpr int G = kCdG;
    const ClassMvArgs& a = A.b;
    extern __shared__ __align__(16) double cm_smem[];
    if (a.done && *reinterpret_cast<const volatile int*>(a.done)) return;
    const int npad = kCdSlices * A.L;                           // padded source positions
    const int ndp = kCdBlk * G + npad;                          // padded diagonals
    const int ts = a.ts, ts2 = ts / 2;
    double* Td = cm_smem;                                       // [ndp][ts]
    double* xs = Td + (size_t)ndp * ts;                         // [npad][6]
    const double* x = a.x + consumer_parity(a.pw) * a.x_stride;
    const int row = a.cta_row[blockIdx.x], begin = a.cta_begin[blockIdx.x], count = a.cta_count[blockIdx.x];
    const int p0 = A.rpos[A.rg_items_pos[begin]];
    const int t = threadIdx.x, sl = t / (kCdBlk * 6), u = t % (kCdBlk * 6), blk = u / 6, k = u % 6;
    const bool active = blk * G < count;
    const int* d23row = a.D23 + (size_t)row * a.ns23;
    const int* order = a.sg_order + (size_t)row * a.ns23;
    double acc[G];
#pragma unroll
    for (int g = 0; g < G; ++g) acc[g] = 0.0;
    const int jb = sl * A.L;
    for (int it = 0; it < a.ns23; ++it) {
        const int sg = order[it];
        if (a.sg_ptr[sg + 1] == a.sg_ptr[sg]) continue;
        const int c23 = d23row[sg];
        __syncthreads();                                        // the previous group has been consumed
        {
            const double2* slab = reinterpret_cast<const double2*>(a.Tm + (size_t)c23 * a.n1 * ts);
            double2* dst = reinterpret_cast<double2*>(Td);
            for (int q = t; q < ndp * ts2; q += kCdThreads) {
                const int dd = q / ts2, w = q - dd * ts2;
                const int off = p0 + dd - (npad - 1) + A.npos - 1;          // receiver position - source position + npos - 1
                int cls = -1;
                if (off >= 0 && off < 2 * A.npos - 1) cls = __ldg(A.diag + off);
                dst[q] = cls >= 0 ? __ldg(slab + (size_t)cls * ts2 + w) : make_double2(0.0, 0.0);
            }
            const int* bypos = A.sg_bypos + (size_t)sg * A.npos;
            for (int q = t; q < npad * 6; q += kCdThreads) {
                const int j = q / 6, p = q - j * 6;
                const int s = j < A.npos ? __ldg(bypos + j) : -1;
                xs[q] = s >= 0 ? x[(size_t)p * a.ns + s] : 0.0;
            }
        }
        __syncthreads();
        if (active) {
            // window: logical g at source position j sits at dd = blk*G + g - j + npad - 1; physical slot (g - jj) mod G
            double W[G][6];
            const double* trow = Td + (size_t)k * 6;
            int dd0 = blk * G - jb + npad - 1;                  // logical 0 at j = jb
#pragma unroll
            for (int g = 1; g < G; ++g) {
                const double2* s2 = reinterpret_cast<const double2*>(trow + (size_t)(dd0 + g) * ts);
                const double2 v0 = s2[0], v1 = s2[1], v2 = s2[2];
                W[g][0] = v0.x; W[g][1] = v0.y; W[g][2] = v1.x; W[g][3] = v1.y; W[g][4] = v2.x; W[g][5] = v2.y;
            }
            for (int j = jb; j < jb + A.L; j += G) {
#pragma unroll
                for (int jj = 0; jj < G; ++jj) {
                    {
                        constexpr int dummy = 0; (void)dummy;
                        const int slot = (G - jj) % G;
                        const double2* s2 = reinterpret_cast<const double2*>(trow + (size_t)(dd0 - jj) * ts);
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
    }
    // fold the slices in order
    __syncthreads();
    double* red = cm_smem;                                      // [slices][blk][G][6]
    if (sl > 0 && active) {
#pragma unroll
        for (int g = 0; g < G; ++g) red[(((size_t)sl * kCdBlk + blk) * G + g) * 6 + k] = acc[g];
    }
    __syncthreads();
    if (sl == 0 && active) {
        for (int s = 1; s < kCdSlices; ++s) {
#pragma unroll
            for (int g = 0; g < G; ++g) acc[g] += red[(((size_t)s * kCdBlk + blk) * G + g) * 6 + k];
        }
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int m = blk * G + g;
            if (m < count) {
                const size_t o = (size_t)k * a.nr + A.rg_items_pos[begin + m];
                a.y_out[o] = (a.y_in ? a.y_in[o] : 0.0) + acc[g];
            }
        }
    }
}

// y_out = (y_in) + A x for a class-form operand; x / x_stride as in MatOperand
static int class_matvec(const OqMatrix* A, const double* x, size_t x_stride, const double* y_in, double* y_out,
                        const PeerWait& pw, const int* done, cudaStream_t st)
{
    const ClassOperand& c = *A->cls;
    if (c.nctas == 0) return 0;
    ClassMvArgs a{};
    a.Tm = c.Tm.p; a.ts = c.ts; a.n1 = c.n1; a.ns1 = c.ns1; a.ns23 = c.ns23;
    a.rc1 = c.rc1.p; a.sc1 = c.sc1.p; a.D1 = c.D1.p; a.D23 = c.D23.p;
    a.rg_items = c.rg_items.p; a.sg_ptr = c.sg_ptr.p; a.sg_items = c.sg_items.p; a.sg_order = c.sg_order.p;
    a.cta_row = c.cta_row.p; a.cta_begin = c.cta_begin.p; a.cta_count = c.cta_count.p;
    a.rb = c.rb; a.max_sg = c.max_sg; a.nr = c.nr; a.ns = c.ns;
    a.x = x; a.x_stride = x_stride; a.y_in = y_in; a.y_out = y_out; a.pw = pw; a.done = done;
    // OQ_CLASSMV=generic keeps the general kernel (validation twin of the diagonal fast path)
    const char* env = getenv("OQ_CLASSMV");
    if (c.diag_ok && c.ndctas > 0 && !(env && strcmp(env, "generic") == 0)) {
        ClassDiagArgs d{};
        d.b = a;
        d.b.cta_row = c.dcta_row.p; d.b.cta_begin = c.dcta_begin.p; d.b.cta_count = c.dcta_count.p;
        d.diag = c.diag.p; d.rpos = c.rpos.p; d.sg_bypos = c.sg_bypos.p; d.rg_items_pos = c.rg_items_pos.p;
        d.npos = c.npos; d.L = c.dL;
        static size_t dsmem_set = 48 * 1024;
        if (c.dsmem > dsmem_set) {
            OQ_CUDA(cudaFuncSetAttribute(class_matvec_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.dsmem));
            dsmem_set = c.dsmem;
        }
        class_matvec_diag_kernel<<<c.ndctas, kCdThreads, c.dsmem, st>>>(d);
        OQ_LAUNCHED();
        return 0;
    }
    void (*kern)(const ClassMvArgs) = nullptr;
    int which = -1;
    if (c.K == 6 && c.P == 6) { kern = class_matvec_kernel<6, 6>; which = 0; }
    else if (c.K == 6 && c.P == 1) { kern = class_matvec_kernel<6, 1>; which = 1; }
    else if (c.K == 1 && c.P == 6) { kern = class_matvec_kernel<1, 6>; which = 2; }
    OQ_CHECK(kern, "class-form operand with %dx%d blocks is not supported", c.K, c.P);
    static size_t smem_set[3] = {48 * 1024, 48 * 1024, 48 * 1024};
    if (c.smem > smem_set[which]) {
        OQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
        smem_set[which] = c.smem;
    }
    kern<<<c.nctas, kCmThreads, c.smem, st>>>(a);
    OQ_LAUNCHED();
    return 0;
}

}  // namespace oq
