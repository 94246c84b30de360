// common.cuh -- runtime plumbing shared by the translation units of liboetqf_b200.so.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/oetqf_b200.h"

namespace oq {

// ---- error reporting (never throw across the ABI) -------------------------------------------
std::string& last_error();
int fail(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;
int current_device();

#define OQ_CUDA(expr)                                                                          \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return ::oq::fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

#define OQ_CHECK(cond, ...)                                   \
    do {                                                      \
        if (!(cond)) return ::oq::fail(__VA_ARGS__);          \
    } while (0)

#define OQ_TRY(expr)                   \
    do {                               \
        int rc__ = (expr);             \
        if (rc__ != 0) return rc__;    \
    } while (0)

// counts a kernel launch of this library and checks the launch itself
#define OQ_LAUNCHED()                                   \
    do {                                                \
        ::oq::g_launches.fetch_add(1);                  \
        OQ_CUDA(cudaGetLastError());                    \
    } while (0)

// Makes sure the calling host thread targets the library's device (Julia tasks migrate between
// threads; CUDA's current device is per-thread).
int enter();

// ---- RAII device memory ---------------------------------------------------------------------
template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    int alloc(size_t count) {
        release();
        if (count == 0) return 0;
        cudaError_t e = cudaMalloc(&p, count * sizeof(T));
        if (e != cudaSuccess)
            return fail("cudaMalloc of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(e));
        n = count;
        return 0;
    }
    int upload(const T* host, size_t count) {
        OQ_TRY(alloc(count));
        if (count) OQ_CUDA(cudaMemcpy(p, host, count * sizeof(T), cudaMemcpyHostToDevice));
        return 0;
    }
    int zero() {
        if (n) OQ_CUDA(cudaMemset(p, 0, n * sizeof(T)));
        return 0;
    }
};

struct EventTimer {
    cudaEvent_t a = nullptr, b = nullptr;
    ~EventTimer() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
    int start(cudaStream_t s = 0) {
        if (!a) { OQ_CUDA(cudaEventCreate(&a)); OQ_CUDA(cudaEventCreate(&b)); }
        OQ_CUDA(cudaEventRecord(a, s));
        return 0;
    }
    int stop(double* ms, cudaStream_t s = 0) {
        OQ_CUDA(cudaEventRecord(b, s));
        OQ_CUDA(cudaEventSynchronize(b));
        float f = 0;
        OQ_CUDA(cudaEventElapsedTime(&f, a, b));
        if (ms) *ms = f;
        return 0;
    }
};

inline size_t round_up(size_t v, size_t m) { return (v + m - 1) / m * m; }

// sind/cosd on the host with exact values at multiples of 90° (Julia's sincosd)
void sincosd(double deg, double* s, double* c);

constexpr int kCdG = 8, kCdBlk = 8, kCdSlices = 4;     // class_matvec_diag_kernel: window length, receiver blocks, source slices

// row stride (16-bit entries) of the D1 rows class_matvec_kernel keeps in shared memory: smallest value >= ns1 that is 2 mod 4
inline __host__ __device__ int cm_d1_stride(int ns1) { return ns1 + ((6 - (ns1 & 3)) & 3); }

// Class form of a Green's operand (classmat.cuh): the table of DISTINCT kernels of a matrix whose (receiver, source)
// pairs fall into translation classes, and the maps from a pair to its class.  An OqMatrix in this form has no dense
// storage; the RHS multiplies straight from the table (class_matvec_kernel).
struct ClassOperand {
    int K = 1, P = 1;                      // rows per receiver unit, columns per source unit (6x6, 6x1, 1x6)
    int nr = 0, ns = 0;                    // local receiver units, source units
    int n1 = 0, n23 = 0, ns1 = 0, ns23 = 0;
    int ts = 0;                            // doubles per class in Tm (K*P padded so that 128-bit loads of 8 consecutive classes hit 32 distinct banks)
    DevBuf<double> Tm;                     // [n23][n1][ts]
    DevBuf<int> rc1, sc1, D1, D23;         // x class of every local receiver / every source; pair-class maps [nr1*ns1], [nr23*ns23]
    DevBuf<int> rc23, sc23;                // (y,z) classes (dense expansion only)
    DevBuf<int> sg_order;                  // [nr23][ns23] per row of D23: source groups in ascending class order
    DevBuf<int> rg_items, sg_ptr;          // receivers ordered by (y,z) class; sources grouped by (y,z) class (CSR over ns23 groups)
    DevBuf<int> xmap, csg;                 // [ns23][xstride]: source of a group slot (-1: padding) and its x class
    DevBuf<double> xg;                     // [ns23][xstride][PX] forcing values in group order (scratch of an evaluation)
    DevBuf<double> part;                   // [4][K][nr] partial sums of the source-group quarters (scratch of an evaluation)
    int xstride = 0;
    DevBuf<int> cta_row, cta_begin, cta_count;   // work list: one CTA = a run of <= 64 receivers of one (y,z) class (its row of D23)
    int nctas = 0, max_sg = 0;
    size_t smem = 0;
    bool d1_smem = false;                  // the D1 rows of a CTA's receivers fit shared memory beside the stages
    double table_bytes = 0;
    // sliding-window kernel (class_matvec_diag_kernel): operands whose x classes are a function of (residue, coarse
    // position difference) -- receivers and sources on commensurate equidistant grids along x (the Toeplitz structure of
    // GF.jl:31-71); see OffsetPlan in greens_classes.cuh
    bool diag_ok = false;
    int dmode = 0, dQ = 1;                 // 0: 6x6 on one grid, 1: 1x6 (receivers finer), 2: 6x1 (sources finer); residues
    int dblk = kCdBlk;                     // receiver blocks of 8 per CTA run (8, or 4 on shards with few runs)
    int npos = 0, dL = 0, noff = 0;        // coarse source positions; positions per slice (multiple of the window length); padded offsets
    DevBuf<double> Td;                     // [n23][noff][38] the table in offset order, zero padded
    DevBuf<int> dxmap;                     // [ns23][4 dL][6]: flat index into x of (source group, coarse position, column), -1: none
    DevBuf<double> dxg;                    // [ns23][4 dL][6]
    DevBuf<int> dout_map;                  // [entries][6]: flat index into y of (coarse receiver position of a run, row), -1: none
    DevBuf<int> dcta_row, dcta_begin, dcta_count, dcta_m0;
    int ndctas = 0;
    size_t dsmem = 0;
};

}  // namespace oq

// ---- the opaque handles -----------------------------------------------------------------------
struct OqMatrix {
    int row_kind = OQ_ROWS_FAULT;   // how [row_begin,row_end) maps to rows
    int row_begin = 0, row_end = 0; // fault cells or mantle elements
    int global_rows = 0;            // nf or 6*ne
    int local_rows = 0;             // (row_end-row_begin) or 6*(row_end-row_begin)
    int cols = 0;
    size_t ld = 0;                  // leading dimension in doubles (multiple of 16)
    oq::DevBuf<double> d;           // [local_rows * ld], row-major, padding zeroed
    double kernel_ms = 0.0;         // device time of the assembly kernel(s)
    int path = -1;                  // hex8 builders: 0 pair, 1 tile, 2 class-table kernels (-1: not a hex8 matrix)
    long long pairs = 0, unique_pairs = 0;       // (receiver, source) pairs of the shard / closed-form evaluations made
    double table_ms = 0.0, expand_ms = 0.0;      // class path: evaluation of the classes / dense expansion
    std::unique_ptr<oq::ClassOperand> cls;       // non-null: class form (no dense storage, d is empty)
};
