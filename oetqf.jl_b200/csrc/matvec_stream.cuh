// matvec_stream.cuh -- the HBM-streaming fused matvec: persistent, warp-specialised, TMA-fed.
//
// One CTA per SM.  The work of a launch is the flat sequence of CHUNKS (kStR rows x kStCH columns of one
// operand) of every row block of both row sets (fault rows: G11|G21, mantle rows: G12|G22); CTA b owns the
// contiguous span [b*T/G, (b+1)*T/G) of it, so the load is balanced to one chunk.  Three roles per CTA:
//   producer warp   streams each chunk -- kStR row pieces of the matrix plus the matching piece of the forcing
//                   vector -- into a ring of kStStages shared-memory stages with 1-D bulk TMA copies
//                   (cp.async.bulk -> SASS UBLKCP) completing on mbarriers;
//   8 consumer warps multiply-accumulate out of shared memory; at the end of a row block they only drop their
//                   partial sums into a double-buffered slot and keep streaming;
//   epilogue warp   folds the 8 partial sums (fixed order), merges row blocks that straddle two CTAs (the last
//                   arriver sums the slots in a fixed order: bitwise deterministic) and applies the pointwise
//                   physics (rhs.cu: update_fault_row / stress-rate store) off the streaming path.
// In-flight HBM bytes live in shared memory (kStStages x 40 KB per SM), not in registers, and neither the
// reduction nor the physics ever drains the pipeline.
//
// Traversal direction.  A CTA walks the row blocks of its span forwards on even launches and backwards on odd ones
// (chunks inside a row block always forwards, so every sum is formed in the same order: results are bitwise
// independent of the direction).  The integrator evaluates the same matrices again and again; an evaluation that
// starts where the previous one ended finds the last ~L2-size worth of matrix still in the 126 MB L2 instead of
// fetching it from HBM -- a few per cent of a 2 GB shard, a third of a 268 MB one (8-GPU shards of configs[2]).
#pragma once

namespace oq {

constexpr int kStR = 4;            // rows per row block
constexpr int kStCH = 1024;        // columns per chunk
constexpr int kStStages = 5;
constexpr int kStConsumers = 256;  // 8 consumer warps; warp 8 = producer, warp 9 = epilogue
constexpr int kStCWarps = kStConsumers / 32;
constexpr int kStThreads = kStConsumers + 64;
constexpr int kStStageDoubles = (kStR + 1) * kStCH;
constexpr size_t kStSmemBytes = (size_t)kStStages * kStStageDoubles * sizeof(double) + 1024;

// The span [g_begin, g_end) of a CTA, cut at row-block boundaries into SEGMENTS, in processing order.
struct SpanWalk {
    long long g_begin, g_end;
    int grb_first, nseg;       // first global row block touched (row sets concatenated), number of segments
    bool reverse;
    __device__ __forceinline__ static int global_rb(const MatvecArgs& a, long long g)
    {
        const int job = ((g >= a.job[1].chunk_begin && a.job[1].nrb > 0) || a.job[0].nrb == 0) ? 1 : 0;
        const MatvecJob& j = a.job[job];
        return (job ? a.job[0].nrb : 0) + (int)((g - j.chunk_begin) / j.chunks_per_rb);
    }
    __device__ __forceinline__ void init(const MatvecArgs& a, long long gb, long long ge, bool rev)
    {
        g_begin = gb; g_end = ge; reverse = rev;
        grb_first = global_rb(a, gb);
        nseg = global_rb(a, ge - 1) - grb_first + 1;
    }
    // segment k: row set, row block, first chunk inside the row block, number of chunks, chunk range of the row block
    __device__ __forceinline__ void get(const MatvecArgs& a, int k, int& job, int& rb, int& rem0, int& n,
                                        long long& rb_g0, long long& rb_g1) const
    {
        const int grb = reverse ? grb_first + nseg - 1 - k : grb_first + k;
        job = grb >= a.job[0].nrb ? 1 : 0;
        const MatvecJob& j = a.job[job];
        rb = grb - (job ? a.job[0].nrb : 0);
        rb_g0 = j.chunk_begin + (long long)rb * j.chunks_per_rb;
        rb_g1 = rb_g0 + j.chunks_per_rb;
        const long long lo = rb_g0 > g_begin ? rb_g0 : g_begin;
        const long long hi = rb_g1 < g_end ? rb_g1 : g_end;
        rem0 = (int)(lo - rb_g0);
        n = (int)(hi - lo);
    }
};

// chunk-by-chunk iteration over a span in processing order
struct ChunkIter {
    int k, c, job, rb, rem0, n;
    __device__ __forceinline__ void load(const MatvecArgs& a, const SpanWalk& w)
    {
        long long g0, g1;
        if (k < w.nseg) w.get(a, k, job, rb, rem0, n, g0, g1);
    }
    __device__ __forceinline__ void start(const MatvecArgs& a, const SpanWalk& w) { k = 0; c = 0; load(a, w); }
    __device__ __forceinline__ bool valid(const SpanWalk& w) const { return k < w.nseg; }
    __device__ __forceinline__ int rem() const { return rem0 + c; }
    __device__ __forceinline__ bool last_of_segment() const { return c + 1 == n; }
    __device__ __forceinline__ void next(const MatvecArgs& a, const SpanWalk& w)
    {
        if (++c == n) { ++k; c = 0; load(a, w); }
    }
};

__device__ __forceinline__ long long span_begin(long long total, int grid, int b)
{
    return (total * b) / grid;
}

// the CTA that owns global chunk g
__device__ __forceinline__ int owner_of(long long total, int grid, long long g)
{
    int b = (int)((g * grid) / total);
    while (b + 1 < grid && span_begin(total, grid, b + 1) <= g) ++b;
    while (b > 0 && span_begin(total, grid, b) > g) --b;
    return b;
}

__global__ void __launch_bounds__(kStThreads, 1)
matvec_stream_kernel(const __grid_constant__ MatvecArgs args)
{
    extern __shared__ __align__(128) double smem[];
    __shared__ __align__(8) uint64_t full_bar[kStStages];
    __shared__ __align__(8) uint64_t empty_bar[kStStages];
    __shared__ __align__(8) uint64_t red_full[2];
    __shared__ __align__(8) uint64_t red_empty[2];
    __shared__ double red[2][kStCWarps][kStR];
    __shared__ int reverse_s;

    if (args.done && *reinterpret_cast<const volatile int*>(args.done)) return;   // integration already complete
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const long long total = args.total_chunks;
    const int grid = gridDim.x;
    const long long g_begin = span_begin(total, grid, blockIdx.x);
    const long long g_end = span_begin(total, grid, blockIdx.x + 1);

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kStCWarps);
        }
        mbar_init(&red_full[0], kStCWarps); mbar_init(&red_full[1], kStCWarps);
        mbar_init(&red_empty[0], 1); mbar_init(&red_empty[1], 1);
        fence_mbar_init();
        // (the previous launch of this plan has completed: its last CTA bumped pass[0] before the kernel ended)
        reverse_s = args.pass ? (int)(*reinterpret_cast<const volatile unsigned long long*>(args.pass) & 1ull) : 0;
    }
    __syncthreads();
    // every CTA reports once per launch, after it has read the direction; the last one flips it for the next launch
    auto report_done = [&]() {
        if (!args.pass) return;
        __threadfence();
        const unsigned long long prev = atomicAdd(args.pass + 1, 1ull);
        if (prev == (unsigned long long)gridDim.x - 1ull) {
            args.pass[1] = 0ull;
            __threadfence();
            atomicAdd(args.pass, 1ull);
        }
    };
    if (g_begin >= g_end) {
        if (tid == 0) report_done();
        return;
    }
    SpanWalk walk;
    walk.init(args, g_begin, g_end, reverse_s != 0);

    if (warp == kStCWarps) {
        // ------------------------------------------------------------------ producer warp
        // The matrix does not depend on this evaluation's forcing vector: the first ring of matrix pieces is
        // requested BEFORE waiting for the forcing kernel / the peers' publication.
        auto issue = [&](const ChunkIter& c, int stg, bool matrix, bool vector, size_t par) {
            const MatvecJob& j = args.job[c.job];
            const int rem = c.rem();
            const int osel = rem < j.nch[0] ? 0 : 1;
            const MatOperand& op = j.op[osel];
            const int c0 = (osel ? rem - j.nch[0] : rem) * kStCH;
            const int ncol = min(kStCH, (int)op.ld - c0);
            const unsigned bytes = (unsigned)(ncol * sizeof(double));
            double* dst = smem + (size_t)stg * kStStageDoubles;
            if (matrix) {
                mbar_arrive_expect_tx(&full_bar[stg], bytes * (kStR + 1));
                // L2 residency: the first `keep_chunks` chunks of every span are asked to stay in L2 (evicted
                // last), everything else to leave first -- the next evaluation finds the kept part on chip
                const long long g = j.chunk_begin + (long long)c.rb * j.chunks_per_rb + rem;
                const unsigned long long pol = args.keep_chunks < 0 ? kL2EvictNormal
                                               : (g - g_begin < args.keep_chunks ? kL2EvictLast : kL2EvictFirst);
#pragma unroll
                for (int r = 0; r < kStR; ++r) {
                    const int row = min(c.rb * kStR + r, j.nrows - 1);
                    tma_load_1d_hint(dst + r * kStCH, op.G + (size_t)row * op.ld + c0, bytes, &full_bar[stg], pol);
                }
            }
            if (vector)
                tma_load_1d_hint(dst + kStR * kStCH, op.x + par * op.x_stride + c0, bytes, &full_bar[stg],
                                 args.keep_chunks < 0 ? kL2EvictNormal : kL2EvictLast);
        };
        ChunkIter cur, pre;
        cur.start(args, walk);
        pre = cur;
        const int npre = (int)min((long long)kStStages, g_end - g_begin);
        if (lane == 0)
            for (int s = 0; s < npre; ++s) { issue(cur, s, true, false, 0); cur.next(args, walk); }
        __syncwarp();
        pdl_wait();                                   // the forcing kernel (predecessor) is complete from here on
        size_t par = 0;
        if (args.pw.epochs) {
            const unsigned long long ep = *(volatile unsigned long long*)(args.pw.epochs + kEpForcing);
            if (args.pw.world > 1) {
                if (args.pw.warp_poll)                // all peers polled at once, one lane each
                    wait_peers_warp(args.pw.flags, args.pw.world, args.pw.rank, ep, args.pw.epochs + kEpError, lane);
                else if (lane == 0)
                    wait_peers(args.pw.flags, args.pw.world, args.pw.rank, ep, args.pw.epochs + kEpError);
                __syncwarp();
                fence_proxy_async();                  // peer stores -> async-proxy (TMA) reads
            }
            par = (size_t)((ep - 1ull) & 1ull);
        }
        if (lane != 0) return;
        for (int s = 0; s < npre; ++s) { issue(pre, s, false, true, par); pre.next(args, walk); }
        int stage = npre % kStStages;
        unsigned phase = npre >= kStStages ? 1u : 0u;
        for (long long g = g_begin + npre; g < g_end; ++g) {
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            issue(cur, stage, true, true, par);
            cur.next(args, walk);
            if (++stage == kStStages) { stage = 0; phase ^= 1u; }
        }
        return;
    }

    if (warp == kStCWarps + 1) {
        // ------------------------------------------------------------------ epilogue warp
        pdl_wait();                                   // the physics reads the predecessor's state
        int buf = 0;
        unsigned rphase[2] = {0u, 0u};
        for (int k = 0; k < walk.nseg; ++k) {
            int jb, rb, rem0, nch;
            long long rb_g0, rb_g1;
            walk.get(args, k, jb, rb, rem0, nch, rb_g0, rb_g1);   // the part of this row block inside my span
            const MatvecJob& j = args.job[jb];
            const int myrow = rb * kStR + lane;
            const bool active = lane < kStR && myrow < j.nrows;
            // the row's state and properties are fetched while the consumers still stream its chunks
            FaultRowInputs fin{};
            if (active && j.epilogue == kEpiFault) fin = load_fault_row(args.fe, myrow);
            mbar_wait(&red_full[buf], rphase[buf]);   // all consumer warps have dropped their partial sums
            rphase[buf] ^= 1u;
            double mine = 0.0;
            if (lane < kStR) {
#pragma unroll
                for (int w = 0; w < kStCWarps; ++w) mine += red[buf][w][lane];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&red_empty[buf]);          // consumers may reuse the slot
            buf ^= 1;
            // contributors: only a row block cut by a span boundary has more than one
            int first = blockIdx.x, last = blockIdx.x;
            if (rb_g0 < g_begin) first = owner_of(total, grid, rb_g0);
            if (rb_g1 > g_end) last = owner_of(total, grid, rb_g1 - 1);
            const int ncontrib = last - first + 1;
            bool do_epilogue = true;
            if (ncontrib > 1) {
                const int slot = (int)blockIdx.x - first;
                if (active) j.partial[((size_t)myrow) * j.slots + slot] = mine;
                __threadfence();
                __syncwarp();
                unsigned prev = 0;
                if (lane == 0) {
                    prev = atomicAdd(&j.counters[rb], 1u);
                    if (prev == (unsigned)ncontrib - 1u) j.counters[rb] = 0u;   // re-arm for the next evaluation
                }
                prev = __shfl_sync(0xffffffffu, prev, 0);
                do_epilogue = (prev == (unsigned)ncontrib - 1u);
                if (do_epilogue) {
                    __threadfence();
                    if (active) {
                        mine = 0.0;
                        const double* pp = j.partial + (size_t)myrow * j.slots;
                        for (int q = 0; q < ncontrib; ++q) mine += ld_cg(pp + q);   // fixed order: deterministic
                    }
                }
            }
            if (do_epilogue && active) {
                if (j.y0) mine += j.y0[myrow];
                if (j.epilogue == kEpiFault) update_fault_row(args.fe, myrow, mine, fin);
                else j.yout[myrow] = mine;
            }
        }
        if (lane == 0) report_done();
        return;
    }

    // ---------------------------------------------------------------------- consumer warps
    double acc[kStR];
#pragma unroll
    for (int r = 0; r < kStR; ++r) acc[r] = 0.0;
    ChunkIter c;
    c.start(args, walk);
    int stage = 0, buf = 0;
    unsigned phase = 0;
    unsigned ephase[2] = {0u, 0u};

    for (; c.valid(walk); c.next(args, walk)) {
        const MatvecJob& j = args.job[c.job];
        const int rem = c.rem();
        const int osel = rem < j.nch[0] ? 0 : 1;
        const MatOperand& op = j.op[osel];
        const int ncol = min(kStCH, (int)op.ld - (osel ? rem - j.nch[0] : rem) * kStCH);
        mbar_wait(&full_bar[stage], phase);
        const double2* s2 = reinterpret_cast<const double2*>(smem + (size_t)stage * kStStageDoubles);
#pragma unroll
        for (int it = 0; it < kStCH / (2 * kStConsumers); ++it) {
            const int c2 = tid + it * kStConsumers;               // double2 index inside the chunk
            if (2 * c2 < ncol) {
                const double2 xv = s2[kStR * (kStCH / 2) + c2];
#pragma unroll
                for (int r = 0; r < kStR; ++r) {
                    const double2 gv = s2[r * (kStCH / 2) + c2];
                    acc[r] = fma(gv.x, xv.x, acc[r]);
                    acc[r] = fma(gv.y, xv.y, acc[r]);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == kStStages) { stage = 0; phase ^= 1u; }
        // end of the row block (or of my part of it): hand the partial sums to the epilogue warp and keep streaming
        if (c.last_of_segment()) {
#pragma unroll
            for (int r = 0; r < kStR; ++r) {
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], off);
            }
            if (lane == 0) {
                mbar_wait(&red_empty[buf], ephase[buf] ^ 1u);     // slot free (always, except pathologically)
#pragma unroll
                for (int r = 0; r < kStR; ++r) red[buf][warp][r] = acc[r];
                mbar_arrive(&red_full[buf]);                      // release: the stores above are visible
            }
            ephase[buf] ^= 1u;
            buf ^= 1;
#pragma unroll
            for (int r = 0; r < kStR; ++r) acc[r] = 0.0;
        }
    }
}

}  // namespace oq
