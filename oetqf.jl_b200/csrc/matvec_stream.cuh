// matvec_stream.cuh -- the HBM-streaming fused matvec: persistent, warp-specialised, TMA-fed.
//
// One CTA per SM.  The work of a launch is the flat sequence of CHUNKS (kStR rows x kStCH columns of one
// operand) of every row block of both row sets (fault rows: G11|G21, mantle rows: G12|G22); CTA b owns the
// contiguous span [b*T/G, (b+1)*T/G) of it, so the load is balanced to one chunk.  A producer warp streams each
// chunk -- kStR row pieces of the matrix plus the matching piece of the forcing vector -- into a ring of
// kStStages shared-memory stages with 1-D bulk TMA copies (cp.async.bulk -> SASS UBLKCP) completing on
// mbarriers; 8 consumer warps multiply-accumulate out of shared memory.  In-flight HBM bytes live in shared
// memory (kStStages x 40 KB per SM), not in registers, and the pipeline never drains between row blocks.
// Row blocks that straddle two CTAs write partial sums; the last arriver folds them in a fixed order
// (deterministic) and applies the pointwise physics (rhs.cu: update_fault_row / stress-rate store).
#pragma once

namespace oq {

constexpr int kStR = 4;            // rows per row block
constexpr int kStCH = 1024;        // columns per chunk
constexpr int kStStages = 5;
constexpr int kStConsumers = 256;  // 8 consumer warps; warp 8 is the producer
constexpr int kStThreads = kStConsumers + 32;
constexpr int kStStageDoubles = (kStR + 1) * kStCH;
constexpr size_t kStSmemBytes = (size_t)kStStages * kStStageDoubles * sizeof(double) + 1024;

struct ChunkRef {
    int job, rb, op, ch;       // row set, row block, operand, chunk index inside the operand
};

__device__ __forceinline__ ChunkRef decode_chunk(const MatvecArgs& a, long long g)
{
    ChunkRef c;
    c.job = g >= a.job[1].chunk_begin && a.job[1].nrb > 0 ? 1 : 0;
    const MatvecJob& j = a.job[c.job];
    const long long loc = g - j.chunk_begin;
    c.rb = (int)(loc / j.chunks_per_rb);
    int rem = (int)(loc - (long long)c.rb * j.chunks_per_rb);
    c.op = rem < j.nch[0] ? 0 : 1;
    c.ch = c.op ? rem - j.nch[0] : rem;
    return c;
}

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

__device__ __forceinline__ void consumer_sync()
{
    asm volatile("bar.sync 1, %0;" ::"n"(kStConsumers) : "memory");
}

__global__ void __launch_bounds__(kStThreads, 1)
matvec_stream_kernel(const __grid_constant__ MatvecArgs args)
{
    extern __shared__ __align__(128) double smem[];
    __shared__ __align__(8) uint64_t full_bar[kStStages];
    __shared__ __align__(8) uint64_t empty_bar[kStStages];
    __shared__ double red[kStConsumers / 32][kStR];
    __shared__ int is_last;

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
            mbar_init(&empty_bar[s], kStConsumers / 32);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == kStConsumers / 32) {
        // ------------------------------------------------------------------ producer warp
        if (lane == 0) {
            // The matrix does not depend on this evaluation's forcing vector: the first ring of matrix pieces is
            // requested BEFORE waiting for the peers' publication, so the flag latency hides behind HBM traffic.
            int stage = 0;
            unsigned phase = 0;
            auto issue = [&](long long g, int stg, bool matrix, bool vector, size_t par) {
                const ChunkRef c = decode_chunk(args, g);
                const MatvecJob& j = args.job[c.job];
                const MatOperand& op = j.op[c.op];
                const int c0 = c.ch * kStCH;
                const int ncol = min(kStCH, (int)op.ld - c0);
                const unsigned bytes = (unsigned)(ncol * sizeof(double));
                double* dst = smem + (size_t)stg * kStStageDoubles;
                if (matrix) {
                    mbar_arrive_expect_tx(&full_bar[stg], bytes * (kStR + 1));
#pragma unroll
                    for (int r = 0; r < kStR; ++r) {
                        const int row = min(c.rb * kStR + r, j.nrows - 1);
                        tma_load_1d(dst + r * kStCH, op.G + (size_t)row * op.ld + c0, bytes, &full_bar[stg]);
                    }
                }
                if (vector) tma_load_1d(dst + kStR * kStCH, op.x + par * op.x_stride + c0, bytes, &full_bar[stg]);
            };
            long long g = g_begin;
            const long long g_pre = min(g_end, g_begin + kStStages);
            for (; g < g_pre; ++g) issue(g, (int)(g - g_begin), true, false, 0);
            size_t par = 0;
            if (args.pw.epochs) {
                const unsigned long long ep = *(volatile unsigned long long*)(args.pw.epochs + kEpForcing);
                if (args.pw.world > 1) {
                    wait_peers(args.pw.flags, args.pw.world, args.pw.rank, ep, args.pw.epochs + kEpError);
                    fence_proxy_async();              // peer stores -> async-proxy (TMA) reads
                }
                par = (size_t)((ep - 1ull) & 1ull);
            }
            for (long long h = g_begin; h < g_pre; ++h) issue(h, (int)(h - g_begin), false, true, par);
            stage = (int)((g_pre - g_begin) % kStStages);
            phase = (g_pre - g_begin) >= kStStages ? 1u : 0u;
            for (; g < g_end; ++g) {
                mbar_wait(&empty_bar[stage], phase ^ 1u);
                issue(g, stage, true, true, par);
                if (++stage == kStStages) { stage = 0; phase ^= 1u; }
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumer warps
    double acc[kStR];
#pragma unroll
    for (int r = 0; r < kStR; ++r) acc[r] = 0.0;
    int cur_job = -1, cur_rb = -1;
    int stage = 0;
    unsigned phase = 0;

    auto finalize = [&](int jb, int rb) {
        const MatvecJob& j = args.job[jb];
#pragma unroll
        for (int r = 0; r < kStR; ++r) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], off);
        }
        if (lane == 0) {
#pragma unroll
            for (int r = 0; r < kStR; ++r) red[warp][r] = acc[r];
        }
        consumer_sync();
        const int myrow = rb * kStR + tid;
        const bool active = tid < kStR && myrow < j.nrows;
        double mine = 0.0;
        if (active) {
#pragma unroll
            for (int w = 0; w < kStConsumers / 32; ++w) mine += red[w][tid];
        }
        // which CTAs contribute to this row block?
        const long long rb_g0 = j.chunk_begin + (long long)rb * j.chunks_per_rb;
        const long long rb_g1 = rb_g0 + j.chunks_per_rb;
        const int first = owner_of(total, grid, rb_g0), last = owner_of(total, grid, rb_g1 - 1);
        const int ncontrib = last - first + 1;
        bool do_epilogue = true;
        if (ncontrib > 1) {
            const int slot = (int)blockIdx.x - first;
            if (active) j.partial[((size_t)myrow) * j.slots + slot] = mine;
            __threadfence();
            consumer_sync();
            if (tid == 0) {
                const unsigned prev = atomicAdd(&j.counters[rb], 1u);
                is_last = (prev == (unsigned)ncontrib - 1u);
                if (is_last) j.counters[rb] = 0u;                 // re-arm for the next evaluation
            }
            consumer_sync();
            do_epilogue = is_last != 0;
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
            if (j.epilogue == kEpiFault) update_fault_row(args.fe, myrow, mine);
            else j.yout[myrow] = mine;
        }
        consumer_sync();                                           // red[] / is_last are reused
#pragma unroll
        for (int r = 0; r < kStR; ++r) acc[r] = 0.0;
    };

    for (long long g = g_begin; g < g_end; ++g) {
        const ChunkRef c = decode_chunk(args, g);
        if (c.job != cur_job || c.rb != cur_rb) {
            if (cur_rb >= 0) finalize(cur_job, cur_rb);
            cur_job = c.job; cur_rb = c.rb;
        }
        const MatOperand& op = args.job[c.job].op[c.op];
        const int ncol = min(kStCH, (int)op.ld - c.ch * kStCH);
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
    }
    if (cur_rb >= 0) finalize(cur_job, cur_rb);
}

}  // namespace oq
