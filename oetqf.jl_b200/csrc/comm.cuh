// comm.cuh -- multi-GPU exchange over peer-mapped device memory (one process per GPU).
//
// Every rank owns one "window" allocation that its peers map through CUDA IPC.  A rank's pointwise
// forcing kernel stores its slice of (v - vpl) and (dϵ - dϵ0) straight into every rank's window over
// NVLink (the all-gather the RHS needs is fused into the producing kernel), then publishes an epoch
// number; consumers spin on their local flags with acquire loads.  The step error norm is summed the
// same way (one double per rank per step, combined in rank order so every rank takes the same decision).
#pragma once
#include <cstddef>
#include <cstdint>

struct OqProblem;

namespace oq {

constexpr int kMaxWorld = 16;

// local device-side counters (unsigned long long each)
enum : int {
    kEpForcing = 0,     // number of forcing publications so far (parity selects the buffer copy)
    kEpReduce = 1,      // number of error-norm publications so far
    kEpBlocksF = 2,     // block-done counter of the forcing kernel
    kEpBlocksR = 3,     // block-done counter of the error-norm kernel
    kEpError = 4,       // sticky error flag (peer wait timed out)
    kEpCount = 8
};

// Offsets (in doubles) inside a window; identical on every rank (they depend on global sizes only).
struct WindowLayout {
    size_t relv_len = 0, reldeps_len = 0;
    size_t off_relv = 0, off_reldeps = 0, off_red = 0, off_flags = 0, off_epochs = 0, total = 0;
};

// Pointers a producer kernel needs to publish into every rank's window.
struct PeerTargets {
    int world = 1, rank = 0;
    double* base[kMaxWorld] = {};          // window base of each rank in this process's address space
};

struct PeerWindow {
    PeerTargets t;
    void* ipc_base[kMaxWorld] = {};        // what cudaIpcOpenMemHandle returned (allocation base)
    bool opened[kMaxWorld] = {};
};

int comm_alloc_window(OqProblem* p);
void comm_release(OqProblem* p);
PeerTargets comm_targets(const OqProblem* p);

#ifdef __CUDACC__
// spin until every peer's flag reaches `epoch` (bounded: sets the sticky error flag after ~4 s)
__device__ __forceinline__ void wait_peers(const unsigned long long* flags, int world, int self,
                                           unsigned long long epoch, unsigned long long* err)
{
    unsigned long long t0 = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int r = 0; r < world; ++r) {
        if (r == self) continue;
        for (;;) {
            unsigned long long v;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + r) : "memory");
            if (v >= epoch) break;
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 4000000000ull) { atomicExch(err, 1ull); return; }
        }
    }
}

// Flag store to a peer.  The caller issues ONE system-scope fence before the loop over peers (a release
// store per peer would pay the fence once per peer, ~microseconds each over NVLink).
__device__ __forceinline__ void publish_flag(unsigned long long* remote_flag, unsigned long long epoch)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(remote_flag), "l"(epoch) : "memory");
}
#endif

}  // namespace oq
