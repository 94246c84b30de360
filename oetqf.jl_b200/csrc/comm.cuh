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
    kEpBlocksF = 16,    // block-done counter of the forcing front end  (own 128-byte line: every CTA adds to it while
    kEpBlocksR = 17,    // block-done counter of the error-norm kernel   the producers of the grid poll kEpForcing)
    kEpError = 4,       // sticky error flag (peer wait timed out); cleared by oq_solve / oq_rhs* on entry
    kEpTimeoutNs = 5,   // how long a kernel waits for a peer before giving up (0: for ever); OQ_PEER_TIMEOUT_S
    kEpHostErr = 6,     // device address of a page-locked host word that receives 1 on a timeout (0: none)
    kEpCount = 32
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

// who owns which columns of the two forcing vectors (row shards are contiguous in rank order): rank r owns fault
// cells [fb[r], fb[r+1]) and mantle elements [eb[r], eb[r+1])
struct ColOwners {
    int world = 1;
    int fb[kMaxWorld + 1] = {};
    int eb[kMaxWorld + 1] = {};
};

struct PeerWindow {
    PeerTargets t;
    ColOwners own;
    void* ipc_base[kMaxWorld] = {};        // what cudaIpcOpenMemHandle returned (allocation base)
    bool opened[kMaxWorld] = {};
};

int comm_alloc_window(OqProblem* p);
void comm_release(OqProblem* p);
PeerTargets comm_targets(const OqProblem* p);
ColOwners comm_owners(const OqProblem* p);
// 0 if no kernel of this problem gave up waiting for a peer since the last comm_clear_error
int comm_check_error(OqProblem* p, const char* where);
void comm_clear_error(OqProblem* p);

#ifdef __CUDACC__
// A peer did not deliver in time: raise the sticky device flag (the step controller turns it into done / retcode 3,
// so the device stops stepping) and the host-visible word (checked by the host after every synchronisation, so
// oq_rhs / oq_rhs_resident / oq_solve return an error instead of numbers computed from stale data).
__device__ __forceinline__ void raise_peer_timeout(unsigned long long* epochs)
{
    atomicExch(epochs + kEpError, 1ull);
    unsigned long long* host = reinterpret_cast<unsigned long long*>(epochs[kEpHostErr]);
    if (host) asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(host), "l"(1ull) : "memory");
}

// spin until *flag >= target.  sys: the flag is written by a peer GPU (acquire at system scope), else by another
// CTA of this GPU.  Bounded by epochs[kEpTimeoutNs]; returns false after raising the error flags.
__device__ __forceinline__ bool spin_until(const unsigned long long* flag, unsigned long long target, bool sys,
                                           unsigned long long* epochs)
{
    unsigned long long t0 = 0;
    const unsigned long long limit = epochs[kEpTimeoutNs];
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        unsigned long long v;
        if (sys) asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
        else asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
        if (v >= target) return true;
        __nanosleep(40);
        if (limit) {
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > limit) { raise_peer_timeout(epochs); return false; }
        }
    }
}

// spin until every peer's flag reaches `epoch` (err = epochs + kEpError, kept for the callers' convenience)
__device__ __forceinline__ void wait_peers(const unsigned long long* flags, int world, int self,
                                           unsigned long long epoch, unsigned long long* err)
{
    unsigned long long* epochs = err - kEpError;
    for (int r = 0; r < world; ++r) {
        if (r == self) continue;
        if (!spin_until(flags + r, epoch, true, epochs)) return;
    }
}

// the same wait executed by a whole warp: lane r polls the flag of peer r with relaxed loads (all peers in flight at
// once instead of one acquire round trip after the other) and confirms with ONE acquire load once the value is there
// (a system-scope fence here costs ~3 us on the critical path of every evaluation: measured, 49.2 -> 53.6 us at 8 GPUs)
__device__ __forceinline__ void wait_peers_warp(const unsigned long long* flags, int world, int self,
                                                unsigned long long epoch, unsigned long long* err, int lane)
{
    unsigned long long* epochs = err - kEpError;
    if (lane < world && lane != self) {
        unsigned long long t0 = 0;
        const unsigned long long limit = epochs[kEpTimeoutNs];
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (;;) {
            unsigned long long v;
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + lane) : "memory");
            if (v >= epoch) {
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + lane) : "memory");
                break;
            }
            if (limit) {
                unsigned long long t1;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > limit) { raise_peer_timeout(epochs); break; }
            }
        }
    }
    __syncwarp();
}

// Flag store to a peer.  The caller issues ONE system-scope fence before the loop over peers (a release
// store per peer would pay the fence once per peer, ~microseconds each over NVLink).
__device__ __forceinline__ void publish_flag(unsigned long long* remote_flag, unsigned long long epoch)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(remote_flag), "l"(epoch) : "memory");
}
#endif

}  // namespace oq
