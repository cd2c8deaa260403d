// peer.cuh -- direction sharding without a collective library: the ranks of one NVLink / NVSwitch box exchange
// through peer-mapped memory, fused into the kernels that produce and consume the data.
//
// Replaces the MPI flux messages of the reference (src/sweep/communicator.rs:59-95) for the single 8-GPU box of
// BASELINE.json's north_star.  Every rank owns one ARENA (one cudaMalloc) holding its per-cell state and its receive
// buffers; the arenas of all ranks are mapped into every rank (cudaIpc between processes, plain pointers between
// handles of one process).  Cells are dealt in contiguous slices of n_per = ceil(N / W) cells; a cell's chemistry state
// (x, T, timescales, previous rate, ionization time) is valid on its OWNER only, the absorption factor and the timestep
// level of every cell are valid everywhere.
//
//   rate partials   the rate kernels store a cell's partial sum over the local directions straight into the owner's
//                   receive buffer recv[source rank][cell - first] (remote stores over NVLink: a one-shot
//                   reduce-scatter fused into the producer);
//   chemistry       the owner folds the W partials in rank order (deterministic), updates the cell and stores the new
//                   absorption factor into the `att` array of EVERY rank (the all-gather, fused into the consumer);
//   levels          the owner derives the new timestep level and stores it into every rank's `level` array; the
//                   per-level histograms travel the same way;
//   read-back       the reading rank pulls the other owners' slices.
//
// Ordering between ranks: every rank has ONE arrival counter in its arena.  A signal is an atomic increment with
// release semantics at system scope of every rank's counter (red.release.sys over NVLink), issued behind a
// system-scope fence by the last thread block of the kernel that produced the data (peer_signal_tail) -- no extra
// launch -- or by a one-warp kernel.  The wait for synchronisation point e is "own counter >= e * W": one stream memory
// operation (cuStreamWaitValue32 -- no SM is held while waiting), a host poll where ranks share a device (tests), or a
// one-warp polling kernel where stream memory operations are unavailable.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace ssw {

constexpr int kMaxPeers = 16;
constexpr uint32_t kPeerFlagStride = 128;   // bytes reserved per flag word
constexpr uint32_t kPeerHistWords = 40;     // u64 per source rank: 32 level counts, [32] = cells whose level changed

struct PeerLayout {   // byte offsets into an arena; identical on every rank
    uint64_t flags, hist, recv, scratch, att, level, x, T, ts, tau, prev, ion, photon, bytes;
};

struct PeerTable {
    unsigned char *base[kMaxPeers];   // arena of every rank as mapped into this rank (base[rank] = own arena)
    int32_t world, rank;
    uint32_t n_per, n_cells;          // cells per rank slice, all cells
    PeerLayout L;
    __host__ __device__ double *f64(int p, uint64_t off) const { return reinterpret_cast<double *>(base[p] + off); }
    __host__ __device__ uint32_t first() const { return (uint32_t)rank * n_per < n_cells ? (uint32_t)rank * n_per : n_cells; }
    __host__ __device__ uint32_t n_own() const { const uint32_t f = first(); return n_cells - f < n_per ? n_cells - f : n_per; }
};

inline uint64_t peer_align(uint64_t x) { return (x + 255u) & ~(uint64_t)255u; }
inline PeerLayout peer_layout(uint32_t n_cells, int world) {
    const uint64_t n_per = ((uint64_t)n_cells + world - 1) / world;
    PeerLayout L;
    uint64_t o = 0;
    L.flags = o;   o += peer_align((uint64_t)kMaxPeers * kPeerFlagStride);
    L.hist = o;    o += peer_align((uint64_t)kMaxPeers * kPeerHistWords * 8u);
    L.recv = o;    o += peer_align(n_per * world * 8u);
    L.scratch = o; o += peer_align((uint64_t)n_cells * 8u);
    L.att = o;     o += peer_align((uint64_t)n_cells * 8u);
    L.level = o;   o += peer_align((uint64_t)n_cells);
    L.x = o;       o += peer_align((uint64_t)n_cells * 8u);
    L.T = o;       o += peer_align((uint64_t)n_cells * 8u);
    L.ts = o;      o += peer_align((uint64_t)n_cells * 8u);
    L.tau = o;     o += peer_align((uint64_t)n_cells * 8u);
    L.prev = o;    o += peer_align((uint64_t)n_cells * 8u);
    L.ion = o;     o += peer_align((uint64_t)n_cells * 8u);
    L.photon = o;  o += peer_align((uint64_t)n_cells * 8u);
    L.bytes = o;
    return L;
}

__device__ __forceinline__ void peer_arrive(const PeerTable &t, int p) {
    unsigned int *counter = reinterpret_cast<unsigned int *>(t.base[p] + t.L.flags);
    asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
}

// signal: "everything this rank queued on its stream before this kernel is visible" -> every rank's arrival counter
__global__ void __launch_bounds__(32)
peer_signal_kernel(PeerTable t) {
    const int p = threadIdx.x;
    if (p >= t.world) return;
    __threadfence_system();
    peer_arrive(t, p);
}

// The same signal from the tail of the kernel that produced the data (called by every thread of every block, at the
// end): each block fences its stores at system scope and checks in; the last one to do so signals all ranks.
// `blocks_done` is a zero-initialised device counter of the handle (reset here for the next kernel).
__device__ __forceinline__ void peer_signal_tail(const PeerTable &t, unsigned int *blocks_done) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();   // cumulative: orders the block's stores (seen through the barrier) before the check-in
        const unsigned int total = gridDim.x * gridDim.y * gridDim.z;
        if (atomicAdd(blocks_done, 1u) == total - 1u) {
            *blocks_done = 0u;
            __threadfence_system();
            for (int p = 0; p < t.world; ++p) peer_arrive(t, p);
        }
    }
}

// fallback wait (no stream memory operations): one thread polls the rank's own counter; bounded, a lost peer traps
__global__ void __launch_bounds__(32)
peer_wait_kernel(PeerTable t, uint32_t target) {
    if (threadIdx.x != 0) return;
    const unsigned int *counter = reinterpret_cast<const unsigned int *>(t.base[t.rank] + t.L.flags);
    const long long t0 = clock64();
    for (;;) {
        unsigned int v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        if ((int32_t)(v - target) >= 0) break;
        if (clock64() - t0 > 60000000000ll) __trap();   // ~30 s
        __nanosleep(200);
    }
}

// partial rates of the active cells -> the owners' receive buffers
__global__ void __launch_bounds__(256)
peer_push_rates_kernel(PeerTable t, const uint32_t *__restrict__ act_list, uint32_t n_act, const double *__restrict__ rate_act,
                       unsigned int *blocks_done) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n_act) {
        const uint32_t c = act_list ? act_list[k] : k;
        const uint32_t owner = c / t.n_per;
        t.f64((int)owner, t.L.recv)[(uint64_t)t.rank * t.n_per + (c - owner * t.n_per)] = rate_act[k];
    }
    peer_signal_tail(t, blocks_done);   // the reduce-scatter and its completion signal in one launch
}

// per-level histogram of this rank's slice -> every rank
__global__ void __launch_bounds__(64)
peer_hist_push_kernel(PeerTable t, const unsigned long long *__restrict__ hist, unsigned int *blocks_done) {
    const uint32_t i = threadIdx.x;
    if (i < 33) {
        const unsigned long long v = hist[i];
        for (int p = 0; p < t.world; ++p)
            reinterpret_cast<unsigned long long *>(t.base[p] + t.L.hist)[(uint64_t)t.rank * kPeerHistWords + i] = v;
    }
    peer_signal_tail(t, blocks_done);   // also publishes the levels the kernel before this one stored into every rank
}

// a replicated-per-cell value of this rank's slice -> every other rank (absorption factors after ssw_set_inputs)
__global__ void __launch_bounds__(256)
peer_broadcast_slice_kernel(PeerTable t, uint64_t off) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= t.n_own()) return;
    const uint32_t c = t.first() + k;
    const double v = t.f64(t.rank, off)[c];
    for (int p = 0; p < t.world; ++p)
        if (p != t.rank) t.f64(p, off)[c] = v;
}

// owner-only field: fetch the other owners' slices into this rank's copy (read-back)
__global__ void __launch_bounds__(256)
peer_pull_kernel(PeerTable t, uint64_t off) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= t.n_cells) return;
    const uint32_t owner = c / t.n_per;
    if ((int)owner == t.rank) return;
    t.f64(t.rank, off)[c] = t.f64((int)owner, off)[c];
}

// dst[i] = sum over ranks (in rank order) of their array at `off`
__global__ void __launch_bounds__(256)
peer_sum_kernel(PeerTable t, uint64_t off, double *__restrict__ dst, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int p = 0; p < t.world; ++p) s += t.f64(p, off)[i];
    dst[i] = s;
}

}  // namespace ssw
