// stream.cuh -- the all-cells sweep as a stream of TMA-fed tile packets ("compiled" schedule).
//
// The generic kernels of kernels.cuh walk the CSR grid for every task and pay one grid barrier
// plus four dependent memory round trips per wavefront level.  The all-cells single sweep (the
// one that dominates a step) has static level sets, so they are compiled once into a form the
// hardware can stream:
//
//   * slot s = position of task (c, dl) in the level-sorted task list; inside a level the tasks
//     are ordered by (cell, direction), so the directions of one cell that share a wavefront
//     level are neighbours ("segment").
//   * the flux state lives in slot order: out_slot[s] = outgoing_total_rate of that task
//     (src/sweep/site.rs:15); every level writes one contiguous range and gathers from ranges of
//     earlier levels (the previous level's range is still in L2).
//   * a level is cut into tiles of <= THREADS slots at segment boundaries.  Tile g is processed
//     by block g % n_blocks; all static data of a tile -- per slot: cell index and entry
//     offset, per upwind face ("entry"): source slot and the geometric share
//     A_rev (-n.d) / sum_downwind(A n.d) of the donor (src/sweep/mod.rs:453-461) -- is packed
//     into ONE contiguous packet, and the packets of a block are laid out back to back in the
//     order the block consumes them.  A block therefore reads one sequential byte stream, which
//     it prefetches STAGES tiles ahead with TMA bulk copies (cp.async.bulk + mbarrier
//     complete_tx) into a shared-memory ring, independent of the wavefront barriers.
//   * periodic upwind faces are ordinary entries flagged kPeriodicBit.  The donor of a periodic
//     face is normally solved in a later level than its target, so reading out_slot[donor]
//     directly yields last sweep's value: the reference's lag (src/sweep/mod.rs:505-513,
//     site.rs:53-56).  Donors in the same or an earlier level are
//     redirected to a snapshot slot behind the task slots that is refreshed before each sweep.
//   * per-cell photon rate: sum_d (incoming[d] + source / D) (src/sweep/mod.rs:554-558) is
//     accumulated per segment through shared memory and added to rate_cell[c] by the one thread
//     that owns the segment: no global atomics, deterministic order.  Different levels are
//     ordered by the level barrier, and a cell has at most one segment per level.
//   * level barrier: one counter per wavefront level, release-add by the blocks that own a tile
//     of the level, acquire-poll by the first tile of the next level.
//
// Algorithmic bytes per cell-direction update as streamed from HBM: 12 B per entry (4 B slot +
// 8 B share), 6 B per task of packet (cell 4 + entry offset 2), 8 B outgoing store; the 8 B
// gathers per entry, the 16-B cell record and the rate accumulator are L2 traffic in the
// steady state.  DESIGN.md section 5 compares this with B_alg = 20 F_up + 24.
#pragma once
#include <cub/cub.cuh>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "kernels.cuh"

namespace ssw {

constexpr uint32_t kPeriodicBit = 0x80000000u;
constexpr int kStreamThreads = 512;
constexpr int kMaxStages = 4;

struct TileDesc {       // 16 B, one per tile, stored per block in consumption order
    uint32_t slot0;     // first slot of the tile
    uint32_t off16;     // packet offset inside the block's stream, in 16-B units
    uint16_t n;         // slots in the tile (<= kStreamThreads)
    uint16_t n_entries; // upwind entries in the tile
    uint32_t level;     // wavefront level
};

struct TileLayout {
    uint32_t w, src, cell, eoff, bytes;
};
__host__ __device__ inline uint32_t align16(uint32_t x) { return (x + 15u) & ~15u; }
// packet = [f64 share[E]] [u32 source slot[E]] [u32 cell[n]] [u16 entry offset[n + 1]], each padded to 16 B
__host__ __device__ inline TileLayout tile_layout(uint32_t n, uint32_t E) {
    TileLayout L;
    uint32_t o = 0;
    L.w = o;    o += align16(8u * E);
    L.src = o;  o += align16(4u * E);
    L.cell = o; o += align16(4u * n);
    L.eoff = o; o += align16(2u * (n + 1u));
    L.bytes = o;
    return L;
}

struct Compiled {
    bool valid = false;
    uint64_t n_tasks = 0;
    uint64_t n_entries = 0;
    uint32_t n_levels = 0;
    uint32_t n_tiles = 0;
    uint32_t n_lag = 0;             // snapshot slots behind the task slots
    uint32_t n_blocks = 0, stages = 0, stage_bytes = 0;
    uint64_t stream_bytes = 0;
    double mean_entries = 0.0;
    uint32_t *slot_of = nullptr;    // [dl*N + c] -> slot
    double *out_slot = nullptr;     // n_tasks + n_lag
    double *ttot_slot = nullptr;    // n_tasks, sum over downwind faces of A * n.d
    uint32_t *lag_src = nullptr;    // n_lag donor slots
    unsigned char *stream = nullptr;
    TileDesc *tab = nullptr;        // n_tiles, block-major
    uint32_t *tab_off = nullptr;    // n_blocks + 1
    uint64_t *stream_off = nullptr; // n_blocks
    uint32_t *lvl_target = nullptr; // n_levels: blocks that own a tile of the level
    unsigned int *lvl_count = nullptr; // n_levels arrival counters
    Compiled() = default;
    Compiled(const Compiled &) = delete;
    Compiled &operator=(const Compiled &) = delete;
    ~Compiled() { release(); }
    void release() {
        cudaFree(slot_of); cudaFree(out_slot); cudaFree(ttot_slot); cudaFree(lag_src); cudaFree(stream);
        cudaFree(tab); cudaFree(tab_off); cudaFree(stream_off); cudaFree(lvl_target); cudaFree(lvl_count);
        slot_of = lag_src = tab_off = lvl_target = nullptr;
        out_slot = ttot_slot = nullptr;
        stream = nullptr;
        tab = nullptr;
        stream_off = nullptr;
        lvl_count = nullptr;
        valid = false;
    }
};

inline bool compiled_supported() { return true; }

// ---- construction kernels (run once per compiled schedule) -----------------------------------

// task id (dl * N + c) -> cell-major sort key (c * Dl + dl)
__global__ void __launch_bounds__(256)
s_key_kernel(const uint32_t *__restrict__ tasks, uint32_t n, uint32_t n_cells, uint32_t n_dl,
             uint32_t *__restrict__ keys) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint32_t t = tasks[s];
    const uint32_t dl = t / n_cells, c = t - dl * n_cells;
    keys[s] = c * n_dl + dl;
}

__global__ void __launch_bounds__(256)
s_slot_scatter_kernel(const uint32_t *__restrict__ keys, uint32_t n, uint32_t n_cells, uint32_t n_dl,
                      uint32_t *__restrict__ slot_of) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint32_t k = keys[s];
    const uint32_t c = k / n_dl, dl = k - c * n_dl;
    slot_of[(size_t)dl * n_cells + c] = s;
}

__device__ __forceinline__ uint32_t level_end_of_slot(const uint32_t *__restrict__ level_off, uint32_t n_levels,
                                                      uint32_t s) {
    // smallest level_off[l + 1] > s
    uint32_t lo = 0, hi = n_levels;  // invariant: level_off[lo] <= s < level_off[hi]
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (level_off[mid] <= s) lo = mid;
        else hi = mid;
    }
    return level_off[hi];
}

// entries per slot (Local and LocalPeriodic upwind faces), total downwind effective area,
// number of periodic entries whose donor is not in a later level (need a snapshot slot)
__global__ void __launch_bounds__(256)
s_count_kernel(GridView g, const uint32_t *__restrict__ keys, uint32_t n, uint32_t n_dl,
               const uint32_t *__restrict__ slot_of, const uint32_t *__restrict__ level_off, uint32_t n_levels,
               uint32_t *__restrict__ cnt, double *__restrict__ ttot_slot, unsigned int *__restrict__ n_lag) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint32_t k = keys[s];
    const uint32_t c = k / n_dl, dl = k - c * n_dl;
    const double dx = c_dirs[3 * dl], dy = c_dirs[3 * dl + 1], dz = c_dirs[3 * dl + 2];
    uint32_t m = 0, lag = 0, lvl_end = 0;
    double ttot = 0.0;
    for (uint32_t f = g.face_off[c]; f < g.face_off[c + 1]; ++f) {
        const double4 geo = ld_geo(g.face_geo + f);
        const double d = dot_dir(geo, dx, dy, dz);
        const int kind = g.face_kind[f];
        if (d < 0.0) {
            if (kind == 0) ++m;
            else if (kind == 2) {
                ++m;
                if (lvl_end == 0) lvl_end = level_end_of_slot(level_off, n_levels, s);
                if (slot_of[(size_t)dl * g.n_cells + (uint32_t)g.face_nb[f]] < lvl_end) ++lag;
            }
        } else if (d > 0.0) {
            ttot += geo.w * d;
        }
    }
    cnt[s] = m;
    ttot_slot[s] = ttot;
    if (lag) atomicAdd(n_lag, lag);
}

// Greedy tile cutting, one warp per wavefront level: a tile is the longest run of <= max_slots
// slots that ends at a segment (cell) boundary.  mode 0 counts, mode 1 writes the tile starts.
__global__ void __launch_bounds__(128)
s_cut_kernel(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ level_off, uint32_t n_levels,
             uint32_t n_dl, uint32_t max_slots, const uint32_t *__restrict__ tile_off,
             uint32_t *__restrict__ tile_cnt, uint32_t *__restrict__ tile_start) {
    const uint32_t l = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (l >= n_levels) return;
    const uint32_t begin = level_off[l], end = level_off[l + 1];
    uint32_t start = begin, count = 0;
    uint32_t out = tile_start ? tile_off[l] : 0;
    while (start < end) {
        if (tile_start && lane == 0) tile_start[out + count] = start;
        ++count;
        uint32_t pos = start + max_slots;   // candidate end (exclusive)
        if (pos >= end) break;
        // largest p in (start, pos] with cell(p) != cell(p - 1); segments are <= n_dl <= 128 long
        uint32_t best = 0;
        for (uint32_t j = lane; j < 160 && j < pos - start; j += 32) {
            const uint32_t p = pos - j;
            if (keys[p] / n_dl != keys[p - 1] / n_dl) best = max(best, p);
        }
        for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
        start = best ? best : pos;   // best == 0 cannot happen (n_dl <= 128 < 160 <= max_slots)
    }
    if (!tile_start && lane == 0) tile_cnt[l] = count;
}

// entry offsets at the tile starts (+ the total at index m)
__global__ void __launch_bounds__(256)
s_gather_offsets_kernel(const unsigned long long *__restrict__ upoff, const uint32_t *__restrict__ idx, uint32_t m,
                        uint32_t n_last, unsigned long long *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) out[i] = upoff[idx[i]];
    else if (i == m) out[i] = upoff[n_last];
}

struct FillArgs {
    GridView g;
    const uint32_t *keys;
    const uint32_t *slot_of;
    const unsigned long long *upoff;   // n + 1
    const double *ttot_slot;
    const uint32_t *level_off;
    uint32_t n_levels, n_dl, n_tasks;
    const TileDesc *tab;          // block-major
    const uint32_t *tab_block;    // block of tile i (block-major index)
    const uint64_t *stream_off;   // per block
    unsigned char *stream;
    uint32_t *lag_src;
    unsigned int *lag_counter;
};

// one thread block per tile: writes the tile's packet
__global__ void __launch_bounds__(kStreamThreads)
s_fill_kernel(FillArgs a) {
    const TileDesc d = a.tab[blockIdx.x];
    unsigned char *pkt = a.stream + a.stream_off[a.tab_block[blockIdx.x]] + (size_t)d.off16 * 16u;
    const TileLayout L = tile_layout(d.n, d.n_entries);
    double *w = reinterpret_cast<double *>(pkt + L.w);
    uint32_t *es = reinterpret_cast<uint32_t *>(pkt + L.src);
    uint32_t *cell = reinterpret_cast<uint32_t *>(pkt + L.cell);
    uint16_t *eoff = reinterpret_cast<uint16_t *>(pkt + L.eoff);
    const unsigned long long e_base = a.upoff[d.slot0];
    const uint32_t tid = threadIdx.x;
    if (tid == 0) eoff[d.n] = (uint16_t)d.n_entries;
    // zero the padding so the stream is fully initialised
    if (tid < 4) {
        if (tid == 0 && (d.n_entries & 1u)) w[d.n_entries] = 0.0;
        const uint32_t pad_src = (align16(4u * d.n_entries) - 4u * d.n_entries) / 4u;
        if (tid < pad_src) es[d.n_entries + tid] = 0u;
        const uint32_t pad_cell = (align16(4u * d.n) - 4u * d.n) / 4u;
        if (tid < pad_cell) cell[d.n + tid] = 0xffffffffu;
    }
    if (tid < 8) {
        const uint32_t used = d.n + 1u, pad = (align16(2u * used) - 2u * used) / 2u;
        if (tid < pad) eoff[used + tid] = 0;
    }
    if (tid >= d.n) return;
    const uint32_t s = d.slot0 + tid;
    const uint32_t k = a.keys[s];
    const uint32_t c = k / a.n_dl, dl = k - c * a.n_dl;
    cell[tid] = c;
    uint32_t e = (uint32_t)(a.upoff[s] - e_base);
    eoff[tid] = (uint16_t)e;
    const double dx = c_dirs[3 * dl], dy = c_dirs[3 * dl + 1], dz = c_dirs[3 * dl + 2];
    uint32_t lvl_end = 0;
    for (uint32_t f = a.g.face_off[c]; f < a.g.face_off[c + 1]; ++f) {
        const int kind = a.g.face_kind[f];
        if (kind != 0 && kind != 2) continue;
        const double dd = dot_dir(ld_geo(a.g.face_geo + f), dx, dy, dz);
        if (!(dd < 0.0)) continue;
        uint32_t src = a.slot_of[(size_t)dl * a.g.n_cells + (uint32_t)a.g.face_nb[f]];
        const double tt = a.ttot_slot[src];
        const double share = tt > 0.0 ? (a.g.face_rev[f] * (-dd)) / tt : 0.0;
        if (kind == 2) {
            if (lvl_end == 0) lvl_end = level_end_of_slot(a.level_off, a.n_levels, s);
            if (src < lvl_end) {   // donor not in a later level: read its pre-sweep snapshot
                const unsigned int j = atomicAdd(a.lag_counter, 1u);
                a.lag_src[j] = src;
                src = a.n_tasks + j;
            }
            src |= kPeriodicBit;
        }
        es[e] = src;
        w[e] = share;
        ++e;
    }
}

__global__ void __launch_bounds__(256)
s_convert_state_kernel(const uint32_t *__restrict__ keys, uint32_t n, uint32_t n_cells, uint32_t n_dl,
                       const double *__restrict__ q_nat, const double *__restrict__ ttot_slot,
                       double *__restrict__ out_slot) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint32_t k = keys[s];
    const uint32_t c = k / n_dl, dl = k - c * n_dl;
    out_slot[s] = q_nat[(size_t)dl * n_cells + c] * ttot_slot[s];
}

__global__ void __launch_bounds__(256)
s_lag_snapshot_kernel(const uint32_t *__restrict__ lag_src, uint32_t n_lag, uint32_t n_tasks,
                      double *__restrict__ out_slot) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n_lag) out_slot[n_tasks + j] = out_slot[lag_src[j]];
}

// ---- PTX helpers: mbarrier + TMA bulk copy ------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tma_bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar,
                                              uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
__device__ __forceinline__ void red_release_gpu(unsigned int *p) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

struct StreamArgs {
    const unsigned char *stream;
    const uint64_t *stream_off;
    const TileDesc *tab;
    const uint32_t *tab_off;
    const uint32_t *lvl_target;
    unsigned int *lvl_count;
    double *out_slot;
    const double2 *cellrec;   // {exp(-n_HI sigma size), source / D} per cell
    double *rate_cell;        // per cell sum over this rank's directions of incoming + source / D
    double threshold;
    uint32_t stages, stage_bytes;
    int solve;                // 1: sweep; 0: only accumulate sum_d incoming (photon_rate read-out)
};

// Persistent kernel: block b consumes the tiles tab[tab_off[b] .. tab_off[b+1]) in order.
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
sweep_stream_kernel(StreamArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *ring = smem;
    double *inc_s = reinterpret_cast<double *>(smem + (size_t)a.stages * a.stage_bytes);
    uint64_t *full = reinterpret_cast<uint64_t *>(inc_s + THREADS);
    const uint32_t tid = threadIdx.x;
    const uint32_t t0 = a.tab_off[blockIdx.x], n_my = a.tab_off[blockIdx.x + 1] - t0;
    const TileDesc *tab = a.tab + t0;
    const unsigned char *stream = a.stream + a.stream_off[blockIdx.x];
    uint64_t policy = 0;
    if (tid == 0) {
        for (uint32_t s = 0; s < a.stages; ++s) mbar_init(smem_u32(full + s), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        policy = policy_evict_first();
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t pre = min(a.stages, n_my);
        for (uint32_t k = 0; k < pre; ++k) {
            const TileDesc d = tab[k];
            const uint32_t bytes = tile_layout(d.n, d.n_entries).bytes;
            mbar_expect_tx(smem_u32(full + k), bytes);
            tma_bulk_load(smem_u32(ring + (size_t)k * a.stage_bytes), stream + (size_t)d.off16 * 16u, bytes,
                          smem_u32(full + k), policy);
        }
    }
    uint32_t prev_level = 0xffffffffu;
    uint32_t stage = 0, parity = 0;
    for (uint32_t k = 0; k < n_my; ++k) {
        const TileDesc d = tab[k];
        if (d.level != prev_level) {
            // all stores of the previous level were issued before the tile-end barrier below
            if (tid == 0) {
                if (prev_level != 0xffffffffu) red_release_gpu(a.lvl_count + prev_level);
                if (d.level > 0) {
                    const unsigned int target = a.lvl_target[d.level - 1];
                    while (ld_acquire_gpu(a.lvl_count + d.level - 1) < target) {}
                }
            }
            __syncthreads();
            prev_level = d.level;
        }
        mbar_wait(smem_u32(full + stage), parity);
        const unsigned char *pkt = ring + (size_t)stage * a.stage_bytes;
        const TileLayout L = tile_layout(d.n, d.n_entries);
        const double *w = reinterpret_cast<const double *>(pkt + L.w);
        const uint32_t *es = reinterpret_cast<const uint32_t *>(pkt + L.src);
        const uint32_t *cell = reinterpret_cast<const uint32_t *>(pkt + L.cell);
        const uint16_t *eoff = reinterpret_cast<const uint16_t *>(pkt + L.eoff);
        uint32_t c = 0xffffffffu;
        double inc = 0.0;
        if (tid < d.n) {
            c = cell[tid];
            uint32_t e = eoff[tid];
            const uint32_t e1 = eoff[tid + 1];
            const double2 rec = __ldg(a.cellrec + c);
            double in_loc = 0.0, in_per = 0.0;
            // gathers of four entries are issued together (memory-level parallelism), the sums
            // run in face order like the reference's accumulation
            for (; e + 4 <= e1; e += 4) {
                const uint32_t s0 = es[e], s1 = es[e + 1], s2 = es[e + 2], s3 = es[e + 3];
                const double v0 = __ldcg(a.out_slot + (s0 & ~kPeriodicBit));
                const double v1 = __ldcg(a.out_slot + (s1 & ~kPeriodicBit));
                const double v2 = __ldcg(a.out_slot + (s2 & ~kPeriodicBit));
                const double v3 = __ldcg(a.out_slot + (s3 & ~kPeriodicBit));
                const double p0 = v0 * w[e], p1 = v1 * w[e + 1], p2 = v2 * w[e + 2], p3 = v3 * w[e + 3];
                if (s0 & kPeriodicBit) in_per += p0; else in_loc += p0;
                if (s1 & kPeriodicBit) in_per += p1; else in_loc += p1;
                if (s2 & kPeriodicBit) in_per += p2; else in_loc += p2;
                if (s3 & kPeriodicBit) in_per += p3; else in_loc += p3;
            }
            {
                const uint32_t r = e1 - e;   // 0..3 remaining
                const uint32_t s0 = r > 0 ? es[e] : 0u, s1 = r > 1 ? es[e + 1] : 0u, s2 = r > 2 ? es[e + 2] : 0u;
                const double v0 = r > 0 ? __ldcg(a.out_slot + (s0 & ~kPeriodicBit)) : 0.0;
                const double v1 = r > 1 ? __ldcg(a.out_slot + (s1 & ~kPeriodicBit)) : 0.0;
                const double v2 = r > 2 ? __ldcg(a.out_slot + (s2 & ~kPeriodicBit)) : 0.0;
                if (r > 0) { const double p = v0 * w[e]; if (s0 & kPeriodicBit) in_per += p; else in_loc += p; }
                if (r > 1) { const double p = v1 * w[e + 1]; if (s1 & kPeriodicBit) in_per += p; else in_loc += p; }
                if (r > 2) { const double p = v2 * w[e + 2]; if (s2 & kPeriodicBit) in_per += p; else in_loc += p; }
            }
            if (a.solve) {
                inc = in_loc + rec.y;                                   // site.rs:49-56
                const double total = inc + in_per;
                // HydrogenOnly::get_outgoing_rate, hydrogen_only/mod.rs:81-87
                const double out = (total < a.threshold) ? 0.0 : total * rec.x;
                __stcg(a.out_slot + d.slot0 + tid, out);
            } else {
                inc = in_loc;                                           // photon_rate, mod.rs:727-730
            }
        }
        inc_s[tid] = inc;
        __syncthreads();
        // segment heads fold their segment in direction order and own the cell's accumulator
        if (tid < d.n && (tid == 0 || cell[tid - 1] != c)) {
            double sum = inc;
            for (uint32_t j = tid + 1; j < d.n && cell[j] == c; ++j) sum += inc_s[j];
            __stcg(a.rate_cell + c, __ldcg(a.rate_cell + c) + sum);
        }
        __syncthreads();   // the stage and inc_s are free again
        if (tid == 0 && k + a.stages < n_my) {
            const TileDesc nd = tab[k + a.stages];
            const uint32_t bytes = tile_layout(nd.n, nd.n_entries).bytes;
            mbar_expect_tx(smem_u32(full + stage), bytes);
            tma_bulk_load(smem_u32(ring + (size_t)stage * a.stage_bytes), stream + (size_t)nd.off16 * 16u, bytes,
                          smem_u32(full + stage), policy);
        }
        if (++stage == a.stages) { stage = 0; parity ^= 1u; }
    }
    // the last level of this block: nobody waits for the final level, but keep the counters
    // complete so that a later launch (after the memset) and debugging see a consistent state
    if (tid == 0 && prev_level != 0xffffffffu) red_release_gpu(a.lvl_count + prev_level);
}

// ---- host side --------------------------------------------------------------------------------------
struct StreamLaunchConfig {
    uint32_t blocks_per_sm, stages, stage_bytes, smem_bytes;
};

inline size_t stream_smem_bytes(uint32_t stages, uint32_t stage_bytes) {
    return (size_t)stages * stage_bytes + sizeof(double) * kStreamThreads + sizeof(uint64_t) * kMaxStages;
}

inline void cuda_ok(cudaError_t e, const char *what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

// Builds the compiled schedule from the level-sorted task list of the all-cells sweep.
//   tasks          level-sorted task ids (dl * N + c)
//   level_off_*    n_levels + 1 offsets into tasks
//   q_nat          flux state in the natural layout (out / ttot), converted into out_slot
inline void compile_schedule(Compiled &C, const GridView &g, const uint32_t *tasks, const uint32_t *level_off_dev,
                             const std::vector<uint32_t> &level_off_host, uint64_t n_tasks, uint32_t n_levels,
                             int n_local_dirs, const double *q_nat, int num_sms, cudaStream_t stream,
                             uint64_t *launch_counter) {
    C.release();
    if (n_tasks >= 0x7fffff00ull) throw std::runtime_error("compile_schedule: more than 2^31 tasks per rank");
    if (n_local_dirs > 128) throw std::runtime_error("compile_schedule: more than 128 local directions");
    const uint32_t n = (uint32_t)n_tasks;
    const uint32_t n_dl = (uint32_t)n_local_dirs;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    uint32_t *keys_in = nullptr, *keys = nullptr, *cnt = nullptr, *tile_cnt = nullptr, *tile_off = nullptr,
             *tile_start = nullptr, *tab_block = nullptr;
    unsigned long long *upoff = nullptr;
    unsigned int *counters = nullptr;   // [0] n_lag (count pass), [1] lag fill cursor
    void *temp = nullptr;
    auto cleanup = [&]() {
        cudaFree(keys_in); cudaFree(keys); cudaFree(cnt); cudaFree(tile_cnt); cudaFree(tile_off);
        cudaFree(tile_start); cudaFree(tab_block); cudaFree(upoff); cudaFree(counters); cudaFree(temp);
    };
    try {
        // 1. order every level by (cell, direction)
        cuda_ok(cudaMalloc(&keys_in, sizeof(uint32_t) * (size_t)n), "malloc keys_in");
        cuda_ok(cudaMalloc(&keys, sizeof(uint32_t) * (size_t)n), "malloc keys");
        s_key_kernel<<<blocks, 256, 0, stream>>>(tasks, n, g.n_cells, n_dl, keys_in);
        size_t bytes = 0;
        cuda_ok(cub::DeviceSegmentedSort::SortKeys(nullptr, bytes, keys_in, keys, (int64_t)n, (int64_t)n_levels,
                                                   level_off_dev, level_off_dev + 1, stream), "segmented sort size");
        cuda_ok(cudaMalloc(&temp, bytes), "malloc sort temp");
        cuda_ok(cub::DeviceSegmentedSort::SortKeys(temp, bytes, keys_in, keys, (int64_t)n, (int64_t)n_levels,
                                                   level_off_dev, level_off_dev + 1, stream), "segmented sort");
        cuda_ok(cudaStreamSynchronize(stream), "segmented sort sync");
        cudaFree(temp); temp = nullptr;
        cudaFree(keys_in); keys_in = nullptr;

        // 2. slots, entry counts, downwind areas
        cuda_ok(cudaMalloc(&C.slot_of, sizeof(uint32_t) * (size_t)n), "malloc slot_of");
        cuda_ok(cudaMalloc(&C.ttot_slot, sizeof(double) * (size_t)n), "malloc ttot_slot");
        cuda_ok(cudaMalloc(&cnt, sizeof(uint32_t) * ((size_t)n + 1)), "malloc cnt");
        cuda_ok(cudaMalloc(&upoff, sizeof(unsigned long long) * ((size_t)n + 1)), "malloc upoff");
        cuda_ok(cudaMalloc(&counters, 2 * sizeof(unsigned int)), "malloc counters");
        cuda_ok(cudaMemsetAsync(counters, 0, 2 * sizeof(unsigned int), stream), "memset");
        cuda_ok(cudaMemsetAsync(cnt + n, 0, sizeof(uint32_t), stream), "memset");
        s_slot_scatter_kernel<<<blocks, 256, 0, stream>>>(keys, n, g.n_cells, n_dl, C.slot_of);
        s_count_kernel<<<blocks, 256, 0, stream>>>(g, keys, n, n_dl, C.slot_of, level_off_dev, n_levels, cnt,
                                                   C.ttot_slot, counters);
        cuda_ok(cub::DeviceScan::ExclusiveSum(nullptr, bytes, cnt, upoff, (int64_t)n + 1, stream), "scan size");
        cuda_ok(cudaMalloc(&temp, bytes), "malloc scan temp");
        cuda_ok(cub::DeviceScan::ExclusiveSum(temp, bytes, cnt, upoff, (int64_t)n + 1, stream), "scan");
        unsigned long long total_entries = 0;
        unsigned int n_lag = 0;
        cuda_ok(cudaMemcpyAsync(&total_entries, upoff + n, sizeof total_entries, cudaMemcpyDeviceToHost, stream), "copy");
        cuda_ok(cudaMemcpyAsync(&n_lag, counters, sizeof n_lag, cudaMemcpyDeviceToHost, stream), "copy");
        cuda_ok(cudaStreamSynchronize(stream), "count sync");
        cudaFree(temp); temp = nullptr;
        cudaFree(cnt); cnt = nullptr;
        C.n_entries = total_entries;
        C.n_lag = n_lag;
        if ((uint64_t)n + n_lag >= 0x7fffff00ull) throw std::runtime_error("compile_schedule: slot index overflow");

        // 3. tiles: greedy cut at segment boundaries, one warp per level
        cuda_ok(cudaMalloc(&tile_cnt, sizeof(uint32_t) * (size_t)n_levels), "malloc tile_cnt");
        cuda_ok(cudaMalloc(&tile_off, sizeof(uint32_t) * ((size_t)n_levels + 1)), "malloc tile_off");
        const unsigned cut_blocks = (unsigned)(((size_t)n_levels * 32 + 127) / 128);
        s_cut_kernel<<<cut_blocks, 128, 0, stream>>>(keys, level_off_dev, n_levels, n_dl, kStreamThreads, nullptr,
                                                     tile_cnt, nullptr);
        std::vector<uint32_t> tcnt(n_levels), toff(n_levels + 1, 0);
        cuda_ok(cudaMemcpyAsync(tcnt.data(), tile_cnt, sizeof(uint32_t) * n_levels, cudaMemcpyDeviceToHost, stream), "copy");
        cuda_ok(cudaStreamSynchronize(stream), "cut sync");
        for (uint32_t l = 0; l < n_levels; ++l) toff[l + 1] = toff[l] + tcnt[l];
        const uint32_t n_tiles = toff[n_levels];
        cuda_ok(cudaMemcpyAsync(tile_off, toff.data(), sizeof(uint32_t) * (n_levels + 1), cudaMemcpyHostToDevice, stream), "copy");
        cuda_ok(cudaMalloc(&tile_start, sizeof(uint32_t) * ((size_t)n_tiles + 1)), "malloc tile_start");
        s_cut_kernel<<<cut_blocks, 128, 0, stream>>>(keys, level_off_dev, n_levels, n_dl, kStreamThreads, tile_off,
                                                     nullptr, tile_start);
        std::vector<uint32_t> tstart(n_tiles + 1);
        cuda_ok(cudaMemcpyAsync(tstart.data(), tile_start, sizeof(uint32_t) * n_tiles, cudaMemcpyDeviceToHost, stream), "copy");
        cuda_ok(cudaStreamSynchronize(stream), "cut sync 2");
        tstart[n_tiles] = n;
        // entry offsets at the tile starts
        std::vector<unsigned long long> tentry(n_tiles + 1);
        unsigned long long *tentry_dev = nullptr;
        cuda_ok(cudaMalloc(&tentry_dev, sizeof(unsigned long long) * ((size_t)n_tiles + 1)), "malloc tentry");
        s_gather_offsets_kernel<<<(n_tiles + 1 + 255) / 256, 256, 0, stream>>>(upoff, tile_start, n_tiles, n, tentry_dev);
        cudaError_t ce = cudaMemcpyAsync(tentry.data(), tentry_dev, sizeof(unsigned long long) * ((size_t)n_tiles + 1),
                                         cudaMemcpyDeviceToHost, stream);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(stream);
        cudaFree(tentry_dev);
        cuda_ok(ce, "tile entry gather");

        // 4. launch geometry: stage size = largest packet; blocks per SM and stages from the smem budget
        uint32_t max_bytes = 0;
        std::vector<uint32_t> tlevel(n_tiles);
        for (uint32_t l = 0; l < n_levels; ++l)
            for (uint32_t t = toff[l]; t < toff[l + 1]; ++t) tlevel[t] = l;
        for (uint32_t t = 0; t < n_tiles; ++t) {
            const uint32_t ns = tstart[t + 1] - tstart[t];
            const unsigned long long E = tentry[t + 1] - tentry[t];
            if (ns > (uint32_t)kStreamThreads || E > 65535ull)
                throw std::runtime_error("compile_schedule: tile too large (more than 65535 upwind entries in 512 tasks)");
            max_bytes = std::max(max_bytes, tile_layout(ns, (uint32_t)E).bytes);
        }
        const uint32_t stage_bytes = (max_bytes + 127u) & ~127u;
        const size_t smem_cap = 227u * 1024u;
        uint32_t blocks_per_sm = 2, stages = 0;
        for (; blocks_per_sm >= 1; --blocks_per_sm) {
            const size_t per_block = smem_cap / blocks_per_sm - 1024;   // 1 KB reserved per block by the driver
            const size_t fixed = stream_smem_bytes(0, 0);
            if (per_block < fixed + stage_bytes) { if (blocks_per_sm == 1) break; continue; }
            stages = (uint32_t)std::min<size_t>(kMaxStages, (per_block - fixed) / stage_bytes);
            if (stages >= 2 || blocks_per_sm == 1) break;
        }
        if (stages < 1) throw std::runtime_error("compile_schedule: a tile packet does not fit in shared memory");
        C.stages = stages;
        C.stage_bytes = stage_bytes;
        const size_t smem = stream_smem_bytes(stages, stage_bytes);
        cuda_ok(cudaFuncSetAttribute(sweep_stream_kernel<kStreamThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem), "set max dynamic smem");
        int per_sm = 0;
        cuda_ok(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sweep_stream_kernel<kStreamThreads>,
                                                              kStreamThreads, smem), "occupancy");
        if (per_sm < 1) throw std::runtime_error("compile_schedule: stream kernel does not fit on an SM");
        C.n_blocks = (uint32_t)std::min<int>(per_sm, (int)blocks_per_sm) * (uint32_t)num_sms;
        const uint32_t nb = C.n_blocks;

        // 5. tile table (block-major), per-block streams, per-level barrier targets
        std::vector<uint32_t> tab_off(nb + 1, 0);
        for (uint32_t b = 0; b < nb; ++b) tab_off[b + 1] = tab_off[b] + (n_tiles > b ? (n_tiles - b + nb - 1) / nb : 0);
        std::vector<TileDesc> tab(n_tiles);
        std::vector<uint32_t> tab_block_h(n_tiles);
        std::vector<uint64_t> stream_off(nb + 1, 0);
        std::vector<uint64_t> cursor(nb, 0);
        for (uint32_t t = 0; t < n_tiles; ++t) {
            const uint32_t b = t % nb, k = t / nb;
            TileDesc d;
            d.slot0 = tstart[t];
            d.n = (uint16_t)(tstart[t + 1] - tstart[t]);
            d.n_entries = (uint16_t)(tentry[t + 1] - tentry[t]);
            d.level = tlevel[t];
            if ((cursor[b] >> 4) > 0xffffffffull) throw std::runtime_error("compile_schedule: block stream exceeds 64 GB");
            d.off16 = (uint32_t)(cursor[b] >> 4);
            cursor[b] += tile_layout(d.n, d.n_entries).bytes;
            tab[tab_off[b] + k] = d;
            tab_block_h[tab_off[b] + k] = b;
        }
        for (uint32_t b = 0; b < nb; ++b) stream_off[b + 1] = stream_off[b] + ((cursor[b] + 127u) & ~(uint64_t)127u);
        C.stream_bytes = stream_off[nb];
        std::vector<uint32_t> lvl_target(n_levels);
        for (uint32_t l = 0; l < n_levels; ++l) lvl_target[l] = std::min<uint32_t>(tcnt[l], nb);
        C.n_tiles = n_tiles;

        cuda_ok(cudaMalloc(&C.stream, std::max<uint64_t>(C.stream_bytes, 16)), "malloc stream");
        cuda_ok(cudaMalloc(&C.tab, sizeof(TileDesc) * (size_t)n_tiles), "malloc tab");
        cuda_ok(cudaMalloc(&tab_block, sizeof(uint32_t) * (size_t)n_tiles), "malloc tab_block");
        cuda_ok(cudaMalloc(&C.tab_off, sizeof(uint32_t) * ((size_t)nb + 1)), "malloc tab_off");
        cuda_ok(cudaMalloc(&C.stream_off, sizeof(uint64_t) * ((size_t)nb + 1)), "malloc stream_off");
        cuda_ok(cudaMalloc(&C.lvl_target, sizeof(uint32_t) * (size_t)n_levels), "malloc lvl_target");
        cuda_ok(cudaMalloc(&C.lvl_count, sizeof(unsigned int) * (size_t)n_levels), "malloc lvl_count");
        cuda_ok(cudaMalloc(&C.out_slot, sizeof(double) * ((size_t)n + n_lag)), "malloc out_slot");
        cuda_ok(cudaMalloc(&C.lag_src, sizeof(uint32_t) * (size_t)std::max<uint32_t>(n_lag, 1)), "malloc lag_src");
        cuda_ok(cudaMemcpyAsync(C.tab, tab.data(), sizeof(TileDesc) * (size_t)n_tiles, cudaMemcpyHostToDevice, stream), "copy tab");
        cuda_ok(cudaMemcpyAsync(tab_block, tab_block_h.data(), sizeof(uint32_t) * (size_t)n_tiles, cudaMemcpyHostToDevice, stream), "copy");
        cuda_ok(cudaMemcpyAsync(C.tab_off, tab_off.data(), sizeof(uint32_t) * ((size_t)nb + 1), cudaMemcpyHostToDevice, stream), "copy");
        cuda_ok(cudaMemcpyAsync(C.stream_off, stream_off.data(), sizeof(uint64_t) * ((size_t)nb + 1), cudaMemcpyHostToDevice, stream), "copy");
        cuda_ok(cudaMemcpyAsync(C.lvl_target, lvl_target.data(), sizeof(uint32_t) * (size_t)n_levels, cudaMemcpyHostToDevice, stream), "copy");

        // 6. packets and state
        FillArgs fa;
        fa.g = g; fa.keys = keys; fa.slot_of = C.slot_of; fa.upoff = upoff; fa.ttot_slot = C.ttot_slot;
        fa.level_off = level_off_dev; fa.n_levels = n_levels; fa.n_dl = n_dl; fa.n_tasks = n;
        fa.tab = C.tab; fa.tab_block = tab_block; fa.stream_off = C.stream_off; fa.stream = C.stream;
        fa.lag_src = C.lag_src; fa.lag_counter = counters + 1;
        s_fill_kernel<<<n_tiles, kStreamThreads, 0, stream>>>(fa);
        s_convert_state_kernel<<<blocks, 256, 0, stream>>>(keys, n, g.n_cells, n_dl, q_nat, C.ttot_slot, C.out_slot);
        cuda_ok(cudaGetLastError(), "compile kernels");
        unsigned int lag_filled = 0;
        cuda_ok(cudaMemcpyAsync(&lag_filled, counters + 1, sizeof lag_filled, cudaMemcpyDeviceToHost, stream), "copy");
        cuda_ok(cudaStreamSynchronize(stream), "compile sync");   // host vectors go out of scope
        if (lag_filled != n_lag) throw std::runtime_error("compile_schedule: periodic snapshot count mismatch");
        if (launch_counter) *launch_counter += 11;
        (void)level_off_host;
    } catch (...) {
        cleanup();
        C.release();
        throw;
    }
    cleanup();
    C.n_tasks = n_tasks;
    C.n_levels = n_levels;
    C.mean_entries = n_tasks ? (double)C.n_entries / (double)n_tasks : 0.0;
    C.valid = true;
}

// One all-cells sweep (solve = 1) or one evaluation of sum_d incoming (solve = 0) over the compiled schedule.
inline void run_compiled(Compiled &C, const double2 *cellrec, double *rate_cell, uint32_t n_cells, double threshold,
                         int solve, cudaStream_t stream, uint64_t *launch_counter) {
    StreamArgs a;
    a.stream = C.stream;
    a.stream_off = C.stream_off;
    a.tab = C.tab;
    a.tab_off = C.tab_off;
    a.lvl_target = C.lvl_target;
    a.lvl_count = C.lvl_count;
    a.out_slot = C.out_slot;
    a.cellrec = cellrec;
    a.rate_cell = rate_cell;
    a.threshold = threshold;
    a.stages = C.stages;
    a.stage_bytes = C.stage_bytes;
    a.solve = solve;
    cuda_ok(cudaMemsetAsync(C.lvl_count, 0, sizeof(unsigned int) * (size_t)C.n_levels, stream), "memset lvl_count");
    cuda_ok(cudaMemsetAsync(rate_cell, 0, sizeof(double) * (size_t)n_cells, stream), "memset rate_cell");
    uint64_t launches = 1;
    if (solve && C.n_lag) {
        s_lag_snapshot_kernel<<<(C.n_lag + 255) / 256, 256, 0, stream>>>(C.lag_src, C.n_lag, (uint32_t)C.n_tasks, C.out_slot);
        ++launches;
    }
    cuda_ok(cudaFuncSetAttribute(sweep_stream_kernel<kStreamThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)stream_smem_bytes(C.stages, C.stage_bytes)), "set max dynamic smem");
    void *args[] = {&a};
    // cooperative launch only to guarantee co-residency of all blocks (the level barrier spins)
    cuda_ok(cudaLaunchCooperativeKernel((const void *)sweep_stream_kernel<kStreamThreads>, dim3(C.n_blocks),
                                        dim3(kStreamThreads), args, stream_smem_bytes(C.stages, C.stage_bytes), stream),
            "sweep_stream_kernel launch");
    if (launch_counter) *launch_counter += launches;
}

}  // namespace ssw
