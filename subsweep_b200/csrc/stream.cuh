// stream.cuh -- the all-cells sweep as a stream of TMA-fed tile packets ("compiled" schedule).
//
// The generic kernels of kernels.cuh walk the CSR grid for every task and pay one grid barrier
// plus four dependent memory round trips per wavefront level.  The all-cells single sweep (the
// one that dominates a step) has static level sets, so they are compiled once into a form the
// hardware can stream:
//
//   * the local directions are dealt into G direction groups (dl % G).  A group is an
//     independent wavefront pipeline with its own thread blocks (block b serves group b % G),
//     level barriers and rate accumulators: while one group waits at a level barrier the
//     co-resident blocks of the other groups keep the memory system busy.
//   * slot s = position of task (c, dl) in the task list sorted by (wavefront level, group, cell,
//     direction): the directions of one cell that share a (level, group) are neighbours
//     ("segment").  A (level, group) pair is a "pseudo-level" pl = level * G + group.
//   * the flux state lives in slot order: out_slot[s] = outgoing_total_rate of that task
//     (src/sweep/site.rs:15); every pseudo-level writes one contiguous range and gathers from
//     ranges of earlier levels (the previous level's range is still in L2).
//   * a pseudo-level is cut into tiles of <= THREADS slots at segment boundaries.  The tiles of a
//     group are dealt round-robin to the group's blocks; all static data of a tile -- per slot:
//     cell index and entry offset, per upwind face ("entry"): source slot and the geometric share
//     A_rev (-n.d) / sum_downwind(A n.d) of the donor (src/sweep/mod.rs:453-461) -- is packed
//     into ONE contiguous packet, and the packets of a block are laid out back to back in the
//     order the block consumes them.  A block therefore reads one sequential byte stream, which
//     it prefetches STAGES tiles ahead with TMA bulk copies (cp.async.bulk + mbarrier
//     complete_tx) into a shared-memory ring, independent of the wavefront barriers.
//   * periodic upwind faces are ordinary entries (listed behind a slot's Local ones).  The donor of a periodic
//     face is normally solved in a later level than its target, so reading out_slot[donor]
//     directly yields last sweep's value: the reference's lag (src/sweep/mod.rs:505-513,
//     site.rs:53-56).  Donors in the same or an earlier level are
//     redirected to a snapshot slot behind the task slots that is refreshed before each sweep.
//   * per-cell photon rate: sum_d incoming[d] (src/sweep/mod.rs:554-558) is reduced per segment
//     with warp shuffles (+ one shared-memory hop across warps) and added to the group's
//     accumulator acc_cell[g][c] by the one thread that owns the segment: no global atomics,
//     deterministic order.  Levels are ordered by the level barrier, a cell has at most one
//     segment per pseudo-level, and groups have separate accumulators.
//   * the periodic_source terms of the rate (the donors' NEW outgoing rates, site.rs:53-56) are
//     "epilogue" tiles: one extra pseudo-level per group behind its last wavefront level whose
//     tasks are the (periodic cell, direction) pairs; they only accumulate into acc_per[g][p].
//   * level barrier: one counter per pseudo-level, release-add by the blocks that own a tile of
//     it, acquire-poll by the first tile of the dependent pseudo-level.
//
// Bytes per cell-direction update as streamed from HBM: 12 B per entry (4 B slot + 8 B share),
// 6 B per task of packet (cell 4 + entry offset 2), 8 B outgoing store; the 8-B gathers per
// entry, the 16-B cell record and the rate accumulators are L2 traffic in the steady state.
// DESIGN.md section 5 compares this with B_alg = 20 F_up + 24.
#pragma once
#include <cub/cub.cuh>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

#include "kernels.cuh"

namespace ssw {

constexpr uint32_t kNoDep = 0xffffffffu;
constexpr int kMaxStreamThreads = 1024;
constexpr int kMaxStages = 8;
constexpr int kMaxGroups = 8;      // interleaved / dedicated groups; solo mode uses one group per local direction
constexpr int kGroupShift = 40;   // sort key = group << 40 | (cell * Dl + dl)

struct TileDesc {       // 20 B, one per tile, stored per block in consumption order
    uint32_t slot0;     // first slot of the tile
    uint32_t off16;     // packet offset inside the block's stream, in 16-B units
    uint16_t n;         // slots in the tile (<= THREADS)
    uint16_t n_entries; // upwind entries in the tile
    uint32_t level;     // pseudo-level
    uint32_t aux;       // walk form: entries gathered from global memory | epilogue tile << 16
    uint32_t need;      // walk form: index (in its block) of the first tile of the same pseudo-level
};

// first 32 bytes of a packet: what the consuming block needs to know about this tile and about
// the tile it will prefetch into the same ring stage next (no global loads on the critical path)
struct PacketHeader {
    uint32_t slot0;
    uint16_t n, n_entries;
    uint32_t level;
    uint32_t next_off16, next_bytes;   // packet of the tile `stages` positions later (bytes = 0: none)
    uint32_t dep, dep_target;          // pseudo-level to wait for (kNoDep: none) and its arrival count
    uint16_t scan_steps, group;        // shuffle steps the longest segment of the tile needs; direction group
};
static_assert(sizeof(PacketHeader) == 32, "packet header is 32 bytes");

// per-slot info word: entry offset in the tile | periodic entries << 16 | segment head << 31
constexpr uint32_t kInfoHead = 0x80000000u;

struct TileLayout {
    uint32_t w, src, cell, info, bytes;
};
__host__ __device__ inline uint32_t align16(uint32_t x) { return (x + 15u) & ~15u; }
// packet = [header 32 B] [f64 share[E]] [u32 source slot[E]] [u32 cell[n]] [u32 info[n + 1]], each padded to 16 B.
// The entries of a slot are its Local upwind faces in face order followed by its periodic ones.
__host__ __device__ inline TileLayout tile_layout(uint32_t n, uint32_t E) {
    TileLayout L;
    uint32_t o = (uint32_t)sizeof(PacketHeader);
    L.w = o;    o += align16(8u * E);
    L.src = o;  o += align16(4u * E);
    L.cell = o; o += align16(4u * n);
    L.info = o; o += align16(4u * (n + 1u));
    L.bytes = o;
    return L;
}

struct PDesc;   // patch.cuh

struct Compiled {
    bool valid = false;
    uint64_t n_tasks = 0;           // real tasks (slots with an outgoing rate)
    uint64_t n_epilogue = 0;        // (periodic cell, direction) pairs
    uint64_t n_entries = 0;
    uint32_t n_levels = 0;          // wavefront levels
    uint32_t n_groups = 1;
    uint32_t n_pl = 0;              // pseudo-levels: n_levels * G + G
    uint32_t n_tiles = 0;
    uint32_t n_lag = 0;             // snapshot slots behind the task slots
    uint32_t threads = 512, bps = 1, n_blocks = 0, stages = 0, stage_bytes = 0;
    bool solo = false;              // walk form (walk.cuh): one block per local direction, block barriers only
    uint32_t window = 0, walk_groups = 1; // walk form: slots of the shared-memory window, compute groups per block
    uint64_t n_near = 0, n_far = 0; // walk form: upwind entries read from the window / from global memory
    uint32_t n_cells = 0, n_periodic = 0;
    uint64_t stream_bytes = 0;
    double mean_entries = 0.0;
    uint32_t *slot_of = nullptr;    // [dl*N + c] -> slot
    double *out_slot = nullptr;     // n_tasks + n_lag
    double *ttot_slot = nullptr;    // n_tasks, sum over downwind faces of A * n.d
    uint32_t *lag_src = nullptr;    // n_lag donor slots
    unsigned char *stream = nullptr;
    TileDesc *tab = nullptr;        // n_tiles, block-major
    uint32_t *tab_off = nullptr;    // n_blocks + 1
    uint64_t *stream_off = nullptr; // n_blocks + 1
    uint32_t *lvl_target = nullptr; // n_pl: blocks that own a tile of the pseudo-level
    uint32_t *lvl_dep = nullptr;    // n_pl: pseudo-level that must be complete first (kNoDep: none)
    unsigned int *lvl_count = nullptr; // n_pl arrival counters
    double2 *rec_slot = nullptr;    // walk form: {absorption factor, source / D} per slot, refreshed every sweep
    uint32_t *cell_of_slot = nullptr; // walk form: n_tasks
    double *acc_cell = nullptr;     // G x N: sum_d incoming of the group's directions
    double *acc_per = nullptr;      // G x n_periodic: sum_d periodic_source (patch mode: one row)
    // patch-ordered form (patch.cuh): macro-tiles (patch, direction group) with point-to-point done flags
    bool patch_mode = false;
    bool accumulate = false;        // patch mode with phases (cyclic patch graph): rate rows are accumulated
    uint32_t n_phases = 1;
    uint32_t n_groups_per = 1;      // rows of acc_per
    uint32_t n_mt = 0, vmax = 0, pc_max = 0, smax = 0, epoch = 0;
    uint32_t kd = 0, n_patches = 0, patch_levels = 0;
    unsigned int *mt_flag = nullptr;   // n_mt: epoch of the sweep that last completed the macro-tile
    PDesc *ptab = nullptr;             // packet table, block-major
    uint32_t *per_off = nullptr, *per_src = nullptr;   // periodic upwind entries per periodic cell (CSR)
    double *per_w = nullptr;
    Compiled() = default;
    Compiled(const Compiled &) = delete;
    Compiled &operator=(const Compiled &) = delete;
    ~Compiled() { release(); }
    void release() {
        cudaFree(slot_of); cudaFree(out_slot); cudaFree(ttot_slot); cudaFree(lag_src); cudaFree(stream);
        cudaFree(tab); cudaFree(tab_off); cudaFree(stream_off); cudaFree(lvl_target); cudaFree(lvl_dep);
        cudaFree(rec_slot); cudaFree(cell_of_slot);
        rec_slot = nullptr; cell_of_slot = nullptr;
        cudaFree(lvl_count); cudaFree(acc_cell); cudaFree(acc_per); cudaFree(mt_flag); cudaFree(ptab); cudaFree(per_off); cudaFree(per_src); cudaFree(per_w);
        per_off = per_src = nullptr;
        per_w = nullptr;
        mt_flag = nullptr;
        ptab = nullptr;
        patch_mode = false;
        solo = false;
        accumulate = false;
        n_phases = 1;
        slot_of = lag_src = tab_off = lvl_target = lvl_dep = nullptr;
        out_slot = ttot_slot = acc_cell = acc_per = nullptr;
        stream = nullptr;
        tab = nullptr;
        stream_off = nullptr;
        lvl_count = nullptr;
        valid = false;
    }
};

inline bool compiled_supported() { return true; }

// thrown by compile_schedule when the walk form (walk.cuh) cannot hold the grid: the caller compiles the level-barrier
// stream instead
struct WalkUnsupported : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// ---- construction kernels (run once per compiled schedule) -----------------------------------

// task id (dl * N + c) -> sort key: group << 40 | (c * Dl + dl)
__global__ void __launch_bounds__(256)
s_key64_kernel(const uint32_t *__restrict__ tasks, uint32_t n, uint32_t n_cells, uint32_t n_dl, uint32_t n_groups,
               unsigned long long *__restrict__ keys) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint32_t t = tasks[s];
    const uint32_t dl = t / n_cells, c = t - dl * n_cells;
    keys[s] = ((unsigned long long)(dl % n_groups) << kGroupShift) | (unsigned long long)(c * n_dl + dl);
}

// epilogue candidates: i = p * Dl + dl; key = group << 40 | (c * Dl + dl) if the cell has a periodic
// upwind face for dl, else ~0 (sorted to the end)
__global__ void __launch_bounds__(256)
s_epilogue_key_kernel(GridView g, const uint32_t *__restrict__ pcells, uint32_t n_periodic, uint32_t n_dl,
                      uint32_t n_groups, unsigned long long *__restrict__ keys, bool degree_bits = false) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_periodic * n_dl) return;
    const uint32_t p = i / n_dl, dl = i - p * n_dl;
    const uint32_t c = pcells[p];
    const double dx = g.dirs[3 * dl], dy = g.dirs[3 * dl + 1], dz = g.dirs[3 * dl + 2];
    uint32_t deg = 0;
    for (uint32_t f = g.face_off[c]; f < g.face_off[c + 1]; ++f)
        if (g.face_kind[f] == 2 && dot_dir(ld_geo(g.face_geo + f), dx, dy, dz) < 0.0) ++deg;
    // walk form: slots of a pseudo-level are ordered by degree, descending (walk.cuh)
    const unsigned long long dbits = degree_bits ? ((unsigned long long)(63u - min(deg, 63u)) << 32) : 0ull;
    keys[i] = deg ? (((unsigned long long)(dl % n_groups) << kGroupShift) | dbits | (unsigned long long)(c * n_dl + dl))
                  : ~0ull;
}

__device__ __forceinline__ uint32_t lower_bound_u64(const unsigned long long *__restrict__ a, uint32_t lo, uint32_t hi,
                                                    unsigned long long v) {
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (a[mid] < v) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// pseudo-level offsets: pl_off[l * G + g] = first slot of (level l, group g); the epilogue
// pseudo-levels [L * G + g] follow behind the n real tasks; pl_off[(L + 1) * G] = n + m
__global__ void __launch_bounds__(256)
s_pl_off_kernel(const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ level_off, uint32_t n_levels,
                uint32_t n_groups, const unsigned long long *__restrict__ epi_keys, uint32_t n_epi_cand, uint32_t n,
                uint32_t *__restrict__ pl_off) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n_real = n_levels * n_groups;
    if (i < n_real) {
        const uint32_t l = i / n_groups, g = i - l * n_groups;
        pl_off[i] = lower_bound_u64(keys, level_off[l], level_off[l + 1], (unsigned long long)g << kGroupShift);
    } else if (i <= n_real + n_groups) {
        const uint32_t g = i - n_real;   // g == n_groups gives the end
        pl_off[i] = n + lower_bound_u64(epi_keys, 0, n_epi_cand, (unsigned long long)g << kGroupShift);
    }
}

__global__ void __launch_bounds__(256)
s_key32_kernel(const unsigned long long *__restrict__ keys64, uint32_t n, uint32_t *__restrict__ keys) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) keys[s] = (uint32_t)(keys64[s] & ((1ull << kGroupShift) - 1ull));
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

__device__ __forceinline__ uint32_t level_end_of_slot(const uint32_t *__restrict__ pl_off, uint32_t n_pl, uint32_t s) {
    // smallest pl_off[l + 1] > s
    uint32_t lo = 0, hi = n_pl;  // invariant: pl_off[lo] <= s < pl_off[hi]
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (pl_off[mid] <= s) lo = mid;
        else hi = mid;
    }
    return pl_off[hi];
}

// entries per slot, total downwind effective area, number of periodic entries whose donor is not
// in a later level (they need a snapshot slot).  Real slots (s < n) count Local and LocalPeriodic
// upwind faces, epilogue slots only the LocalPeriodic ones.
__global__ void __launch_bounds__(256)
s_count_kernel(GridView g, const uint32_t *__restrict__ keys, uint32_t n, uint32_t n_all, uint32_t n_dl,
               const uint32_t *__restrict__ slot_of, const uint32_t *__restrict__ pl_off, uint32_t n_pl,
               uint32_t *__restrict__ cnt, double *__restrict__ ttot_slot, unsigned int *__restrict__ n_lag) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_all) return;
    const uint32_t k = keys[s];
    const uint32_t c = k / n_dl, dl = k - c * n_dl;
    const double dx = g.dirs[3 * dl], dy = g.dirs[3 * dl + 1], dz = g.dirs[3 * dl + 2];
    const bool real = s < n;
    uint32_t m = 0, lag = 0, lvl_end = 0;
    double ttot = 0.0;
    for (uint32_t f = g.face_off[c]; f < g.face_off[c + 1]; ++f) {
        const double4 geo = ld_geo(g.face_geo + f);
        const double d = dot_dir(geo, dx, dy, dz);
        const int kind = g.face_kind[f];
        if (d < 0.0) {
            if (kind == 0) {
                if (real) ++m;
            } else if (kind == 2) {
                ++m;
                if (real) {
                    if (lvl_end == 0) lvl_end = level_end_of_slot(pl_off, n_pl, s);
                    if (slot_of[(size_t)dl * g.n_cells + (uint32_t)g.face_nb[f]] < lvl_end) ++lag;
                }
            }
        } else if (d > 0.0) {
            ttot += geo.w * d;
        }
    }
    cnt[s] = m;
    if (real) ttot_slot[s] = ttot;
    if (lag) atomicAdd(n_lag, lag);
}

// Greedy tile cutting, one warp per pseudo-level: a tile is the longest run of <= max_slots
// slots that ends at a segment (cell) boundary.  Without tile_start only counts.
__global__ void __launch_bounds__(128)
s_cut_kernel(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ pl_off, uint32_t n_pl, uint32_t n_dl,
             uint32_t max_slots_uniform, const uint32_t *__restrict__ max_slots_pl,
             const uint32_t *__restrict__ tile_off, uint32_t *__restrict__ tile_cnt,
             uint32_t *__restrict__ tile_start) {
    const uint32_t l = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (l >= n_pl) return;
    const uint32_t max_slots = max_slots_pl ? max_slots_pl[l] : max_slots_uniform;
    const uint32_t end = pl_off[l + 1];
    uint32_t start = pl_off[l], count = 0;
    const uint32_t out = tile_start ? tile_off[l] : 0;
    while (start < end) {
        if (tile_start && lane == 0) tile_start[out + count] = start;
        ++count;
        const uint32_t pos = start + max_slots;   // candidate end (exclusive)
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
    const int32_t *pidx;
    const unsigned long long *upoff;   // n_all + 1
    const double *ttot_slot;
    const uint32_t *pl_off;
    uint32_t n_pl, n_real_pl, n_dl, n_tasks, n_groups;
    const TileDesc *tab;          // block-major
    const uint32_t *tab_block;    // block of tile i (block-major index)
    const uint32_t *tab_off;      // per block
    const uint32_t *lvl_dep, *lvl_target;
    const uint64_t *stream_off;   // per block
    unsigned char *stream;
    uint32_t *lag_src;
    unsigned int *lag_counter;
    uint32_t stages;
};

// one thread block per tile: writes the tile's packet
__global__ void __launch_bounds__(kMaxStreamThreads)
s_fill_kernel(FillArgs a) {
    __shared__ uint32_t s_maxlen;
    const TileDesc d = a.tab[blockIdx.x];
    unsigned char *pkt = a.stream + a.stream_off[a.tab_block[blockIdx.x]] + (size_t)d.off16 * 16u;
    const TileLayout L = tile_layout(d.n, d.n_entries);
    double *w = reinterpret_cast<double *>(pkt + L.w);
    uint32_t *es = reinterpret_cast<uint32_t *>(pkt + L.src);
    uint32_t *cell = reinterpret_cast<uint32_t *>(pkt + L.cell);
    uint32_t *info = reinterpret_cast<uint32_t *>(pkt + L.info);
    const unsigned long long e_base = a.upoff[d.slot0];
    const uint32_t tid = threadIdx.x;
    const bool epilogue = d.level >= a.n_real_pl;
    if (tid == 0) s_maxlen = 1;
    __syncthreads();
    // zero the padding so the stream is fully initialised
    if (tid < 4) {
        if (tid == 0 && (d.n_entries & 1u)) w[d.n_entries] = 0.0;
        const uint32_t pad_src = (align16(4u * d.n_entries) - 4u * d.n_entries) / 4u;
        if (tid < pad_src) es[d.n_entries + tid] = 0u;
        const uint32_t pad_cell = (align16(4u * d.n) - 4u * d.n) / 4u;
        if (tid < pad_cell) cell[d.n + tid] = 0xffffffffu;
        const uint32_t used = d.n + 1u, pad_info = (align16(4u * used) - 4u * used) / 4u;
        if (tid < pad_info) info[used + tid] = 0u;
    }
    if (tid < d.n) {
        const uint32_t s = d.slot0 + tid;
        const uint32_t k = a.keys[s];
        const uint32_t c = k / a.n_dl, dl = k - c * a.n_dl;
        cell[tid] = epilogue ? (uint32_t)a.pidx[c] : c;   // epilogue tiles address acc_per by periodic row
        const uint32_t e0 = (uint32_t)(a.upoff[s] - e_base);
        uint32_t e = e0, n_per = 0;
        const double dx = a.g.dirs[3 * dl], dy = a.g.dirs[3 * dl + 1], dz = a.g.dirs[3 * dl + 2];
        uint32_t lvl_end = 0;
        for (int pass = epilogue ? 1 : 0; pass < 2; ++pass) {   // Local faces first, then the periodic ones
            for (uint32_t f = a.g.face_off[c]; f < a.g.face_off[c + 1]; ++f) {
                if (a.g.face_kind[f] != (pass ? 2 : 0)) continue;
                const double dd = dot_dir(ld_geo(a.g.face_geo + f), dx, dy, dz);
                if (!(dd < 0.0)) continue;
                uint32_t src = a.slot_of[(size_t)dl * a.g.n_cells + (uint32_t)a.g.face_nb[f]];
                const double tt = a.ttot_slot[src];
                const double share = tt > 0.0 ? (a.g.face_rev[f] * (-dd)) / tt : 0.0;
                if (pass) {
                    ++n_per;
                    if (!epilogue) {
                        if (lvl_end == 0) lvl_end = level_end_of_slot(a.pl_off, a.n_pl, s);
                        if (src < lvl_end) {   // donor not in a later level: read its pre-sweep snapshot
                            const unsigned int j = atomicAdd(a.lag_counter, 1u);
                            a.lag_src[j] = src;
                            src = a.n_tasks + j;
                        }
                    }
                }
                es[e] = src;
                w[e] = share;
                ++e;
            }
        }
        if (n_per > 255u) atomicAdd(a.lag_counter + 1, 1u);   // reported as an error by the host
        const bool head = tid == 0 || a.keys[s - 1] / a.n_dl != c;
        info[tid] = e0 | (min(n_per, 255u) << 16) | (head ? kInfoHead : 0u);
        if (head) {   // segment length -> shuffle steps
            uint32_t len = 1;
            while (tid + len < d.n && a.keys[s + len] / a.n_dl == c) ++len;
            atomicMax(&s_maxlen, len);
        }
    }
    __syncthreads();
    if (tid == 0) {
        info[d.n] = d.n_entries;
        PacketHeader h;
        h.slot0 = d.slot0; h.n = d.n; h.n_entries = d.n_entries; h.level = d.level;
        h.next_off16 = 0; h.next_bytes = 0;
        h.group = (uint16_t)(epilogue ? d.level - a.n_real_pl : d.level % a.n_groups);
        h.dep = a.lvl_dep[d.level];
        h.dep_target = h.dep != kNoDep ? a.lvl_target[h.dep] : 0u;
        uint32_t steps = 0;
        while ((1u << steps) < min(s_maxlen, 32u)) ++steps;
        h.scan_steps = (uint16_t)steps;
        const uint32_t nxt = blockIdx.x + a.stages;
        if (nxt < a.tab_off[a.tab_block[blockIdx.x] + 1]) {
            const TileDesc nd = a.tab[nxt];
            h.next_off16 = nd.off16;
            h.next_bytes = tile_layout(nd.n, nd.n_entries).bytes;
        }
        *reinterpret_cast<PacketHeader *>(pkt) = h;
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

// rate_act[c] = sum_d ((incoming[d] + source / D) + periodic_source[d]) over this rank's directions
// (src/sweep/mod.rs:554-558, site.rs:53-56) from the group accumulators of the compiled sweep
__global__ void __launch_bounds__(256)
s_rate_finish_kernel(uint32_t n_cells, uint32_t n_groups, uint32_t n_groups_per, uint32_t n_periodic, int n_local_dirs,
                     double n_dirs_total,
                     const double *__restrict__ acc_cell, const double *__restrict__ acc_per,
                     const int32_t *__restrict__ pidx, const double *__restrict__ src, double *__restrict__ rate_act,
                     double *__restrict__ photon) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    double in = 0.0;
    for (uint32_t g = 0; g < n_groups; ++g) in += acc_cell[(size_t)g * n_cells + c];
    if (photon) photon[c] = in;                              // photon_rate, mod.rs:727-730
    if (!rate_act) return;
    double rate = in;
    const int32_t p = pidx[c];
    if (p >= 0) {
        double per = 0.0;
        for (uint32_t g = 0; g < n_groups_per; ++g) per += acc_per[(size_t)g * n_periodic + p];
        rate += per;
    }
    rate_act[c] = rate + (src[c] / n_dirs_total) * (double)n_local_dirs;
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
__device__ __forceinline__ unsigned int ld_relaxed_gpu(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// level counters sit one per 32-byte sector so that pollers of different pseudo-levels do not share a sector
constexpr uint32_t kCountStride = 8;

struct StreamArgs {
    const unsigned char *stream;
    const uint64_t *stream_off;
    const TileDesc *tab;
    const uint32_t *tab_off;
    const uint32_t *lvl_target;
    const uint32_t *lvl_dep;
    unsigned int *lvl_count;
    double *out_slot;
    const double2 *cellrec;   // {exp(-n_HI sigma size), source / D} per cell
    double *acc_cell;         // G x N
    double *acc_per;          // G x n_periodic
    double threshold;
    uint32_t stages, stage_bytes;
    uint32_t n_groups, n_real_pl, n_cells, n_periodic;
    uint32_t poll_ns;         // back-off between polls of a level counter
    unsigned long long *prof; // optional per-block cycle counters {total, wait behind arrive, packet wait, tiles, dependency poll, arrive}
};

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace ssw
#include "walk.cuh"
namespace ssw {

// Persistent kernel: block b consumes the tiles tab[tab_off[b] .. tab_off[b+1]) in order.  Per tile:
//   phase 1 (entry-parallel)  prod[e] = out_slot[source slot[e]] * share[e], written over share[e] in the
//                             ring stage: every gather of the tile is independent and in flight at once
//   phase 2 (task-parallel)   a thread sums the products of its slot (Local faces in face order, then
//                             the periodic ones), applies the absorption and stores the outgoing rate;
//                             the per-cell rate is reduced over the segment with warp shuffles
//
// SOLO: the block owns one local direction (group = direction, grid = number of local directions).  Every
// dependency of a tile was then produced by this very block: a level change is the block barrier behind the
// previous tile, there are no level counters, the gathers may hit the SM's own L1 (plain loads of values the
// same SM stored), and every (direction, cell) rate term is written exactly once (plain store, no accumulator
// read, no segment reduction).
template <int THREADS, int MIN_BLOCKS, bool PROFILE, bool SOLO>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
sweep_stream_kernel(StreamArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int WARPS = THREADS / 32;
    __shared__ double s_wsum[2][WARPS];
    __shared__ uint32_t s_wfirst[2][WARPS], s_wlast[2][WARPS];
    unsigned char *const ring = smem;
    const uint32_t stages = a.stages, stage_bytes = a.stage_bytes;
    uint64_t *const full = reinterpret_cast<uint64_t *>(smem + (size_t)stages * stage_bytes);
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t n_my = a.tab_off[blockIdx.x + 1] - a.tab_off[blockIdx.x];
    const unsigned char *const stream = a.stream + a.stream_off[blockIdx.x];
    double *const out_slot = a.out_slot;
    const uint32_t n_real_pl = a.n_real_pl;
    const double threshold = a.threshold;
    uint64_t policy = 0;
    if (tid == 0) {
        for (uint32_t s = 0; s < stages; ++s) mbar_init(smem_u32(full + s), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        policy = policy_evict_first();
    }
    __syncthreads();
    if (tid == 0) {
        const TileDesc *tab = a.tab + a.tab_off[blockIdx.x];
        const uint32_t pre = min(stages, n_my);
        for (uint32_t k = 0; k < pre; ++k) {
            const TileDesc d = tab[k];
            const uint32_t bytes = tile_layout(d.n, d.n_entries).bytes;
            mbar_expect_tx(smem_u32(full + k), bytes);
            tma_bulk_load(smem_u32(ring + (size_t)k * stage_bytes), stream + (size_t)d.off16 * 16u, bytes,
                          smem_u32(full + k), policy);
        }
    }
    uint32_t prev_level = kNoDep;
    uint32_t stage = 0, parity = 0;
    long long t_begin = 0, t_bar = 0, t_pkt = 0, t_rel = 0, tp = 0;
    if (PROFILE && tid == 0) t_begin = clock64();
    for (uint32_t k = 0; k < n_my; ++k) {
        if (PROFILE && tid == 0) tp = clock64();
        mbar_wait(smem_u32(full + stage), parity);
        if (PROFILE && tid == 0) t_pkt += clock64() - tp;
        unsigned char *const pkt = ring + (size_t)stage * stage_bytes;
        const PacketHeader d = *reinterpret_cast<const PacketHeader *>(pkt);
        if (!SOLO && d.level != prev_level) {
            __syncthreads();   // every thread has issued all its stores of the previous pseudo-level
            // arrive (thread 0) and wait (thread 32) run side by side; the arrival must not wait for the
            // dependency, or two blocks could wait for each other
            if (tid == 0 && prev_level != kNoDep) {
                if (PROFILE) tp = clock64();
                red_release_gpu(a.lvl_count + (size_t)prev_level * kCountStride);
                if (PROFILE) t_rel += clock64() - tp;
            }
            if (tid == 32 && d.dep != kNoDep) {
                long long t0 = 0;
                if (PROFILE) t0 = clock64();
                const unsigned int *cnt = a.lvl_count + (size_t)d.dep * kCountStride;
                while (ld_acquire_gpu(cnt) < d.dep_target) __nanosleep(a.poll_ns);
                if (PROFILE) a.prof[6 * blockIdx.x + 4] += (unsigned long long)(clock64() - t0);
            }
            if (PROFILE && tid == 0) tp = clock64();
            __syncthreads();
            if (PROFILE && tid == 0) t_bar += clock64() - tp;
            prev_level = d.level;
        }
        const bool epilogue = d.level >= n_real_pl;
        double *const acc = epilogue ? a.acc_per + (size_t)d.group * a.n_periodic : a.acc_cell + (size_t)d.group * a.n_cells;
        const uint32_t n = d.n, E = d.n_entries;
        const TileLayout L = tile_layout(n, E);
        double *const prod = reinterpret_cast<double *>(pkt + L.w);
        const uint32_t *const es = reinterpret_cast<const uint32_t *>(pkt + L.src);
        const uint32_t *const cell = reinterpret_cast<const uint32_t *>(pkt + L.cell);
        const uint32_t *const info = reinterpret_cast<const uint32_t *>(pkt + L.info);

        // per-slot data whose latency hides behind phase 1
        uint32_t c = 0xffffffffu, inf = 0, e1 = 0;
        double acc_old = 0.0;
        double2 rec = make_double2(0.0, 0.0);
        if (tid < n) {
            c = cell[tid];
            inf = info[tid];
            e1 = info[tid + 1] & 0xffffu;
            if (!SOLO && (inf & kInfoHead)) acc_old = __ldcg(acc + c);
            if (!epilogue) rec = __ldg(a.cellrec + c);
        }
        if (PROFILE && SOLO && tid == 0) tp = clock64();
        // phase 1: eight independent gathers per thread and round (an unstructured tile holds ~8 entries per slot: one
        // L2 round trip for the whole tile instead of two)
        for (uint32_t i = tid; i < E; i += 8u * THREADS) {
            double v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t ij = i + (uint32_t)j * THREADS;
                v[j] = 0.0;
                if (ij < E) v[j] = SOLO ? out_slot[es[ij]] : __ldcg(out_slot + es[ij]);   // SOLO: the SM's own stores, L1 is coherent
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t ij = i + (uint32_t)j * THREADS;
                if (ij < E) prod[ij] = v[j] * prod[ij];
            }
        }
        __syncthreads();
        if (PROFILE && SOLO && tid == 0) { const long long now = clock64(); t_bar += now - tp; tp = now; }
        // phase 2
        double inc = 0.0;
        if (tid < n) {
            uint32_t e = inf & 0xffffu;
            const uint32_t em = e1 - ((inf >> 16) & 0xffu);
            double in_loc = 0.0, in_per = 0.0;
            // same left-to-right order as one addition per iteration; four loads in flight per round
#pragma unroll 1
            for (; e + 4u <= em; e += 4u) {
                const double p0 = prod[e], p1 = prod[e + 1], p2 = prod[e + 2], p3 = prod[e + 3];
                in_loc = (((in_loc + p0) + p1) + p2) + p3;
            }
#pragma unroll 1
            for (; e < em; ++e) in_loc += prod[e];
#pragma unroll 1
            for (; e < e1; ++e) in_per += prod[e];
            if (epilogue) {
                inc = in_per;                                           // periodic_source of this sweep
            } else {
                inc = in_loc;                                           // incoming_total_rate[d]
                const double total = (in_loc + rec.y) + in_per;         // site.rs:49-56
                // HydrogenOnly::get_outgoing_rate, hydrogen_only/mod.rs:81-87
                const double out = (total < threshold) ? 0.0 : total * rec.x;
                if (SOLO) out_slot[d.slot0 + tid] = out;
                else __stcg(out_slot + d.slot0 + tid, out);
            }
        }
        if (SOLO) {
            if (tid < n) __stcs(acc + c, inc);   // the one term (direction, cell): read once by s_rate_finish_kernel
            __syncthreads();   // all threads are done with the stage; the stores above are visible to the block
            if (PROFILE && tid == 0) t_rel += clock64() - tp;
            if (tid == 0 && d.next_bytes) {
                fence_proxy_async_smem();
                mbar_expect_tx(smem_u32(full + stage), d.next_bytes);
                tma_bulk_load(smem_u32(pkt), stream + (size_t)d.next_off16 * 16u, d.next_bytes, smem_u32(full + stage),
                              policy);
            }
            if (++stage == stages) { stage = 0; parity ^= 1u; }
            continue;
        }
        // segmented suffix sums over the lanes of a warp: a lane ends up with the sum of its segment
        // from itself to the segment's end inside the warp
        // (a lane's segment ends at the next head lane, at the first lane without a slot, or at the warp's end)
        const uint32_t stops = __ballot_sync(0xffffffffu, (inf & kInfoHead) != 0 || tid >= n);
        const uint32_t rest = lane == 31u ? 0u : (stops >> (lane + 1u));
        const uint32_t seg_end = rest ? lane + (uint32_t)__ffs((int)rest) : 32u;
        double sum = inc;
        for (uint32_t st = 0, o = 1; st < d.scan_steps; ++st, o <<= 1) {
            const double v = __shfl_down_sync(0xffffffffu, sum, o);
            if (lane + o < seg_end) sum += v;
        }
        const uint32_t c_last = __shfl_sync(0xffffffffu, c, 31);
        const uint32_t buf = k & 1u;
        if (lane == 0) {
            s_wsum[buf][warp] = sum;
            s_wfirst[buf][warp] = c;
            s_wlast[buf][warp] = c_last;
        }
        __syncthreads();   // all threads are done with the stage; warp partial sums are visible
        if (tid == 0 && d.next_bytes) {
            fence_proxy_async_smem();   // the products were written through the generic proxy
            mbar_expect_tx(smem_u32(full + stage), d.next_bytes);
            tma_bulk_load(smem_u32(pkt), stream + (size_t)d.next_off16 * 16u, d.next_bytes, smem_u32(full + stage),
                          policy);
        }
        if (inf & kInfoHead) {
            if (c_last == c) {   // the segment runs on into the following warps
                for (uint32_t ww = warp + 1; ww < (uint32_t)WARPS && s_wfirst[buf][ww] == c; ++ww) {
                    sum += s_wsum[buf][ww];
                    if (s_wlast[buf][ww] != c) break;
                }
            }
            __stcg(acc + c, acc_old + sum);
        }
        if (++stage == stages) { stage = 0; parity ^= 1u; }
    }
    __syncthreads();
    if (!SOLO && tid == 0 && prev_level != kNoDep) red_release_gpu(a.lvl_count + (size_t)prev_level * kCountStride);
    if (PROFILE && tid == 0) {
        a.prof[6 * blockIdx.x + 0] = (unsigned long long)(clock64() - t_begin);
        a.prof[6 * blockIdx.x + 1] = (unsigned long long)t_bar;
        a.prof[6 * blockIdx.x + 2] = (unsigned long long)t_pkt;
        a.prof[6 * blockIdx.x + 3] = n_my;
        a.prof[6 * blockIdx.x + 5] = (unsigned long long)t_rel;
    }
}

// ---- host side --------------------------------------------------------------------------------------
typedef void (*StreamKernel)(StreamArgs);

inline StreamKernel stream_kernel_for(uint32_t threads, uint32_t blocks_per_sm, bool profile = false, bool solo = false) {
    (void)solo;   // the solo form has its own kernel (walk.cuh)
    if (profile) return threads == 256 ? sweep_stream_kernel<256, 4, true, false> : sweep_stream_kernel<512, 2, true, false>;
    if (threads == 256) {
        if (blocks_per_sm >= 8) return sweep_stream_kernel<256, 8, false, false>;
        if (blocks_per_sm >= 6) return sweep_stream_kernel<256, 6, false, false>;
        return sweep_stream_kernel<256, 4, false, false>;
    }
    if (blocks_per_sm >= 4) return sweep_stream_kernel<512, 4, false, false>;
    if (blocks_per_sm >= 3) return sweep_stream_kernel<512, 3, false, false>;
    return sweep_stream_kernel<512, 2, false, false>;
}

inline size_t stream_smem_bytes(uint32_t stages, uint32_t stage_bytes) {
    return (size_t)stages * stage_bytes + sizeof(uint64_t) * kMaxStages;
}

inline void cuda_ok(cudaError_t e, const char *what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

// Opt-in shared memory per block (227 KB on sm_100).  The limit is a per-function attribute of the whole process: it is
// always raised to the maximum the function can have (227 KB minus its static shared memory), never to the size of one
// launch, so that handles on different threads cannot lower it under each other's launches.
inline void raise_smem_limit(const void *func) {
    cudaFuncAttributes fa;
    cuda_ok(cudaFuncGetAttributes(&fa, func), "function attributes");
    cuda_ok(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - (int)fa.sharedSizeBytes),
            "set max dynamic smem");
}

inline uint32_t env_u32(const char *name, uint32_t fallback) {
    const char *v = std::getenv(name);
    if (!v || !*v) return fallback;
    return (uint32_t)std::strtoul(v, nullptr, 10);
}

// Builds the compiled schedule from the level-sorted task list of the all-cells sweep.
//   tasks          level-sorted task ids (dl * N + c)
//   level_off_*    n_levels + 1 offsets into tasks
//   q_nat          flux state in the natural layout (out / ttot), converted into out_slot
inline void compile_schedule(Compiled &C, const GridView &g, const uint32_t *tasks, const uint32_t *level_off_dev,
                             uint64_t n_tasks, uint32_t n_levels, int n_local_dirs, const uint32_t *pcells,
                             uint32_t n_periodic, const int32_t *pidx, const double *q_nat, int num_sms,
                             cudaStream_t stream, uint64_t *launch_counter, bool allow_walk = false, uint32_t default_groups = 2) {
    C.release();
    if (n_tasks >= 0x7fffff00ull) throw std::runtime_error("compile_schedule: more than 2^31 tasks per rank");
    if (n_local_dirs > 128) throw std::runtime_error("compile_schedule: more than 128 local directions");
    const uint32_t n = (uint32_t)n_tasks;
    const uint32_t n_dl = (uint32_t)n_local_dirs;
    // tunables (environment overrides are for experiments; DESIGN.md section 5 lists the defaults)
    // few local directions (a direction shard of a multi-GPU job): the sweep is bound by the latency of its
    // level barriers, fewer and larger blocks with a single direction group cross them fastest (measured at
    // 10 and 21 directions); otherwise 256-thread blocks, 4 per SM, two interleaved groups
    const bool few_dirs = n_dl <= 24;
    // solo: one block per local direction walks that direction's whole wavefront; block barriers only
    const bool solo = allow_walk;
    uint32_t threads, tile_slots, G, want_bps, want_stages, window = 0, walk_groups = 1, walk_entries = 0;
    if (solo) {
        // compute groups x threads per group (= slots per tile); the pairs walk_kernel_for knows
        walk_groups = env_u32("SSW_WALK_GROUPS", 2);
        threads = env_u32("SSW_WALK_THREADS", 256);
        if (!walk_kernel_for(walk_groups, threads, false)) { walk_groups = 2; threads = 256; }
        tile_slots = threads;
        walk_entries = std::max<uint32_t>(env_u32("SSW_WALK_ENTRIES", 6u * threads), (uint32_t)kWalkMaxDeg);
        G = n_dl;
        want_bps = 1;
        want_stages = std::max<uint32_t>(2u, std::min<uint32_t>(env_u32("SSW_WALK_STAGES", kWalkMaxStages), kWalkMaxStages));
        window = std::min<uint32_t>(env_u32("SSW_WALK_WINDOW", 16384), 24576u) & ~1023u;
    } else {
        threads = env_u32("SSW_STREAM_THREADS", few_dirs ? 512 : 256) == 256 ? 256u : 512u;
        tile_slots = threads;
        G = std::min<uint32_t>(env_u32("SSW_STREAM_GROUPS", few_dirs ? 1 : default_groups), kMaxGroups);
        G = std::max<uint32_t>(1u, std::min<uint32_t>(G, n_dl));
        want_bps = env_u32("SSW_STREAM_BPS", threads == 256 ? 4 : 2);
        want_stages = std::max<uint32_t>(1u, std::min<uint32_t>(env_u32("SSW_STREAM_STAGES", 3), kMaxStages));
    }
    const uint32_t n_real_pl = n_levels * G, n_pl = n_real_pl + G;
    const uint32_t n_epi_cand = n_periodic * n_dl;
    const unsigned blocks = (unsigned)((n + 255) / 256);

    unsigned long long *keys64_in = nullptr, *keys64 = nullptr, *epi_in = nullptr, *epi_sorted = nullptr, *upoff = nullptr,
                       *tentry_dev = nullptr;
    uint32_t *keys = nullptr, *cnt = nullptr, *pl_off = nullptr, *tile_cnt = nullptr, *tile_off = nullptr,
             *tile_start = nullptr, *tab_block = nullptr, *pl_max_dev = nullptr;
    unsigned int *counters = nullptr;   // [0] n_lag (count pass), [1] lag fill cursor, [2] error flag
    void *temp = nullptr;
    auto cleanup = [&]() {
        cudaFree(keys64_in); cudaFree(keys64); cudaFree(epi_in); cudaFree(epi_sorted); cudaFree(upoff);
        cudaFree(tentry_dev); cudaFree(keys); cudaFree(cnt); cudaFree(pl_off); cudaFree(tile_cnt); cudaFree(tile_off);
        cudaFree(tile_start); cudaFree(tab_block); cudaFree(counters); cudaFree(temp); cudaFree(pl_max_dev);
    };
    try {
        size_t bytes = 0;
        // 1. order every level by (group, cell, direction)
        cuda_ok(cudaMalloc(&keys64_in, sizeof(unsigned long long) * (size_t)n), "malloc keys64_in");
        cuda_ok(cudaMalloc(&keys64, sizeof(unsigned long long) * (size_t)n), "malloc keys64");
        if (solo) w_key64_kernel<<<blocks, 256, 0, stream>>>(g, tasks, n, n_dl, keys64_in);
        else s_key64_kernel<<<blocks, 256, 0, stream>>>(tasks, n, g.n_cells, n_dl, G, keys64_in);
        cuda_ok(cub::DeviceSegmentedSort::SortKeys(nullptr, bytes, keys64_in, keys64, (int64_t)n, (int64_t)n_levels,
                                                   level_off_dev, level_off_dev + 1, stream), "segmented sort size");
        cuda_ok(cudaMalloc(&temp, bytes), "malloc sort temp");
        cuda_ok(cub::DeviceSegmentedSort::SortKeys(temp, bytes, keys64_in, keys64, (int64_t)n, (int64_t)n_levels,
                                                   level_off_dev, level_off_dev + 1, stream), "segmented sort");
        cuda_ok(cudaStreamSynchronize(stream), "segmented sort sync");
        cudaFree(temp); temp = nullptr;
        cudaFree(keys64_in); keys64_in = nullptr;

        // 2. epilogue tasks: (periodic cell, direction) pairs with a periodic upwind face
        cuda_ok(cudaMalloc(&epi_in, sizeof(unsigned long long) * (size_t)std::max<uint32_t>(n_epi_cand, 1)), "malloc epi");
        cuda_ok(cudaMalloc(&epi_sorted, sizeof(unsigned long long) * (size_t)std::max<uint32_t>(n_epi_cand, 1)), "malloc epi");
        if (n_epi_cand) {
            s_epilogue_key_kernel<<<(n_epi_cand + 255) / 256, 256, 0, stream>>>(g, pcells, n_periodic, n_dl, G, epi_in, solo);
            cuda_ok(cub::DeviceRadixSort::SortKeys(nullptr, bytes, epi_in, epi_sorted, (int64_t)n_epi_cand, 0, 64, stream), "radix size");
            cuda_ok(cudaMalloc(&temp, bytes), "malloc radix temp");
            cuda_ok(cub::DeviceRadixSort::SortKeys(temp, bytes, epi_in, epi_sorted, (int64_t)n_epi_cand, 0, 64, stream), "radix sort");
        }
        // 3. pseudo-level offsets
        cuda_ok(cudaMalloc(&pl_off, sizeof(uint32_t) * ((size_t)n_pl + 1)), "malloc pl_off");
        s_pl_off_kernel<<<(n_pl + 1 + 255) / 256, 256, 0, stream>>>(keys64, level_off_dev, n_levels, G, epi_sorted, n_epi_cand,
                                                                    n, pl_off);
        std::vector<uint32_t> pl_off_h(n_pl + 1);
        cuda_ok(cudaMemcpyAsync(pl_off_h.data(), pl_off, sizeof(uint32_t) * ((size_t)n_pl + 1), cudaMemcpyDeviceToHost, stream), "copy pl_off");
        cuda_ok(cudaStreamSynchronize(stream), "pl_off sync");
        cudaFree(temp); temp = nullptr;
        const uint32_t n_all = pl_off_h[n_pl];
        const uint32_t m = n_all - n;
        if (pl_off_h[n_real_pl] != n) throw std::runtime_error("compile_schedule: pseudo-level offsets inconsistent");
        // 32-bit keys (cell * Dl + dl) of real and epilogue slots
        cuda_ok(cudaMalloc(&keys, sizeof(uint32_t) * (size_t)std::max<uint32_t>(n_all, 1)), "malloc keys");
        if (solo) {
            // walk form: renumber the slots direction-major -- pseudo-level (direction g, level l) = g * L + l -- so that a
            // block's slots are one contiguous range in the order it solves them
            std::vector<uint32_t> new_off(n_pl + 1), delta(n_real_pl);
            uint32_t o = 0;
            for (uint32_t gg = 0; gg < G; ++gg)
                for (uint32_t l = 0; l < n_levels; ++l) {
                    new_off[gg * n_levels + l] = o;
                    delta[l * G + gg] = o - pl_off_h[l * G + gg];   // modulo 2^32
                    o += pl_off_h[l * G + gg + 1] - pl_off_h[l * G + gg];
                }
            for (uint32_t pl = n_real_pl; pl <= n_pl; ++pl) new_off[pl] = pl_off_h[pl];
            uint32_t *delta_dev = nullptr;
            cuda_ok(cudaMalloc(&delta_dev, sizeof(uint32_t) * (size_t)std::max<uint32_t>(n_real_pl, 1)), "malloc delta");
            cuda_ok(cudaMemcpyAsync(delta_dev, delta.data(), sizeof(uint32_t) * (size_t)n_real_pl, cudaMemcpyHostToDevice, stream), "copy delta");
            w_permute_kernel<<<blocks, 256, 0, stream>>>(keys64, n, pl_off, n_real_pl, delta_dev, keys);
            cuda_ok(cudaStreamSynchronize(stream), "permute sync");
            cudaFree(delta_dev);
            pl_off_h = new_off;
            cuda_ok(cudaMemcpyAsync(pl_off, pl_off_h.data(), sizeof(uint32_t) * ((size_t)n_pl + 1), cudaMemcpyHostToDevice, stream), "copy pl_off");
        } else
        s_key32_kernel<<<blocks, 256, 0, stream>>>(keys64, n, keys);
        if (m) s_key32_kernel<<<(m + 255) / 256, 256, 0, stream>>>(epi_sorted, m, keys + n);
        cuda_ok(cudaStreamSynchronize(stream), "key32 sync");
        cudaFree(keys64); keys64 = nullptr;
        cudaFree(epi_in); epi_in = nullptr;
        cudaFree(epi_sorted); epi_sorted = nullptr;

        // 4. slots, entry counts, downwind areas
        cuda_ok(cudaMalloc(&C.slot_of, sizeof(uint32_t) * (size_t)n), "malloc slot_of");
        cuda_ok(cudaMalloc(&C.ttot_slot, sizeof(double) * (size_t)n), "malloc ttot_slot");
        cuda_ok(cudaMalloc(&cnt, sizeof(uint32_t) * ((size_t)n_all + 1)), "malloc cnt");
        cuda_ok(cudaMalloc(&upoff, sizeof(unsigned long long) * ((size_t)n_all + 1)), "malloc upoff");
        cuda_ok(cudaMalloc(&counters, 8 * sizeof(unsigned int)), "malloc counters");
        cuda_ok(cudaMemsetAsync(counters, 0, 8 * sizeof(unsigned int), stream), "memset");
        cuda_ok(cudaMemsetAsync(cnt + n_all, 0, sizeof(uint32_t), stream), "memset");
        s_slot_scatter_kernel<<<blocks, 256, 0, stream>>>(keys, n, g.n_cells, n_dl, C.slot_of);
        s_count_kernel<<<(n_all + 255) / 256, 256, 0, stream>>>(g, keys, n, n_all, n_dl, C.slot_of, pl_off, n_pl, cnt,
                                                                C.ttot_slot, counters);
        cuda_ok(cub::DeviceScan::ExclusiveSum(nullptr, bytes, cnt, upoff, (int64_t)n_all + 1, stream), "scan size");
        cuda_ok(cudaMalloc(&temp, bytes), "malloc scan temp");
        cuda_ok(cub::DeviceScan::ExclusiveSum(temp, bytes, cnt, upoff, (int64_t)n_all + 1, stream), "scan");
        unsigned long long total_entries = 0;
        unsigned int n_lag = 0;
        cuda_ok(cudaMemcpyAsync(&total_entries, upoff + n_all, sizeof total_entries, cudaMemcpyDeviceToHost, stream), "copy");
        cuda_ok(cudaMemcpyAsync(&n_lag, counters, sizeof n_lag, cudaMemcpyDeviceToHost, stream), "copy");
        cuda_ok(cudaStreamSynchronize(stream), "count sync");
        cudaFree(temp); temp = nullptr;
        cudaFree(cnt); cnt = nullptr;
        C.n_entries = total_entries;
        C.n_lag = n_lag;
        if ((uint64_t)n + n_lag >= 0x7fffff00ull) throw std::runtime_error("compile_schedule: slot index overflow");

        // 5. tiles: greedy cut at segment boundaries, one warp per pseudo-level.  First pass: uniform
        //    tiles of <= threads slots (sizes the ring stages and hence the grid); second pass (after 6):
        //    per pseudo-level tile size chosen so that every block of the group gets the same number of
        //    tiles of the level (no block idles at the level barrier because of a remainder).
        cuda_ok(cudaMalloc(&tile_cnt, sizeof(uint32_t) * (size_t)n_pl), "malloc tile_cnt");
        cuda_ok(cudaMalloc(&tile_off, sizeof(uint32_t) * ((size_t)n_pl + 1)), "malloc tile_off");
        const unsigned cut_blocks = (unsigned)(((size_t)n_pl * 32 + 127) / 128);
        std::vector<uint32_t> tcnt(n_pl), toff(n_pl + 1, 0), tstart;
        std::vector<unsigned long long> tentry;
        uint32_t n_tiles = 0;
        auto cut = [&](const uint32_t *pl_max_dev) {
            if (solo) w_cut_kernel<<<(n_pl + 127) / 128, 128, 0, stream>>>(upoff, pl_off, n_pl, tile_slots, walk_entries, nullptr, tile_cnt, nullptr);
            else
            s_cut_kernel<<<cut_blocks, 128, 0, stream>>>(keys, pl_off, n_pl, n_dl, tile_slots, pl_max_dev, nullptr, tile_cnt, nullptr);
            cuda_ok(cudaMemcpyAsync(tcnt.data(), tile_cnt, sizeof(uint32_t) * n_pl, cudaMemcpyDeviceToHost, stream), "copy");
            cuda_ok(cudaStreamSynchronize(stream), "cut sync");
            for (uint32_t l = 0; l < n_pl; ++l) toff[l + 1] = toff[l] + tcnt[l];
            n_tiles = toff[n_pl];
            cuda_ok(cudaMemcpyAsync(tile_off, toff.data(), sizeof(uint32_t) * ((size_t)n_pl + 1), cudaMemcpyHostToDevice, stream), "copy");
            cudaFree(tile_start); tile_start = nullptr;
            cudaFree(tentry_dev); tentry_dev = nullptr;
            cuda_ok(cudaMalloc(&tile_start, sizeof(uint32_t) * ((size_t)n_tiles + 1)), "malloc tile_start");
            if (solo) w_cut_kernel<<<(n_pl + 127) / 128, 128, 0, stream>>>(upoff, pl_off, n_pl, tile_slots, walk_entries, tile_off, nullptr, tile_start);
            else
            s_cut_kernel<<<cut_blocks, 128, 0, stream>>>(keys, pl_off, n_pl, n_dl, tile_slots, pl_max_dev, tile_off, nullptr, tile_start);
            tstart.assign(n_tiles + 1, 0);
            tentry.assign(n_tiles + 1, 0);
            cuda_ok(cudaMalloc(&tentry_dev, sizeof(unsigned long long) * ((size_t)n_tiles + 1)), "malloc tentry");
            s_gather_offsets_kernel<<<(n_tiles + 1 + 255) / 256, 256, 0, stream>>>(upoff, tile_start, n_tiles, n_all, tentry_dev);
            cuda_ok(cudaMemcpyAsync(tstart.data(), tile_start, sizeof(uint32_t) * n_tiles, cudaMemcpyDeviceToHost, stream), "copy");
            cuda_ok(cudaMemcpyAsync(tentry.data(), tentry_dev, sizeof(unsigned long long) * ((size_t)n_tiles + 1),
                                    cudaMemcpyDeviceToHost, stream), "copy");
            cuda_ok(cudaStreamSynchronize(stream), "cut sync 2");
            tstart[n_tiles] = n_all;
        };
        cut(nullptr);

        // 6. launch geometry: stage size = largest packet; blocks per SM and stages from the smem budget
        uint32_t max_bytes = 0;
        for (uint32_t t = 0; t < n_tiles && !solo; ++t) {
            const uint32_t ns = tstart[t + 1] - tstart[t];
            const unsigned long long E = tentry[t + 1] - tentry[t];
            if (ns > tile_slots || E > 65535ull)
                throw std::runtime_error("compile_schedule: tile too large (more than 65535 upwind entries in one tile)");
            max_bytes = std::max(max_bytes, tile_layout(ns, (uint32_t)E).bytes);
        }
        uint32_t stage_bytes = std::max<uint32_t>(128u, (max_bytes + 127u) & ~127u);
        const size_t smem_sm = 227u * 1024u;
        uint32_t bps = std::max<uint32_t>(1u, want_bps), stages = 0;
        for (; !solo; --bps) {
            const size_t per_block = smem_sm / bps - 1024 - 1024;   // driver reserve + static shared memory
            const size_t fixed = stream_smem_bytes(0, 0);
            if (per_block >= fixed + stage_bytes) {
                stages = (uint32_t)std::min<size_t>(want_stages, (per_block - fixed) / stage_bytes);
                if (stages >= 2 || bps == 1) break;
            }
            if (bps == 1) break;
        }
        if (!solo && stages < 1) throw std::runtime_error("compile_schedule: a tile packet does not fit in shared memory");
        int per_sm = 0;
        if (solo) {
            per_sm = 1;   // the ring is sized once the far entries of the tiles are counted (step 7)
        } else {
            StreamKernel kernel = stream_kernel_for(threads, bps, false, false);
            const size_t smem = stream_smem_bytes(stages, stage_bytes);
            raise_smem_limit((const void *)kernel);
            cuda_ok(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, (int)threads, smem), "occupancy");
        }
        if (per_sm < 1) throw std::runtime_error("compile_schedule: stream kernel does not fit on an SM");
        uint32_t nb = (uint32_t)std::min<int>(per_sm, (int)bps) * (uint32_t)num_sms;
        if (solo) nb = G;
        // interleaved (default): every block serves all direction groups in turn -- group A's tiles of level l,
        // group B's tiles of level l, group A's of level l + 1, ... -- so the latency of one group's level
        // barrier hides behind the other groups' tiles.  Otherwise block b is dedicated to group b % G.
        const bool interleave = !solo && env_u32("SSW_STREAM_INTERLEAVE", 1) != 0;
        if (!interleave) {
            nb -= nb % G;
            if (nb < G) throw std::runtime_error("compile_schedule: fewer blocks than direction groups");
        }
        const uint32_t nb_g = interleave ? nb : nb / G;
        C.threads = threads;
        C.solo = solo;
        C.bps = bps;
        C.stages = stages;            // walk form: set in step 7
        C.stage_bytes = stage_bytes;
        C.n_blocks = nb;
        if (!solo && env_u32("SSW_STREAM_BALANCED", 1)) {
            // k tiles per block and level: tile size = ceil(n_l / (nb_g k)) plus one segment of slack for
            // the cuts at segment boundaries, so the level never needs more than nb_g * k tiles
            std::vector<uint32_t> pl_max(n_pl);
            const uint64_t seg_max = solo ? 1 : n_dl;   // longest run of slots that must stay in one tile
            for (uint32_t pl = 0; pl < n_pl; ++pl) {
                const uint64_t n_l = pl_off_h[pl + 1] - pl_off_h[pl];
                const uint64_t k = std::max<uint64_t>(1, (n_l + (uint64_t)nb_g * tile_slots - 1) / ((uint64_t)nb_g * tile_slots));
                uint64_t sz = (n_l + nb_g * k - 1) / (nb_g * k) + (solo ? 0 : seg_max);
                sz = std::max<uint64_t>(sz, std::min<uint64_t>(tile_slots, 2 * seg_max + 32));   // tiny levels: few tiles
                pl_max[pl] = (uint32_t)std::min<uint64_t>(sz, tile_slots);
            }
            cuda_ok(cudaMalloc(&pl_max_dev, sizeof(uint32_t) * (size_t)n_pl), "malloc pl_max");
            cuda_ok(cudaMemcpyAsync(pl_max_dev, pl_max.data(), sizeof(uint32_t) * (size_t)n_pl, cudaMemcpyHostToDevice, stream), "copy");
            cut(pl_max_dev);
            bool fits = true;
            for (uint32_t t = 0; t < n_tiles && fits; ++t) {
                const uint32_t ns = tstart[t + 1] - tstart[t];
                const unsigned long long E = tentry[t + 1] - tentry[t];
                fits = ns <= tile_slots && E <= 65535ull && (solo || tile_layout(ns, (uint32_t)E).bytes <= stage_bytes);
            }
            if (!fits) cut(nullptr);   // a shifted tile outgrew the ring stage: keep the uniform cut
        }

        // 7. tile table (block-major), per-block streams, barrier targets and dependencies
        std::vector<uint32_t> tile_block(n_tiles), tile_level(n_tiles), per_block_count(nb, 0), rr(G, 0);
        uint32_t rr_all = 0;
        for (uint32_t pl = 0; pl < n_pl; ++pl) {
            const uint32_t grp = pl >= n_real_pl ? pl - n_real_pl : (solo ? pl / n_levels : pl % G);
            for (uint32_t t = toff[pl]; t < toff[pl + 1]; ++t) {
                const uint32_t b = interleave ? (rr_all++ % nb) : grp + (rr[grp]++ % nb_g) * G;
                tile_block[t] = b;
                tile_level[t] = pl;
                per_block_count[b]++;
            }
        }
        std::vector<uint32_t> tile_far(n_tiles, 0);
        uint32_t *tile_level_dev = nullptr, *tile_far_dev = nullptr, *tile_start_dev = nullptr;
        if (solo && n_tiles) {
            // entries outside the window (gathered from global memory) per tile: sizes the packets and the ring stages
            cuda_ok(cudaMalloc(&tile_level_dev, sizeof(uint32_t) * (size_t)n_tiles), "malloc tile_level");
            cuda_ok(cudaMalloc(&tile_far_dev, sizeof(uint32_t) * (size_t)n_tiles), "malloc tile_far");
            cuda_ok(cudaMalloc(&tile_start_dev, sizeof(uint32_t) * ((size_t)n_tiles + 1)), "malloc tile_start");
            cuda_ok(cudaMemcpyAsync(tile_level_dev, tile_level.data(), sizeof(uint32_t) * (size_t)n_tiles, cudaMemcpyHostToDevice, stream), "copy");
            cuda_ok(cudaMemcpyAsync(tile_start_dev, tstart.data(), sizeof(uint32_t) * ((size_t)n_tiles + 1), cudaMemcpyHostToDevice, stream), "copy");
            WalkFillArgs fa{};
            fa.g = g; fa.keys = keys; fa.slot_of = C.slot_of; fa.pidx = pidx; fa.upoff = upoff; fa.ttot_slot = C.ttot_slot;
            fa.pl_off = pl_off; fa.n_pl = n_pl; fa.n_real_pl = n_real_pl; fa.n_dl = n_dl; fa.n_tasks = n; fa.n_levels = n_levels;
            fa.tile_start = tile_start_dev; fa.tile_level = tile_level_dev; fa.tile_far = tile_far_dev;
            fa.lag_counter = counters + 1;
            // The window competes with the ring for shared memory.  The number of ring stages is a multiple of both the
            // compute groups and the gather warps (a waiter then sees every phase of the mbarriers it waits on: the
            // previous tile of the same stage was its own), at least groups + 2.  The window shrinks in steps of 2048
            // slots until that many stages fit; it never goes below what keeps every source outside the usable window
            // complete in global memory when a tile's gather starts: (stages + 3 groups) tiles.
            uint32_t unit = walk_groups;
            while (unit % kWalkGatherWarps) unit += walk_groups;
            const uint32_t stages_min = ((walk_groups + 2u + unit - 1u) / unit) * unit;
            const uint32_t window_min = std::max<uint32_t>(2048u, (kWalkMaxStages + 3u * walk_groups) * tile_slots);
            window = std::max(window, window_min);
            for (;; window -= 2048u) {
                fa.window = window;
                fa.window_usable = window - walk_groups * tile_slots;   // tiles of one level in flight side by side
                w_fill_kernel<true><<<n_tiles, 512, 0, stream>>>(fa);
                cuda_ok(cudaMemcpyAsync(tile_far.data(), tile_far_dev, sizeof(uint32_t) * (size_t)n_tiles, cudaMemcpyDeviceToHost, stream), "copy");
                cuda_ok(cudaStreamSynchronize(stream), "far count sync");
                uint32_t mx = 0;
                for (uint32_t t = 0; t < n_tiles; ++t) {
                    const uint32_t ns = tstart[t + 1] - tstart[t];
                    const unsigned long long E = tentry[t + 1] - tentry[t];
                    if (ns > tile_slots || E > 65535ull || tile_far[t] > 65535u) throw WalkUnsupported("walk form: tile too large");
                    mx = std::max(mx, walk_layout(ns, (uint32_t)E, tile_far[t]).stage_bytes);
                }
                stage_bytes = std::max<uint32_t>(128u, (mx + 127u) & ~127u);
                const size_t per_block = smem_sm - 1024 - 1024;
                const size_t fixed = walk_smem_bytes(0, 0, window);
                stages = per_block > fixed ? (uint32_t)std::min<size_t>(want_stages, (per_block - fixed) / stage_bytes) : 0;
                stages -= stages % unit;
                if (stages >= stages_min || window < window_min + 2048u) break;
            }
            cudaFree(tile_level_dev); cudaFree(tile_far_dev); cudaFree(tile_start_dev);
            if (stages < unit) throw WalkUnsupported("walk form: too few tile packets fit beside the window");
            WalkKernel wk = walk_kernel_for(walk_groups, threads, false);
            const size_t smem = walk_smem_bytes(stages, stage_bytes, window);
            raise_smem_limit((const void *)wk);
            int occ = 0;
            cuda_ok(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, wk, (int)(walk_groups * threads) + 32 + 32 * kWalkGatherWarps, smem), "occupancy");
            if (occ < 1) throw WalkUnsupported("walk form: the kernel does not fit on an SM");
            C.stages = stages;
            C.stage_bytes = stage_bytes;
            C.walk_groups = walk_groups;
        }
        std::vector<uint32_t> tab_off(nb + 1, 0);
        for (uint32_t b = 0; b < nb; ++b) tab_off[b + 1] = tab_off[b] + per_block_count[b];
        std::vector<TileDesc> tab(std::max<uint32_t>(n_tiles, 1));
        std::vector<uint32_t> tab_block_h(std::max<uint32_t>(n_tiles, 1));
        std::vector<uint64_t> stream_off(nb + 1, 0), cursor(nb, 0);
        std::vector<uint32_t> fill_pos(nb, 0), first_of_level(nb, 0), level_of_block(nb, 0xffffffffu);
        for (uint32_t t = 0; t < n_tiles; ++t) {
            const uint32_t b = tile_block[t];
            if (level_of_block[b] != tile_level[t]) { level_of_block[b] = tile_level[t]; first_of_level[b] = fill_pos[b]; }
            TileDesc d;
            d.need = first_of_level[b];
            d.slot0 = tstart[t];
            d.n = (uint16_t)(tstart[t + 1] - tstart[t]);
            d.n_entries = (uint16_t)(tentry[t + 1] - tentry[t]);
            d.level = tile_level[t];
            d.aux = tile_far[t] | (d.level >= n_real_pl ? 1u << 16 : 0u);
            if ((cursor[b] >> 4) > 0xffffffffull) throw std::runtime_error("compile_schedule: block stream exceeds 64 GB");
            d.off16 = (uint32_t)(cursor[b] >> 4);
            cursor[b] += solo ? walk_layout(d.n, d.n_entries, tile_far[t]).bytes : tile_layout(d.n, d.n_entries).bytes;
            tab[tab_off[b] + fill_pos[b]] = d;
            tab_block_h[tab_off[b] + fill_pos[b]] = b;
            fill_pos[b]++;
        }
        for (uint32_t b = 0; b < nb; ++b) stream_off[b + 1] = stream_off[b] + ((cursor[b] + 127u) & ~(uint64_t)127u);
        C.stream_bytes = stream_off[nb];
        std::vector<uint32_t> lvl_target(n_pl), lvl_dep(n_pl, kNoDep);
        for (uint32_t pl = 0; pl < n_pl; ++pl) lvl_target[pl] = std::min<uint32_t>(tcnt[pl], nb_g);
        for (uint32_t pl = G; pl < n_real_pl && !solo; ++pl) lvl_dep[pl] = pl - G;
        for (uint32_t grp = 0; grp < G && !solo; ++grp) {   // epilogue: behind the group's last non-empty wavefront level
            for (uint32_t l = n_levels; l-- > 0;) {
                if (tcnt[l * G + grp] > 0) { lvl_dep[n_real_pl + grp] = l * G + grp; break; }
            }
        }
        C.n_tiles = n_tiles;

        cuda_ok(cudaMalloc(&C.stream, std::max<uint64_t>(C.stream_bytes, 128)), "malloc stream");
        cuda_ok(cudaMalloc(&C.tab, sizeof(TileDesc) * tab.size()), "malloc tab");
        cuda_ok(cudaMalloc(&tab_block, sizeof(uint32_t) * tab_block_h.size()), "malloc tab_block");
        cuda_ok(cudaMalloc(&C.tab_off, sizeof(uint32_t) * ((size_t)nb + 1)), "malloc tab_off");
        cuda_ok(cudaMalloc(&C.stream_off, sizeof(uint64_t) * ((size_t)nb + 1)), "malloc stream_off");
        cuda_ok(cudaMalloc(&C.lvl_target, sizeof(uint32_t) * (size_t)n_pl), "malloc lvl_target");
        cuda_ok(cudaMalloc(&C.lvl_dep, sizeof(uint32_t) * (size_t)n_pl), "malloc lvl_dep");
        cuda_ok(cudaMalloc(&C.lvl_count, sizeof(unsigned int) * (solo ? (size_t)kCountStride : (size_t)n_pl * kCountStride)), "malloc lvl_count");
        cuda_ok(cudaMalloc(&C.out_slot, sizeof(double) * ((size_t)n + n_lag)), "malloc out_slot");
        cuda_ok(cudaMalloc(&C.lag_src, sizeof(uint32_t) * (size_t)std::max<uint32_t>(n_lag, 1)), "malloc lag_src");
        cuda_ok(cudaMalloc(&C.acc_cell, sizeof(double) * (size_t)G * g.n_cells), "malloc acc_cell");
        if (solo) {
            cuda_ok(cudaMalloc(&C.rec_slot, sizeof(double2) * (size_t)std::max<uint32_t>(n, 1)), "malloc rec_slot");
            cuda_ok(cudaMalloc(&C.cell_of_slot, sizeof(uint32_t) * (size_t)std::max<uint32_t>(n, 1)), "malloc cell_of_slot");
            w_cell_of_slot_kernel<<<blocks, 256, 0, stream>>>(keys, n, n_dl, C.cell_of_slot);
        }
        cuda_ok(cudaMalloc(&C.acc_per, sizeof(double) * (size_t)G * std::max<uint32_t>(n_periodic, 1)), "malloc acc_per");
        cuda_ok(cudaMemcpyAsync(C.tab, tab.data(), sizeof(TileDesc) * tab.size(), cudaMemcpyHostToDevice, stream), "copy tab");
        cuda_ok(cudaMemcpyAsync(tab_block, tab_block_h.data(), sizeof(uint32_t) * tab_block_h.size(), cudaMemcpyHostToDevice, stream), "copy");
        cuda_ok(cudaMemcpyAsync(C.tab_off, tab_off.data(), sizeof(uint32_t) * ((size_t)nb + 1), cudaMemcpyHostToDevice, stream), "copy");
        cuda_ok(cudaMemcpyAsync(C.stream_off, stream_off.data(), sizeof(uint64_t) * ((size_t)nb + 1), cudaMemcpyHostToDevice, stream), "copy");
        cuda_ok(cudaMemcpyAsync(C.lvl_target, lvl_target.data(), sizeof(uint32_t) * (size_t)n_pl, cudaMemcpyHostToDevice, stream), "copy");
        cuda_ok(cudaMemcpyAsync(C.lvl_dep, lvl_dep.data(), sizeof(uint32_t) * (size_t)n_pl, cudaMemcpyHostToDevice, stream), "copy");

        // 8. packets and state
        if (n_tiles && solo) {
            WalkFillArgs fa{};
            fa.g = g; fa.keys = keys; fa.slot_of = C.slot_of; fa.pidx = pidx; fa.upoff = upoff; fa.ttot_slot = C.ttot_slot;
            fa.pl_off = pl_off; fa.n_pl = n_pl; fa.n_real_pl = n_real_pl; fa.n_dl = n_dl; fa.n_tasks = n; fa.n_levels = n_levels;
            fa.window = window; fa.window_usable = window - walk_groups * tile_slots;
            fa.tab = C.tab; fa.tab_block = tab_block; fa.stream_off = C.stream_off; fa.stream = C.stream;
            fa.lag_src = C.lag_src; fa.lag_counter = counters + 1;
            w_fill_kernel<false><<<n_tiles, 512, 0, stream>>>(fa);
        } else if (n_tiles) {
            FillArgs fa;
            fa.g = g; fa.keys = keys; fa.slot_of = C.slot_of; fa.pidx = pidx; fa.upoff = upoff; fa.ttot_slot = C.ttot_slot;
            fa.pl_off = pl_off; fa.n_pl = n_pl; fa.n_real_pl = n_real_pl; fa.n_dl = n_dl; fa.n_tasks = n; fa.n_groups = G;
            fa.tab = C.tab; fa.tab_block = tab_block; fa.tab_off = C.tab_off; fa.stream_off = C.stream_off;
            fa.stream = C.stream; fa.lag_src = C.lag_src; fa.lag_counter = counters + 1; fa.stages = stages;
            fa.lvl_dep = C.lvl_dep; fa.lvl_target = C.lvl_target;
            s_fill_kernel<<<n_tiles, kMaxStreamThreads, 0, stream>>>(fa);
        }
        s_convert_state_kernel<<<blocks, 256, 0, stream>>>(keys, n, g.n_cells, n_dl, q_nat, C.ttot_slot, C.out_slot);
        cuda_ok(cudaGetLastError(), "compile kernels");
        unsigned int lag_filled = 0;
        cuda_ok(cudaMemcpyAsync(&lag_filled, counters + 1, sizeof lag_filled, cudaMemcpyDeviceToHost, stream), "copy");
        cuda_ok(cudaStreamSynchronize(stream), "compile sync");   // host vectors go out of scope
        if (lag_filled != n_lag) throw std::runtime_error("compile_schedule: periodic snapshot count mismatch");
        unsigned int too_many = 0;
        cuda_ok(cudaMemcpy(&too_many, counters + 2, sizeof too_many, cudaMemcpyDeviceToHost), "copy");
        if (too_many && solo) throw WalkUnsupported("walk form: a task has more than 62 upwind faces");
        if (too_many) throw std::runtime_error("compile_schedule: a task has more than 255 periodic upwind faces");
        if (solo) {
            unsigned int nf[2] = {0, 0};
            cuda_ok(cudaMemcpy(nf, counters + 3, sizeof nf, cudaMemcpyDeviceToHost), "copy");
            C.n_near = nf[0];
            C.n_far = nf[1];
            C.window = window;
        }
        if (launch_counter) *launch_counter += 14;
        C.n_epilogue = m;
    } catch (...) {
        cleanup();
        C.release();
        throw;
    }
    cleanup();
    C.n_tasks = n_tasks;
    C.n_levels = n_levels;
    C.n_groups = G;
    C.n_groups_per = G;
    C.n_pl = n_pl;
    C.n_cells = g.n_cells;
    C.n_periodic = n_periodic;
    C.mean_entries = n_tasks ? (double)C.n_entries / (double)n_tasks : 0.0;
    C.valid = true;
}

// One all-cells sweep over the compiled schedule.  Leaves the per-group sums in C.acc_cell / C.acc_per (s_rate_finish_kernel folds them).
// One all-cells sweep in the walk form.  Leaves the incoming rates in C.inc_slot / C.acc_per (w_rate_finish_kernel folds them).
inline void run_walk(Compiled &C, const double *att, const double *src, double n_dirs_total, double threshold,
                     cudaStream_t stream, uint64_t *launch_counter) {
    WalkArgs a;
    a.stream = C.stream;
    a.stream_off = C.stream_off;
    a.tab = C.tab;
    a.tab_off = C.tab_off;
    a.out_slot = C.out_slot;
    a.rec_slot = C.rec_slot;
    a.acc_cell = C.acc_cell;
    a.acc_per = C.acc_per;
    a.n_cells = C.n_cells;
    a.threshold = threshold;
    a.stages = C.stages;
    a.stage_bytes = C.stage_bytes;
    a.window = C.window;
    a.n_periodic = C.n_periodic;
    a.l2_ahead = env_u32("SSW_WALK_L2_AHEAD", 0);   // measured: no gain (the copies are not DRAM-latency bound)
    a.prof = nullptr;
    unsigned long long *prof_dev = nullptr;
    if (env_u32("SSW_STREAM_PROFILE", 0)) {
        cuda_ok(cudaMalloc(&prof_dev, sizeof(unsigned long long) * 8 * (size_t)C.n_blocks), "malloc prof");
        a.prof = prof_dev;
    }
    // every slot's incoming rate is stored exactly once per sweep; only the periodic rows have gaps
    if (C.n_periodic)
        cuda_ok(cudaMemsetAsync(C.acc_per, 0, sizeof(double) * (size_t)C.n_groups * C.n_periodic, stream), "memset acc_per");
    w_rec_kernel<<<(unsigned)((C.n_tasks + 255) / 256), 256, 0, stream>>>(C.cell_of_slot, att, src, n_dirs_total, (uint32_t)C.n_tasks, C.rec_slot);
    uint64_t launches = 2;
    if (C.n_lag) {
        s_lag_snapshot_kernel<<<(C.n_lag + 255) / 256, 256, 0, stream>>>(C.lag_src, C.n_lag, (uint32_t)C.n_tasks, C.out_slot);
        ++launches;
    }
    WalkKernel kernel = walk_kernel_for(C.walk_groups, C.threads, prof_dev != nullptr);
    const size_t smem = walk_smem_bytes(C.stages, C.stage_bytes, C.window);
    raise_smem_limit((const void *)kernel);
    kernel<<<C.n_blocks, C.walk_groups * C.threads + 32 + 32 * kWalkGatherWarps, smem, stream>>>(a);
    cuda_ok(cudaGetLastError(), "walk_kernel launch");
    if (prof_dev) {
        std::vector<unsigned long long> h(8 * (size_t)C.n_blocks);
        cudaMemcpyAsync(h.data(), prof_dev, sizeof(unsigned long long) * h.size(), cudaMemcpyDeviceToHost, stream);
        cudaStreamSynchronize(stream);
        cudaFree(prof_dev);
        double tot = 0, pkt = 0, tiles = 0, tmax = 0, loop = 0, outp = 0, bar = 0, gat = 0, dep = 0;
        for (uint32_t b = 0; b < C.n_blocks; ++b) {
            tot += (double)h[8 * b]; pkt += (double)h[8 * b + 1]; tiles += (double)h[8 * b + 2];
            loop += (double)h[8 * b + 3]; outp += (double)h[8 * b + 4]; bar += (double)h[8 * b + 5]; gat += (double)h[8 * b + 6];
            dep += (double)h[8 * b + 7];
            tmax = std::max(tmax, (double)h[8 * b]);
        }
        fprintf(stderr, "[walk profile] blocks %u groups %u x %u threads tiles %.0f levels %u  cycles/block mean %.0f max %.0f  thread 0: packet wait %.1f%% "
                        "level wait %.1f%% gather wait %.1f%% entries %.1f%% output %.1f%% barrier %.1f%%  cycles per tile %.0f  stage bytes %u stages %u window %u  near %.1f%% of %llu entries\n",
                C.n_blocks, C.walk_groups, C.threads, tiles, C.n_levels, tot / C.n_blocks, tmax, 100.0 * pkt / tot, 100.0 * dep / tot, 100.0 * gat / tot, 100.0 * loop / tot, 100.0 * outp / tot,
                100.0 * bar / tot, tot / tiles, C.stage_bytes, C.stages,
                C.window, 100.0 * (double)C.n_near / (double)std::max<uint64_t>(1, C.n_near + C.n_far),
                (unsigned long long)(C.n_near + C.n_far));
    }
    if (launch_counter) *launch_counter += launches;
}

inline void run_compiled(Compiled &C, const double2 *cellrec, double threshold, cudaStream_t stream,
                         uint64_t *launch_counter) {
    if (C.solo) throw std::runtime_error("run_compiled: the walk form is run by run_walk");
    StreamArgs a;
    a.stream = C.stream;
    a.stream_off = C.stream_off;
    a.tab = C.tab;
    a.tab_off = C.tab_off;
    a.lvl_target = C.lvl_target;
    a.lvl_dep = C.lvl_dep;
    a.lvl_count = C.lvl_count;
    a.out_slot = C.out_slot;
    a.cellrec = cellrec;
    a.acc_cell = C.acc_cell;
    a.acc_per = C.acc_per;
    a.threshold = threshold;
    a.stages = C.stages;
    a.stage_bytes = C.stage_bytes;
    a.n_groups = C.n_groups;
    a.n_real_pl = C.n_levels * C.n_groups;
    a.n_cells = C.n_cells;
    a.n_periodic = C.n_periodic;
    a.poll_ns = env_u32("SSW_STREAM_POLL_NS", 20);
    a.prof = nullptr;
    unsigned long long *prof_dev = nullptr;
    if (env_u32("SSW_STREAM_PROFILE", 0)) {
        cuda_ok(cudaMalloc(&prof_dev, sizeof(unsigned long long) * 6 * (size_t)C.n_blocks), "malloc prof");
        cuda_ok(cudaMemsetAsync(prof_dev, 0, sizeof(unsigned long long) * 6 * (size_t)C.n_blocks, stream), "memset prof");
        a.prof = prof_dev;
    }
    if (!C.solo) {   // solo: no level counters, and every (direction, cell) term is stored exactly once per sweep
        cuda_ok(cudaMemsetAsync(C.lvl_count, 0, sizeof(unsigned int) * (size_t)C.n_pl * kCountStride, stream), "memset lvl_count");
        cuda_ok(cudaMemsetAsync(C.acc_cell, 0, sizeof(double) * (size_t)C.n_groups * C.n_cells, stream), "memset acc_cell");
    }
    if (C.n_periodic)
        cuda_ok(cudaMemsetAsync(C.acc_per, 0, sizeof(double) * (size_t)C.n_groups * C.n_periodic, stream), "memset acc_per");
    uint64_t launches = 1;
    if (C.n_lag) {
        s_lag_snapshot_kernel<<<(C.n_lag + 255) / 256, 256, 0, stream>>>(C.lag_src, C.n_lag, (uint32_t)C.n_tasks, C.out_slot);
        ++launches;
    }
    StreamKernel kernel = stream_kernel_for(C.threads, C.bps, prof_dev != nullptr, C.solo);
    const size_t smem = stream_smem_bytes(C.stages, C.stage_bytes);
    raise_smem_limit((const void *)kernel);
    void *args[] = {&a};
    // cooperative launch only to guarantee co-residency of all blocks (the level barriers spin)
    cuda_ok(cudaLaunchCooperativeKernel((const void *)kernel, dim3(C.n_blocks), dim3(C.threads), args, smem, stream),
            "sweep_stream_kernel launch");
    if (prof_dev) {   // experiments only: where do the blocks spend their cycles?
        std::vector<unsigned long long> h(6 * (size_t)C.n_blocks);
        cudaMemcpyAsync(h.data(), prof_dev, sizeof(unsigned long long) * h.size(), cudaMemcpyDeviceToHost, stream);
        cudaStreamSynchronize(stream);
        cudaFree(prof_dev);
        double tot = 0, bar = 0, pkt = 0, tiles = 0, tmax = 0, poll = 0, rel = 0;
        for (uint32_t b = 0; b < C.n_blocks; ++b) {
            tot += (double)h[6 * b]; bar += (double)h[6 * b + 1]; pkt += (double)h[6 * b + 2]; tiles += (double)h[6 * b + 3];
            poll += (double)h[6 * b + 4]; rel += (double)h[6 * b + 5];
            tmax = std::max(tmax, (double)h[6 * b]);
        }
        if (C.solo)
            fprintf(stderr, "[solo profile] blocks %u tiles %.0f levels %u  cycles/block mean %.0f max %.0f  packet wait %.1f%%  phase 1 (gathers) %.1f%%  "
                            "phase 2 (sums, stores) %.1f%%  cycles per tile %.0f  stage bytes %u stages %u\n",
                    C.n_blocks, tiles, C.n_levels, tot / C.n_blocks, tmax, 100.0 * pkt / tot, 100.0 * bar / tot, 100.0 * rel / tot, tot / tiles,
                    C.stage_bytes, C.stages);
        else
        fprintf(stderr, "[stream profile] blocks %u tiles %.0f  cycles/block mean %.0f max %.0f  level change: arrive %.1f%% + wait behind it %.1f%% "
                        "(dependency poll %.1f%%)  packet wait %.1f%%  cycles per tile (excl. level changes) %.0f  pseudo-levels %u\n",
                C.n_blocks, tiles, tot / C.n_blocks, tmax, 100.0 * rel / tot, 100.0 * bar / tot, 100.0 * poll / tot, 100.0 * pkt / tot,
                (tot - bar - rel) / tiles, C.n_pl);
    }
    if (launch_counter) *launch_counter += launches;
}

}  // namespace ssw
