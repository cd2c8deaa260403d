// patch.cuh -- the all-cells sweep as a dataflow of macro-tiles (spatial patch x direction group).
//
// stream.cuh crosses one device-wide barrier per wavefront level: 383 of them for a 128^3 Cartesian
// grid, each costing a release -> L2 -> acquire-poll -> gather round trip (about 3.5 us) however
// little work the level holds.  That latency floor bounds the single-GPU sweep and, far more, a
// direction shard of a multi-GPU job.  This file removes it the way structured-grid transport
// sweeps do (KBA): cells are grouped into spatial patches (about 8^3 cells, from the cell centres
// handed in through ssw_set_cell_positions), directions into groups of <= KD directions of one
// octant, and a MACRO-TILE = all tasks (c, d) with c in one patch and d in one group.
//
//   * inside a macro-tile the wavefront is walked by ONE thread block: its tasks are ordered by
//     sub-level (wavefront level minus the smallest level of that direction inside the patch), the
//     outgoing rates of the macro-tile live in shared memory, and consecutive sub-levels are
//     separated by a block barrier (tens of cycles) instead of a device-wide one;
//   * fluxes entering the patch ("external" entries: Local faces to other patches, periodic faces)
//     are gathered from global memory once, when the macro-tile starts;
//   * macro-tiles depend on each other through the quotient graph patch -> patch of their direction
//     group.  Every macro-tile owns a done flag (the epoch of the sweep that completed it); a
//     macro-tile acquire-polls the flags of its (<= 64) upwind macro-tiles and releases its own.
//     No device-wide barrier at all: a 128^3 grid has 46 dependent macro-tile levels instead of 383
//     wavefront levels.  Blocks consume macro-tiles in a global topological order (rank), dealt
//     round-robin, all blocks co-resident (cooperative launch): the lowest unfinished rank is always
//     runnable, so the polling cannot deadlock.
//   * everything a block needs is, as in stream.cuh, ONE sequential byte stream per block
//     (macro-tile head packet, then one tile packet per <= THREADS slots of a sub-level), prefetched
//     with TMA bulk copies into a shared-memory ring.
//   * per-cell photon rate: a task writes its incoming rate into a shared-memory row [patch cell][direction of the
//     group]; when the macro-tile is done one thread per cell sums its row in direction order (sum_d incoming[d],
//     src/sweep/mod.rs:554-558) and stores it once per (group, cell): no atomics, no read-modify-write in global
//     memory, and nothing of the reduction sits on the dependent chain of the tiles.
//   * periodic faces never carry a dependency (src/sweep/mod.rs:505-513): every periodic entry reads
//     a snapshot of its donor taken before the sweep (the lag of DESIGN.md section 4); the
//     periodic_source term of the rate (donors' NEW rates, site.rs:53-56) is one small kernel after
//     the sweep.
//
// The form needs an acyclic quotient graph.  Cartesian grids always have one; a jittered Voronoi
// grid usually has mutual patch dependencies for directions nearly parallel to a patch face -- then
// compile_patch_schedule throws PatchUnsupported and the caller keeps the level-barrier stream.
#pragma once
#include <cub/cub.cuh>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>

#include "stream.cuh"

namespace ssw {

struct PatchUnsupported : std::runtime_error {
    using std::runtime_error::runtime_error;
};

constexpr uint32_t kMaxPatchDeps = 64;
constexpr uint32_t kMaxPatchCells = 1024;
constexpr uint32_t kEmptyDep = 0xffffffffu;
constexpr int kRankShift = 44;       // sort key = rank << 44 | sub-level << 32 | (cell * Dl + dl)
constexpr int kSubShift = 32;
constexpr uint32_t kMaxSub = 4096;
constexpr uint32_t kMaxRank = 1u << 20;

struct PatchGrid {   // device view of the cell -> patch map
    const uint32_t *patch_of;     // N
    const uint16_t *lidx;         // N: index of the cell inside its patch
    const uint32_t *patch_off;    // P + 1
    const uint32_t *patch_cells;  // N, patch-major, ascending cell index inside a patch
    uint32_t n_patches;
    uint32_t max_cells;
    uint32_t dims[3];             // boxes per axis: patch = (bx * dims[1] + by) * dims[2] + bz
};

// packet descriptor, block-major, in consumption order
struct PDesc {
    uint32_t off16;   // packet offset inside the block's stream, 16-B units
    uint32_t bytes;
    uint32_t id;      // head: macro-tile rank; tile: tile index
    uint32_t flags;   // bit 0: head packet; bit 1: last tile of its macro-tile
};
constexpr uint32_t kPHead = 1u, kPLast = 2u;

// first 32 bytes of every packet
struct __align__(16) PHdr {
    uint16_t kind;     // 0 tile, 1 head
    uint16_t n;        // tile: slots; head: upwind macro-tiles to wait for
    uint16_t a16;      // tile: entries; head: cells of the patch
    uint16_t b16;      // tile: bit 0 = last tile of the macro-tile, bits 1-3 = K if every slot has exactly K Local entries and
                       // no periodic ones (else 0); head: direction group | directions in the group << 10
    uint32_t c32;      // tile: first slot of the tile inside the macro-tile; head: external entries
    uint32_t next_off16, next_bytes;   // packet that goes into this ring stage next (bytes = 0: none)
    uint32_t gslot0;   // head: first global slot of the macro-tile; tile: idx offset | lcell offset << 16 (16-B units)
    uint32_t n_slots;  // head: slots of the macro-tile; tile: info offset (16-B units)
    uint32_t rank;     // head: index of the macro-tile's done flag
};
static_assert(sizeof(PHdr) == 32, "packet header is 32 bytes");

struct HeadLayout { uint32_t dep, ext, cells, bytes; };
// head packet = [header][u32 flag index of every upwind macro-tile][u32 global source slot of every external entry][u32 cell ids]
__host__ __device__ inline HeadLayout head_layout(uint32_t n_dep, uint32_t n_ext, uint32_t n_cells) {
    HeadLayout L;
    uint32_t o = (uint32_t)sizeof(PHdr);
    L.dep = o;   o += align16(4u * n_dep);
    L.ext = o;   o += align16(4u * n_ext);
    L.cells = o; o += align16(4u * n_cells);
    L.bytes = o;
    return L;
}
struct PTileLayout { uint32_t w, idx, lcell, info, bytes; };
// tile packet = [header][f64 share[E]][u16 value index[E]][u16 patch-local cell[n]][u32 info[n + 1]]
// value index: < n_slots -> outgoing rate of a slot of this macro-tile, else n_slots + external entry
__host__ __device__ inline PTileLayout ptile_layout(uint32_t n, uint32_t E) {
    PTileLayout L;
    uint32_t o = (uint32_t)sizeof(PHdr);
    L.w = o;     o += align16(8u * E);
    L.idx = o;   o += align16(2u * E);
    L.lcell = o; o += align16(2u * n);
    L.info = o;  o += align16(4u * (n + 1u));
    L.bytes = o;
    return L;
}

struct PatchSmem { uint32_t bars, val, rec, inc, cellid, total; };
__host__ __device__ inline PatchSmem patch_smem(uint32_t stages, uint32_t stage_bytes, uint32_t vmax, uint32_t pc_max, uint32_t smax) {
    PatchSmem L;
    uint32_t o = stages * stage_bytes;
    L.bars = o;   o += 8u * kMaxStages;
    L.val = o;    o += align16(8u * vmax);
    L.rec = o;    o += 16u * pc_max;
    L.inc = o;    o += align16(8u * smax);   // incoming_total_rate per (patch cell, direction of the group)
    L.cellid = o; o += 2u * align16(4u * pc_max);   // double-buffered: the flush of one macro-tile overlaps the next head
    L.total = o;
    return L;
}

// ---- construction kernels ---------------------------------------------------------------------------

// wavefront level of every task (from its position in the level-sorted list) and the smallest level of
// every (patch, direction)
__global__ void __launch_bounds__(256)
p_level_kernel(const uint32_t *__restrict__ tasks, uint32_t n, const uint32_t *__restrict__ level_off, uint32_t n_levels,
               uint32_t n_cells, uint32_t n_dl, const uint32_t *__restrict__ patch_of, uint32_t *__restrict__ tlevel,
               unsigned int *__restrict__ minlev) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint32_t t = tasks[s];
    uint32_t lo = 0, hi = n_levels;   // invariant: level_off[lo] <= s < level_off[hi]
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (level_off[mid] <= s) lo = mid;
        else hi = mid;
    }
    tlevel[t] = lo;
    const uint32_t dl = t / n_cells, c = t - dl * n_cells;
    atomicMin(minlev + (size_t)patch_of[c] * n_dl + dl, lo);
}

// quotient graph: dep_tab[(group * P + patch) * 64 ..] = set of patches with a Local face into `patch`
// that is upwind for a direction of `group` (init_counts, src/sweep/mod.rs:346-386, lifted to patches)
__global__ void __launch_bounds__(256)
p_edges_kernel(GridView g, uint32_t n_dl, const uint32_t *__restrict__ patch_of, const uint16_t *__restrict__ group_of,
               uint32_t n_patches, unsigned int *dep_tab, unsigned int *err) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)g.n_cells * n_dl) return;
    const uint32_t dl = (uint32_t)(t / g.n_cells), c = (uint32_t)(t - (size_t)dl * g.n_cells);
    const uint32_t pc = patch_of[c];
    const double dx = g.dirs[3 * dl], dy = g.dirs[3 * dl + 1], dz = g.dirs[3 * dl + 2];
    unsigned int *row = dep_tab + ((size_t)group_of[dl] * n_patches + pc) * kMaxPatchDeps;
    for (uint32_t f = g.face_off[c]; f < g.face_off[c + 1]; ++f) {
        if (g.face_kind[f] != 0) continue;
        if (!(dot_dir(ld_geo(g.face_geo + f), dx, dy, dz) < 0.0)) continue;
        const uint32_t pn = patch_of[(uint32_t)g.face_nb[f]];
        if (pn == pc) continue;
        bool done = false;
        for (uint32_t i = 0; i < kMaxPatchDeps && !done; ++i) {
            unsigned int v = *((volatile unsigned int *)(row + i));
            if (v == kEmptyDep) v = atomicCAS(row + i, kEmptyDep, pn);
            done = v == pn || v == kEmptyDep;
        }
        if (!done) atomicExch(err, 1u);
    }
}

// ---- phases: the patch form for grids whose patch graph is cyclic (Voronoi) ----------------------------------------
// pi[grp * P + p] is a total order of the patches of a group.  A dependency edge is "back" if it runs from a later to
// an earlier patch; phase(task) = max over its upwind tasks of (their phase + [edge is back]).  Macro-tiles
// (patch, group, phase) ordered by (phase, pi) are a topological order of their quotient graph by construction.
// One launch per wavefront level, in level order (the upwind tasks of a level are final).
__global__ void __launch_bounds__(256)
p_phase_kernel(GridView g, const uint32_t *__restrict__ tasks, uint32_t s0, uint32_t s1, const uint32_t *__restrict__ patch_of,
               const uint16_t *__restrict__ group_of, uint32_t n_patches, const uint32_t *__restrict__ pi,
               uint8_t *phase, unsigned int *counters) {
    const uint32_t i = s0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s1) return;
    const uint32_t t = tasks[i];
    const uint32_t N = g.n_cells;
    const uint32_t dl = t / N, c = t - dl * N;
    const uint32_t *const pig = pi + (size_t)group_of[dl] * n_patches;
    const uint32_t my_pi = pig[patch_of[c]];
    const double dx = g.dirs[3 * dl], dy = g.dirs[3 * dl + 1], dz = g.dirs[3 * dl + 2];
    uint32_t ph = 0;
    for (uint32_t f = g.face_off[c]; f < g.face_off[c + 1]; ++f) {
        if (g.face_kind[f] != 0) continue;
        if (!(dot_dir(ld_geo(g.face_geo + f), dx, dy, dz) < 0.0)) continue;
        const uint32_t nb = (uint32_t)g.face_nb[f];
        const uint32_t q = (uint32_t)phase[(size_t)dl * N + nb] + (pig[patch_of[nb]] > my_pi ? 1u : 0u);
        ph = max(ph, q);
    }
    if (ph > 254u) { atomicExch(counters + 7, 1u); ph = 254u; }
    phase[t] = (uint8_t)ph;
    atomicMax(counters + 6, ph);
}

// which (group, patch, phase) triples exist
__global__ void __launch_bounds__(256)
p_present_kernel(uint32_t n_cells, uint32_t n_dl, const uint32_t *__restrict__ patch_of, const uint16_t *__restrict__ group_of,
                 uint32_t n_patches, uint32_t n_phase, const uint8_t *__restrict__ phase, uint32_t *__restrict__ present) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)n_cells * n_dl) return;
    const uint32_t dl = (uint32_t)(t / n_cells), c = (uint32_t)(t - (size_t)dl * n_cells);
    present[((size_t)group_of[dl] * n_patches + patch_of[c]) * n_phase + phase[t]] = 1u;
}

// dense macro-tile ids from the exclusive scan of `present`; mt_key[id] = (group * P + patch) * n_phase + phase
__global__ void __launch_bounds__(256)
p_mtid_kernel(const uint32_t *__restrict__ present, const uint32_t *__restrict__ scan, uint32_t n_keys,
              uint32_t *__restrict__ mtid_of, uint32_t *__restrict__ mt_key) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_keys) return;
    if (present[k]) { mtid_of[k] = scan[k]; mt_key[scan[k]] = k; }
    else mtid_of[k] = 0xffffffffu;
}

// quotient graph over macro-tiles with phases: dep_tab[id * 64 ..] = upwind macro-tiles; smallest wavefront level of
// every (macro-tile, direction of the group)
__global__ void __launch_bounds__(256)
p_edges2_kernel(GridView g, uint32_t n_dl, const uint32_t *__restrict__ patch_of, const uint16_t *__restrict__ group_of,
                const uint16_t *__restrict__ group_rank, uint32_t n_patches, uint32_t n_phase, uint32_t kd_max,
                const uint8_t *__restrict__ phase, const uint32_t *__restrict__ mtid_of, const uint32_t *__restrict__ tlevel,
                unsigned int *dep_tab, unsigned int *minlev, unsigned int *err) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t N = g.n_cells;
    if (t >= (size_t)N * n_dl) return;
    const uint32_t dl = (uint32_t)(t / N), c = (uint32_t)(t - (size_t)dl * N);
    const size_t gbase = (size_t)group_of[dl] * n_patches;
    const uint32_t me = mtid_of[(gbase + patch_of[c]) * n_phase + phase[t]];
    atomicMin(minlev + (size_t)me * kd_max + group_rank[dl], tlevel[t]);
    const double dx = g.dirs[3 * dl], dy = g.dirs[3 * dl + 1], dz = g.dirs[3 * dl + 2];
    unsigned int *row = dep_tab + (size_t)me * kMaxPatchDeps;
    for (uint32_t f = g.face_off[c]; f < g.face_off[c + 1]; ++f) {
        if (g.face_kind[f] != 0) continue;
        if (!(dot_dir(ld_geo(g.face_geo + f), dx, dy, dz) < 0.0)) continue;
        const uint32_t nb = (uint32_t)g.face_nb[f];
        const uint32_t up = mtid_of[(gbase + patch_of[nb]) * n_phase + phase[(size_t)dl * N + nb]];
        if (up == me) continue;
        bool done = false;
        for (uint32_t i = 0; i < kMaxPatchDeps && !done; ++i) {
            unsigned int v = *((volatile unsigned int *)(row + i));
            if (v == kEmptyDep) v = atomicCAS(row + i, kEmptyDep, up);
            done = v == up || v == kEmptyDep;
        }
        if (!done) atomicExch(err, 1u);
    }
}

__global__ void __launch_bounds__(256)
p_key2_kernel(uint32_t n_cells, uint32_t n_dl, const uint32_t *__restrict__ patch_of, const uint16_t *__restrict__ group_of,
              const uint16_t *__restrict__ group_rank, uint32_t n_patches, uint32_t n_phase, uint32_t kd_max,
              const uint8_t *__restrict__ phase, const uint32_t *__restrict__ mtid_of, const uint32_t *__restrict__ rank_of,
              const uint32_t *__restrict__ tlevel, const unsigned int *__restrict__ minlev,
              unsigned long long *__restrict__ keys, unsigned int *err) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)n_cells * n_dl) return;
    const uint32_t dl = (uint32_t)(t / n_cells), c = (uint32_t)(t - (size_t)dl * n_cells);
    const uint32_t me = mtid_of[((size_t)group_of[dl] * n_patches + patch_of[c]) * n_phase + phase[t]];
    uint32_t sub = tlevel[t] - minlev[(size_t)me * kd_max + group_rank[dl]];
    if (sub >= kMaxSub) { atomicExch(err, 1u); sub = kMaxSub - 1; }
    keys[t] = ((unsigned long long)rank_of[me] << kRankShift) | ((unsigned long long)sub << kSubShift) |
              (unsigned long long)(c * n_dl + dl);
}

__global__ void __launch_bounds__(256)
p_key_kernel(uint32_t n_cells, uint32_t n_dl, const uint32_t *__restrict__ patch_of, const uint16_t *__restrict__ group_of,
             uint32_t n_patches, const uint32_t *__restrict__ rank_of, const uint32_t *__restrict__ tlevel,
             const unsigned int *__restrict__ minlev, unsigned long long *__restrict__ keys, unsigned int *err) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)n_cells * n_dl) return;
    const uint32_t dl = (uint32_t)(t / n_cells), c = (uint32_t)(t - (size_t)dl * n_cells);
    const uint32_t pc = patch_of[c];
    uint32_t sub = tlevel[t] - minlev[(size_t)pc * n_dl + dl];
    if (sub >= kMaxSub) { atomicExch(err, 1u); sub = kMaxSub - 1; }
    const uint32_t r = rank_of[(size_t)group_of[dl] * n_patches + pc];
    keys[t] = ((unsigned long long)r << kRankShift) | ((unsigned long long)sub << kSubShift) |
              (unsigned long long)(c * n_dl + dl);
}

// boundaries in the sorted key list: pseudo-level = (macro-tile, sub-level) starts, macro-tile starts; 32-bit keys
__global__ void __launch_bounds__(256)
p_bounds_kernel(const unsigned long long *__restrict__ keys, uint32_t n, uint8_t *__restrict__ pl_flag,
                uint32_t *__restrict__ mt_slot0, uint32_t *__restrict__ k32) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const unsigned long long k = keys[s];
    const unsigned long long prev = s ? keys[s - 1] : ~k;
    pl_flag[s] = (k >> kSubShift) != (prev >> kSubShift) ? 1 : 0;
    if ((k >> kRankShift) != (prev >> kRankShift)) mt_slot0[(uint32_t)(k >> kRankShift)] = s;
    k32[s] = (uint32_t)k;
}

__global__ void __launch_bounds__(256)
p_pl_rank_kernel(const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ pl_start, uint32_t n_pl,
                 uint32_t *__restrict__ pl_rank) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_pl) pl_rank[i] = (uint32_t)(keys[pl_start[i]] >> kRankShift);
}

// per slot: upwind entries (Local + periodic), external entries (Local into another patch + periodic),
// total downwind effective area; counters[0] += periodic entries, counters[2] = error (> 255 periodic faces)
__global__ void __launch_bounds__(256)
p_count_kernel(GridView g, const uint32_t *__restrict__ k32, uint32_t n, uint32_t n_dl,
               const uint32_t *__restrict__ patch_of, const uint8_t *__restrict__ phase, uint32_t *__restrict__ cnt_e,
               uint32_t *__restrict__ cnt_x, double *__restrict__ ttot_slot, unsigned int *counters) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint32_t k = k32[s];
    const uint32_t c = k / n_dl, dl = k - c * n_dl;
    const uint32_t pc = patch_of[c];
    const uint8_t *const ph = phase ? phase + (size_t)dl * g.n_cells : nullptr;   // phases of this direction's tasks
    const double dx = g.dirs[3 * dl], dy = g.dirs[3 * dl + 1], dz = g.dirs[3 * dl + 2];
    uint32_t m = 0, x = 0, np = 0;
    double ttot = 0.0;
    for (uint32_t f = g.face_off[c]; f < g.face_off[c + 1]; ++f) {
        const double4 geo = ld_geo(g.face_geo + f);
        const double d = dot_dir(geo, dx, dy, dz);
        const int kind = g.face_kind[f];
        if (d < 0.0) {
            if (kind == 0) {
                ++m;
                const uint32_t nb = (uint32_t)g.face_nb[f];
                if (patch_of[nb] != pc || (ph && ph[nb] != ph[c])) ++x;   // not the same macro-tile: external
            } else if (kind == 2) {
                ++m; ++x; ++np;
            }
        } else if (d > 0.0) {
            ttot += geo.w * d;
        }
    }
    cnt_e[s] = m;
    cnt_x[s] = x;
    ttot_slot[s] = ttot;
    if (np) atomicAdd(counters, np);
    if (np > 255u) atomicExch(counters + 2, 1u);
}

struct CastU64 {
    __host__ __device__ unsigned long long operator()(uint32_t v) const { return (unsigned long long)v; }
};

struct PFillArgs {
    GridView g;
    PatchGrid pg;
    const uint32_t *k32, *slot_of;
    const unsigned long long *upoff, *xoff;   // n + 1 each
    const double *ttot_slot;
    uint32_t n_dl, n_tasks, stages;
    const PDesc *ptab;            // block-major
    const uint32_t *ptab_block;   // block of packet i
    const uint32_t *tab_off;      // per block
    const uint64_t *stream_off;   // per block
    unsigned char *stream;
    const uint32_t *tile_start;   // n_tiles + 1
    const uint32_t *tile_rank;    // n_tiles
    const uint32_t *mt_slot0;     // n_mt + 1
    const uint32_t *mt_group, *mt_patch, *mt_ndep;   // n_mt
    const uint16_t *group_rank;   // n_dl: index of the direction inside its group
    const uint32_t *group_kd;     // G: directions in the group
    const uint64_t *head_off;     // n_mt: byte offset of the macro-tile's head packet in the stream
    const unsigned int *dep_tab;
    const uint32_t *rank_of;
    const uint32_t *mt_row;       // n_mt: row of the macro-tile in dep_tab
    const uint32_t *dep_base;     // n_mt: added to a dep_tab entry before the rank_of lookup
    const uint8_t *phase;         // per task (dl * N + c), or null: no phases
    uint32_t *lag_src;
    unsigned int *counters;       // [1] lag cursor, [3] ordering violations
};

// one thread block per packet
__global__ void __launch_bounds__(256)
p_fill_kernel(PFillArgs a) {
    const PDesc d = a.ptab[blockIdx.x];
    const uint32_t blk = a.ptab_block[blockIdx.x];
    unsigned char *pkt = a.stream + a.stream_off[blk] + (size_t)d.off16 * 16u;
    const uint32_t tid = threadIdx.x;
    PHdr h;
    h.next_off16 = 0; h.next_bytes = 0; h.gslot0 = 0; h.n_slots = 0; h.rank = 0;
    const uint32_t nxt = blockIdx.x + a.stages;
    if (nxt < a.tab_off[blk + 1]) {
        const PDesc nd = a.ptab[nxt];
        h.next_off16 = nd.off16;
        h.next_bytes = nd.bytes;
    }
    if (d.flags & kPHead) {
        const uint32_t r = d.id;
        const uint32_t grp = a.mt_group[r], p = a.mt_patch[r], n_dep = a.mt_ndep[r];
        const uint32_t c0 = a.pg.patch_off[p], n_cells = a.pg.patch_off[p + 1] - c0;
        const uint32_t mt0 = a.mt_slot0[r], mt1 = a.mt_slot0[r + 1];
        const uint32_t n_ext = (uint32_t)(a.xoff[mt1] - a.xoff[mt0]);
        const HeadLayout L = head_layout(n_dep, n_ext, n_cells);
        uint32_t *dep = reinterpret_cast<uint32_t *>(pkt + L.dep);
        uint32_t *ext = reinterpret_cast<uint32_t *>(pkt + L.ext);
        uint32_t *cells = reinterpret_cast<uint32_t *>(pkt + L.cells);
        const unsigned int *row = a.dep_tab + (size_t)a.mt_row[r] * kMaxPatchDeps;
        const uint32_t dep_pad = align16(4u * n_dep) / 4u;
        for (uint32_t i = tid; i < dep_pad; i += blockDim.x)
            dep[i] = i < n_dep ? a.rank_of[(size_t)a.dep_base[r] + row[i]] : 0u;
        const uint32_t ext_pad = align16(4u * n_ext) / 4u;
        for (uint32_t i = n_ext + tid; i < ext_pad; i += blockDim.x) ext[i] = 0u;   // entries proper: tile blocks
        const uint32_t cell_pad = align16(4u * n_cells) / 4u;
        for (uint32_t i = tid; i < cell_pad; i += blockDim.x) cells[i] = i < n_cells ? a.pg.patch_cells[c0 + i] : 0u;
        if (tid == 0) {
            h.kind = 1; h.n = (uint16_t)n_dep; h.a16 = (uint16_t)n_cells; h.b16 = (uint16_t)(grp | (a.group_kd[grp] << 10)); h.c32 = n_ext;
            h.gslot0 = mt0; h.n_slots = mt1 - mt0; h.rank = r;
            *reinterpret_cast<PHdr *>(pkt) = h;
        }
        return;
    }
    const uint32_t t = d.id;
    const uint32_t slot0 = a.tile_start[t], n = a.tile_start[t + 1] - slot0;
    const unsigned long long e_base = a.upoff[slot0];
    const uint32_t E = (uint32_t)(a.upoff[slot0 + n] - e_base);
    const uint32_t r = a.tile_rank[t];
    const uint32_t mt0 = a.mt_slot0[r], ns_mt = a.mt_slot0[r + 1] - mt0;
    const PTileLayout L = ptile_layout(n, E);
    double *w = reinterpret_cast<double *>(pkt + L.w);
    uint16_t *idx = reinterpret_cast<uint16_t *>(pkt + L.idx);
    uint16_t *lcell = reinterpret_cast<uint16_t *>(pkt + L.lcell);
    uint32_t *info = reinterpret_cast<uint32_t *>(pkt + L.info);
    uint32_t *hext = reinterpret_cast<uint32_t *>(a.stream + a.head_off[r] + sizeof(PHdr) + align16(4u * a.mt_ndep[r]));
    // zero the padding so the stream is fully initialised
    if (tid < 8) {
        if (tid == 0 && (E & 1u)) w[E] = 0.0;
        const uint32_t pad_idx = (align16(2u * E) - 2u * E) / 2u;
        if (tid < pad_idx) idx[E + tid] = 0;
        const uint32_t pad_lc = (align16(2u * n) - 2u * n) / 2u;
        if (tid < pad_lc) lcell[n + tid] = 0xffffu;
        const uint32_t used = n + 1u, pad_info = (align16(4u * used) - 4u * used) / 4u;
        if (tid < pad_info) info[used + tid] = 0u;
    }
    // A tile whose slots all have the same number K <= 4 of Local upwind entries and no periodic ones (every interior
    // tile of a Cartesian grid: K = 3) is flagged in its header: the kernel then indexes the entries as K * slot and
    // never reads the info words.
    __shared__ uint32_t s_deg0;
    uint32_t my_deg = 0, my_per = 0;
    if (tid < n) {
        const uint32_t s = slot0 + tid;
        const uint32_t k = a.k32[s];
        const uint32_t c = k / a.n_dl, dl = k - c * a.n_dl;
        const uint32_t pc = a.pg.patch_of[c];
        const uint8_t *const ph = a.phase ? a.phase + (size_t)dl * a.g.n_cells : nullptr;
        lcell[tid] = (uint16_t)(a.pg.lidx[c] | ((uint32_t)a.group_rank[dl] << 10));
        const uint32_t e0 = (uint32_t)(a.upoff[s] - e_base);
        uint32_t e = e0, n_per = 0;
        uint32_t x = (uint32_t)(a.xoff[s] - a.xoff[mt0]);
        const double dx = a.g.dirs[3 * dl], dy = a.g.dirs[3 * dl + 1], dz = a.g.dirs[3 * dl + 2];
        for (int pass = 0; pass < 2; ++pass) {   // Local faces first, then the periodic ones
            for (uint32_t f = a.g.face_off[c]; f < a.g.face_off[c + 1]; ++f) {
                if (a.g.face_kind[f] != (pass ? 2 : 0)) continue;
                const double dd = dot_dir(ld_geo(a.g.face_geo + f), dx, dy, dz);
                if (!(dd < 0.0)) continue;
                const uint32_t nb = (uint32_t)a.g.face_nb[f];
                const uint32_t src = a.slot_of[(size_t)dl * a.g.n_cells + nb];
                const double tt = a.ttot_slot[src];
                const double share = tt > 0.0 ? (a.g.face_rev[f] * (-dd)) / tt : 0.0;
                uint32_t vi;
                if (pass == 0 && a.pg.patch_of[nb] == pc && (!ph || ph[nb] == ph[c])) {
                    vi = src - mt0;   // a slot of this macro-tile in an earlier sub-level
                    if (src < mt0 || src >= slot0) atomicExch(a.counters + 3, 1u);
                } else {
                    uint32_t gsrc = src;
                    if (pass) {   // periodic: the donor's pre-sweep snapshot (the reference's lag)
                        ++n_per;
                        const unsigned int j = atomicAdd(a.counters + 1, 1u);
                        a.lag_src[j] = src;
                        gsrc = a.n_tasks + j;
                    }
                    hext[x] = gsrc;
                    vi = ns_mt + x;
                    ++x;
                }
                idx[e] = (uint16_t)vi;
                w[e] = share;
                ++e;
            }
        }
        info[tid] = e0 | (min(n_per, 255u) << 16);
        my_deg = e - e0;
        my_per = n_per;
        if (tid == 0) s_deg0 = my_deg;
    }
    __syncthreads();
    const bool uniform = __syncthreads_and(tid >= n || (my_per == 0u && my_deg == s_deg0)) != 0;
    if (tid == 0) {
        info[n] = E;
        h.kind = 0; h.n = (uint16_t)n; h.a16 = (uint16_t)E;
        const uint32_t K = uniform && s_deg0 >= 1u && s_deg0 <= 4u ? s_deg0 : 0u;
        h.b16 = (uint16_t)(((d.flags & kPLast) ? 1u : 0u) | (K << 1));
        h.c32 = slot0 - mt0;
        h.gslot0 = (L.idx >> 4) | ((L.lcell >> 4) << 16);   // tile packets: section offsets in 16-B units
        h.n_slots = L.info >> 4;
        *reinterpret_cast<PHdr *>(pkt) = h;
    }
}

// sum_d periodic_source of the NEW outgoing rates (site.rs:53-56; handle_local_periodic_neighbour,
// src/sweep/mod.rs:505-513, in gather form).  The periodic upwind entries of every periodic cell are static:
// they are listed once (donor slot, share), ordered by (periodic cell, direction, face), and summed after
// every sweep by one warp per periodic cell (strided partial sums, one xor-shuffle tree: deterministic).
// fill == false: cnt[p] = number of entries; fill == true: writes the entries at off[p]
__global__ void __launch_bounds__(128)
p_periodic_list_kernel(GridView g, const uint32_t *__restrict__ pcells, uint32_t n_periodic, uint32_t n_dl,
                       const uint32_t *__restrict__ slot_of, const double *__restrict__ ttot_slot, bool fill,
                       uint32_t *__restrict__ cnt, const uint32_t *__restrict__ off, uint32_t *__restrict__ src_out,
                       double *__restrict__ w_out) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_periodic) return;
    const uint32_t c = pcells[p];
    const uint32_t f0 = g.face_off[c], f1 = g.face_off[c + 1];
    uint32_t m = fill ? off[p] : 0u;
    for (uint32_t dl = 0; dl < n_dl; ++dl) {
        const double dx = g.dirs[3 * dl], dy = g.dirs[3 * dl + 1], dz = g.dirs[3 * dl + 2];
        for (uint32_t f = f0; f < f1; ++f) {
            if (g.face_kind[f] != 2) continue;
            const double dd = dot_dir(ld_geo(g.face_geo + f), dx, dy, dz);
            if (!(dd < 0.0)) continue;
            if (fill) {
                const uint32_t src = slot_of[(size_t)dl * g.n_cells + (uint32_t)g.face_nb[f]];
                const double tt = ttot_slot[src];
                src_out[m] = src;
                w_out[m] = tt > 0.0 ? (g.face_rev[f] * (-dd)) / tt : 0.0;
            }
            ++m;
        }
    }
    if (!fill) cnt[p] = m;
}

__global__ void __launch_bounds__(256)
p_periodic_rate_kernel(const uint32_t *__restrict__ off, const uint32_t *__restrict__ src, const double *__restrict__ w,
                       uint32_t n_periodic, const double *__restrict__ out_slot, double *__restrict__ acc_per) {
    const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (wid >= n_periodic) return;
    const uint32_t e0 = off[wid], e1 = off[wid + 1];
    double sum = 0.0;
    for (uint32_t e = e0 + lane; e < e1; e += 32) sum += __ldcg(out_slot + src[e]) * w[e];
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) acc_per[wid] = sum;
}

// ---- the sweep kernel ---------------------------------------------------------------------------------
__device__ __forceinline__ void st_release_gpu(unsigned int *p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

struct PatchArgs {
    const unsigned char *stream;
    const uint64_t *stream_off;
    const PDesc *ptab;
    const uint32_t *tab_off;
    unsigned int *mt_flag;
    double *out_slot;
    const double2 *cellrec;   // {exp(-n_HI sigma size), source / D} per cell
    double *acc_cell;         // G x N
    double threshold;
    uint32_t stages, stage_bytes, vmax, pc_max, smax;
    uint32_t n_cells, epoch, poll_ns;
    uint32_t accumulate;      // phases: a (group, cell) rate row is split over several macro-tiles -> add instead of store
    unsigned long long *prof; // optional per-block cycle counters, 10 per block (SSW_STREAM_PROFILE)
};

// One-warp blocks (THREADS == 32) separate the sub-levels of a macro-tile with a warp barrier instead of a block barrier:
// nothing on the dependent chain of a direction shard waits for another warp.
template <int THREADS>
__device__ __forceinline__ void patch_block_sync() {
    if constexpr (THREADS == 32) __syncwarp(); else __syncthreads();
}

template <int THREADS, int MIN_BLOCKS, bool PROFILE>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
patch_sweep_kernel(PatchArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t stages = a.stages, stage_bytes = a.stage_bytes;
    const PatchSmem SL = patch_smem(stages, stage_bytes, a.vmax, a.pc_max, a.smax);
    unsigned char *const ring = smem;
    uint64_t *const full = reinterpret_cast<uint64_t *>(smem + SL.bars);
    double *const val = reinterpret_cast<double *>(smem + SL.val);
    double2 *const s_rec = reinterpret_cast<double2 *>(smem + SL.rec);
    double *const s_inc = reinterpret_cast<double *>(smem + SL.inc);
    uint32_t *s_cellid = reinterpret_cast<uint32_t *>(smem + SL.cellid);
    const uint32_t cellid_stride = align16(4u * a.pc_max) / 4u;
    uint32_t cellid_buf = 0;
    const uint32_t tid = threadIdx.x;
    const uint32_t n_my = a.tab_off[blockIdx.x + 1] - a.tab_off[blockIdx.x];
    const unsigned char *const stream = a.stream + a.stream_off[blockIdx.x];
    double *const out_slot = a.out_slot;
    const double threshold = a.threshold;
    uint64_t policy = 0;
    if (tid == THREADS - 32) {   // one thread issues every TMA copy: lane 0 of the last warp, the warp most often without slots
        for (uint32_t s = 0; s < stages; ++s) mbar_init(smem_u32(full + s), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        policy = policy_evict_first();
    }
    patch_block_sync<THREADS>();
    if (tid == THREADS - 32) {
        const PDesc *tab = a.ptab + a.tab_off[blockIdx.x];
        const uint32_t pre = min(stages, n_my);
        for (uint32_t k = 0; k < pre; ++k) {
            const PDesc d = tab[k];
            mbar_expect_tx(smem_u32(full + k), d.bytes);
            tma_bulk_load(smem_u32(ring + (size_t)k * stage_bytes), stream + (size_t)d.off16 * 16u, d.bytes,
                          smem_u32(full + k), policy);
        }
    }
    uint32_t gslot0 = 0, n_slots = 0, group = 0, rank = 0, n_cells = 0, kdg = 1;
    uint32_t stage = 0, parity = 0;
    long long t_begin = 0, t_poll = 0, t_pkt = 0, tp = 0, t_cmp = 0, t_bar = 0, t_post = 0, t_head = 0, tq = 0;
    if (PROFILE && tid == 0) t_begin = clock64();
    for (uint32_t k = 0; k < n_my; ++k) {
        if (PROFILE && tid == 0) tp = clock64();
        mbar_wait(smem_u32(full + stage), parity);
        if (PROFILE && tid == 0) t_pkt += clock64() - tp;
        unsigned char *const pkt = ring + (size_t)stage * stage_bytes;
        const PHdr h = *reinterpret_cast<const PHdr *>(pkt);
        if (PROFILE && tid == 0) tq = clock64();
        if (h.kind) {
            // ---- macro-tile head: stage the patch's cell data, wait for the upwind macro-tiles, gather the
            //      fluxes that enter the patch
            gslot0 = h.gslot0; n_slots = h.n_slots; rank = h.rank; group = h.b16 & 0x3ffu; kdg = h.b16 >> 10; n_cells = h.a16;
            const uint32_t n_dep = h.n, n_ext = h.c32;
            const HeadLayout L = head_layout(n_dep, n_ext, n_cells);
            const uint32_t *const dep = reinterpret_cast<const uint32_t *>(pkt + L.dep);
            const uint32_t *const ext = reinterpret_cast<const uint32_t *>(pkt + L.ext);
            const uint32_t *const cells = reinterpret_cast<const uint32_t *>(pkt + L.cells);
            cellid_buf ^= 1u;
            s_cellid = reinterpret_cast<uint32_t *>(smem + SL.cellid) + cellid_buf * cellid_stride;
            if constexpr (THREADS == 32) {
                // one warp: stage first (the cell records do not depend on the upwind macro-tiles; the loads of a round
                // are independent), then poll
                for (uint32_t i0 = 0; i0 < n_cells; i0 += 8u * 32u) {
                    uint32_t cc[8];
                    double2 rr[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint32_t i = i0 + (uint32_t)j * 32u + tid;
                        cc[j] = i < n_cells ? cells[i] : 0u;
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) rr[j] = __ldg(a.cellrec + cc[j]);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint32_t i = i0 + (uint32_t)j * 32u + tid;
                        if (i < n_cells) { s_cellid[i] = cc[j]; s_rec[i] = rr[j]; }
                    }
                }
                if (a.accumulate)
                    for (uint32_t i = tid; i < n_cells * kdg; i += 32u) s_inc[i] = 0.0;
                if (tid < n_dep) {
                    if (PROFILE && tid == 0) tp = clock64();
                    for (uint32_t i = tid; i < n_dep; i += 32u) {
                        const unsigned int *flag = a.mt_flag + dep[i];
                        while (ld_relaxed_gpu(flag) != a.epoch) __nanosleep(a.poll_ns);
                    }
                    fence_acq_rel_gpu();
                    if (PROFILE && tid == 0) t_poll += clock64() - tp;
                }
            } else if (tid < 32u) {
                // warp 0 only polls: the flag round trips start at once and run beside the staging of the other warps
                if (tid < n_dep) {
                    if (PROFILE && tid == 0) tp = clock64();
                    // relaxed polls (no L1 invalidation per round trip), one acquire fence once the flags are there
                    for (uint32_t i = tid; i < n_dep; i += 32u) {
                        const unsigned int *flag = a.mt_flag + dep[i];
                        while (ld_relaxed_gpu(flag) != a.epoch) __nanosleep(a.poll_ns);
                    }
                    fence_acq_rel_gpu();
                    if (PROFILE && tid == 0) t_poll += clock64() - tp;
                }
            } else {
                for (uint32_t i = tid - 32u; i < n_cells; i += THREADS - 32u) {
                    const uint32_t c = cells[i];
                    s_cellid[i] = c;
                    s_rec[i] = __ldg(a.cellrec + c);
                }
                if (a.accumulate)   // ragged macro-tile: not every (cell, direction) of the patch has a task here
                    for (uint32_t i = tid - 32u; i < n_cells * kdg; i += THREADS - 32u) s_inc[i] = 0.0;
            }
            patch_block_sync<THREADS>();
            double *const vx = val + n_slots;
            for (uint32_t i = tid; i < n_ext; i += 8u * THREADS) {   // eight independent gathers in flight per thread
                double v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t ij = i + (uint32_t)j * THREADS;
                    v[j] = ij < n_ext ? __ldcg(out_slot + ext[ij]) : 0.0;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t ij = i + (uint32_t)j * THREADS;
                    if (ij < n_ext) vx[ij] = v[j];
                }
            }
            patch_block_sync<THREADS>();   // external values visible; every thread is done with the stage
            if (tid == THREADS - 32 && h.next_bytes) {
                mbar_expect_tx(smem_u32(full + stage), h.next_bytes);
                tma_bulk_load(smem_u32(pkt), stream + (size_t)h.next_off16 * 16u, h.next_bytes, smem_u32(full + stage), policy);
            }
        } else {
            // ---- one tile: <= THREADS tasks of one sub-level; every value it reads is in shared memory
            const uint32_t n = h.n, lslot0 = h.c32;
            const double *const w = reinterpret_cast<const double *>(pkt + sizeof(PHdr));
            const uint16_t *const idx = reinterpret_cast<const uint16_t *>(pkt + ((h.gslot0 & 0xffffu) << 4));
            const uint16_t *const lcell = reinterpret_cast<const uint16_t *>(pkt + ((h.gslot0 >> 16) << 4));
            const uint32_t *const info = reinterpret_cast<const uint32_t *>(pkt + (h.n_slots << 4));
            const uint32_t uniform_k = (h.b16 >> 1) & 7u;
            if (tid < n) {   // (warps beyond the tile's slots only keep the barrier)
                const uint32_t lcj = lcell[tid];   // patch-local cell | index of the direction inside its group << 10
                const uint32_t lc = lcj & 0x3ffu;
                const double2 rec = s_rec[lc];
                double in_loc = 0.0, in_per = 0.0;
                // product and sum rounded separately, Local faces in face order, then the periodic ones: the
                // arithmetic of stream.cuh bit for bit.
                if (uniform_k == 3u) {
                    // every slot of the tile has exactly three Local entries (interior of a Cartesian grid): no info
                    // words, no predicates, all six loads independent
                    const uint32_t e = 3u * tid;
                    const uint32_t i0 = idx[e], i1 = idx[e + 1u], i2 = idx[e + 2u];
                    const double w0 = w[e], w1 = w[e + 1u], w2 = w[e + 2u];
                    const double v0 = val[i0], v1 = val[i1], v2 = val[i2];
                    in_loc = __dadd_rn(__dadd_rn(__dadd_rn(0.0, __dmul_rn(v0, w0)), __dmul_rn(v1, w1)), __dmul_rn(v2, w2));
                } else if (uniform_k) {
                    uint32_t e = uniform_k * tid;
                    const uint32_t em = e + uniform_k;
#pragma unroll 1
                    for (; e < em; ++e) in_loc = __dadd_rn(in_loc, __dmul_rn(val[idx[e]], w[e]));
                } else {
                const uint32_t inf = info[tid];
                const uint32_t e1 = info[tid + 1] & 0xffffu;
                uint32_t e = inf & 0xffffu;
                const uint32_t em = e1 - ((inf >> 16) & 0xffu);
                // Four entries per round: their loads are independent, only the additions form a chain.
#pragma unroll 1
                for (; e < em; e += 4u) {
                    uint32_t vi[4];
                    double wv[4], vv[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const bool ok = e + j < em;
                        vi[j] = ok ? idx[e + j] : 0u;
                        wv[j] = ok ? w[e + j] : 0.0;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) vv[j] = e + j < em ? val[vi[j]] : 0.0;   // (no dummy reads: racecheck-clean)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (e + j < em) in_loc = __dadd_rn(in_loc, __dmul_rn(vv[j], wv[j]));
                }
#pragma unroll 1
                for (uint32_t ep = em; ep < e1; ++ep) in_per = __dadd_rn(in_per, __dmul_rn(val[idx[ep]], w[ep]));
                }
                const double total = (in_loc + rec.y) + in_per;         // site.rs:49-56
                // HydrogenOnly::get_outgoing_rate, hydrogen_only/mod.rs:81-87
                const double out = (total < threshold) ? 0.0 : total * rec.x;
                val[lslot0 + tid] = out;
                __stcg(out_slot + gslot0 + lslot0 + tid, out);
                s_inc[lc * kdg + (lcj >> 10)] = in_loc;                 // incoming_total_rate[d], summed per cell when the macro-tile is done
            }
            if (PROFILE && tid == 0) { const long long t = clock64(); t_cmp += t - tq; tq = t; }
            patch_block_sync<THREADS>();   // this sub-level's rates are visible; every thread is done with the stage
            if (PROFILE && tid == 0) { const long long t = clock64(); t_bar += t - tq; tq = t; }
            // the ring stage was only read (generic proxy), never written: no proxy fence before the refill
            if ((tid >> 5) == (uint32_t)(THREADS / 32 - 1)) {   // warp-uniform: the other warps branch around the issue code
                if (tid == THREADS - 32 && h.next_bytes) {
                    mbar_expect_tx(smem_u32(full + stage), h.next_bytes);
                    tma_bulk_load(smem_u32(pkt), stream + (size_t)h.next_off16 * 16u, h.next_bytes, smem_u32(full + stage), policy);
                }
            }
            if (h.b16 & 1u) {
                // ---- macro-tile done.  Every outgoing rate of the macro-tile was stored before the barrier above, so the
                //      done flag goes out first (the release is cumulative over the block's stores; downwind macro-tiles
                //      are waiting for it).  Then sum_d incoming of the group's directions per cell, in direction order
                //      (src/sweep/mod.rs:554-558): plain stores, every (group, cell) is written exactly once per sweep
                double *const acc = a.acc_cell + (size_t)group * a.n_cells;
                if (!a.accumulate) {
                    if (tid == 0) st_release_gpu(a.mt_flag + rank, a.epoch);
                    for (uint32_t i = tid; i < n_cells; i += THREADS) {
                        const double *const row = s_inc + i * kdg;
                        double sum = 0.0;
                        for (uint32_t j = 0; j < kdg; ++j) sum += row[j];
                        __stcg(acc + s_cellid[i], sum);
                    }
                } else {
                    // phases of one (patch, group) are chained through their done flags, so the read-modify-write of
                    // the row is ordered; it has to be complete before the flag goes out
                    for (uint32_t i = tid; i < n_cells; i += THREADS) {
                        const double *const row = s_inc + i * kdg;
                        double sum = 0.0;
                        for (uint32_t j = 0; j < kdg; ++j) sum += row[j];
                        double *const dst = acc + s_cellid[i];
                        __stcg(dst, __ldcg(dst) + sum);
                    }
                    patch_block_sync<THREADS>();
                    if (tid == 0) st_release_gpu(a.mt_flag + rank, a.epoch);
                }
                // no barrier: the next head fills the other s_cellid buffer, and two barriers separate it from the
                // next write to s_inc
            }
        }
        if (PROFILE && tid == 0) { if (h.kind) t_head += clock64() - tq; else t_post += clock64() - tq; }
        if (++stage == stages) { stage = 0; parity ^= 1u; }
    }
    if (PROFILE && tid == 0) {
        a.prof[10 * blockIdx.x + 0] = (unsigned long long)(clock64() - t_begin);
        a.prof[10 * blockIdx.x + 1] = (unsigned long long)t_poll;
        a.prof[10 * blockIdx.x + 2] = (unsigned long long)t_pkt;
        a.prof[10 * blockIdx.x + 3] = n_my;
        a.prof[10 * blockIdx.x + 4] = (unsigned long long)t_cmp;
        a.prof[10 * blockIdx.x + 6] = (unsigned long long)t_bar;
        a.prof[10 * blockIdx.x + 7] = (unsigned long long)t_post;
        a.prof[10 * blockIdx.x + 8] = (unsigned long long)t_head;
    }
}

typedef void (*PatchKernel)(PatchArgs);
inline PatchKernel patch_kernel_for(uint32_t threads, bool profile) {
    // minimum blocks per SM chosen so that the register file never limits residency below 1024 threads (<= 64 registers)
    if (threads == 32) return profile ? patch_sweep_kernel<32, 16, true> : patch_sweep_kernel<32, 16, false>;
    if (profile) return threads == 64 ? patch_sweep_kernel<64, 16, true> : threads == 128 ? patch_sweep_kernel<128, 8, true> : patch_sweep_kernel<256, 4, true>;
    return threads == 64 ? patch_sweep_kernel<64, 16, false> : threads == 128 ? patch_sweep_kernel<128, 8, false> : patch_sweep_kernel<256, 4, false>;
}

template <class T>
struct DTmp {
    T *p = nullptr;
    DTmp() = default;
    DTmp(const DTmp &) = delete;
    DTmp &operator=(const DTmp &) = delete;
    ~DTmp() { reset(); }
    void reset() { if (p) cudaFree(p); p = nullptr; }
    void alloc(size_t n, const char *what) {
        reset();
        cuda_ok(cudaMalloc(&p, sizeof(T) * std::max<size_t>(n, 1)), what);
    }
    void upload(const std::vector<T> &v, cudaStream_t s, const char *what) {
        alloc(v.size(), what);
        if (!v.empty()) cuda_ok(cudaMemcpyAsync(p, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice, s), what);
    }
};

// direction groups: directions of one octant (sign pattern of the direction vector), at most kd per group
inline uint32_t make_direction_groups(const double *dirs_local, uint32_t n_dl, uint32_t kd, std::vector<uint16_t> &group_of,
                                      std::vector<uint16_t> &group_rank, std::vector<uint32_t> &group_kd) {
    group_of.assign(n_dl, 0);
    group_rank.assign(n_dl, 0);
    group_kd.clear();
    std::vector<std::vector<uint32_t>> cls(27);
    for (uint32_t dl = 0; dl < n_dl; ++dl) {
        int code = 0;
        for (int k = 0; k < 3; ++k) {
            const double v = dirs_local[3 * dl + k];
            code = code * 3 + (v > 0.0 ? 2 : (v < 0.0 ? 0 : 1));
        }
        cls[code].push_back(dl);
    }
    uint32_t G = 0;
    for (auto &c : cls) {
        if (c.empty()) continue;
        const uint32_t chunks = ((uint32_t)c.size() + kd - 1) / kd;
        group_kd.resize(G + chunks, 0);
        for (uint32_t i = 0; i < c.size(); ++i) {
            const uint32_t grp = G + (uint32_t)((uint64_t)i * chunks / c.size());
            group_of[c[i]] = (uint16_t)grp;
            group_rank[c[i]] = (uint16_t)group_kd[grp]++;   // c is ascending in dl: rank = direction order inside the group
        }
        G += chunks;
    }
    return G;
}

// Levels of the macro-tiles: Kahn's algorithm over the quotient graph of every group.  dep[(grp * P + p) * 64 ..] lists
// the upwind patches of patch p (kEmptyDep-terminated).  level = 1 + max level of the upwind macro-tiles (0: none).
// Returns false if some group's graph has a cycle.  Pure host code (also behind ssw_patch_levels for the CPU tests).
inline bool level_macro_tiles(const unsigned int *dep, uint32_t G, uint32_t P, std::vector<uint32_t> &mt_level,
                              std::vector<uint32_t> &ndep, uint32_t &max_level) {
    mt_level.assign((size_t)G * P, 0);
    ndep.assign((size_t)G * P, 0);
    max_level = 0;
    for (uint32_t grp = 0; grp < G; ++grp) {
        std::vector<uint32_t> indeg(P, 0), succ_off(P + 1, 0), succ, order;
        for (uint32_t p = 0; p < P; ++p) {
            const unsigned int *row = dep + ((size_t)grp * P + p) * kMaxPatchDeps;
            uint32_t k = 0;
            while (k < kMaxPatchDeps && row[k] != kEmptyDep) {
                if (row[k] >= P) return false;
                succ_off[row[k] + 1]++;
                ++k;
            }
            indeg[p] = k;
            ndep[(size_t)grp * P + p] = k;
        }
        for (uint32_t p = 0; p < P; ++p) succ_off[p + 1] += succ_off[p];
        succ.resize(succ_off[P]);
        std::vector<uint32_t> cur(succ_off.begin(), succ_off.end() - 1);
        for (uint32_t p = 0; p < P; ++p) {
            const unsigned int *row = dep + ((size_t)grp * P + p) * kMaxPatchDeps;
            for (uint32_t k = 0; k < indeg[p]; ++k) succ[cur[row[k]]++] = p;
        }
        order.reserve(P);
        for (uint32_t p = 0; p < P; ++p) if (indeg[p] == 0) order.push_back(p);
        for (size_t i = 0; i < order.size(); ++i) {
            const uint32_t p = order[i];
            for (uint32_t j = succ_off[p]; j < succ_off[p + 1]; ++j) {
                const uint32_t q = succ[j];
                uint32_t &lv = mt_level[(size_t)grp * P + q];
                lv = std::max(lv, mt_level[(size_t)grp * P + p] + 1);
                if (--indeg[q] == 0) order.push_back(q);
            }
        }
        if (order.size() != P) return false;
        for (uint32_t p = 0; p < P; ++p) max_level = std::max(max_level, mt_level[(size_t)grp * P + p]);
    }
    return true;
}

// Builds the patch-ordered schedule from the level-sorted task list of the all-cells sweep.  Throws
// PatchUnsupported when the grid does not admit the form (the caller falls back to compile_schedule).
inline void compile_patch_schedule(Compiled &C, const GridView &g, const PatchGrid &pg, const double *dirs_local,
                                   const uint32_t *tasks, const uint32_t *level_off_dev, uint64_t n_tasks,
                                   uint32_t n_levels, int n_local_dirs, const uint32_t *pcells, uint32_t n_periodic,
                                   const double *q_nat, const std::vector<uint32_t> *level_off_host,
                                   int num_sms, cudaStream_t stream, uint64_t *launch_counter) {
    C.release();
    if (n_tasks >= 0x7fffff00ull) throw PatchUnsupported("more than 2^31 tasks per rank");
    if (n_local_dirs > 128) throw PatchUnsupported("more than 128 local directions");
    if (n_tasks != (uint64_t)g.n_cells * (uint64_t)n_local_dirs) throw PatchUnsupported("not an all-cells schedule");
    const uint32_t n = (uint32_t)n_tasks, n_dl = (uint32_t)n_local_dirs, N = g.n_cells, P = pg.n_patches;
    const uint32_t threads_env = env_u32("SSW_PATCH_THREADS", 128);
    const uint32_t threads = threads_env == 256 ? 256u : (threads_env == 64 ? 64u : (threads_env == 32 ? 32u : 128u));
    const uint32_t kd_default = n_dl <= 24 ? 3u : 11u;   // many directions: one group per octant (10-11 of the 84)
    const uint32_t kd = std::max<uint32_t>(1u, std::min<uint32_t>(env_u32("SSW_PATCH_KD", kd_default), 32u));
    const uint32_t want_stages = std::max<uint32_t>(2u, std::min<uint32_t>(env_u32("SSW_PATCH_STAGES", 3), kMaxStages));
    const unsigned blocks_n = (unsigned)((n + 255) / 256);
    uint64_t launches = 0;
    try {
        // 1. direction groups, wavefront level of every task, quotient graph
        std::vector<uint16_t> group_of, group_rank;
        std::vector<uint32_t> group_kd;
        const uint32_t G = make_direction_groups(dirs_local, n_dl, kd, group_of, group_rank, group_kd);
        if ((uint64_t)G * P > kMaxRank) throw PatchUnsupported("more than 2^20 macro-tiles");
        if (G > 1023) throw PatchUnsupported("more than 1023 direction groups");
        DTmp<uint16_t> group_dev;  group_dev.upload(group_of, stream, "group_of");
        DTmp<uint16_t> group_rank_dev; group_rank_dev.upload(group_rank, stream, "group_rank");
        DTmp<uint32_t> group_kd_dev;   group_kd_dev.upload(group_kd, stream, "group_kd");
        DTmp<uint32_t> tlevel;     tlevel.alloc(n, "tlevel");
        DTmp<unsigned int> minlev; minlev.alloc((size_t)P * n_dl, "minlev");
        DTmp<unsigned int> dep_tab; dep_tab.alloc((size_t)G * P * kMaxPatchDeps, "dep_tab");
        DTmp<unsigned int> counters; counters.alloc(8, "counters");   // [0] periodic entries [1] lag cursor [2] >255 periodic [3] order [4] dep overflow [5] sub overflow
        cuda_ok(cudaMemsetAsync(minlev.p, 0xff, sizeof(unsigned int) * (size_t)P * n_dl, stream), "memset");
        cuda_ok(cudaMemsetAsync(dep_tab.p, 0xff, sizeof(unsigned int) * (size_t)G * P * kMaxPatchDeps, stream), "memset");
        cuda_ok(cudaMemsetAsync(counters.p, 0, sizeof(unsigned int) * 8, stream), "memset");
        p_level_kernel<<<blocks_n, 256, 0, stream>>>(tasks, n, level_off_dev, n_levels, N, n_dl, pg.patch_of, tlevel.p, minlev.p);
        p_edges_kernel<<<blocks_n, 256, 0, stream>>>(g, n_dl, pg.patch_of, group_dev.p, P, dep_tab.p, counters.p + 4);
        launches += 2;
        std::vector<unsigned int> dep_h((size_t)G * P * kMaxPatchDeps);
        unsigned int cnt_h[8];
        cuda_ok(cudaMemcpyAsync(dep_h.data(), dep_tab.p, sizeof(unsigned int) * dep_h.size(), cudaMemcpyDeviceToHost, stream), "copy dep_tab");
        cuda_ok(cudaMemcpyAsync(cnt_h, counters.p, sizeof cnt_h, cudaMemcpyDeviceToHost, stream), "copy counters");
        cuda_ok(cudaStreamSynchronize(stream), "quotient graph sync");
        if (cnt_h[4]) throw PatchUnsupported("a patch has more than 64 upwind patches");

        // 2. macro-tile levels (Kahn over the quotient graph of every group) -> global order
        std::vector<uint32_t> patch_size(P);
        {
            std::vector<uint32_t> po(P + 1);
            cuda_ok(cudaMemcpy(po.data(), pg.patch_off, sizeof(uint32_t) * ((size_t)P + 1), cudaMemcpyDeviceToHost), "copy patch_off");
            for (uint32_t p = 0; p < P; ++p) patch_size[p] = po[p + 1] - po[p];
        }
        std::vector<uint32_t> mt_level, ndep;
        uint32_t max_level = 0;
        // per macro-tile (rank order): group, patch, number of upwind macro-tiles, row in dep_tab, offset of its dep_tab
        // entries in rank_of
        std::vector<uint32_t> rank_of, mt_group, mt_patch, mt_ndep, mt_row, dep_base;
        uint32_t n_mt = 0, n_phase = 1, kd_max = 1;
        for (uint32_t k : group_kd) kd_max = std::max(kd_max, k);
        DTmp<uint8_t> phase_dev;          // phases only: phase of every task
        DTmp<uint32_t> mtid_dev;          // phases only: (group, patch, phase) -> dense macro-tile id
        DTmp<unsigned int> minlev2;       // phases only: smallest level per (macro-tile, direction of the group)
        const bool phases = !level_macro_tiles(dep_h.data(), G, P, mt_level, ndep, max_level);
        if (!phases) {
            std::vector<uint32_t> mt_list;   // grp * P + p of every non-empty macro-tile, in rank order
            mt_list.reserve((size_t)G * P);
            for (uint32_t grp = 0; grp < G; ++grp)
                for (uint32_t p = 0; p < P; ++p)
                    if (patch_size[p]) mt_list.push_back(grp * P + p);
            std::stable_sort(mt_list.begin(), mt_list.end(), [&](uint32_t x, uint32_t y) {
                if (mt_level[x] != mt_level[y]) return mt_level[x] < mt_level[y];
                return (x % P) != (y % P) ? (x % P) < (y % P) : x < y;   // same level: neighbouring patches of all groups together
            });
            n_mt = (uint32_t)mt_list.size();
            rank_of.assign((size_t)G * P, 0xffffffffu);
            mt_group.resize(n_mt); mt_patch.resize(n_mt); mt_ndep.resize(n_mt); mt_row.resize(n_mt); dep_base.resize(n_mt);
            for (uint32_t r = 0; r < n_mt; ++r) {
                rank_of[mt_list[r]] = r;
                mt_group[r] = mt_list[r] / P;
                mt_patch[r] = mt_list[r] % P;
                mt_ndep[r] = ndep[mt_list[r]];
                mt_row[r] = mt_list[r];
                dep_base[r] = mt_group[r] * P;
            }
        } else {
            // The patch graph of some group is cyclic (a jagged Voronoi patch boundary carries flux both ways for
            // directions nearly parallel to it).  Split the macro-tiles into phases (see p_phase_kernel).
            // Opt-in (SSW_PATCH_PHASES=1): correct (tests/test_gpu_patch.py) but not yet profitable -- with the simple
            // lattice order used for pi a 32^3 Voronoi box needs 137 phases and 581 dependent macro-tile levels, more
            // than its 525 wavefront levels (6.1 ms against 2.4 ms for the level-barrier stream).
            if (!env_u32("SSW_PATCH_PHASES", 0))
                throw PatchUnsupported("patches depend on each other cyclically for a direction group");
            if (!level_off_host || level_off_host->size() != (size_t)n_levels + 1)
                throw PatchUnsupported("cyclic patch graph and no host copy of the level offsets");
            // a. total order pi of the patches per group: along the octant of the group's directions
            std::vector<uint32_t> pi((size_t)G * P);
            {
                std::vector<int> sgn((size_t)G * 3, 1);
                for (uint32_t dl = 0; dl < n_dl; ++dl)
                    for (int k = 0; k < 3; ++k)
                        if (dirs_local[3 * dl + k] < 0.0) sgn[(size_t)group_of[dl] * 3 + k] = -1;
                std::vector<uint64_t> key(P);
                std::vector<uint32_t> order(P);
                for (uint32_t grp = 0; grp < G; ++grp) {
                    for (uint32_t p = 0; p < P; ++p) {
                        uint32_t b[3] = {p / (pg.dims[1] * pg.dims[2]), (p / pg.dims[2]) % pg.dims[1], p % pg.dims[2]};
                        uint64_t o[3];
                        for (int k = 0; k < 3; ++k) o[k] = sgn[(size_t)grp * 3 + k] > 0 ? b[k] : pg.dims[k] - 1 - b[k];
                        key[p] = ((o[0] + o[1] + o[2]) << 40) | (o[0] << 26) | (o[1] << 13) | o[2];
                        order[p] = p;
                    }
                    std::sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return key[x] < key[y]; });
                    for (uint32_t i = 0; i < P; ++i) pi[(size_t)grp * P + order[i]] = i;
                }
            }
            DTmp<uint32_t> pi_dev; pi_dev.upload(pi, stream, "pi");
            // b. phases, level by level
            phase_dev.alloc(n, "phase");
            cuda_ok(cudaMemsetAsync(phase_dev.p, 0, n, stream), "memset phase");
            for (uint32_t l = 0; l < n_levels; ++l) {
                const uint32_t s0 = (*level_off_host)[l], s1 = (*level_off_host)[l + 1];
                if (s1 > s0)
                    p_phase_kernel<<<(s1 - s0 + 255) / 256, 256, 0, stream>>>(g, tasks, s0, s1, pg.patch_of, group_dev.p, P, pi_dev.p,
                                                                            phase_dev.p, counters.p);
            }
            launches += n_levels;
            cuda_ok(cudaMemcpyAsync(cnt_h, counters.p, sizeof cnt_h, cudaMemcpyDeviceToHost, stream), "copy counters");
            cuda_ok(cudaStreamSynchronize(stream), "phase sync");
            if (cnt_h[7]) throw PatchUnsupported("more than 255 phases");
            n_phase = cnt_h[6] + 1;
            const uint64_t n_keys64 = (uint64_t)G * P * n_phase;
            if (n_keys64 > (64ull << 20)) throw PatchUnsupported("too many (group, patch, phase) triples");
            const uint32_t n_keys = (uint32_t)n_keys64;
            // c. dense macro-tile ids
            DTmp<uint32_t> present, scan, mt_key_dev;
            present.alloc((size_t)n_keys + 1, "present");
            scan.alloc((size_t)n_keys + 1, "scan");
            mtid_dev.alloc(n_keys, "mtid_of");
            cuda_ok(cudaMemsetAsync(present.p, 0, sizeof(uint32_t) * ((size_t)n_keys + 1), stream), "memset present");
            p_present_kernel<<<blocks_n, 256, 0, stream>>>(N, n_dl, pg.patch_of, group_dev.p, P, n_phase, phase_dev.p, present.p);
            {
                size_t bytes = 0;
                DTmp<unsigned char> temp;
                cuda_ok(cub::DeviceScan::ExclusiveSum(nullptr, bytes, present.p, scan.p, (int)n_keys + 1, stream), "scan size");
                temp.alloc(bytes, "scan temp");
                cuda_ok(cub::DeviceScan::ExclusiveSum(temp.p, bytes, present.p, scan.p, (int)n_keys + 1, stream), "scan");
                cuda_ok(cudaMemcpyAsync(&n_mt, scan.p + n_keys, sizeof n_mt, cudaMemcpyDeviceToHost, stream), "copy n_mt");
                cuda_ok(cudaStreamSynchronize(stream), "scan sync");
            }
            if (n_mt == 0 || n_mt > kMaxRank) throw PatchUnsupported("more than 2^20 macro-tiles");
            mt_key_dev.alloc(n_mt, "mt_key");
            p_mtid_kernel<<<(n_keys + 255) / 256, 256, 0, stream>>>(present.p, scan.p, n_keys, mtid_dev.p, mt_key_dev.p);
            std::vector<uint32_t> mt_key(n_mt);
            cuda_ok(cudaMemcpyAsync(mt_key.data(), mt_key_dev.p, sizeof(uint32_t) * (size_t)n_mt, cudaMemcpyDeviceToHost, stream), "copy mt_key");
            // d. quotient graph over the macro-tiles, smallest level per (macro-tile, direction of the group)
            dep_tab.alloc((size_t)n_mt * kMaxPatchDeps, "dep_tab2");
            minlev2.alloc((size_t)n_mt * kd_max, "minlev2");
            cuda_ok(cudaMemsetAsync(dep_tab.p, 0xff, sizeof(unsigned int) * (size_t)n_mt * kMaxPatchDeps, stream), "memset");
            cuda_ok(cudaMemsetAsync(minlev2.p, 0xff, sizeof(unsigned int) * (size_t)n_mt * kd_max, stream), "memset");
            p_edges2_kernel<<<blocks_n, 256, 0, stream>>>(g, n_dl, pg.patch_of, group_dev.p, group_rank_dev.p, P, n_phase, kd_max,
                                                       phase_dev.p, mtid_dev.p, tlevel.p, dep_tab.p, minlev2.p, counters.p + 4);
            dep_h.assign((size_t)n_mt * kMaxPatchDeps, kEmptyDep);
            cuda_ok(cudaMemcpyAsync(dep_h.data(), dep_tab.p, sizeof(unsigned int) * dep_h.size(), cudaMemcpyDeviceToHost, stream), "copy dep_tab2");
            cuda_ok(cudaMemcpyAsync(cnt_h, counters.p, sizeof cnt_h, cudaMemcpyDeviceToHost, stream), "copy counters");
            cuda_ok(cudaStreamSynchronize(stream), "quotient graph 2 sync");
            launches += 6;
            if (cnt_h[4]) throw PatchUnsupported("a macro-tile has more than 64 upwind macro-tiles");
            // chain the phases of one (group, patch): their rate rows are accumulated in phase order.  Ids grow with the
            // key (group, patch, phase), so the previous phase of the same (group, patch) is the previous id.
            for (uint32_t id = 1; id < n_mt; ++id) {
                if (mt_key[id] / n_phase != mt_key[id - 1] / n_phase) continue;
                unsigned int *row = dep_h.data() + (size_t)id * kMaxPatchDeps;
                uint32_t k = 0;
                while (k < kMaxPatchDeps && row[k] != kEmptyDep && row[k] != id - 1) ++k;
                if (k == kMaxPatchDeps) throw PatchUnsupported("a macro-tile has more than 64 upwind macro-tiles");
                row[k] = id - 1;
            }
            cuda_ok(cudaMemcpyAsync(dep_tab.p, dep_h.data(), sizeof(unsigned int) * dep_h.size(), cudaMemcpyHostToDevice, stream), "copy dep_tab2");
            // e. levels and global order of the macro-tiles
            if (!level_macro_tiles(dep_h.data(), 1, n_mt, mt_level, ndep, max_level))
                throw std::runtime_error("compile_patch_schedule: macro-tile graph with phases is cyclic");
            std::vector<uint32_t> mt_list(n_mt);
            std::iota(mt_list.begin(), mt_list.end(), 0u);
            std::stable_sort(mt_list.begin(), mt_list.end(), [&](uint32_t x, uint32_t y) { return mt_level[x] < mt_level[y]; });
            rank_of.assign(n_mt, 0xffffffffu);
            mt_group.resize(n_mt); mt_patch.resize(n_mt); mt_ndep.resize(n_mt); mt_row.resize(n_mt); dep_base.assign(n_mt, 0u);
            for (uint32_t r = 0; r < n_mt; ++r) {
                const uint32_t id = mt_list[r], gp = mt_key[id] / n_phase;
                rank_of[id] = r;
                mt_group[r] = gp / P;
                mt_patch[r] = gp % P;
                mt_ndep[r] = ndep[id];
                mt_row[r] = id;
            }
        }
        DTmp<uint32_t> rank_dev; rank_dev.upload(rank_of, stream, "rank_of");

        // 3. slot order: (macro-tile rank, sub-level, cell, direction)
        DTmp<unsigned long long> keys_in, keys;
        keys_in.alloc(n, "keys_in");
        keys.alloc(n, "keys");
        if (!phases)
            p_key_kernel<<<blocks_n, 256, 0, stream>>>(N, n_dl, pg.patch_of, group_dev.p, P, rank_dev.p, tlevel.p, minlev.p,
                                                      keys_in.p, counters.p + 5);
        else
            p_key2_kernel<<<blocks_n, 256, 0, stream>>>(N, n_dl, pg.patch_of, group_dev.p, group_rank_dev.p, P, n_phase, kd_max,
                                                       phase_dev.p, mtid_dev.p, rank_dev.p, tlevel.p, minlev2.p, keys_in.p,
                                                       counters.p + 5);
        ++launches;
        {
            size_t bytes = 0;
            DTmp<unsigned char> temp;
            cuda_ok(cub::DeviceRadixSort::SortKeys(nullptr, bytes, keys_in.p, keys.p, (int64_t)n, 0, 64, stream), "radix size");
            temp.alloc(bytes, "radix temp");
            cuda_ok(cub::DeviceRadixSort::SortKeys(temp.p, bytes, keys_in.p, keys.p, (int64_t)n, 0, 64, stream), "radix sort");
            cuda_ok(cudaStreamSynchronize(stream), "radix sync");
            launches += 8;
        }
        keys_in.reset();
        tlevel.reset();
        cuda_ok(cudaMemcpy(cnt_h, counters.p, sizeof cnt_h, cudaMemcpyDeviceToHost), "copy counters");
        if (cnt_h[5]) throw PatchUnsupported("a patch spans more than 4096 wavefront levels");

        DTmp<uint8_t> pl_flag;    pl_flag.alloc(n, "pl_flag");
        DTmp<uint32_t> mt_slot0;  mt_slot0.alloc((size_t)n_mt + 1, "mt_slot0");
        DTmp<uint32_t> k32;       k32.alloc(n, "k32");
        DTmp<uint32_t> pl_start;  pl_start.alloc((size_t)n + 1, "pl_start");
        DTmp<uint32_t> n_sel;     n_sel.alloc(1, "n_sel");
        cuda_ok(cudaMemsetAsync(mt_slot0.p, 0xff, sizeof(uint32_t) * ((size_t)n_mt + 1), stream), "memset");
        p_bounds_kernel<<<blocks_n, 256, 0, stream>>>(keys.p, n, pl_flag.p, mt_slot0.p, k32.p);
        cuda_ok(cudaMemcpyAsync(mt_slot0.p + n_mt, &n, sizeof(uint32_t), cudaMemcpyHostToDevice, stream), "copy");
        uint32_t n_pl = 0;
        {
            cub::CountingInputIterator<uint32_t> iota(0);
            size_t bytes = 0;
            DTmp<unsigned char> temp;
            cuda_ok(cub::DeviceSelect::Flagged(nullptr, bytes, iota, pl_flag.p, pl_start.p, n_sel.p, (int)n, stream), "select size");
            temp.alloc(bytes, "select temp");
            cuda_ok(cub::DeviceSelect::Flagged(temp.p, bytes, iota, pl_flag.p, pl_start.p, n_sel.p, (int)n, stream), "select");
            cuda_ok(cudaMemcpyAsync(&n_pl, n_sel.p, sizeof n_pl, cudaMemcpyDeviceToHost, stream), "copy n_pl");
            cuda_ok(cudaStreamSynchronize(stream), "select sync");
            launches += 3;
        }
        pl_flag.reset();
        cuda_ok(cudaMemcpyAsync(pl_start.p + n_pl, &n, sizeof(uint32_t), cudaMemcpyHostToDevice, stream), "copy");
        std::vector<uint32_t> mt_slot0_h((size_t)n_mt + 1);
        cuda_ok(cudaMemcpyAsync(mt_slot0_h.data(), mt_slot0.p, sizeof(uint32_t) * ((size_t)n_mt + 1), cudaMemcpyDeviceToHost, stream), "copy");
        DTmp<uint32_t> pl_rank; pl_rank.alloc(n_pl, "pl_rank");
        p_pl_rank_kernel<<<(n_pl + 255) / 256, 256, 0, stream>>>(keys.p, pl_start.p, n_pl, pl_rank.p);
        std::vector<uint32_t> pl_rank_h(n_pl);
        cuda_ok(cudaMemcpyAsync(pl_rank_h.data(), pl_rank.p, sizeof(uint32_t) * (size_t)n_pl, cudaMemcpyDeviceToHost, stream), "copy");
        cuda_ok(cudaStreamSynchronize(stream), "bounds sync");
        keys.reset();
        pl_rank.reset();
        for (uint32_t r = 0; r <= n_mt; ++r)
            if (mt_slot0_h[r] == 0xffffffffu || (r && mt_slot0_h[r] <= mt_slot0_h[r - 1]))
                throw std::runtime_error("compile_patch_schedule: macro-tile offsets inconsistent");

        // 4. slots, entry counts, downwind areas
        cuda_ok(cudaMalloc(&C.slot_of, sizeof(uint32_t) * (size_t)n), "malloc slot_of");
        cuda_ok(cudaMalloc(&C.ttot_slot, sizeof(double) * (size_t)n), "malloc ttot_slot");
        DTmp<uint32_t> cnt_e, cnt_x;
        cnt_e.alloc((size_t)n + 1, "cnt_e");
        cnt_x.alloc((size_t)n + 1, "cnt_x");
        DTmp<unsigned long long> upoff, xoff;
        upoff.alloc((size_t)n + 1, "upoff");
        xoff.alloc((size_t)n + 1, "xoff");
        cuda_ok(cudaMemsetAsync(cnt_e.p + n, 0, sizeof(uint32_t), stream), "memset");
        cuda_ok(cudaMemsetAsync(cnt_x.p + n, 0, sizeof(uint32_t), stream), "memset");
        s_slot_scatter_kernel<<<blocks_n, 256, 0, stream>>>(k32.p, n, N, n_dl, C.slot_of);
        p_count_kernel<<<blocks_n, 256, 0, stream>>>(g, k32.p, n, n_dl, pg.patch_of, phases ? phase_dev.p : nullptr, cnt_e.p, cnt_x.p,
                                                    C.ttot_slot, counters.p);
        {
            cub::TransformInputIterator<unsigned long long, CastU64, const uint32_t *> in_e(cnt_e.p, CastU64()), in_x(cnt_x.p, CastU64());
            size_t bytes = 0;
            DTmp<unsigned char> temp;
            cuda_ok(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in_e, upoff.p, (int64_t)n + 1, stream), "scan size");
            temp.alloc(bytes, "scan temp");
            cuda_ok(cub::DeviceScan::ExclusiveSum(temp.p, bytes, in_e, upoff.p, (int64_t)n + 1, stream), "scan");
            cuda_ok(cub::DeviceScan::ExclusiveSum(temp.p, bytes, in_x, xoff.p, (int64_t)n + 1, stream), "scan");
            cuda_ok(cudaStreamSynchronize(stream), "scan sync");
            launches += 6;
        }
        cnt_e.reset();
        cnt_x.reset();
        unsigned long long total_entries = 0;
        cuda_ok(cudaMemcpy(&total_entries, upoff.p + n, sizeof total_entries, cudaMemcpyDeviceToHost), "copy");
        cuda_ok(cudaMemcpy(cnt_h, counters.p, sizeof cnt_h, cudaMemcpyDeviceToHost), "copy counters");
        if (cnt_h[2]) throw PatchUnsupported("a task has more than 255 periodic upwind faces");
        const uint32_t n_lag = cnt_h[0];
        if ((uint64_t)n + n_lag >= 0x7fffff00ull) throw PatchUnsupported("slot index overflow");
        C.n_entries = total_entries;
        C.n_lag = n_lag;

        // 5. tiles: <= threads slots of one pseudo-level, cut at cell boundaries
        DTmp<uint32_t> tile_cnt, tile_off, tile_start;
        tile_cnt.alloc(n_pl, "tile_cnt");
        tile_off.alloc((size_t)n_pl + 1, "tile_off");
        const unsigned cut_blocks = (unsigned)(((size_t)n_pl * 32 + 127) / 128);
        std::vector<uint32_t> tcnt(n_pl), toff((size_t)n_pl + 1, 0);
        s_cut_kernel<<<cut_blocks, 128, 0, stream>>>(k32.p, pl_start.p, n_pl, n_dl, threads, nullptr, nullptr, tile_cnt.p, nullptr);
        cuda_ok(cudaMemcpyAsync(tcnt.data(), tile_cnt.p, sizeof(uint32_t) * (size_t)n_pl, cudaMemcpyDeviceToHost, stream), "copy");
        cuda_ok(cudaStreamSynchronize(stream), "cut sync");
        for (uint32_t l = 0; l < n_pl; ++l) toff[l + 1] = toff[l] + tcnt[l];
        const uint32_t n_tiles = toff[n_pl];
        cuda_ok(cudaMemcpyAsync(tile_off.p, toff.data(), sizeof(uint32_t) * ((size_t)n_pl + 1), cudaMemcpyHostToDevice, stream), "copy");
        tile_start.alloc((size_t)n_tiles + 1, "tile_start");
        s_cut_kernel<<<cut_blocks, 128, 0, stream>>>(k32.p, pl_start.p, n_pl, n_dl, threads, nullptr, tile_off.p, nullptr, tile_start.p);
        cuda_ok(cudaMemcpyAsync(tile_start.p + n_tiles, &n, sizeof(uint32_t), cudaMemcpyHostToDevice, stream), "copy");
        DTmp<unsigned long long> tentry_dev, mtx_dev;
        tentry_dev.alloc((size_t)n_tiles + 1, "tentry");
        mtx_dev.alloc((size_t)n_mt + 1, "mtx");
        s_gather_offsets_kernel<<<(n_tiles + 1 + 255) / 256, 256, 0, stream>>>(upoff.p, tile_start.p, n_tiles, n, tentry_dev.p);
        s_gather_offsets_kernel<<<(n_mt + 1 + 255) / 256, 256, 0, stream>>>(xoff.p, mt_slot0.p, n_mt, n, mtx_dev.p);
        std::vector<uint32_t> tstart((size_t)n_tiles + 1);
        std::vector<unsigned long long> tentry((size_t)n_tiles + 1), mtx((size_t)n_mt + 1);
        cuda_ok(cudaMemcpyAsync(tstart.data(), tile_start.p, sizeof(uint32_t) * ((size_t)n_tiles + 1), cudaMemcpyDeviceToHost, stream), "copy");
        cuda_ok(cudaMemcpyAsync(tentry.data(), tentry_dev.p, sizeof(unsigned long long) * ((size_t)n_tiles + 1), cudaMemcpyDeviceToHost, stream), "copy");
        cuda_ok(cudaMemcpyAsync(mtx.data(), mtx_dev.p, sizeof(unsigned long long) * ((size_t)n_mt + 1), cudaMemcpyDeviceToHost, stream), "copy");
        cuda_ok(cudaStreamSynchronize(stream), "tile sync");
        launches += 4;
        tentry_dev.reset();
        mtx_dev.reset();
        tile_cnt.reset();
        tile_off.reset();

        // 6. packets: sizes, shared-memory budget, launch geometry
        std::vector<uint32_t> tile_rank(n_tiles), mt_tile0((size_t)n_mt + 1, 0);
        for (uint32_t pl = 0; pl < n_pl; ++pl)
            for (uint32_t t = toff[pl]; t < toff[pl + 1]; ++t) tile_rank[t] = pl_rank_h[pl];
        {
            uint32_t t = 0;
            for (uint32_t r = 0; r < n_mt; ++r) {
                mt_tile0[r] = t;
                while (t < n_tiles && tile_rank[t] == r) ++t;
                if (t == mt_tile0[r]) throw std::runtime_error("compile_patch_schedule: macro-tile without tiles");
            }
            mt_tile0[n_mt] = t;
            if (t != n_tiles) throw std::runtime_error("compile_patch_schedule: tiles out of rank order");
        }
        uint32_t max_bytes = 0, vmax = 0, pc_max = 0, smax = 0;
        std::vector<uint32_t> head_bytes(n_mt), tile_bytes(n_tiles);
        for (uint32_t r = 0; r < n_mt; ++r) {
            const uint64_t n_ext = mtx[r + 1] - mtx[r];
            const uint64_t ns = mt_slot0_h[r + 1] - mt_slot0_h[r];
            if (ns + n_ext > 65535ull) throw PatchUnsupported("a macro-tile holds more than 65535 values (smaller patches or direction groups needed)");
            vmax = std::max<uint32_t>(vmax, (uint32_t)(ns + n_ext));
            const uint64_t full = (uint64_t)patch_size[mt_patch[r]] * group_kd[mt_group[r]];   // rows of the rate reduction
            smax = std::max<uint32_t>(smax, (uint32_t)full);
            if (phases ? ns > full : ns != full)
                throw std::runtime_error("compile_patch_schedule: macro-tile is not (patch cells) x (group directions)");
            pc_max = std::max(pc_max, patch_size[mt_patch[r]]);
            head_bytes[r] = head_layout(mt_ndep[r], (uint32_t)n_ext, patch_size[mt_patch[r]]).bytes;
            max_bytes = std::max(max_bytes, head_bytes[r]);
        }
        for (uint32_t t = 0; t < n_tiles; ++t) {
            const uint32_t ns = tstart[t + 1] - tstart[t];
            const unsigned long long E = tentry[t + 1] - tentry[t];
            if (ns > threads || ns == 0 || E > 65535ull) throw PatchUnsupported("tile too large");
            tile_bytes[t] = ptile_layout(ns, (uint32_t)E).bytes;
            max_bytes = std::max(max_bytes, tile_bytes[t]);
        }
        const uint32_t stage_bytes = std::max<uint32_t>(128u, (max_bytes + 127u) & ~127u);
        PatchKernel kernel = patch_kernel_for(threads, false);
        uint32_t stages = want_stages;
        size_t smem = patch_smem(stages, stage_bytes, vmax, pc_max, smax).total;
        const size_t smem_block_max = 227u * 1024u - 1024u;
        while (smem > smem_block_max && stages > 2) { --stages; smem = patch_smem(stages, stage_bytes, vmax, pc_max, smax).total; }
        if (smem > smem_block_max) throw PatchUnsupported("a macro-tile does not fit in shared memory");
        raise_smem_limit((const void *)kernel);
        int per_sm = 0;
        cuda_ok(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, (int)threads, smem), "occupancy");
        if (per_sm < 1) throw PatchUnsupported("patch kernel does not fit on an SM");
        const uint32_t want_bps = env_u32("SSW_PATCH_BPS", 16);
        uint32_t nb = (uint32_t)std::min<int>(per_sm, (int)std::max<uint32_t>(1u, want_bps)) * (uint32_t)num_sms;
        nb = std::max<uint32_t>(1u, std::min<uint32_t>(nb, n_mt));

        // 7. packet table (block-major): macro-tile r goes to block r % nb
        std::vector<uint32_t> per_block(nb, 0);
        for (uint32_t r = 0; r < n_mt; ++r) per_block[r % nb] += 1 + (mt_tile0[r + 1] - mt_tile0[r]);
        std::vector<uint32_t> tab_off((size_t)nb + 1, 0);
        for (uint32_t b = 0; b < nb; ++b) tab_off[b + 1] = tab_off[b] + per_block[b];
        const uint32_t n_packets = tab_off[nb];
        std::vector<PDesc> ptab(n_packets);
        std::vector<uint32_t> ptab_block(n_packets), fill_pos(nb, 0);
        std::vector<uint64_t> cursor(nb, 0), head_rel(n_mt);
        for (uint32_t r = 0; r < n_mt; ++r) {
            const uint32_t b = r % nb;
            auto push = [&](uint32_t bytes, uint32_t id, uint32_t flags) {
                if ((cursor[b] >> 4) > 0xffffffffull) throw std::runtime_error("compile_patch_schedule: block stream exceeds 64 GB");
                PDesc d;
                d.off16 = (uint32_t)(cursor[b] >> 4);
                d.bytes = bytes; d.id = id; d.flags = flags;
                cursor[b] += bytes;
                ptab[tab_off[b] + fill_pos[b]] = d;
                ptab_block[tab_off[b] + fill_pos[b]] = b;
                fill_pos[b]++;
            };
            head_rel[r] = cursor[b];
            push(head_bytes[r], r, kPHead);
            for (uint32_t t = mt_tile0[r]; t < mt_tile0[r + 1]; ++t) push(tile_bytes[t], t, t + 1 == mt_tile0[r + 1] ? kPLast : 0u);
        }
        std::vector<uint64_t> stream_off((size_t)nb + 1, 0), head_off(n_mt);
        for (uint32_t b = 0; b < nb; ++b) stream_off[b + 1] = stream_off[b] + ((cursor[b] + 127u) & ~(uint64_t)127u);
        for (uint32_t r = 0; r < n_mt; ++r) head_off[r] = stream_off[r % nb] + head_rel[r];
        C.stream_bytes = stream_off[nb];

        cuda_ok(cudaMalloc(&C.stream, std::max<uint64_t>(C.stream_bytes, 128)), "malloc stream");
        cuda_ok(cudaMalloc(&C.ptab, sizeof(PDesc) * (size_t)std::max<uint32_t>(n_packets, 1)), "malloc ptab");
        cuda_ok(cudaMalloc(&C.tab_off, sizeof(uint32_t) * ((size_t)nb + 1)), "malloc tab_off");
        cuda_ok(cudaMalloc(&C.stream_off, sizeof(uint64_t) * ((size_t)nb + 1)), "malloc stream_off");
        cuda_ok(cudaMalloc(&C.out_slot, sizeof(double) * ((size_t)n + n_lag)), "malloc out_slot");
        cuda_ok(cudaMalloc(&C.lag_src, sizeof(uint32_t) * (size_t)std::max<uint32_t>(n_lag, 1)), "malloc lag_src");
        cuda_ok(cudaMalloc(&C.acc_cell, sizeof(double) * (size_t)G * N), "malloc acc_cell");
        cuda_ok(cudaMalloc(&C.acc_per, sizeof(double) * (size_t)std::max<uint32_t>(n_periodic, 1)), "malloc acc_per");
        cuda_ok(cudaMalloc(&C.mt_flag, sizeof(unsigned int) * (size_t)n_mt), "malloc mt_flag");
        cuda_ok(cudaMemsetAsync(C.mt_flag, 0, sizeof(unsigned int) * (size_t)n_mt, stream), "memset");
        cuda_ok(cudaMemsetAsync(C.acc_cell, 0, sizeof(double) * (size_t)G * N, stream), "memset");
        cuda_ok(cudaMemsetAsync(C.acc_per, 0, sizeof(double) * (size_t)std::max<uint32_t>(n_periodic, 1), stream), "memset");
        cuda_ok(cudaMemcpyAsync(C.ptab, ptab.data(), sizeof(PDesc) * (size_t)n_packets, cudaMemcpyHostToDevice, stream), "copy ptab");
        cuda_ok(cudaMemcpyAsync(C.tab_off, tab_off.data(), sizeof(uint32_t) * ((size_t)nb + 1), cudaMemcpyHostToDevice, stream), "copy");
        cuda_ok(cudaMemcpyAsync(C.stream_off, stream_off.data(), sizeof(uint64_t) * ((size_t)nb + 1), cudaMemcpyHostToDevice, stream), "copy");
        DTmp<uint32_t> ptab_block_dev, tile_rank_dev, mt_group_dev, mt_patch_dev, mt_ndep_dev, mt_row_dev, dep_base_dev;
        mt_row_dev.upload(mt_row, stream, "mt_row");
        dep_base_dev.upload(dep_base, stream, "dep_base");
        DTmp<uint64_t> head_off_dev;
        ptab_block_dev.upload(ptab_block, stream, "ptab_block");
        tile_rank_dev.upload(tile_rank, stream, "tile_rank");
        mt_group_dev.upload(mt_group, stream, "mt_group");
        mt_patch_dev.upload(mt_patch, stream, "mt_patch");
        mt_ndep_dev.upload(mt_ndep, stream, "mt_ndep");
        head_off_dev.upload(head_off, stream, "head_off");

        // 8. packets and state
        PFillArgs fa;
        fa.g = g; fa.pg = pg; fa.k32 = k32.p; fa.slot_of = C.slot_of; fa.upoff = upoff.p; fa.xoff = xoff.p;
        fa.ttot_slot = C.ttot_slot; fa.n_dl = n_dl; fa.n_tasks = n; fa.stages = stages;
        fa.ptab = C.ptab; fa.ptab_block = ptab_block_dev.p; fa.tab_off = C.tab_off; fa.stream_off = C.stream_off;
        fa.stream = C.stream; fa.tile_start = tile_start.p; fa.tile_rank = tile_rank_dev.p; fa.mt_slot0 = mt_slot0.p;
        fa.mt_group = mt_group_dev.p; fa.mt_patch = mt_patch_dev.p; fa.mt_ndep = mt_ndep_dev.p; fa.head_off = head_off_dev.p;
        fa.group_rank = group_rank_dev.p; fa.group_kd = group_kd_dev.p;
        fa.dep_tab = dep_tab.p; fa.rank_of = rank_dev.p; fa.lag_src = C.lag_src; fa.counters = counters.p;
        fa.mt_row = mt_row_dev.p; fa.dep_base = dep_base_dev.p; fa.phase = phases ? phase_dev.p : nullptr;
        if (n_packets) p_fill_kernel<<<n_packets, 256, 0, stream>>>(fa);
        s_convert_state_kernel<<<blocks_n, 256, 0, stream>>>(k32.p, n, N, n_dl, q_nat, C.ttot_slot, C.out_slot);
        cuda_ok(cudaGetLastError(), "patch compile kernels");
        cuda_ok(cudaMemcpyAsync(cnt_h, counters.p, sizeof cnt_h, cudaMemcpyDeviceToHost, stream), "copy");
        cuda_ok(cudaStreamSynchronize(stream), "patch compile sync");   // host vectors go out of scope
        launches += 2;
        if (cnt_h[1] != n_lag) throw std::runtime_error("compile_patch_schedule: periodic snapshot count mismatch");
        if (cnt_h[3]) throw std::runtime_error("compile_patch_schedule: a task reads a slot of a later sub-level");

        // static list of the periodic upwind entries per periodic cell (the periodic_source term of the rate)
        if (n_periodic) {
            DTmp<uint32_t> pcnt;
            pcnt.alloc((size_t)n_periodic + 1, "pcnt");
            cuda_ok(cudaMalloc(&C.per_off, sizeof(uint32_t) * ((size_t)n_periodic + 1)), "malloc per_off");
            cuda_ok(cudaMemsetAsync(pcnt.p + n_periodic, 0, sizeof(uint32_t), stream), "memset");
            const unsigned pb = (n_periodic + 127) / 128;
            p_periodic_list_kernel<<<pb, 128, 0, stream>>>(g, pcells, n_periodic, n_dl, C.slot_of, C.ttot_slot, false, pcnt.p,
                                                          nullptr, nullptr, nullptr);
            size_t bytes = 0;
            DTmp<unsigned char> temp;
            cuda_ok(cub::DeviceScan::ExclusiveSum(nullptr, bytes, pcnt.p, C.per_off, (int)n_periodic + 1, stream), "scan size");
            temp.alloc(bytes, "scan temp");
            cuda_ok(cub::DeviceScan::ExclusiveSum(temp.p, bytes, pcnt.p, C.per_off, (int)n_periodic + 1, stream), "scan");
            uint32_t n_pe = 0;
            cuda_ok(cudaMemcpyAsync(&n_pe, C.per_off + n_periodic, sizeof n_pe, cudaMemcpyDeviceToHost, stream), "copy");
            cuda_ok(cudaStreamSynchronize(stream), "periodic list sync");
            if (n_pe != n_lag) throw std::runtime_error("compile_patch_schedule: periodic entry count mismatch");
            cuda_ok(cudaMalloc(&C.per_src, sizeof(uint32_t) * (size_t)std::max<uint32_t>(n_pe, 1)), "malloc per_src");
            cuda_ok(cudaMalloc(&C.per_w, sizeof(double) * (size_t)std::max<uint32_t>(n_pe, 1)), "malloc per_w");
            p_periodic_list_kernel<<<pb, 128, 0, stream>>>(g, pcells, n_periodic, n_dl, C.slot_of, C.ttot_slot, true, nullptr,
                                                          C.per_off, C.per_src, C.per_w);
            cuda_ok(cudaStreamSynchronize(stream), "periodic list sync");
            launches += 4;
        }

        C.threads = threads;
        C.bps = (uint32_t)per_sm;
        C.stages = stages;
        C.stage_bytes = stage_bytes;
        C.n_blocks = nb;
        C.n_tiles = n_tiles;
        C.n_mt = n_mt;
        C.vmax = vmax;
        C.pc_max = pc_max;
        C.smax = smax;
        C.kd = kd;
        C.n_patches = P;
        C.patch_levels = max_level + 1;
        C.n_pl = n_pl;
        C.n_groups = G;
        C.n_groups_per = 1;
        C.epoch = 0;
        C.patch_mode = true;
        C.accumulate = phases;
        C.n_phases = n_phase;
    } catch (...) {
        C.release();
        if (launch_counter) *launch_counter += launches;
        throw;
    }
    if (launch_counter) *launch_counter += launches;
    C.n_tasks = n_tasks;
    C.n_levels = n_levels;
    C.n_cells = g.n_cells;
    C.n_periodic = n_periodic;
    C.n_epilogue = 0;
    C.mean_entries = n_tasks ? (double)C.n_entries / (double)n_tasks : 0.0;
    C.valid = true;
}

// One all-cells sweep over the patch-ordered schedule.  Leaves sum_d incoming per (group, cell) in C.acc_cell
// and sum_d periodic_source per periodic cell in C.acc_per (s_rate_finish_kernel folds them).
inline void run_patch(Compiled &C, const double2 *cellrec, double threshold, cudaStream_t stream, uint64_t *launch_counter) {
    PatchArgs a;
    a.stream = C.stream;
    a.stream_off = C.stream_off;
    a.ptab = C.ptab;
    a.tab_off = C.tab_off;
    a.mt_flag = C.mt_flag;
    a.out_slot = C.out_slot;
    a.cellrec = cellrec;
    a.acc_cell = C.acc_cell;
    a.threshold = threshold;
    a.stages = C.stages;
    a.stage_bytes = C.stage_bytes;
    a.vmax = C.vmax;
    a.pc_max = C.pc_max;
    a.smax = C.smax;
    a.n_cells = C.n_cells;
    a.epoch = ++C.epoch;
    a.accumulate = C.accumulate ? 1u : 0u;
    if (C.accumulate) cuda_ok(cudaMemsetAsync(C.acc_cell, 0, sizeof(double) * (size_t)C.n_groups * C.n_cells, stream), "memset acc_cell");
    a.poll_ns = env_u32("SSW_STREAM_POLL_NS", 20);
    a.prof = nullptr;
    unsigned long long *prof_dev = nullptr;
    if (env_u32("SSW_STREAM_PROFILE", 0)) {
        cuda_ok(cudaMalloc(&prof_dev, sizeof(unsigned long long) * 10 * (size_t)C.n_blocks), "malloc prof");
        cuda_ok(cudaMemsetAsync(prof_dev, 0, sizeof(unsigned long long) * 10 * (size_t)C.n_blocks, stream), "memset prof");
        a.prof = prof_dev;
    }
    uint64_t launches = 1;
    if (C.n_lag) {
        s_lag_snapshot_kernel<<<(C.n_lag + 255) / 256, 256, 0, stream>>>(C.lag_src, C.n_lag, (uint32_t)C.n_tasks, C.out_slot);
        ++launches;
    }
    PatchKernel kernel = patch_kernel_for(C.threads, prof_dev != nullptr);
    const size_t smem = patch_smem(C.stages, C.stage_bytes, C.vmax, C.pc_max, C.smax).total;
    raise_smem_limit((const void *)kernel);
    void *args[] = {&a};
    // cooperative launch only to guarantee co-residency of all blocks (the done flags are polled)
    cuda_ok(cudaLaunchCooperativeKernel((const void *)kernel, dim3(C.n_blocks), dim3(C.threads), args, smem, stream),
            "patch_sweep_kernel launch");
    if (C.n_periodic) {
        p_periodic_rate_kernel<<<(unsigned)(((size_t)C.n_periodic * 32 + 255) / 256), 256, 0, stream>>>(
            C.per_off, C.per_src, C.per_w, C.n_periodic, C.out_slot, C.acc_per);
        ++launches;
    }
    if (prof_dev) {
        std::vector<unsigned long long> h(10 * (size_t)C.n_blocks);
        cudaMemcpyAsync(h.data(), prof_dev, sizeof(unsigned long long) * h.size(), cudaMemcpyDeviceToHost, stream);
        cudaStreamSynchronize(stream);
        cudaFree(prof_dev);
        double tot = 0, poll = 0, pkt = 0, packets = 0, tmax = 0, cmp = 0, bar = 0, post = 0, head = 0;
        for (uint32_t b = 0; b < C.n_blocks; ++b) {
            tot += (double)h[10 * b]; poll += (double)h[10 * b + 1]; pkt += (double)h[10 * b + 2]; packets += (double)h[10 * b + 3];
            cmp += (double)h[10 * b + 4]; bar += (double)h[10 * b + 6]; post += (double)h[10 * b + 7];
            head += (double)h[10 * b + 8];
            tmax = std::max(tmax, (double)h[10 * b]);
        }
        fprintf(stderr, "[patch phases] thread 0, cycles per tile: compute %.0f  barrier %.0f  post %.0f  mbar wait (all packets) %.0f;  "
                        "cycles per head (incl. poll) %.0f\n",
                cmp / C.n_tiles, bar / C.n_tiles, post / C.n_tiles, pkt / packets, head / C.n_mt);
        fprintf(stderr, "[patch profile] blocks %u (%u/SM x %u thr) macro-tiles %u tiles %u  cycles/block mean %.0f max %.0f  "
                        "dependency poll %.1f%%  packet wait %.1f%%  cycles per packet %.0f  patch levels %u  vmax %u stage %u B x %u\n",
                C.n_blocks, C.bps, C.threads, C.n_mt, C.n_tiles, tot / C.n_blocks, tmax, 100.0 * poll / tot, 100.0 * pkt / tot,
                (tot - poll) / packets, C.patch_levels, C.vmax, C.stage_bytes, C.stages);
    }
    if (launch_counter) *launch_counter += launches;
}

}  // namespace ssw
