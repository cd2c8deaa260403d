// kernels.cuh -- sm_100a kernels of the sweep hot path (general path: any active set).
//
// Data layout in HBM:
//   grid, static      face_geo[F]   double4 {nx, ny, nz, area}       one 32-B sector per face
//                     face_rev[F]   double   area of the same face as stored by the neighbour
//                     face_nb[F]    int32    neighbour cell, -1 boundary
//                     face_kind[F]  uint8    SSW_FACE_*
//                     face_off[N+1] uint32
//   per cell          rho, x, T, ts, tau, prev_rate, src, size, volume, att (= exp(-n_HI sigma size)),
//                     level (uint8), pidx (int32: row in the periodic tables or -1)
//   per (dir, cell)   q[dl*N + c]        outgoing rate / total downwind effective area
//                     incoming[dl*N + c] incoming_total_rate + source / D as read by the task
//                     missing[dl*N + c]  int32, Kahn counters (build mode only)
//   periodic rows     per_lag[dl*P + p], per_new[dl*P + p]
//
// The reference keeps incoming / outgoing / periodic_source accumulators per (cell, dir) and
// scatters *corrections* (src/sweep/mod.rs:423-513).  Because every accumulator starts at zero
// and every correction is (new_out - old_out) * share, the accumulators telescope to
//     incoming[c][d]        = sum over Local    upwind faces  out[nb][d] * share(nb -> c, d)
//     periodic_source[c][d] = sum over Periodic upwind faces  out[nb][d] * share(nb -> c, d)
// at all times, with share = A * (n.d) / sum_downwind(A * (n.d)) evaluated on the donor's
// faces.  The kernels therefore keep only q = out / sum_downwind(A n.d) per (cell, dir) and
// GATHER: no atomics on flux, results independent of the order tasks run in.
#pragma once
#include <cooperative_groups.h>
#include <cstdint>

#include "chemistry.cuh"

namespace ssw {
namespace cg = cooperative_groups;

constexpr int kMaxDirs = 128;

struct GridView {
    const double4 *face_geo;
    const double *face_rev;
    const int32_t *face_nb;
    const uint8_t *face_kind;
    const uint32_t *face_off;
    const double *dirs;   // this rank's directions, local order, 3 per direction (per handle: several handles may share a process)
    uint32_t n_cells;
};

struct CellView {
    double *rho, *x, *T, *ts, *tau, *prev_rate, *src, *att, *ion_time;
    const double *size, *volume;
    uint8_t *level;
    const int32_t *pidx;
};

// Flux state access.  Before the all-cells schedule is compiled the state is q = out / ttot in
// the natural [dl][c] layout; afterwards it is out_slot (patch.cuh / stream.cuh) reached through slot_of.
struct StateView {
    double *q;
    const uint32_t *slot_of;
    double *out_slot;
    const double *ttot_slot;
    uint32_t n_cells;
    __device__ __forceinline__ double load_q(uint32_t dl, uint32_t c) const {
        const size_t t = (size_t)dl * n_cells + c;
        if (slot_of) {
            const uint32_t sl = __ldg(slot_of + t);
            const double tt = __ldg(ttot_slot + sl);
            return tt > 0.0 ? __ldcg(out_slot + sl) / tt : 0.0;
        }
        return __ldcg(q + t);
    }
    __device__ __forceinline__ double load_out(uint32_t dl, uint32_t c, double ttot) const {
        const size_t t = (size_t)dl * n_cells + c;
        if (slot_of) return __ldcg(out_slot + __ldg(slot_of + t));
        return __ldcg(q + t) * ttot;
    }
    __device__ __forceinline__ void store(uint32_t dl, uint32_t c, double out, double ttot) const {
        const size_t t = (size_t)dl * n_cells + c;
        if (slot_of) __stcg(out_slot + __ldg(slot_of + t), out);
        else __stcg(q + t, ttot > 0.0 ? out / ttot : 0.0);
    }
};

// glam DVec3::dot without FMA contraction: the sign decides upwind / downwind
// (src/sweep/grid/cell.rs:126-132) and must match the CPU bit for bit.
__device__ __forceinline__ double dot_dir(const double4 &g, const double dx, const double dy,
                                          const double dz) {
    return __dadd_rn(__dadd_rn(__dmul_rn(g.x, dx), __dmul_rn(g.y, dy)), __dmul_rn(g.z, dz));
}

__device__ __forceinline__ double4 ld_geo(const double4 *p) {
    // two 16-byte read-only loads (the grid is immutable after create)
    const double2 a = __ldg(reinterpret_cast<const double2 *>(p));
    const double2 b = __ldg(reinterpret_cast<const double2 *>(p) + 1);
    return make_double4(a.x, a.y, b.x, b.y);
}

// ------------------------------------------------------------------------------------------
// Kahn counters: Sweep::init_counts (src/sweep/mod.rs:346-386) + get_initial_tasks (:388-398)
// one thread per (active cell, local direction); blockIdx.y = local direction.
// ------------------------------------------------------------------------------------------
struct QueueCtl {       // device-resident control block of one build
    unsigned int cnt[3];  // pushes of level l go to cnt[(l+1)%3]; cnt[0] = initial tasks
    unsigned int n_levels;
    unsigned int solved;
    unsigned int overflow;
};

__device__ __forceinline__ void warp_push(uint32_t *queue, unsigned int base, unsigned int *counter,
                                          bool pred, uint32_t task) {
    const unsigned mask = __activemask();
    const unsigned ballot = __ballot_sync(mask, pred);
    if (ballot == 0) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(ballot) - 1;
    unsigned int pos = 0;
    if (lane == leader) pos = atomicAdd(counter, __popc(ballot));
    pos = __shfl_sync(mask, pos, leader);
    if (pred) queue[base + pos + __popc(ballot & ((1u << lane) - 1))] = task;
}

__global__ void __launch_bounds__(256)
init_counts_kernel(GridView g, const uint8_t *__restrict__ level, int cur,
                   const uint32_t *__restrict__ act_list, uint32_t n_act, int32_t *missing,
                   uint32_t *queue, QueueCtl *ctl, int32_t *wlevel_dbg, int dl_base) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    const int dl = dl_base + blockIdx.y;
    const bool valid = k < n_act;
    uint32_t c = 0;
    int m = 0;
    if (valid) {
        c = act_list ? act_list[k] : k;
        const double dx = g.dirs[3 * dl], dy = g.dirs[3 * dl + 1], dz = g.dirs[3 * dl + 2];
        const uint32_t f0 = g.face_off[c], f1 = g.face_off[c + 1];
        for (uint32_t f = f0; f < f1; ++f) {
            if (g.face_kind[f] != 0) continue;  // only Local faces carry dependencies
            const double d = dot_dir(ld_geo(g.face_geo + f), dx, dy, dz);
            if (!(d < 0.0)) continue;
            if (level[g.face_nb[f]] >= cur) ++m;
        }
        missing[(size_t)dl * g.n_cells + c] = m;
        if (wlevel_dbg && m == 0) wlevel_dbg[c] = 0;
    }
    warp_push(queue, 0, &ctl->cnt[0], valid && m == 0, (uint32_t)dl * g.n_cells + c);
}

// ------------------------------------------------------------------------------------------
// one task: Sweep::solve_task (src/sweep/mod.rs:423-485) in gather form
// ------------------------------------------------------------------------------------------
struct SweepArgs {
    GridView g;
    const double *att;      // per cell
    const double *src;      // per cell
    const int32_t *pidx;    // per cell
    const uint8_t *level;   // per cell
    StateView st;           // flux state
    int solve;              // 0: only peel the level sets (no flux)
    double *incoming;       // [dl][c]
    const double *per_lag;  // [dl][p]
    int32_t *missing;       // [dl][c]
    uint32_t n_periodic;
    double inv_threshold_unused;
    double threshold;
    double n_dirs_total;    // D as f64 (source / D, src/sweep/site.rs:49-51)
    int cur;
};

template <bool BUILD>
__device__ __forceinline__ void solve_task(const SweepArgs &a, uint32_t task, uint32_t *queue,
                                           unsigned int push_base, unsigned int *push_counter,
                                           int32_t *wlevel_dbg, int wave, bool valid) {
    const uint32_t N = a.g.n_cells;
    uint32_t c = 0, dl = 0, f0 = 0, f1 = 0;
    double dx = 0, dy = 0, dz = 0;
    if (valid) {
        dl = task / N;
        c = task - dl * N;
        dx = a.g.dirs[3 * dl];
        dy = a.g.dirs[3 * dl + 1];
        dz = a.g.dirs[3 * dl + 2];
        f0 = a.g.face_off[c];
        f1 = a.g.face_off[c + 1];
    }
    if (valid && a.solve) {
        double in = 0.0, ttot = 0.0;
        // faces in chunks of kChunk: all loads of a chunk are issued before the first use, so a task costs a few
        // dependent memory round trips instead of a few per face; the sums still run in face order
        constexpr int kChunk = 8;
        for (uint32_t fb = f0; fb < f1; fb += kChunk) {
            double dd[kChunk], area[kChunk], w[kChunk], qv[kChunk];
            uint32_t nb[kChunk];
            bool up[kChunk];
#pragma unroll
            for (int j = 0; j < kChunk; ++j) {
                const uint32_t f = fb + j;
                dd[j] = 0.0; area[j] = 0.0; w[j] = 0.0; nb[j] = 0; up[j] = false;
                if (f < f1) {
                    const double4 geo = ld_geo(a.g.face_geo + f);
                    dd[j] = dot_dir(geo, dx, dy, dz);
                    area[j] = geo.w;
                    up[j] = dd[j] < 0.0 && a.g.face_kind[f] == 0;
                    nb[j] = (uint32_t)a.g.face_nb[f];
                    w[j] = __ldg(a.g.face_rev + f);
                }
            }
#pragma unroll
            for (int j = 0; j < kChunk; ++j) qv[j] = up[j] ? a.st.load_q(dl, nb[j]) : 0.0;
#pragma unroll
            for (int j = 0; j < kChunk; ++j) {
                if (up[j]) in += qv[j] * (w[j] * (-dd[j]));
                else if (dd[j] > 0.0) ttot += area[j] * dd[j];
            }
        }
        const double inc = in + a.src[c] / a.n_dirs_total;           // site.rs:49-56
        const int32_t p = a.pidx[c];
        const double total = p >= 0 ? inc + a.per_lag[(size_t)dl * a.n_periodic + p] : inc + 0.0;
        // HydrogenOnly::get_outgoing_rate, hydrogen_only/mod.rs:81-87
        const double out = (total < a.threshold) ? 0.0 : total * a.att[c];
        a.st.store(dl, c, out, ttot);
        __stcg(a.incoming + (size_t)dl * N + c, inc);
    }
    if (BUILD && valid && wlevel_dbg) wlevel_dbg[c] = wave;
    if (BUILD) {
        // handle_local_neighbour (src/sweep/mod.rs:487-503): release downwind active neighbours.  Faces in chunks of
        // kChunk: geometry and neighbour of the whole chunk are loaded together, then the neighbours' timestep levels,
        // then the counters are decremented and the released tasks pushed -- three dependent memory round trips per
        // chunk instead of three per face (a wavefront level of the build costs a few microseconds, not tens).  The
        // pushes are warp-synchronous: every lane of the warp runs the same number of rounds.
        __syncwarp();
        constexpr int kChunk = 8;
        uint32_t n_faces = valid ? f1 - f0 : 0u;
        for (int o = 16; o > 0; o >>= 1) n_faces = max(n_faces, __shfl_xor_sync(0xffffffffu, n_faces, o));
        for (uint32_t fb = 0; fb < n_faces; fb += kChunk) {
            uint32_t nb[kChunk];
            bool cand[kChunk];
#pragma unroll
            for (int j = 0; j < kChunk; ++j) {
                const uint32_t f = f0 + fb + j;
                nb[j] = 0;
                cand[j] = false;
                if (valid && f < f1) {
                    const double d = dot_dir(ld_geo(a.g.face_geo + f), dx, dy, dz);
                    nb[j] = (uint32_t)a.g.face_nb[f];
                    cand[j] = a.g.face_kind[f] == 0 && d > 0.0;
                }
            }
#pragma unroll
            for (int j = 0; j < kChunk; ++j)
                if (cand[j]) cand[j] = a.level[nb[j]] >= a.cur;
            bool ready[kChunk];
#pragma unroll
            for (int j = 0; j < kChunk; ++j) ready[j] = cand[j] && atomicSub(a.missing + (size_t)dl * N + nb[j], 1) == 1;
#pragma unroll
            for (int j = 0; j < kChunk; ++j) {
                if (__ballot_sync(0xffffffffu, ready[j]) == 0) continue;   // uniform across the warp
                warp_push(queue, push_base, push_counter, ready[j], dl * N + nb[j]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// BUILD: fused Kahn peeling + solve, persistent cooperative kernel.  Level l of the wavefront
// is queue[start_l, end_l); tasks released while solving it are appended behind end_l, so when
// the kernel ends `queue` is the level-sorted task list and ctl->n_levels / level_off describe
// the level sets (north_star: "upwind dependencies are resolved on device into wavefront level
// sets").  One grid barrier per level.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sweep_build_kernel(SweepArgs a, uint32_t *queue, QueueCtl *ctl, uint32_t *level_off,
                   uint32_t level_off_cap, int32_t *wlevel_dbg) {
    cg::grid_group grid = cg::this_grid();
    const unsigned int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int gsz = gridDim.x * blockDim.x;
    unsigned int start = 0;
    unsigned int end = ctl->cnt[0];
    unsigned int lvl = 0;
    while (start < end) {
        if (gtid == 0) {
            if (lvl < level_off_cap) level_off[lvl] = start;
            else ctl->overflow = 1;
            ctl->cnt[(lvl + 2) % 3] = 0;  // slot used by level lvl+1's pushes; nobody reads it now
        }
        unsigned int *counter = &ctl->cnt[(lvl + 1) % 3];
        // every lane of every warp runs the same number of rounds (warp-synchronous pushes)
        const unsigned int span = end - start;
        const unsigned int rounds = (span + gsz - 1) / gsz;
        for (unsigned int r = 0; r < rounds; ++r) {
            const unsigned int i = start + r * gsz + gtid;
            const bool valid = i < end;
            const uint32_t task = valid ? __ldcg(queue + i) : 0u;
            solve_task<true>(a, task, queue, end, counter, wlevel_dbg, (int)lvl, valid);
        }
        grid.sync();
        start = end;
        end = end + *((volatile unsigned int *)counter);
        ++lvl;
    }
    if (gtid == 0) {
        if (lvl < level_off_cap) level_off[lvl] = start;
        else ctl->overflow = 1;
        ctl->n_levels = lvl;
        ctl->solved = start;
    }
}

// REPLAY: the level sets are known (cached); same gather, no counters.  Persistent cooperative
// kernel, one grid barrier per wavefront level.
__global__ void __launch_bounds__(256)
sweep_replay_kernel(SweepArgs a, const uint32_t *__restrict__ queue,
                    const uint32_t *__restrict__ level_off, uint32_t n_levels) {
    cg::grid_group grid = cg::this_grid();
    const unsigned int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int gsz = gridDim.x * blockDim.x;
    for (uint32_t lvl = 0; lvl < n_levels; ++lvl) {
        const uint32_t s = level_off[lvl], e = level_off[lvl + 1];
        for (uint32_t i = s + gtid; i < e; i += gsz)
            solve_task<false>(a, queue[i], nullptr, 0, nullptr, nullptr, 0, true);
        if (lvl + 1 < n_levels) grid.sync();
    }
}

// REPLAY of a small active set: every wavefront level fits one block, so a block-wide barrier
// (tens of cycles) replaces the grid barrier (microseconds).  Same arithmetic as sweep_replay_kernel.
__global__ void __launch_bounds__(512)
sweep_replay_small_kernel(SweepArgs a, const uint32_t *__restrict__ queue,
                          const uint32_t *__restrict__ level_off, uint32_t n_levels) {
    for (uint32_t lvl = 0; lvl < n_levels; ++lvl) {
        const uint32_t s = level_off[lvl], e = level_off[lvl + 1];
        for (uint32_t i = s + threadIdx.x; i < e; i += blockDim.x)
            solve_task<false>(a, queue[i], nullptr, 0, nullptr, nullptr, 0, true);
        __syncthreads();   // global writes of this level are visible to the block's next level
    }
}

// ------------------------------------------------------------------------------------------
// Task records of a cached partial schedule ("mini" form): what solve_task derives from the CSR grid
// on every replay -- own state index, sum_downwind(A n.d), and per Local upwind face the neighbour's
// state index and weight -- is stored once per task in level order, so a replay costs three
// dependent memory round trips per wavefront level instead of one or two per face.
//   natural state (q = out / ttot, [dl][c]):  src = dl*N + nb,  w = A_rev (-n.d)        (same arithmetic as solve_task)
//   slot-ordered state (out_slot, stream.cuh): src = slot(nb),   w = A_rev (-n.d) / ttot(nb)
// ------------------------------------------------------------------------------------------
struct MiniView {
    const uint32_t *off;    // n_tasks + 1
    const uint32_t *src;    // entries
    const double *w;        // entries
    const uint32_t *self;   // n_tasks: own state index
    const double *ttot;     // n_tasks
};

__global__ void __launch_bounds__(256)
mini_count_kernel(GridView g, const uint32_t *__restrict__ tasks, uint32_t n_tasks, uint32_t *__restrict__ cnt) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_tasks) return;
    const uint32_t t = tasks[i];
    const uint32_t dl = t / g.n_cells, c = t - dl * g.n_cells;
    const double dx = g.dirs[3 * dl], dy = g.dirs[3 * dl + 1], dz = g.dirs[3 * dl + 2];
    uint32_t m = 0;
    for (uint32_t f = g.face_off[c]; f < g.face_off[c + 1]; ++f)
        if (g.face_kind[f] == 0 && dot_dir(ld_geo(g.face_geo + f), dx, dy, dz) < 0.0) ++m;
    cnt[i] = m;
}

__global__ void __launch_bounds__(256)
mini_fill_kernel(GridView g, StateView st, const uint32_t *__restrict__ tasks, uint32_t n_tasks,
                 const uint32_t *__restrict__ off, uint32_t *__restrict__ src, double *__restrict__ w,
                 uint32_t *__restrict__ self, double *__restrict__ ttot_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_tasks) return;
    const uint32_t t = tasks[i];
    const uint32_t N = g.n_cells;
    const uint32_t dl = t / N, c = t - dl * N;
    const double dx = g.dirs[3 * dl], dy = g.dirs[3 * dl + 1], dz = g.dirs[3 * dl + 2];
    uint32_t e = off[i];
    double ttot = 0.0;
    for (uint32_t f = g.face_off[c]; f < g.face_off[c + 1]; ++f) {
        const double4 geo = ld_geo(g.face_geo + f);
        const double d = dot_dir(geo, dx, dy, dz);
        if (d < 0.0) {
            if (g.face_kind[f] == 0) {
                const uint32_t nb = (uint32_t)g.face_nb[f];
                const double wf = __ldg(g.face_rev + f) * (-d);
                if (st.slot_of) {
                    const uint32_t sl = st.slot_of[(size_t)dl * N + nb];
                    const double tt = st.ttot_slot[sl];
                    src[e] = sl;
                    w[e] = tt > 0.0 ? wf / tt : 0.0;
                } else {
                    src[e] = dl * N + nb;
                    w[e] = wf;
                }
                ++e;
            }
        } else if (d > 0.0) {
            ttot += geo.w * d;
        }
    }
    self[i] = st.slot_of ? st.slot_of[(size_t)dl * N + c] : t;
    ttot_out[i] = ttot;
}

__device__ __forceinline__ void mini_task(const SweepArgs &a, const MiniView &m, uint32_t i, uint32_t task) {
    const uint32_t N = a.g.n_cells;
    const uint32_t dl = task / N, c = task - dl * N;
    const uint32_t e0 = m.off[i], e1 = m.off[i + 1];
    const double *state = a.st.slot_of ? a.st.out_slot : a.st.q;
    double in = 0.0;
    constexpr int kChunk = 8;
    for (uint32_t eb = e0; eb < e1; eb += kChunk) {
        uint32_t sidx[kChunk];
        double wv[kChunk], v[kChunk];
#pragma unroll
        for (int j = 0; j < kChunk; ++j) {
            const bool ok = eb + j < e1;
            sidx[j] = ok ? m.src[eb + j] : 0u;
            wv[j] = ok ? m.w[eb + j] : 0.0;
        }
#pragma unroll
        for (int j = 0; j < kChunk; ++j) v[j] = eb + j < e1 ? __ldcg(state + sidx[j]) : 0.0;
#pragma unroll
        for (int j = 0; j < kChunk; ++j)
            if (eb + j < e1) in += v[j] * wv[j];
    }
    const double inc = in + a.src[c] / a.n_dirs_total;           // site.rs:49-56
    const int32_t p = a.pidx[c];
    const double total = p >= 0 ? inc + a.per_lag[(size_t)dl * a.n_periodic + p] : inc + 0.0;
    const double out = (total < a.threshold) ? 0.0 : total * a.att[c];
    if (a.st.slot_of) {
        __stcg(a.st.out_slot + m.self[i], out);
    } else {
        const double ttot = m.ttot[i];
        __stcg(a.st.q + m.self[i], ttot > 0.0 ? out / ttot : 0.0);
    }
    __stcg(a.incoming + (size_t)dl * N + c, inc);
}

// every wavefront level fits one block: block barrier between levels.  The records of a task are static, so a
// thread loads everything its task of the NEXT level needs (offsets, own state index, the first four upwind entries,
// the cell's absorption, source and lagged periodic term) while the gathers of the current level are in flight:
// a level then costs one dependent round trip (the state gather) instead of three.
constexpr int kMiniSmallThreads = 768;
struct MiniPre {
    uint32_t valid, i, task, e0, e1, self, src[4];
    double ttot, w[4], att, add, per;   // add = source / D, per = lagged periodic_source
};
__device__ __forceinline__ MiniPre mini_prefetch(const SweepArgs &a, const MiniView &m, const uint32_t *__restrict__ queue,
                                                 uint32_t i, bool valid) {
    MiniPre p;
    p.valid = valid ? 1u : 0u;
    p.i = i;
    if (!valid) return p;
    const uint32_t N = a.g.n_cells;
    p.task = queue[i];
    const uint32_t dl = p.task / N, c = p.task - dl * N;
    p.e0 = m.off[i];
    p.e1 = m.off[i + 1];
    p.self = m.self[i];
    p.ttot = m.ttot[i];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const bool ok = p.e0 + j < p.e1;
        p.src[j] = ok ? m.src[p.e0 + j] : 0u;
        p.w[j] = ok ? m.w[p.e0 + j] : 0.0;
    }
    p.att = a.att[c];
    p.add = a.src[c] / a.n_dirs_total;                           // site.rs:49-56
    const int32_t pp = a.pidx[c];
    p.per = pp >= 0 ? a.per_lag[(size_t)dl * a.n_periodic + pp] : 0.0;
    return p;
}
__device__ __forceinline__ void mini_solve(const SweepArgs &a, const MiniView &m, const MiniPre &p) {
    if (!p.valid) return;
    const uint32_t N = a.g.n_cells;
    const double *state = a.st.slot_of ? a.st.out_slot : a.st.q;
    double v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = p.e0 + j < p.e1 ? __ldcg(state + p.src[j]) : 0.0;
    double in = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (p.e0 + j < p.e1) in += v[j] * p.w[j];
    for (uint32_t e = p.e0 + 4; e < p.e1; ++e) in += __ldcg(state + m.src[e]) * m.w[e];   // same order as mini_task
    const double inc = in + p.add;
    const double total = inc + p.per;
    const double out = (total < a.threshold) ? 0.0 : total * p.att;
    if (a.st.slot_of) __stcg(a.st.out_slot + p.self, out);
    else __stcg(a.st.q + p.self, p.ttot > 0.0 ? out / p.ttot : 0.0);
    const uint32_t dl = p.task / N, c = p.task - dl * N;
    __stcg(a.incoming + (size_t)dl * N + c, inc);
}
__device__ __forceinline__ void mini_replay_small_body(const SweepArgs &a, const MiniView &m, const uint32_t *__restrict__ queue,
                                                       const uint32_t *__restrict__ level_off, uint32_t n_levels) {
    if (n_levels == 0) return;
    uint32_t s = level_off[0], e = level_off[1];
    MiniPre cur = mini_prefetch(a, m, queue, s + threadIdx.x, s + threadIdx.x < e);
    for (uint32_t lvl = 0; lvl < n_levels; ++lvl) {
        uint32_t s2 = e, e2 = e;
        if (lvl + 1 < n_levels) e2 = level_off[lvl + 2];
        const MiniPre nxt = mini_prefetch(a, m, queue, s2 + threadIdx.x, s2 + threadIdx.x < e2);
        mini_solve(a, m, cur);
        for (uint32_t i = s + threadIdx.x + blockDim.x; i < e; i += blockDim.x) mini_task(a, m, i, queue[i]);
        __syncthreads();   // global writes of this level are visible to the block's next level
        cur = nxt;
        s = s2;
        e = e2;
    }
}
__global__ void __launch_bounds__(kMiniSmallThreads)
mini_replay_small_kernel(SweepArgs a, MiniView m, const uint32_t *__restrict__ queue,
                         const uint32_t *__restrict__ level_off, uint32_t n_levels) {
    mini_replay_small_body(a, m, queue, level_off, n_levels);
}

__global__ void __launch_bounds__(256)
mini_replay_kernel(SweepArgs a, MiniView m, const uint32_t *__restrict__ queue,
                   const uint32_t *__restrict__ level_off, uint32_t n_levels) {
    cg::grid_group grid = cg::this_grid();
    const unsigned int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int gsz = gridDim.x * blockDim.x;
    for (uint32_t lvl = 0; lvl < n_levels; ++lvl) {
        const uint32_t s = level_off[lvl], e = level_off[lvl + 1];
        for (uint32_t i = s + gtid; i < e; i += gsz) mini_task(a, m, i, queue[i]);
        if (lvl + 1 < n_levels) grid.sync();
    }
}

// ------------------------------------------------------------------------------------------
// periodic_source rows: sum over LocalPeriodic upwind faces (handle_local_periodic_neighbour,
// src/sweep/mod.rs:505-513, in gather form).  One thread per (periodic cell, local direction).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
periodic_gather_kernel(GridView g, const uint32_t *__restrict__ pcells, uint32_t n_periodic,
                       const uint32_t *__restrict__ act_list, uint32_t n_act,
                       const int32_t *__restrict__ pidx, StateView st, double *__restrict__ dst) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    const int dl = blockIdx.y;
    uint32_t c, p;
    if (act_list) {   // partial active set: only the rows the tasks of this sweep read
        if (k >= n_act) return;
        c = act_list[k];
        const int32_t pp = pidx[c];
        if (pp < 0) return;
        p = (uint32_t)pp;
    } else {
        if (k >= n_periodic) return;
        p = k;
        c = pcells[p];
    }
    const double dx = g.dirs[3 * dl], dy = g.dirs[3 * dl + 1], dz = g.dirs[3 * dl + 2];
    double acc = 0.0;
    for (uint32_t f = g.face_off[c]; f < g.face_off[c + 1]; ++f) {
        if (g.face_kind[f] != 2) continue;
        const double d = dot_dir(ld_geo(g.face_geo + f), dx, dy, dz);
        if (d < 0.0) acc += st.load_q(dl, (uint32_t)g.face_nb[f]) * (g.face_rev[f] * (-d));
    }
    dst[(size_t)dl * n_periodic + p] = acc;
}

// ssw_set_directions: the flux state for a new direction set.  New direction dl takes the outgoing rate of the old
// direction map[dl] it is best aligned with; q = out / sum_downwind(A n.d) with the NEW directions (g.dirs).
__global__ void __launch_bounds__(256)
remap_directions_kernel(GridView g, const double *__restrict__ out_old_cell_major, const int32_t *__restrict__ map,
                        int n_local_dirs, double *__restrict__ q_new) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= g.n_cells) return;
    for (int dl = 0; dl < n_local_dirs; ++dl) {
        const double dx = g.dirs[3 * dl], dy = g.dirs[3 * dl + 1], dz = g.dirs[3 * dl + 2];
        double ttot = 0.0;
        for (uint32_t f = g.face_off[c]; f < g.face_off[c + 1]; ++f) {
            const double4 geo = ld_geo(g.face_geo + f);
            const double d = dot_dir(geo, dx, dy, dz);
            if (d > 0.0) ttot += geo.w * d;
        }
        const double out = out_old_cell_major[(size_t)c * n_local_dirs + map[dl]];
        q_new[(size_t)dl * g.n_cells + c] = ttot > 0.0 ? out / ttot : 0.0;
    }
}

// incoming_total_rate for every (cell, dir) from the current q (read-back / photon_rate,
// src/sweep/mod.rs:727-730).  which: 0 incoming (Local faces), 2 periodic_source, 1 outgoing.
__global__ void __launch_bounds__(256)
dir_state_kernel(GridView g, StateView st, int which, int n_local_dirs,
                 double *__restrict__ out_cell_major, double *__restrict__ photon_rate) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= g.n_cells) return;
    double total = 0.0;
    for (int dl = 0; dl < n_local_dirs; ++dl) {
        const double dx = g.dirs[3 * dl], dy = g.dirs[3 * dl + 1], dz = g.dirs[3 * dl + 2];
        double acc = 0.0, ttot = 0.0;
        for (uint32_t f = g.face_off[c]; f < g.face_off[c + 1]; ++f) {
            const double4 geo = ld_geo(g.face_geo + f);
            const double d = dot_dir(geo, dx, dy, dz);
            const int kind = g.face_kind[f];
            if (d < 0.0) {
                if ((which == 0 && kind == 0) || (which == 2 && kind == 2))
                    acc += st.load_q(dl, (uint32_t)g.face_nb[f]) * (g.face_rev[f] * (-d));
            } else if (d > 0.0) {
                ttot += geo.w * d;
            }
        }
        if (which == 1) acc = st.load_out(dl, c, ttot);
        if (out_cell_major) out_cell_major[(size_t)c * n_local_dirs + dl] = acc;
        total += acc;
    }
    if (photon_rate) photon_rate[c] = total;
}

// photon_rate bookkeeping for partial active sets (src/sweep/mod.rs:487-503, 727-730): a sweep over
// the active set A changes incoming_total_rate of every Local neighbour of A, active or not.
// mark_touched_kernel flags those neighbours; photon_patch_kernel re-evaluates sum_d incoming[d]
// for them -- one block per touched cell, one thread per local direction, summed in direction order.
__global__ void __launch_bounds__(256)
mark_touched_kernel(GridView g, const uint32_t *__restrict__ act_list, uint32_t n_act, uint8_t *__restrict__ flags) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_act) return;
    const uint32_t c = act_list[k];
    for (uint32_t f = g.face_off[c]; f < g.face_off[c + 1]; ++f)
        if (g.face_kind[f] == 0) flags[g.face_nb[f]] = 1;
}

__global__ void __launch_bounds__(kMaxDirs)
photon_patch_kernel(GridView g, StateView st, const uint32_t *__restrict__ touch_list, int n_local_dirs,
                    double *__restrict__ photon) {
    __shared__ double s_in[kMaxDirs];
    const uint32_t c = touch_list[blockIdx.x];
    const int dl = threadIdx.x;
    if (dl < n_local_dirs) {
        const double dx = g.dirs[3 * dl], dy = g.dirs[3 * dl + 1], dz = g.dirs[3 * dl + 2];
        double acc = 0.0;
        constexpr int kChunk = 8;
        const uint32_t f0 = g.face_off[c], f1 = g.face_off[c + 1];
        for (uint32_t fb = f0; fb < f1; fb += kChunk) {
            double wd[kChunk], qv[kChunk];
            uint32_t nb[kChunk];
            bool up[kChunk];
#pragma unroll
            for (int j = 0; j < kChunk; ++j) {
                const uint32_t f = fb + j;
                wd[j] = 0.0; nb[j] = 0; up[j] = false;
                if (f < f1) {
                    const double d = dot_dir(ld_geo(g.face_geo + f), dx, dy, dz);
                    up[j] = d < 0.0 && g.face_kind[f] == 0;
                    nb[j] = (uint32_t)g.face_nb[f];
                    wd[j] = __ldg(g.face_rev + f) * (-d);
                }
            }
#pragma unroll
            for (int j = 0; j < kChunk; ++j) qv[j] = up[j] ? st.load_q(dl, nb[j]) : 0.0;
#pragma unroll
            for (int j = 0; j < kChunk; ++j)
                if (up[j]) acc += qv[j] * wd[j];
        }
        s_in[dl] = acc;
    }
    __syncthreads();
    if (dl == 0) {
        double total = 0.0;
        for (int k = 0; k < n_local_dirs; ++k) total += s_in[k];
        photon[c] = total;
    }
}

// ------------------------------------------------------------------------------------------
// rate reduction over this rank's directions: sum_d get_rate(d) (src/sweep/mod.rs:554-558,
// site.rs:53-56), left fold in direction order.  One thread per active cell; consecutive
// threads read consecutive cells of incoming[dl][*] (coalesced).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rate_kernel(const uint32_t *__restrict__ act_list, uint32_t n_act, uint32_t n_cells,
            int n_local_dirs, const double *__restrict__ incoming, const int32_t *__restrict__ pidx,
            const double *__restrict__ per_new, uint32_t n_periodic, double *__restrict__ rate_act) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_act) return;
    const uint32_t c = act_list ? act_list[k] : k;
    const int32_t p = pidx[c];
    double rate = 0.0;
    if (p >= 0) {
        for (int dl = 0; dl < n_local_dirs; ++dl)
            rate += __ldcs(incoming + (size_t)dl * n_cells + c) + per_new[(size_t)dl * n_periodic + p];
    } else {
        for (int dl = 0; dl < n_local_dirs; ++dl) rate += __ldcs(incoming + (size_t)dl * n_cells + c) + 0.0;
    }
    rate_act[k] = rate;
}

// the same fold for a small active set: one block per active cell, one thread per direction loads its
// term, thread 0 folds them in direction order (bit-identical to rate_kernel, without its serial loads)
__global__ void __launch_bounds__(kMaxDirs)
rate_small_kernel(const uint32_t *__restrict__ act_list, uint32_t n_cells, int n_local_dirs,
                  const double *__restrict__ incoming, const int32_t *__restrict__ pidx,
                  const double *__restrict__ per_new, uint32_t n_periodic, double *__restrict__ rate_act) {
    __shared__ double s_term[kMaxDirs];
    const uint32_t k = blockIdx.x;
    const uint32_t c = act_list[k];
    const int dl = threadIdx.x;
    if (dl < n_local_dirs) {
        const int32_t p = pidx[c];
        const double per = p >= 0 ? per_new[(size_t)dl * n_periodic + p] : 0.0;
        s_term[dl] = __ldcs(incoming + (size_t)dl * n_cells + c) + per;
    }
    __syncthreads();
    if (dl == 0) {
        double rate = 0.0;
        for (int d = 0; d < n_local_dirs; ++d) rate += s_term[d];
        rate_act[k] = rate;
    }
}

// {exp(-n_HI sigma size), source / D} per cell: the 16-byte record the compiled sweep gathers
__global__ void __launch_bounds__(256)
cellrec_kernel(const double *__restrict__ att, const double *__restrict__ src, double n_dirs_total, uint32_t n,
               double2 *__restrict__ cellrec) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n) cellrec[c] = make_double2(att[c], src[c] / n_dirs_total);
}

// ------------------------------------------------------------------------------------------
// chemistry: Sweep::update_chemistry (src/sweep/mod.rs:549-574) for one active cell per thread
// ------------------------------------------------------------------------------------------
struct ChemParams {
    double max_timestep, threshold, scale_factor, safety;
    int prevent_cooling;
};

struct ChemStats {
    unsigned long long cells, failures, attempts;
    unsigned int max_depth;
};

// Direction sharding over peer-mapped memory (peer.cuh): the chemistry of a cell runs on its owner only.  The rate is
// the sum of the W partials the ranks stored into the owner's receive buffer, folded in rank order; the new absorption
// factor goes to the `att` array of every rank.
struct PeerChem {
    int32_t world;              // 0 / 1: not sharded this way
    uint32_t first, n_own, n_per;
    const double *recv;         // W x n_per partial rates of the own slice
    double *att[16];            // the att array of every rank
    unsigned int *counter[16];  // the arrival counter of every rank (peer.cuh)
    unsigned int *blocks_done;  // this handle's block check-in counter
};

// Work balancing: the substep count of a cell spans three decades at an ionization front and a warp runs as long as
// its slowest lane (measured on the front workload: 8.3 of 32 lanes active on average).  `order` (optional) is a
// permutation of the launch indices sorted by the substep count the cell needed LAST time (chem_order_*_kernel +
// radix sort), so the lanes of a warp hold cells of similar cost.  Every cell's arithmetic is independent of which
// thread runs it: results are bit-identical with and without.
__global__ void __launch_bounds__(256)
chem_order_keys_kernel(const uint32_t *__restrict__ act_list, uint32_t n, uint32_t first_cell, const uint16_t *__restrict__ last_attempts,
                       uint32_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t c = act_list ? act_list[k] : first_cell + k;
    keys[k] = 65535u - last_attempts[c];   // heaviest first: the long warps start early
    vals[k] = k;
}

constexpr int kChemThreads = 32;   // one warp per block (see the statistics fold at the end of the kernel)
__global__ void __launch_bounds__(128)
chemistry_kernel(CellView cv, const uint32_t *__restrict__ act_list, uint32_t n_act,
                 const double *__restrict__ rate_act, ChemParams cp, ChemStats *stats, uint32_t first_cell, PeerChem pc,
                 const uint32_t *__restrict__ order, uint16_t *__restrict__ last_attempts) {
    const uint32_t kk = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long attempts = 0;
    unsigned int depth = 0, failed = 0, mine = 0;
    bool active = kk < n_act;
    uint32_t c = 0, k = kk;
    if (active) {
        if (order) k = order[kk];
        c = act_list ? act_list[k] : first_cell + k;   // no list: the contiguous cells [first_cell, first_cell + n_act)
        if (pc.world > 1 && (c < pc.first || c >= pc.first + pc.n_own)) active = false;   // another rank owns the cell
    }
    if (active) {
        mine = 1;
        double rate;
        if (pc.world > 1) {
            rate = 0.0;
            for (int p = 0; p < pc.world; ++p) rate += pc.recv[(size_t)p * pc.n_per + (c - pc.first)];
        } else {
            rate = rate_act[k];
        }
        const int lvl = cv.level[c];
        const double timestep = cp.max_timestep * exp2(-(double)lvl);  // max_timestep * 0.5^level (exact)
        double relative_change;
        if (fabs(rate) < fabs(cp.threshold)) {                        // chemistry/mod.rs:79-81
            relative_change = 0.0;
        } else {                                                      // chemistry/mod.rs:73-77
            relative_change = fabs(fmin(fabs(fabs(rate - cv.prev_rate[c]) / rate), kInvEps));
        }
        cv.prev_rate[c] = rate;
        const double rate_timescale = timestep / relative_change;
        Solver s;
        s.xhii = cv.x[c];
        s.temperature = cv.T[c];
        s.density = cv.rho[c];
        s.volume = cv.volume[c];
        s.length = cv.size[c];
        s.rate = rate;
        s.scale_factor = cp.scale_factor;
        s.has_floor = cp.prevent_cooling != 0;
        s.floor_temperature = s.temperature;
        s.floor_xhii = s.xhii;
        const ChemResult r = perform_timestep(s, timestep, cp.safety);
        cv.T[c] = s.temperature;
        cv.x[c] = s.xhii;
        cv.ts[c] = r.timescale;
        cv.tau[c] = (rate_timescale < r.timescale) ? rate_timescale : r.timescale;  // Timescale::min
        const double att = non_absorbed_fraction(s.density, s.xhii, s.length);
        if (pc.world > 1) {
            for (int p = 0; p < pc.world; ++p) pc.att[p][c] = att;
        } else {
            cv.att[c] = att;
        }
        attempts = r.attempts;
        depth = (unsigned)r.max_depth;
        failed = (unsigned)r.failed;
        last_attempts[c] = (uint16_t)(attempts < 65535ull ? attempts : 65535ull);
    }
    // statistics: folded over the warp, one reduction per warp.  No block barrier: the substep count of a cell spans three
    // decades at an ionization front, and a warp that waits for the slowest warp of its block keeps its registers
    // (measured on the front workload: 31 % of the warp samples sat at the barrier that used to be here).
    for (int o = 16; o > 0; o >>= 1) {
        attempts += __shfl_down_sync(0xffffffffu, attempts, o);
        failed += __shfl_down_sync(0xffffffffu, failed, o);
        mine += __shfl_down_sync(0xffffffffu, mine, o);
        depth = max(depth, __shfl_down_sync(0xffffffffu, depth, o));
    }
    if ((threadIdx.x & 31) == 0 && mine) {
        atomicAdd(&stats->attempts, attempts);
        atomicAdd(&stats->cells, (unsigned long long)mine);
        if (failed) atomicAdd(&stats->failures, (unsigned long long)failed);
        if (depth) atomicMax(&stats->max_depth, depth);
    }
    if (pc.world > 1) {
        // the new absorption factors are in every rank's array: tell them (the all-gather's completion signal, from the
        // last block of this kernel; see peer_signal_tail)
        __syncthreads();
        if (threadIdx.x == 0) __threadfence_system();   // cumulative over the block's stores (seen through the barrier)
        if (threadIdx.x == 0 && atomicAdd(pc.blocks_done, 1u) == gridDim.x - 1u) {
            *pc.blocks_done = 0u;
            __threadfence_system();
            for (int p = 0; p < pc.world; ++p)
                asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(pc.counter[p]) : "memory");
        }
    }
}

// Cell-sliced chemistry of a direction-sharded job: every rank updates the cells
// [rank * n_per, (rank + 1) * n_per) and packs what the other ranks need -- x, T, change_timescale, timestep and
// the all-reduced rate (= previous_incoming_total_rate) -- into its chunk of the all-gather buffer
// pack[rank][field][k]; after the all-gather every rank unpacks the other ranks' chunks.
constexpr int kPackFields = 5;
__global__ void __launch_bounds__(256)
chem_pack_kernel(CellView cv, uint32_t first_cell, uint32_t n_own, uint32_t n_per, double *__restrict__ chunk) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_own) return;
    const uint32_t c = first_cell + k;
    chunk[k] = cv.x[c];
    chunk[(size_t)n_per + k] = cv.T[c];
    chunk[2 * (size_t)n_per + k] = cv.tau[c];
    chunk[3 * (size_t)n_per + k] = cv.ts[c];
    chunk[4 * (size_t)n_per + k] = cv.prev_rate[c];
}

__global__ void __launch_bounds__(256)
chem_unpack_kernel(CellView cv, uint32_t n_cells, uint32_t n_per, uint32_t my_rank, const double *__restrict__ pack) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    const uint32_t r = c / n_per, k = c - r * n_per;
    if (r == my_rank) return;
    const double *chunk = pack + (size_t)r * kPackFields * n_per;
    const double x = chunk[k];
    cv.x[c] = x;
    cv.T[c] = chunk[(size_t)n_per + k];
    cv.tau[c] = chunk[2 * (size_t)n_per + k];
    cv.ts[c] = chunk[3 * (size_t)n_per + k];
    cv.prev_rate[c] = chunk[4 * (size_t)n_per + k];
    cv.att[c] = non_absorbed_fraction(cv.rho[c], x, cv.size[c]);
}

// att = exp(-n_HI sigma size) for all cells (create / set_inputs)
__global__ void __launch_bounds__(256) attenuation_kernel(CellView cv, uint32_t n) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n) cv.att[c] = non_absorbed_fraction(cv.rho[c], cv.x[c], cv.size[c]);
}

// ------------------------------------------------------------------------------------------
// timestep levels: Sweep::update_timestep_levels (src/sweep/mod.rs:576-589),
// TimestepLevel::from_max_timestep_and_desired_timestep (timestep_level.rs:27-36),
// get_desired_level_from_desired_timestep (timestep_state.rs:54-64) + per-level histogram
// ------------------------------------------------------------------------------------------
__host__ __device__ inline int level_rule(int max_num_levels, double max_timestep, double desired) {
    const double ratio = max_timestep / desired;
    const double l = ceil(log2(ratio));
    // Rust `as usize`: NaN -> 0, negative -> 0, huge -> usize::MAX; then clamp(0, L-1)
    int level;
    if (!(l > 0.0)) level = 0;
    else if (l >= (double)(max_num_levels - 1)) level = max_num_levels - 1;
    else level = (int)l;
    return level;
}

// `first` / `n`: the cells [first, first + n) (all cells, or the slice this rank owns); with peers the new level is
// stored into the level array of every rank (peer.cuh)
struct PeerLevels {
    int32_t world;
    uint8_t *level[16];
};
__global__ void __launch_bounds__(256)
levels_kernel(const double *__restrict__ tau, uint8_t *__restrict__ level, uint32_t first, uint32_t n, int n_levels,
              double max_timestep, double safety, int lowest_allowed,
              unsigned long long *__restrict__ hist /* 32 counts + [32] = number of changed cells */, PeerLevels pl) {
    __shared__ unsigned int s_hist[32];
    if (threadIdx.x < 32) s_hist[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t c = first + k;
    bool changed = false;
    if (k < n) {
        int lv = level_rule(n_levels, max_timestep, safety * tau[c]);
        if (lv < lowest_allowed) lv = lowest_allowed;
        if (level[c] != (uint8_t)lv) {
            if (pl.world > 1) {
                for (int p = 0; p < pl.world; ++p) pl.level[p][c] = (uint8_t)lv;
            } else {
                level[c] = (uint8_t)lv;
            }
            changed = true;
        }
        atomicAdd(&s_hist[lv], 1u);
    }
    const int any_changed = __syncthreads_or(changed ? 1 : 0);
    if (threadIdx.x < 32 && s_hist[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)s_hist[threadIdx.x]);
    if (threadIdx.x == 0 && any_changed) atomicAdd(&hist[32], 1ull);
}

__global__ void __launch_bounds__(256)
histogram_kernel(const uint8_t *__restrict__ level, uint32_t n, unsigned long long *__restrict__ hist) {
    __shared__ unsigned int s_hist[32];
    if (threadIdx.x < 32) s_hist[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n) atomicAdd(&s_hist[level[c] & 31], 1u);
    __syncthreads();
    if (threadIdx.x < 32 && s_hist[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)s_hist[threadIdx.x]);
}

// ionization_time (src/sweep/mod.rs:731-738): first time xHII > 0.5; +inf = IonizationTime::default() = not yet
__global__ void __launch_bounds__(256)
ionization_time_kernel(const double *__restrict__ x, double *__restrict__ ion_time, uint32_t first, uint32_t n, double now) {
    const uint32_t c = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (c < first + n && x[c] > 0.5 && isinf(ion_time[c])) ion_time[c] = now;
}

// optional chemistry outputs, src/sweep/chemistry_output.rs:25-55 with Sweep::get_solver (mod.rs:612-632)
__global__ void __launch_bounds__(256)
chem_output_kernel(CellView cv, uint32_t n, const double *__restrict__ rate, double scale_factor,
                   int field, double *__restrict__ out) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    Solver s;
    s.xhii = cv.x[c];
    s.temperature = cv.T[c];
    s.density = cv.rho[c];
    s.volume = cv.volume[c];
    s.length = cv.size[c];
    s.rate = rate[c];
    s.scale_factor = scale_factor;
    s.has_floor = false;
    s.floor_temperature = 0.0;
    s.floor_xhii = 0.0;
    const double one_year = 1.0 * units::years;
    double v;
    if (field == 5) v = s.photoionization_rate(one_year);
    else if (field == 6) v = s.photoheating_rate(one_year) - s.cooling_rate();
    else if (field == 7) v = s.alpha_b() * s.ne() * s.xhii;
    else v = s.beta() * s.ne() * (1.0 - s.xhii);
    out[c] = v;
}

// standalone chemistry on independent cells (ssw_chemistry_batch)
__global__ void __launch_bounds__(128)
chemistry_batch_kernel(uint64_t n, double *x, double *T, const double *rho, const double *vol,
                       const double *len, const double *rate, const double *dt, double scale_factor,
                       double safety, int prevent_cooling, double *timescale, int *process, int *depth,
                       unsigned long long *attempts) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Solver s;
    s.xhii = x[i];
    s.temperature = T[i];
    s.density = rho[i];
    s.volume = vol[i];
    s.length = len[i];
    s.rate = rate[i];
    s.scale_factor = scale_factor;
    s.has_floor = prevent_cooling != 0;
    s.floor_temperature = s.temperature;
    s.floor_xhii = s.xhii;
    const ChemResult r = perform_timestep(s, dt[i], safety);
    x[i] = s.xhii;
    T[i] = s.temperature;
    timescale[i] = r.timescale;
    process[i] = r.failed ? -1 : r.process;
    depth[i] = r.max_depth;
    attempts[i] = r.attempts;
}

// ------------------------------------------------------------------------------------------
// time series of compute_time_series_system (src/sweep/time_series.rs:61-155): mass- and
// volume-weighted sums over all cells.  Two deterministic passes: per-block partial sums (fixed
// tree), then one block folds the partials in block order.
//   sums[0] = sum m x   [1] = sum m   [2] = sum V x   [3] = sum V   [4] = sum T m   [5] = sum T V
//   [6] = sum Gamma V   [7] = sum Gamma x V        (Gamma = photoionization rate, optional)
// ------------------------------------------------------------------------------------------
constexpr int kSeriesSums = 8;

__global__ void __launch_bounds__(256)
time_series_partial_kernel(CellView cv, uint32_t n, const double *__restrict__ mass, const double *__restrict__ gamma,
                           double *__restrict__ partial) {
    __shared__ double s_red[kSeriesSums][8];
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    double v[kSeriesSums];
#pragma unroll
    for (int i = 0; i < kSeriesSums; ++i) v[i] = 0.0;
    if (c < n) {
        const double vol = cv.volume[c], x = cv.x[c], T = cv.T[c];
        const double m = mass ? mass[c] : cv.rho[c] * vol;
        v[0] = m * x; v[1] = m; v[2] = vol * x; v[3] = vol; v[4] = T * m; v[5] = T * vol;
        if (gamma) { v[6] = gamma[c] * vol; v[7] = gamma[c] * x * vol; }
    }
#pragma unroll
    for (int i = 0; i < kSeriesSums; ++i) {
        for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_down_sync(0xffffffffu, v[i], o);
        if ((threadIdx.x & 31) == 0) s_red[i][threadIdx.x >> 5] = v[i];
    }
    __syncthreads();
    if (threadIdx.x < kSeriesSums) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += s_red[threadIdx.x][w];
        partial[(size_t)blockIdx.x * kSeriesSums + threadIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256)
time_series_final_kernel(const double *__restrict__ partial, uint32_t n_blocks, double *__restrict__ sums) {
    __shared__ double s_red[kSeriesSums][32];
    // thread (i, j): quantity i = tid / 32, lane j folds blocks j, j + 32, ... in order
    const int i = threadIdx.x >> 5, j = threadIdx.x & 31;
    double t = 0.0;
    for (uint32_t b = j; b < n_blocks; b += 32) t += partial[(size_t)b * kSeriesSums + i];
    s_red[i][j] = t;
    __syncthreads();
    if (j == 0) {
        double total = 0.0;
        for (int k = 0; k < 32; ++k) total += s_red[i][k];
        sums[i] = total;
    }
}

__global__ void __launch_bounds__(256)
iota_active_kernel(const uint8_t *__restrict__ level, int cur, uint32_t n, uint8_t *__restrict__ flags) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n) flags[c] = level[c] >= cur ? 1 : 0;
}


// sum_d get_rate(d) for every cell from the per-direction sums (Sweep::get_solver, mod.rs:616-620)
__global__ void __launch_bounds__(256)
combine_rates_kernel(const double *__restrict__ in_sum, const double *__restrict__ per_sum,
                     const double *__restrict__ src, double local_dir_fraction, uint32_t n,
                     double *__restrict__ out) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n) out[c] = in_sum[c] + src[c] * local_dir_fraction + per_sum[c];
}

}  // namespace ssw
