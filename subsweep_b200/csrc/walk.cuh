// walk.cuh -- the all-cells sweep as one wavefront walk per direction ("solo" form of the compiled schedule).
//
// Tasks of different directions never interact inside a sweep (src/sweep/mod.rs:412-513: every array is indexed
// [cell][dir]).  So one thread block owns one local direction for the whole sweep and walks that direction's
// wavefront levels in order; consecutive levels are separated by a BLOCK barrier only -- no device-wide barrier, no
// flags, no spinning, grid = number of local directions.  What makes this fast on unstructured grids:
//
//   * slots are numbered direction-major (direction, wavefront level, upwind degree descending, cell): a block's
//     outgoing rates are one contiguous range of out_slot, written in order;
//   * the last W of them (W = 8192 ... 16384 slots, 64 ... 128 KB) also live in a shared-memory WINDOW.  An upwind
//     neighbour of a Voronoi cell was solved at most a dozen levels earlier, i.e. within the window: its rate is read
//     from shared memory (29-cycle LDS) instead of through L1TEX/L2 (a fully divergent 8-byte gather costs the SM one
//     tag cycle per lane -- the measured bound of the gather-from-global form on 84 SMs, profiles/r2b_*).  Sources
//     outside the window, periodic donors solved later in the sweep (they still hold last sweep's value = the
//     reference's lag, src/sweep/mod.rs:505-513) and the snapshot slots of the other periodic donors are read from
//     global memory;
//   * nothing on the dependent chain of a tile touches global memory.  A dedicated producer lane streams, several
//     tiles ahead, the tile's static packet and its slot-ordered absorption records (written per sweep by a
//     throughput kernel) with TMA bulk copies (cp.async.bulk + mbarrier complete_tx) into a shared-memory ring; a
//     gather warp fetches the tile's sources outside the window with 8-byte cp.async copies into the same stage
//     (completion through cp.async.mbarrier.arrive); the compute warps only wait on the two mbarriers.  Inside a
//     packet the upwind entries are stored by ROUND: round j holds the j-th entry of every slot that has more than j
//     entries; with the slots of a tile sorted by degree (descending) thread t reads entry `round_off[j] + t` --
//     consecutive threads, consecutive addresses, no bank conflicts, and warps leave the loop together;
//   * per slot the arithmetic is that of the other forms bit for bit: product of (donor rate, precomputed share)
//     rounded, summed Local faces in face order, then the periodic ones (src/sweep/mod.rs:453-461, site.rs:49-56).
//
// The incoming rate of every task is stored once per sweep in acc_cell[direction][cell]; s_rate_finish_kernel folds the
// directions in order -- exactly the reference's left fold over d (src/sweep/mod.rs:554-558).  (Storing it in slot order
// and gathering through slot_of in the rate kernel was measured: 176 M random 8-byte gathers cost 2.8 ms at 128^3.)
// Included from stream.cuh behind its TileDesc and the mbarrier / TMA helpers.
#pragma once

namespace ssw {

constexpr uint32_t kWalkNear = 0x80000000u;   // entry code: window index (low bits); else index into the tile's gathered values
constexpr int kWalkMaxDeg = 62;               // upwind entries per task in the walk form (else: level-barrier stream)
constexpr int kWalkMaxStages = 8;
constexpr int kWalkGatherWarps = 2;           // warps that fetch the sources outside the window (tiles dealt round-robin)
constexpr int kWalkMaxGroups = 8;             // compute groups of a block (independent tiles of one level run side by side)

struct WalkHeader {        // first 32 bytes of a packet
    uint32_t slot0;        // first global slot of the tile (real tiles), first epilogue slot otherwise
    uint16_t n, n_entries;
    uint32_t local0;       // (slot0 - first slot of the direction) mod window: window position of the tile's first slot
    uint16_t n_far, epilogue;
    uint32_t group;        // local direction
    uint32_t need;         // tiles of the block that must be complete before this one starts: index of the first tile of its level
    uint32_t pad[2];
};
static_assert(sizeof(WalkHeader) == 32, "walk packet header is 32 bytes");

struct WalkLayout {
    uint32_t roff, w, code, far, cell, info, bytes;   // static packet (streamed from global memory)
    uint32_t rec, val, stage_bytes;                   // behind it in the ring stage: absorption records, gathered values
};
__host__ __device__ inline uint32_t walk_align16(uint32_t x) { return (x + 15u) & ~15u; }
// packet = [header 32 B] [u16 round_off[64]] [f64 share[E]] [u32 code[E]] [u32 far source slot[n_far]] [u32 cell[n]] [u32 info[n]]
// entry (slot t, round j) sits at index round_off[j] + t;  info = degree | periodic entries << 8
__host__ __device__ inline WalkLayout walk_layout(uint32_t n, uint32_t E, uint32_t n_far) {
    WalkLayout L;
    uint32_t o = (uint32_t)sizeof(WalkHeader);
    L.roff = o;  o += 128u;
    L.w = o;     o += walk_align16(8u * E);
    L.code = o;  o += walk_align16(4u * E);
    L.far = o;   o += walk_align16(4u * n_far);
    L.cell = o;  o += walk_align16(4u * n);
    L.info = o;  o += walk_align16(4u * n);
    L.bytes = o;
    L.rec = o;   o += 16u * n;
    L.val = o;   o += walk_align16(8u * n_far);
    L.stage_bytes = o;
    return L;
}

// sort key of the walk form: direction << 40 | (63 - min(degree, 63)) << 32 | (cell * Dl + dl); degree = number of
// flux-carrying upwind faces (Local + LocalPeriodic) of the task
__global__ void __launch_bounds__(256)
w_key64_kernel(GridView g, const uint32_t *__restrict__ tasks, uint32_t n, uint32_t n_dl,
               unsigned long long *__restrict__ keys) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint32_t t = tasks[s];
    const uint32_t dl = t / g.n_cells, c = t - dl * g.n_cells;
    const double dx = g.dirs[3 * dl], dy = g.dirs[3 * dl + 1], dz = g.dirs[3 * dl + 2];
    uint32_t deg = 0;
    for (uint32_t f = g.face_off[c]; f < g.face_off[c + 1]; ++f) {
        const int kind = g.face_kind[f];
        if ((kind == 0 || kind == 2) && dot_dir(ld_geo(g.face_geo + f), dx, dy, dz) < 0.0) ++deg;
    }
    keys[s] = ((unsigned long long)dl << 40) | ((unsigned long long)(63u - min(deg, 63u)) << 32) |
              (unsigned long long)(c * n_dl + dl);
}

// level-major -> direction-major: slot s of pseudo-level (level l, direction g) moves by delta[l * G + g]
__global__ void __launch_bounds__(256)
w_permute_kernel(const unsigned long long *__restrict__ keys64, uint32_t n, const uint32_t *__restrict__ pl_off_old,
                 uint32_t n_real_pl, const uint32_t *__restrict__ delta, uint32_t *__restrict__ keys_new) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    uint32_t lo = 0, hi = n_real_pl;   // pl_off_old[lo] <= s < pl_off_old[hi]
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (pl_off_old[mid] <= s) lo = mid;
        else hi = mid;
    }
    keys_new[s + delta[lo]] = (uint32_t)(keys64[s] & 0xffffffffull);
}

// Tile cutting of the walk form, one thread per pseudo-level: a tile is the longest run of slots with at most
// max_slots slots and max_entries upwind entries (the slots of a pseudo-level are sorted by degree, so the first
// tiles of a level hold fewer, heavier slots).  Without tile_start only counts.
__global__ void __launch_bounds__(128)
w_cut_kernel(const unsigned long long *__restrict__ upoff, const uint32_t *__restrict__ pl_off, uint32_t n_pl,
             uint32_t max_slots, uint32_t max_entries, const uint32_t *__restrict__ tile_off,
             uint32_t *__restrict__ tile_cnt, uint32_t *__restrict__ tile_start) {
    const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_pl) return;
    const uint32_t end = pl_off[l + 1];
    uint32_t start = pl_off[l], count = 0;
    const uint32_t out = tile_start ? tile_off[l] : 0;
    while (start < end) {
        if (tile_start) tile_start[out + count] = start;
        ++count;
        const unsigned long long lim = upoff[start] + max_entries;
        uint32_t lo = start + 1, hi = min(start + max_slots, end);   // the tile ends in [lo, hi]
        while (lo < hi) {
            const uint32_t mid = (lo + hi + 1) >> 1;
            if (upoff[mid] <= lim) lo = mid;
            else hi = mid - 1;
        }
        start = lo;
    }
    if (!tile_start) tile_cnt[l] = count;
}

__global__ void __launch_bounds__(256)
w_cell_of_slot_kernel(const uint32_t *__restrict__ keys, uint32_t n, uint32_t n_dl, uint32_t *__restrict__ cell_of_slot) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) cell_of_slot[s] = keys[s] / n_dl;
}

// is the upwind source `src` of a task in the tile [slot0, slot0 + n) read from the shared-memory window?  It was
// solved by this block before the tile started and is still inside the window when the tile runs; `window` is the
// usable part: the physical window minus the slots of the tiles that may be in flight at the same time.
__device__ __forceinline__ bool walk_is_near(uint32_t src, uint32_t slot0, uint32_t n, uint32_t n_tasks, uint32_t window) {
    return src < n_tasks && src < slot0 && slot0 + n - src <= window;
}

struct WalkFillArgs {
    GridView g;
    const uint32_t *keys;
    const uint32_t *slot_of;
    const int32_t *pidx;
    const unsigned long long *upoff;   // n_all + 1
    const double *ttot_slot;
    const uint32_t *pl_off;            // direction-major pseudo-level offsets
    uint32_t n_pl, n_real_pl, n_dl, n_tasks, n_levels, window, window_usable;
    // count pass (tile order): first slot and pseudo-level of every tile -> far entries per tile
    const uint32_t *tile_start, *tile_level;
    uint32_t *tile_far;
    // fill pass (block-major tile table)
    const TileDesc *tab;
    const uint32_t *tab_block;
    const uint64_t *stream_off;        // per block
    unsigned char *stream;
    uint32_t *lag_src;
    unsigned int *lag_counter;         // [0] fill cursor, [1] error flag, [2] near / [3] far entry counts
};

// One thread block per tile, slot t of the tile = thread t.  COUNT: far entries of the tile (tile order; sizes the
// packets).  Otherwise writes the tile's packet.  Both passes enumerate the entries identically, except that only the
// fill pass allocates snapshot slots: a periodic donor in the same or an earlier level is far in both.
template <bool COUNT>
__global__ void __launch_bounds__(512)
w_fill_kernel(WalkFillArgs a) {
    __shared__ uint32_t s_cnt[64];    // s_cnt[j] = slots with more than j entries
    __shared__ uint32_t s_roff[65];
    __shared__ uint32_t s_far;
    uint32_t slot0, n, n_entries = 0, level, off_far_total = 0, need = 0;
    unsigned char *pkt = nullptr;
    if (COUNT) {
        slot0 = a.tile_start[blockIdx.x];
        n = a.tile_start[blockIdx.x + 1] - slot0;
        level = a.tile_level[blockIdx.x];
    } else {
        const TileDesc d = a.tab[blockIdx.x];
        slot0 = d.slot0; n = d.n; n_entries = d.n_entries; level = d.level; off_far_total = d.aux & 0xffffu; need = d.need;
        pkt = a.stream + a.stream_off[a.tab_block[blockIdx.x]] + (size_t)d.off16 * 16u;
    }
    const WalkLayout L = walk_layout(n, n_entries, off_far_total);
    uint16_t *roff = reinterpret_cast<uint16_t *>(pkt + L.roff);
    double *w = reinterpret_cast<double *>(pkt + L.w);
    uint32_t *code = reinterpret_cast<uint32_t *>(pkt + L.code);
    uint32_t *far = reinterpret_cast<uint32_t *>(pkt + L.far);
    uint32_t *cell = reinterpret_cast<uint32_t *>(pkt + L.cell);
    uint32_t *info = reinterpret_cast<uint32_t *>(pkt + L.info);
    const uint32_t tid = threadIdx.x;
    const bool epilogue = level >= a.n_real_pl;
    const uint32_t dir = epilogue ? level - a.n_real_pl : level / a.n_levels;
    const uint32_t dir_base = a.pl_off[dir * a.n_levels];
    if (tid < 64) s_cnt[tid] = 0;
    if (tid == 0) s_far = 0;
    __syncthreads();
    uint32_t c = 0, dl = 0, deg = 0, n_per = 0;
    double dx = 0, dy = 0, dz = 0;
    if (tid < n) {
        const uint32_t k = a.keys[slot0 + tid];
        c = k / a.n_dl;
        dl = k - c * a.n_dl;
        dx = a.g.dirs[3 * dl]; dy = a.g.dirs[3 * dl + 1]; dz = a.g.dirs[3 * dl + 2];
        deg = (uint32_t)(a.upoff[slot0 + tid + 1] - a.upoff[slot0 + tid]);
        if (deg > (uint32_t)kWalkMaxDeg) {
            atomicAdd(a.lag_counter + 1, 1u);
            deg = kWalkMaxDeg;
        }
        // the round layout needs the slots of a tile in descending order of degree (the sort key provides it)
        if (tid > 0 && (uint32_t)(a.upoff[slot0 + tid] - a.upoff[slot0 + tid - 1]) < deg) atomicAdd(a.lag_counter + 1, 1u);
        if (!COUNT)
            for (uint32_t j = 0; j < deg; ++j) atomicAdd(&s_cnt[j], 1u);
    }
    __syncthreads();
    if (!COUNT) {
        if (tid == 0) {
            uint32_t o = 0;
            for (int j = 0; j < 64; ++j) { s_roff[j] = o; o += s_cnt[j]; }
            s_roff[64] = o;
        }
        __syncthreads();
        if (tid < 64) roff[tid] = (uint16_t)s_roff[tid];
        // zero the padding so the stream is fully initialised
        if (tid < 4) {
            if (tid == 0 && (n_entries & 1u)) w[n_entries] = 0.0;
            const uint32_t pad_e = (walk_align16(4u * n_entries) - 4u * n_entries) / 4u;
            if (tid < pad_e) code[n_entries + tid] = kWalkNear;
            const uint32_t pad_f = (walk_align16(4u * off_far_total) - 4u * off_far_total) / 4u;
            if (tid < pad_f) far[off_far_total + tid] = 0u;
            const uint32_t pad_n = (walk_align16(4u * n) - 4u * n) / 4u;
            if (tid < pad_n) { cell[n + tid] = 0xffffffffu; info[n + tid] = 0u; }
        }
    }
    // entries: Local faces in face order first, then the periodic ones
    if (tid < n) {
        const uint32_t s = slot0 + tid;
        if (!COUNT) cell[tid] = epilogue ? (uint32_t)a.pidx[c] : c;   // epilogue tiles address acc_per by periodic row
        uint32_t j = 0, lvl_end = 0;
        for (int pass = epilogue ? 1 : 0; pass < 2; ++pass) {
            for (uint32_t f = a.g.face_off[c]; f < a.g.face_off[c + 1] && j < deg; ++f) {
                if (a.g.face_kind[f] != (pass ? 2 : 0)) continue;
                const double dd = dot_dir(ld_geo(a.g.face_geo + f), dx, dy, dz);
                if (!(dd < 0.0)) continue;
                uint32_t src = a.slot_of[(size_t)dl * a.g.n_cells + (uint32_t)a.g.face_nb[f]];
                bool snapshot = false;
                if (pass) {
                    ++n_per;
                    if (!epilogue) {
                        if (lvl_end == 0) lvl_end = level_end_of_slot(a.pl_off, a.n_pl, s);
                        snapshot = src < lvl_end;   // donor not in a later level: read its pre-sweep snapshot
                    }
                }
                const bool near = !epilogue && !snapshot && walk_is_near(src, slot0, n, a.n_tasks, a.window_usable);
                if (COUNT) {
                    if (!near) atomicAdd(&s_far, 1u);
                } else {
                    const double tt = a.ttot_slot[src];
                    const double share = tt > 0.0 ? (a.g.face_rev[f] * (-dd)) / tt : 0.0;
                    if (snapshot) {
                        const unsigned int q = atomicAdd(a.lag_counter, 1u);
                        a.lag_src[q] = src;
                        src = a.n_tasks + q;
                    }
                    const uint32_t pos = s_roff[j] + tid;
                    if (near) {
                        code[pos] = kWalkNear | ((src - dir_base) % a.window);
                        atomicAdd(a.lag_counter + 2, 1u);
                    } else {
                        const uint32_t i = atomicAdd(&s_far, 1u);
                        if (i < off_far_total) far[i] = src;
                        code[pos] = i;
                        atomicAdd(a.lag_counter + 3, 1u);
                    }
                    w[pos] = share;
                }
                ++j;
            }
        }
        if (!COUNT) info[tid] = deg | (n_per << 8);
    }
    __syncthreads();
    if (tid == 0) {
        if (COUNT) {
            a.tile_far[blockIdx.x] = s_far;
        } else {
            if (s_far != off_far_total) atomicAdd(a.lag_counter + 1, 1u);   // the two passes must agree
            WalkHeader h;
            h.slot0 = slot0; h.n = (uint16_t)n; h.n_entries = (uint16_t)n_entries;
            h.local0 = epilogue ? 0u : (slot0 - dir_base) % a.window;   // window position of the tile's first slot
            h.n_far = (uint16_t)off_far_total;
            h.epilogue = epilogue ? 1 : 0;
            h.group = dir;
            h.need = need;
            h.pad[0] = h.pad[1] = 0;
            *reinterpret_cast<WalkHeader *>(pkt) = h;
        }
    }
}

// per sweep, before the walk: absorption record of every slot in slot order, {exp(-n_HI sigma size), source / D}
__global__ void __launch_bounds__(256)
w_rec_kernel(const uint32_t *__restrict__ cell_of_slot, const double *__restrict__ att, const double *__restrict__ src,
             double n_dirs_total, uint32_t n, double2 *__restrict__ rec_slot) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint32_t c = cell_of_slot[s];
    __stcs(rec_slot + s, make_double2(__ldg(att + c), __ldg(src + c) / n_dirs_total));
}

// the walk kernel's arguments
struct WalkArgs {
    const unsigned char *stream;
    const uint64_t *stream_off;
    const TileDesc *tab;
    const uint32_t *tab_off;
    double *out_slot;
    const double2 *rec_slot;     // {exp(-n_HI sigma size), source / D} per slot
    double *acc_cell;            // Dl x N: incoming_total_rate per (direction, cell)
    double *acc_per;             // Dl x n_periodic
    double threshold;
    uint32_t stages, stage_bytes, window;
    uint32_t n_periodic, n_cells;
    uint32_t l2_ahead;           // prefetch the packets of the next 32 tiles into L2
    unsigned long long *prof;    // optional per-block cycles
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_cta_shared(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_cta_shared_add(uint32_t *p, uint32_t v) {
    asm volatile("red.release.cta.shared::cta.add.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
// The waits of the walk kernel are bounded: a broken schedule traps (the launch fails with an error) instead of
// hanging the device.
__device__ __forceinline__ void walk_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok, spins = 0;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (!ok && ++spins > (1u << 24)) __trap();
    } while (!ok);
}
__device__ __forceinline__ void walk_wait_done(const uint32_t *done, uint32_t need) {
    uint32_t spins = 0;
    while (ld_acquire_cta_shared(done) < need)
        if (++spins > (1u << 28)) __trap();
}
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_8(uint32_t dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
// arrive on the mbarrier once all cp.async of this thread so far have landed (the pending count is not raised)
__device__ __forceinline__ void cp_async_mbar_arrive(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

// Block b walks local direction b: tiles tab[tab_off[b] .. tab_off[b+1]) in order.
//   compute groups (NG x NTG threads): group g solves tiles g, g + NG, ...; thread t of the group = slot t of the tile.
//     Tiles of one wavefront level are independent and run side by side; the first tile of a level waits until every
//     earlier tile is complete (header.need against the block's `done` counter).
//   producer warp: one lane keeps the ring full (static packet + the tile's absorption records, TMA bulk copies).
//   gather warps: fetch a tile's sources outside the window into its stage with 8-byte cp.async copies.
template <int NG, int NTG, bool PROFILE>
__global__ void __launch_bounds__(NG * NTG + 32 + 32 * kWalkGatherWarps, 1)
walk_kernel(WalkArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int NT = NG * NTG;
    const uint32_t stages = a.stages, stage_bytes = a.stage_bytes, W = a.window;
    unsigned char *const ring = smem;
    double *const win = reinterpret_cast<double *>(smem + (size_t)stages * stage_bytes);
    uint64_t *const full = reinterpret_cast<uint64_t *>(smem + (size_t)stages * stage_bytes + (size_t)W * 8u);
    uint64_t *const empty = full + kWalkMaxStages;
    uint64_t *const gathered = empty + kWalkMaxStages;
    uint32_t *const done = reinterpret_cast<uint32_t *>(gathered + kWalkMaxStages);
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const uint32_t n_my = a.tab_off[blockIdx.x + 1] - a.tab_off[blockIdx.x];
    const unsigned char *const stream = a.stream + a.stream_off[blockIdx.x];
    if (tid == 0) {
        for (uint32_t s = 0; s < stages; ++s) {
            mbar_init(smem_u32(full + s), 1);
            mbar_init(smem_u32(empty + s), 1);
            mbar_init(smem_u32(gathered + s), 32);
        }
        *done = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == NT / 32) {
        // ---- producer ----
        const uint64_t policy = policy_evict_first();
        const TileDesc *tab = a.tab + a.tab_off[blockIdx.x];
        uint32_t stage = 0, use = 0;
        for (uint32_t k0 = 0; k0 < n_my; k0 += 32) {
            // the lanes fetch 32 tile descriptors at once; lane 0 issues the copies
            TileDesc mine;
            mine.slot0 = mine.off16 = mine.level = mine.aux = mine.need = 0; mine.n = mine.n_entries = 0;
            if (k0 + lane < n_my) mine = tab[k0 + lane];
            const uint32_t cnt = min(32u, n_my - k0);
            // pull the packets of the batch after this one into L2: the copies into shared memory then find them there
            if (a.l2_ahead && k0 + 32u + lane < n_my) {
                const TileDesc nd = tab[k0 + 32u + lane];
                const WalkLayout NL = walk_layout(nd.n, nd.n_entries, nd.aux & 0xffffu);
                bulk_prefetch_l2(stream + (size_t)nd.off16 * 16u, NL.bytes);
                if ((nd.aux >> 16) == 0) bulk_prefetch_l2(a.rec_slot + nd.slot0, 16u * nd.n);
            }
            for (uint32_t i = 0; i < cnt; ++i) {
                const uint32_t slot0 = __shfl_sync(0xffffffffu, mine.slot0, i), off16 = __shfl_sync(0xffffffffu, mine.off16, i);
                const uint32_t nn = __shfl_sync(0xffffffffu, (uint32_t)mine.n | ((uint32_t)mine.n_entries << 16), i);
                const uint32_t aux = __shfl_sync(0xffffffffu, mine.aux, i);
                if (lane == 0) {
                    if (use > 0) walk_mbar_wait(smem_u32(empty + stage), (use - 1u) & 1u);
                    const uint32_t n = nn & 0xffffu;
                    const WalkLayout L = walk_layout(n, nn >> 16, aux & 0xffffu);
                    const bool real = (aux >> 16) == 0;
                    const uint32_t dst = smem_u32(ring + (size_t)stage * stage_bytes);
                    mbar_expect_tx(smem_u32(full + stage), L.bytes + (real ? 16u * n : 0u));
                    tma_bulk_load(dst, stream + (size_t)off16 * 16u, L.bytes, smem_u32(full + stage), policy);
                    if (real) tma_bulk_load(dst + L.rec, a.rec_slot + slot0, 16u * n, smem_u32(full + stage), policy);
                }
                if (++stage == stages) { stage = 0; ++use; }
            }
        }
        return;
    }
    if (warp > NT / 32) {
        // ---- gather warps: sources outside the window -> the stage, asynchronously ----
        const uint32_t gw = warp - NT / 32 - 1;
        for (uint32_t k = gw; k < n_my; k += kWalkGatherWarps) {
            const uint32_t stage = k % stages, parity = (k / stages) & 1u;
            walk_mbar_wait(smem_u32(full + stage), parity);
            unsigned char *const pkt = ring + (size_t)stage * stage_bytes;
            const WalkHeader h = *reinterpret_cast<const WalkHeader *>(pkt);
            if (h.epilogue) {   // the periodic terms read this sweep's rates: every real tile must be complete
                walk_wait_done(done, h.need);
            }
            const WalkLayout L = walk_layout(h.n, h.n_entries, h.n_far);
            const uint32_t *const far = reinterpret_cast<const uint32_t *>(pkt + L.far);
            const uint32_t val = smem_u32(pkt + L.val);
            for (uint32_t i = lane; i < h.n_far; i += 32u) cp_async_8(val + 8u * i, a.out_slot + far[i]);
            cp_async_mbar_arrive(smem_u32(gathered + stage));
        }
        return;
    }
    // ---- compute groups ----
    const uint32_t grp = tid / NTG, t = tid - grp * NTG;
    const double threshold = a.threshold;
    long long t_begin = 0, t_pkt = 0, t_gat = 0, t_dep = 0, t_loop = 0, t_out = 0, t_bar = 0, tp = 0;
    if (PROFILE && tid == 0) t_begin = clock64();
    for (uint32_t k = grp; k < n_my; k += NG) {
        const uint32_t stage = k % stages, parity = (k / stages) & 1u;
        if (PROFILE && tid == 0) tp = clock64();
        walk_mbar_wait(smem_u32(full + stage), parity);
        if (PROFILE && tid == 0) { const long long now = clock64(); t_pkt += now - tp; tp = now; }
        const unsigned char *const pkt = ring + (size_t)stage * stage_bytes;
        const WalkHeader h = *reinterpret_cast<const WalkHeader *>(pkt);
        if (NG > 1 || h.epilogue) {
            walk_wait_done(done, h.need);
        }
        if (PROFILE && tid == 0) { const long long now = clock64(); t_dep += now - tp; tp = now; }
        walk_mbar_wait(smem_u32(gathered + stage), parity);
        if (PROFILE && tid == 0) { const long long now = clock64(); t_gat += now - tp; tp = now; }
        const uint32_t n = h.n;
        const WalkLayout L = walk_layout(n, h.n_entries, h.n_far);
        const uint16_t *const roff = reinterpret_cast<const uint16_t *>(pkt + L.roff);
        const double *const ws = reinterpret_cast<const double *>(pkt + L.w);
        const uint32_t *const code = reinterpret_cast<const uint32_t *>(pkt + L.code);
        const double *const val = reinterpret_cast<const double *>(pkt + L.val);
        if (t < n) {
            const uint32_t inf = reinterpret_cast<const uint32_t *>(pkt + L.info)[t];
            const uint32_t deg = inf & 0xffu, n_loc = deg - ((inf >> 8) & 0xffu);
            double in_loc = 0.0, in_per = 0.0;
            for (uint32_t j0 = 0; j0 < deg; j0 += 4u) {
                uint32_t cd[4];
                double wv[4], v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const bool ok = j0 + u < deg;
                    const uint32_t pos = ok ? (uint32_t)roff[j0 + u] + t : 0u;
                    cd[u] = ok ? code[pos] : kWalkNear;
                    wv[u] = ok ? ws[pos] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const double *const p = (cd[u] & kWalkNear) ? win + (cd[u] & ~kWalkNear) : val + cd[u];
                    v[u] = *p;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const double p = v[u] * wv[u];
                    if (j0 + u < n_loc) in_loc += p;
                    else if (j0 + u < deg) in_per += p;
                }
            }
            if (PROFILE && tid == 0) { const long long now = clock64(); t_loop += now - tp; tp = now; }
            if (h.epilogue) {
                const uint32_t row = reinterpret_cast<const uint32_t *>(pkt + L.cell)[t];
                __stcs(a.acc_per + (size_t)h.group * a.n_periodic + row, in_per);   // periodic_source of this sweep
            } else {
                const double2 rec = reinterpret_cast<const double2 *>(pkt + L.rec)[t];
                const double total = (in_loc + rec.y) + in_per;                   // site.rs:49-56
                // HydrogenOnly::get_outgoing_rate, hydrogen_only/mod.rs:81-87
                const double out = (total < threshold) ? 0.0 : total * rec.x;
                const uint32_t wpos = h.local0 + t;
                win[wpos >= W ? wpos - W : wpos] = out;
                a.out_slot[h.slot0 + t] = out;
                // incoming_total_rate[d] of the cell: the one term (direction, cell), folded by s_rate_finish_kernel
                __stcs(a.acc_cell + (size_t)h.group * a.n_cells + reinterpret_cast<const uint32_t *>(pkt + L.cell)[t], in_loc);
            }
        }
        if (PROFILE && tid == 0) { const long long now = clock64(); t_out += now - tp; tp = now; }
        named_bar_sync(1 + grp, NTG);   // every read of the stage is done; the window and out_slot hold this tile's rates
        if (PROFILE && tid == 0) { const long long now = clock64(); t_bar += now - tp; tp = now; }
        if (t == 0) {
            red_release_cta_shared_add(done, 1u);
            mbar_arrive(smem_u32(empty + stage));
        }
    }
    if (PROFILE && tid == 0) {
        a.prof[8 * blockIdx.x + 0] = (unsigned long long)(clock64() - t_begin);
        a.prof[8 * blockIdx.x + 1] = (unsigned long long)t_pkt;
        a.prof[8 * blockIdx.x + 2] = n_my;
        a.prof[8 * blockIdx.x + 3] = (unsigned long long)t_loop;
        a.prof[8 * blockIdx.x + 4] = (unsigned long long)t_out;
        a.prof[8 * blockIdx.x + 5] = (unsigned long long)t_bar;
        a.prof[8 * blockIdx.x + 6] = (unsigned long long)t_gat;
        a.prof[8 * blockIdx.x + 7] = (unsigned long long)t_dep;
    }
}

typedef void (*WalkKernel)(WalkArgs);
// (compute groups, threads per group) variants; the block has 32 * (1 + kWalkGatherWarps) more threads
inline WalkKernel walk_kernel_for(uint32_t groups, uint32_t threads, bool profile) {
#define SSW_WALK_CASE(G, T) if (groups == G && threads == T) return profile ? walk_kernel<G, T, true> : walk_kernel<G, T, false>;
    SSW_WALK_CASE(1, 128) SSW_WALK_CASE(2, 128) SSW_WALK_CASE(4, 128) SSW_WALK_CASE(6, 128)
    SSW_WALK_CASE(1, 256) SSW_WALK_CASE(2, 256) SSW_WALK_CASE(3, 256)
    SSW_WALK_CASE(1, 512)
    SSW_WALK_CASE(8, 64) SSW_WALK_CASE(4, 64)
#undef SSW_WALK_CASE
    return nullptr;
}
inline size_t walk_smem_bytes(uint32_t stages, uint32_t stage_bytes, uint32_t window) {
    return (size_t)stages * stage_bytes + (size_t)window * 8u + sizeof(uint64_t) * 3 * kWalkMaxStages + 16;
}

}  // namespace ssw
