// sweep.cu -- host side of libsubsweep_b200.so: the B200 mirror of the reference's
// Sweep<HydrogenOnly> (src/sweep/mod.rs:172-610) and the C ABI of include/subsweep_b200.h.
//
// There is no CPU fallback in this file: every path launches the kernels of kernels.cuh /
// patch.cuh / stream.cuh on the CUDA device and fails with SSW_E_CUDA when that is impossible.
#include <cub/cub.cuh>
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <thread>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/subsweep_b200.h"
#include "kernels.cuh"
#include "stream.cuh"
#include "patch.cuh"
#include "peer.cuh"
#include "small.cuh"

namespace ssw {

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

[[noreturn]] static void fail(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    throw Error(code, buf);
}

#define CUDA_CHECK(expr)                                                                       \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess)                                                                \
            ::ssw::fail(e__ == cudaErrorMemoryAllocation ? SSW_E_NOMEM : SSW_E_CUDA,           \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    bool owned = true;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p && owned) cudaFree(p);
        p = nullptr;
        n = 0;
        owned = true;
    }
    // a window into memory owned by somebody else (the per-cell arrays live in the rank's peer arena)
    void view(void *ptr, size_t count) {
        release();
        p = static_cast<T *>(ptr);
        n = count;
        owned = false;
    }
    void alloc(size_t count) {
        release();
        if (count == 0) count = 1;
        CUDA_CHECK(cudaMalloc(&p, count * sizeof(T)));
        n = count;
    }
    void ensure(size_t count) {
        if (count > n) alloc(count + count / 8);
    }
    void upload(const T *src, size_t count, cudaStream_t s) {
        CUDA_CHECK(cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, s));
    }
    void zero(cudaStream_t s) { CUDA_CHECK(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
};

static inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }
static inline uint32_t stream_env_u32(const char *name, uint32_t fallback) { return env_u32(name, fallback); }

// ------------------------------------------------------------------------------------------
// wavefront level sets of one active set
// ------------------------------------------------------------------------------------------
struct Schedule {
    bool valid = false;
    uint64_t version = 0;       // levels_version it was built for
    uint32_t n_act = 0;
    uint64_t n_tasks = 0;
    uint32_t n_levels = 0;
    uint32_t max_level_tasks = 0;
    DevBuf<uint32_t> act_list;  // ascending cell indices (empty = all cells)
    DevBuf<uint32_t> touch_list; // Local neighbours of the active cells (photon_rate bookkeeping)
    uint32_t n_touch = 0;
    DevBuf<uint32_t> tasks;     // level-sorted, (dl, cell)-sorted inside a level
    DevBuf<uint32_t> level_off; // n_levels + 1
    std::vector<uint32_t> level_off_host;
    Compiled compiled;          // slot-ordered form (stream.cuh), all-cells sweep only
    // task records of a cached partial schedule (kernels.cuh, MiniView)
    bool mini_valid = false;
    bool mini_slot_state = false;   // built against the slot-ordered flux state
    DevBuf<uint32_t> m_off, m_src, m_self;
    DevBuf<double> m_w, m_ttot;
    MiniView mini_view() const { return MiniView{m_off.p, m_src.p, m_w.p, m_self.p, m_ttot.p}; }
};

// SSW_TIME_SCHED=1: wall-clock time of the pieces of a schedule rebuild (synchronises the stream around each piece)
struct SchedProbe {
    cudaStream_t st;
    const char *what;
    bool on;
    std::chrono::steady_clock::time_point t0;
    SchedProbe(cudaStream_t s, const char *w) : st(s), what(w), on(std::getenv("SSW_TIME_SCHED") != nullptr) {
        if (on) { cudaStreamSynchronize(st); t0 = std::chrono::steady_clock::now(); }
    }
    ~SchedProbe() {
        if (on) {
            cudaStreamSynchronize(st);
            fprintf(stderr, "[sched] %-22s %.3f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
        }
    }
};

enum TimerCat { T_SWEEP = 0, T_CHEM, T_LEVELS, T_SCHED, T_ALLREDUCE, T_KERNEL, T_STEP, T_COUNT };

struct Sweep {
    ssw_params P{};
    std::vector<double> dirs_all;  // D x 3
    int D = 0, Dl = 0, d0 = 0;
    uint32_t N = 0;
    uint64_t F = 0;
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;

    // grid
    DevBuf<double4> face_geo;
    DevBuf<double> dirs_dev;       // this rank's directions (3 per direction, local order)
    DevBuf<double> face_rev;
    DevBuf<int32_t> face_nb;
    DevBuf<uint8_t> face_kind;
    DevBuf<uint32_t> face_off;
    DevBuf<double> size, volume;
    // cells
    DevBuf<double> rho, x, T, ts, tau, prev_rate, src, att, ion_time;
    DevBuf<uint8_t> level;
    DevBuf<int32_t> pidx;
    DevBuf<uint32_t> pcells;
    uint32_t n_periodic = 0;
    // spatial patches (ssw_set_cell_positions): cell -> patch map of the patch-ordered all-cells sweep (patch.cuh)
    DevBuf<uint32_t> patch_of, patch_off, patch_cells;
    DevBuf<uint16_t> patch_lidx;
    uint32_t n_patches = 0, patch_max_cells = 0, patch_dims[3] = {1, 1, 1};
    bool have_patches = false;
    std::string patch_note;   // why the patch-ordered form is not in use (empty: it is, or was never tried)
    // per (dir, cell)
    DevBuf<double> q, incoming;
    DevBuf<int32_t> missing;
    DevBuf<double> per_lag, per_new;
    // scratch
    DevBuf<uint32_t> queue_scratch;
    DevBuf<uint32_t> level_off_scratch;
    DevBuf<QueueCtl> ctl;
    DevBuf<uint8_t> flags;
    DevBuf<uint8_t> cub_temp;
    DevBuf<uint32_t> n_selected;
    DevBuf<double> rate_act, cell_tmp, cell_tmp2, photon;
    bool photon_valid = false;  // photon = sum_d incoming[d] of this rank's directions is current
    DevBuf<double2> cellrec;
    DevBuf<unsigned long long> hist;
    DevBuf<ChemStats> chem_stats;
    DevBuf<uint16_t> last_attempts;          // substep attempts of every cell's last chemistry update
    DevBuf<uint32_t> order_keys, order_keys_out, order_vals, order_vals_out;
    bool chem_balance = false;               // sort the cells of a launch by their last substep count
    uint64_t chem_prev_attempts = 0, chem_prev_cells = 0;
    DevBuf<int32_t> wlevel;

    // TimestepState (src/sweep/timestep_state.rs:4-9)
    int lowest_allowed = 0;
    bool first_done = false;
    double sim_time = 0.0;
    std::vector<uint64_t> bin_count;  // cells per level
    uint64_t levels_version = 1;

    std::vector<std::unique_ptr<Schedule>> sched;  // [0..L-1] partial sets by current level, [L] all cells
    Compiled *state = nullptr;  // non-null once the flux state lives in slot order (patch.cuh / stream.cuh)

    ssw_allreduce_fn allreduce = nullptr;
    void *allreduce_ctx = nullptr;
    ssw_collective_fn collective = nullptr;   // optional: reduce-scatter / all-gather -> cell-sliced chemistry
    void *collective_ctx = nullptr;
    DevBuf<double> chem_pack;                 // world_size x kPackFields x cells_per_rank

    // direction sharding over peer-mapped memory (peer.cuh): the arena holds the per-cell state and the receive buffers
    DevBuf<unsigned char> arena;
    PeerTable pt{};
    bool peers = false;            // the arenas of all ranks are attached: no hooks, no collective library
    uint32_t peer_epoch = 0;       // synchronisation points so far; every rank's counter reaches peer_epoch * world_size
    DevBuf<unsigned int> blocks_done;
    std::vector<void *> ipc_opened;
    typedef int (*StreamWaitValue32)(cudaStream_t, unsigned long long, unsigned int, unsigned int);
    StreamWaitValue32 stream_wait_value32 = nullptr;
    int peer_wait_mode = 0;        // 0 stream memory operation, 1 host polling (ranks sharing a device), 2 polling kernel
    cudaStream_t aux_stream = nullptr;
    unsigned char *flags_host = nullptr;   // pinned
    DevBuf<double> series_partial, series_sums, series_mass;   // ssw_time_series_compute

    // statistics / timers
    uint64_t stat[16] = {0};
    struct Pending { cudaEvent_t a, b; int cat; int lvl; };
    std::vector<Pending> pending;
    std::vector<cudaEvent_t> event_pool;
    ssw_timings timings{};
    int coop_blocks_build = 0, coop_blocks_replay = 0, coop_blocks_mini = 0;

    ~Sweep() {
        for (void *m : ipc_opened) cudaIpcCloseMemHandle(m);
        for (auto &p : pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
        for (auto e : event_pool) cudaEventDestroy(e);
        if (stream) cudaStreamDestroy(stream);
        if (aux_stream) cudaStreamDestroy(aux_stream);
        if (flags_host) cudaFreeHost(flags_host);
    }

    // ---- helpers ----
    void bind() { CUDA_CHECK(cudaSetDevice(device)); }
    void launched(uint64_t n = 1) { stat[SSW_STAT_KERNEL_LAUNCHES] += n; }
    cudaEvent_t get_event() {
        if (!event_pool.empty()) {
            cudaEvent_t e = event_pool.back();
            event_pool.pop_back();
            return e;
        }
        cudaEvent_t e;
        CUDA_CHECK(cudaEventCreate(&e));
        return e;
    }
    // Timing level: 0 = the step and the sweep kernels of the largest active sets only (what a production host and the
    // roofline of bench.py need: four event records per step), 1 = every phase of every single sweep (the breakdown
    // ssw_get_timings reports; ~10 event records per single sweep, which cost a direction shard with its microsecond
    // sub-level sweeps a measurable share of the step).  ssw_set_timing_level / SSW_TIMERS.
    int timing_level = 1;
    static constexpr size_t kNoTimer = ~(size_t)0;
    size_t tic(int cat, int lvl = -1) {
        if (timing_level == 0 && !(cat == T_STEP || (cat == T_KERNEL && lvl == lowest_allowed))) return kNoTimer;
        Pending p{get_event(), get_event(), cat, lvl};
        CUDA_CHECK(cudaEventRecord(p.a, stream));
        pending.push_back(p);
        return pending.size() - 1;
    }
    void toc(size_t id) {
        if (id == kNoTimer) return;
        CUDA_CHECK(cudaEventRecord(pending[id].b, stream));
    }
    void resolve_timers() {
        CUDA_CHECK(cudaStreamSynchronize(stream));
        for (auto &p : pending) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
                switch (p.cat) {
                case T_SWEEP:
                    timings.sweep_ms += ms;
                    if (p.lvl >= 0 && p.lvl < 32) timings.sweep_level_ms[p.lvl] += ms;
                    break;
                case T_CHEM: timings.chemistry_ms += ms; break;
                case T_LEVELS: timings.update_levels_ms += ms; break;
                case T_SCHED: timings.schedule_ms += ms; break;
                case T_ALLREDUCE: timings.allreduce_ms += ms; break;
                case T_KERNEL:
                    timings.sweep_kernel_ms += ms;
                    if (p.lvl >= 0 && p.lvl < 32) timings.kernel_level_ms[p.lvl] += ms;
                    break;
                case T_STEP: timings.step_ms += ms; timings.steps += 1; break;
                }
            }
            event_pool.push_back(p.a);
            event_pool.push_back(p.b);
        }
        pending.clear();
    }

    GridView grid_view() const {
        GridView g;
        g.face_geo = face_geo.p;
        g.face_rev = face_rev.p;
        g.face_nb = face_nb.p;
        g.face_kind = face_kind.p;
        g.face_off = face_off.p;
        g.dirs = dirs_dev.p;
        g.n_cells = N;
        return g;
    }
    CellView cell_view() const {
        CellView c;
        c.rho = rho.p; c.x = x.p; c.T = T.p; c.ts = ts.p; c.tau = tau.p;
        c.prev_rate = prev_rate.p; c.src = src.p; c.att = att.p; c.ion_time = ion_time.p;
        c.size = size.p; c.volume = volume.p; c.level = level.p; c.pidx = pidx.p;
        return c;
    }
    StateView state_view() const {
        StateView st;
        st.q = q.p;
        st.slot_of = state ? state->slot_of : nullptr;
        st.out_slot = state ? state->out_slot : nullptr;
        st.ttot_slot = state ? state->ttot_slot : nullptr;
        st.n_cells = N;
        return st;
    }
    SweepArgs sweep_args(int cur) const {
        SweepArgs a;
        a.g = grid_view();
        a.att = att.p;
        a.src = src.p;
        a.pidx = pidx.p;
        a.level = level.p;
        a.st = state_view();
        a.solve = 1;
        a.incoming = incoming.p;
        a.per_lag = per_lag.p;
        a.missing = missing.p;
        a.n_periodic = n_periodic;
        a.inv_threshold_unused = 0.0;
        a.threshold = P.significant_rate_threshold_per_s;
        a.n_dirs_total = (double)D;
        a.cur = cur;
        return a;
    }

    uint64_t count_at_least(int l) const {
        uint64_t n = 0;
        for (int k = l; k < P.n_levels; ++k) n += bin_count[k];
        return n;
    }

    void create(const ssw_params *p, const ssw_grid *g, const double *density, const double *xhii,
                const double *temperature, const double *source);
    void set_positions(const double *xyz);
    void set_directions(const double *dirs_xyz);
    bool rotating = false;   // the directions change between steps (ssw_set_directions): nothing direction-dependent is compiled
    PatchGrid patch_view() const {
        PatchGrid pg;
        pg.patch_of = patch_of.p; pg.lidx = patch_lidx.p; pg.patch_off = patch_off.p; pg.patch_cells = patch_cells.p;
        pg.n_patches = n_patches; pg.max_cells = patch_max_cells;
        for (int k = 0; k < 3; ++k) pg.dims[k] = patch_dims[k];
        return pg;
    }
    void refresh_histogram();
    void build_active_list(Schedule &S, int cur);
    void gather_periodic(double *dst, const uint32_t *act, uint32_t n_act);
    void build_schedule(Schedule &S, int cur, bool solve, int dl_base, int n_dl, int32_t *wl);
    void build_mini(Schedule &S);
    void single_sweep(int cur);
    void update_timestep_levels();
    double run_sweeps();
    void read_field(int field, double *out, bool wait = true);
    void all_rates(double *dev_out);
    void maybe_allreduce(double *buf, uint64_t n);
    uint32_t cells_per_rank() const { return (uint32_t)(((uint64_t)N + P.world_size - 1) / P.world_size); }
    void run_collective(int op, double *buf, uint64_t n_per_rank);
    // peer-mapped exchange (peer.cuh)
    void launch_chemistry(const uint32_t *act, uint32_t n, const double *rate, uint32_t first_cell, const ChemParams &cp,
                          const PeerChem &pc);
    void peer_attach(void *const *bases);
    void peer_sync_point();
    void peer_wait();
    void peer_allreduce(double *buf, uint64_t n);
    void peer_pull(uint64_t off);
    PeerChem peer_chem() const;
    PeerLevels peer_levels() const;
    uint32_t own_first() const { return peers ? pt.first() : 0; }
    uint32_t own_count() const { return peers ? pt.n_own() : N; }
};

static int coop_grid(const void *kernel, int threads, int num_sms) {
    int per_sm = 0;
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0));
    if (per_sm < 1) fail(SSW_E_CUDA, "cooperative kernel does not fit on an SM");
    return per_sm * num_sms;
}

// ------------------------------------------------------------------------------------------
// Sweep::new (src/sweep/mod.rs:194-232) + init_sweep_system (:634-692)
// ------------------------------------------------------------------------------------------
void Sweep::create(const ssw_params *p, const ssw_grid *g, const double *density,
                   const double *xhii, const double *temperature, const double *source) {
    if (!p || !g || !density || !xhii || !temperature || !source) fail(SSW_E_INVALID, "null argument");
    if (p->n_dirs < 1 || p->n_dirs > kMaxDirs || !p->dirs_xyz)
        fail(SSW_E_INVALID, "n_dirs must be in [1, %d]", kMaxDirs);
    if (p->n_levels < 1 || p->n_levels > 31)   // timestep_state.rs:72 asserts < 32
        fail(SSW_E_INVALID, "num_timestep_levels must be in [1, 31]");
    if (g->n_cells < 1 || g->n_cells > 0xfffffff0ull) fail(SSW_E_INVALID, "bad n_cells");
    if (p->world_size < 0 || p->world_size > kMaxPeers || (p->world_size > 1 && (p->rank < 0 || p->rank >= p->world_size)))
        fail(SSW_E_INVALID, "bad rank / world_size (at most %d ranks)", kMaxPeers);
    P = *p;
    if (P.world_size < 1) { P.world_size = 1; P.rank = 0; }
    D = p->n_dirs;
    dirs_all.assign(p->dirs_xyz, p->dirs_xyz + 3 * (size_t)D);
    P.dirs_xyz = nullptr;
    int32_t b = 0, e = D;
    ssw_direction_shard(D, P.world_size, P.rank, &b, &e);
    d0 = b;
    Dl = e - b;
    if (Dl < 1) fail(SSW_E_INVALID, "rank %d of %d has no direction (D = %d)", P.rank, P.world_size, D);
    N = (uint32_t)g->n_cells;
    F = g->face_offsets[N];
    if (F > 0xfffffff0ull) fail(SSW_E_INVALID, "too many faces");
    if ((uint64_t)N * (uint64_t)Dl > 0xfffffff0ull)
        fail(SSW_E_INVALID, "n_cells * local directions must be < 2^32");

    // ---- device ----
    device = p->device_id;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev < 1)
        fail(SSW_E_CUDA, "no CUDA device: this library has no CPU fallback");
    if (device < 0 || device >= n_dev) fail(SSW_E_CUDA, "device %d out of range (%d devices)", device, n_dev);
    CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) fail(SSW_E_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    if (!prop.cooperativeLaunch) fail(SSW_E_CUDA, "device lacks cooperative launch");
    num_sms = prop.multiProcessorCount;
    CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));

    // ---- validate the grid, find the reverse face of every face ----
    std::vector<double4> geo(F);
    std::vector<double> rev(F, 0.0);
    std::vector<uint32_t> off(N + 1);
    std::vector<int32_t> pidx_h(N, -1);
    std::vector<uint32_t> pcells_h;
    for (uint32_t c = 0; c <= N; ++c) {
        if (c > 0 && g->face_offsets[c] < g->face_offsets[c - 1]) fail(SSW_E_INVALID, "face_offsets not monotone at cell %u", c);
        off[c] = (uint32_t)g->face_offsets[c];
    }
    if (off[0] != 0) fail(SSW_E_INVALID, "face_offsets[0] must be 0");
    for (uint32_t c = 0; c < N; ++c) {
        bool has_periodic = false;
        for (uint32_t f = off[c]; f < off[c + 1]; ++f) {
            const double *n = g->face_normal + 3 * (size_t)f;
            geo[f] = make_double4(n[0], n[1], n[2], g->face_area[f]);
            const uint8_t kind = g->face_kind[f];
            const int32_t nb = g->face_neighbour[f];
            if (kind == SSW_FACE_BOUNDARY) continue;
            if (kind != SSW_FACE_LOCAL && kind != SSW_FACE_LOCAL_PERIODIC) fail(SSW_E_INVALID, "face %u: bad kind %d", f, (int)kind);
            if (nb < 0 || (uint32_t)nb >= N) fail(SSW_E_INVALID, "face %u: neighbour %d out of range", f, nb);
            if (kind == SSW_FACE_LOCAL_PERIODIC) has_periodic = true;
            // reverse face: the neighbour's face back to c, same kind, most anti-parallel normal
            double best = 2.0;
            int64_t bestg = -1;
            for (uint32_t gg = off[nb]; gg < off[nb + 1]; ++gg) {
                if (g->face_neighbour[gg] != (int32_t)c || g->face_kind[gg] != kind) continue;
                const double *m = g->face_normal + 3 * (size_t)gg;
                const double dd = n[0] * m[0] + n[1] * m[1] + n[2] * m[2];
                if (dd < best) { best = dd; bestg = gg; }
            }
            if (bestg < 0) fail(SSW_E_INVALID, "face %u of cell %u has no reverse face in cell %d", f, c, nb);
            const double *m = g->face_normal + 3 * (size_t)bestg;
            if (std::fabs(n[0] + m[0]) > 1e-9 || std::fabs(n[1] + m[1]) > 1e-9 || std::fabs(n[2] + m[2]) > 1e-9)
                fail(SSW_E_INVALID, "face %u of cell %u: normal is not the negative of its reverse face's", f, c);
            rev[f] = g->face_area[bestg];
        }
        if (has_periodic) {
            pidx_h[c] = (int32_t)pcells_h.size();
            pcells_h.push_back(c);
        }
    }
    n_periodic = (uint32_t)pcells_h.size();

    // ---- upload ----
    dirs_dev.alloc(3 * (size_t)Dl); dirs_dev.upload(dirs_all.data() + 3 * (size_t)d0, 3 * (size_t)Dl, stream);
    face_geo.alloc(F); face_geo.upload(geo.data(), F, stream);
    face_rev.alloc(F); face_rev.upload(rev.data(), F, stream);
    face_nb.alloc(F); face_nb.upload(g->face_neighbour, F, stream);
    face_kind.alloc(F); face_kind.upload(g->face_kind, F, stream);
    face_off.alloc(N + 1); face_off.upload(off.data(), N + 1, stream);
    size.alloc(N); size.upload(g->cell_size, N, stream);
    volume.alloc(N); volume.upload(g->cell_volume, N, stream);
    // per-cell state: one arena per rank, so that the ranks of a sharded job can map each other's (peer.cuh)
    {
        const PeerLayout L = peer_layout(N, P.world_size);
        arena.alloc(L.bytes);
        arena.zero(stream);
        pt.L = L;
        pt.world = P.world_size;
        pt.rank = P.rank;
        pt.n_per = cells_per_rank();
        pt.n_cells = N;
        for (int r = 0; r < kMaxPeers; ++r) pt.base[r] = nullptr;
        pt.base[P.rank] = arena.p;
        x.view(arena.p + L.x, N); T.view(arena.p + L.T, N); ts.view(arena.p + L.ts, N); tau.view(arena.p + L.tau, N);
        prev_rate.view(arena.p + L.prev, N); att.view(arena.p + L.att, N); ion_time.view(arena.p + L.ion, N);
        photon.view(arena.p + L.photon, N); level.view(arena.p + L.level, N);
    }
    rho.alloc(N); rho.upload(density, N, stream);
    x.upload(xhii, N, stream);
    T.upload(temperature, N, stream);
    src.alloc(N); src.upload(source, N, stream);
    {
        std::vector<double> infv(N, std::numeric_limits<double>::infinity());   // IonizationTime::default(), components.rs:79-83
        ion_time.upload(infv.data(), N, stream);
        CUDA_CHECK(cudaStreamSynchronize(stream));
    }
    pidx.alloc(N); pidx.upload(pidx_h.data(), N, stream);
    pcells.alloc(n_periodic); if (n_periodic) pcells.upload(pcells_h.data(), n_periodic, stream);
    CUDA_CHECK(cudaMemsetAsync(level.p, P.n_levels - 1, N, stream));  // initial_level, mod.rs:206
    const size_t ND = (size_t)N * Dl;
    q.alloc(ND); q.zero(stream);
    incoming.alloc(ND); incoming.zero(stream);
    per_lag.alloc((size_t)n_periodic * Dl); per_lag.zero(stream);
    per_new.alloc((size_t)n_periodic * Dl); per_new.zero(stream);
    ctl.alloc(1);
    flags.alloc(N);
    n_selected.alloc(1);
    rate_act.alloc((size_t)cells_per_rank() * P.world_size);   // >= N: padded to whole per-rank chunks (reduce-scatter)
    rate_act.zero(stream);
    cellrec.alloc(N);
    cell_tmp.alloc(N);
    cell_tmp2.alloc(N);
    hist.alloc(33);
    blocks_done.alloc(1); blocks_done.zero(stream);
    chem_stats.alloc(1); chem_stats.zero(stream);
    last_attempts.alloc(N); last_attempts.zero(stream);
    timing_level = (int)std::min<uint32_t>(1u, stream_env_u32("SSW_TIMERS", 1));
    CUDA_CHECK(cudaStreamSynchronize(stream));  // host vectors go out of scope

    bind();
    attenuation_kernel<<<cdiv(N, 256), 256, 0, stream>>>(cell_view(), N);
    launched();
    CUDA_CHECK(cudaGetLastError());

    lowest_allowed = P.n_levels - 1;  // timestep_state.rs:17
    first_done = false;
    bin_count.assign(P.n_levels, 0);
    bin_count[P.n_levels - 1] = N;
    sched.resize(P.n_levels + 1);
    for (auto &s : sched) s.reset(new Schedule());
    coop_blocks_build = coop_grid((const void *)sweep_build_kernel, 256, num_sms);
    coop_blocks_replay = coop_grid((const void *)sweep_replay_kernel, 256, num_sms);
    coop_blocks_mini = coop_grid((const void *)mini_replay_kernel, 256, num_sms);
    CUDA_CHECK(cudaStreamSynchronize(stream));
}

// Cell centres -> spatial patches: boxes of a regular lattice over the bounding box of the centres, sized for
// about SSW_PATCH_CELLS (512) cells each.  Host preprocessing, once per grid.
// Cell centres -> spatial patches: boxes of a regular lattice over the bounding box of the centres, sized for about
// `target` cells each (halved until no patch exceeds kMaxPatchCells).  Pure host code (also behind ssw_patch_lattice
// for the CPU tests).  Returns the number of patches, 0 if no lattice qualifies (`why` says why).
static uint32_t build_patch_lattice(const double *xyz, uint32_t N, double target, std::vector<uint32_t> &pof,
                                    std::vector<uint32_t> &poff, std::vector<uint32_t> &pcl, std::vector<uint16_t> &lidx,
                                    uint32_t &max_cells, std::string &why, uint32_t lattice[3]) {
    double lo[3], hi[3];
    for (int k = 0; k < 3; ++k) { lo[k] = std::numeric_limits<double>::infinity(); hi[k] = -lo[k]; }
    for (uint32_t c = 0; c < N; ++c)
        for (int k = 0; k < 3; ++k) {
            const double v = xyz[3 * (size_t)c + k];
            if (!std::isfinite(v)) fail(SSW_E_INVALID, "cell %u: position is not finite", c);
            lo[k] = std::min(lo[k], v);
            hi[k] = std::max(hi[k], v);
        }
    int dims = 0;
    double vol = 1.0;
    for (int k = 0; k < 3; ++k)
        if (hi[k] > lo[k]) { ++dims; vol *= hi[k] - lo[k]; }
    target = std::min<double>(std::max<double>(target, 8.0), (double)kMaxPatchCells);
    pof.assign(N, 0);
    pcl.assign(N, 0);
    lidx.assign(N, 0);
    for (int attempt = 0; attempt < 8; ++attempt, target *= 0.5) {
        uint32_t nbx[3] = {1, 1, 1};
        double org[3] = {lo[0], lo[1], lo[2]}, width[3] = {1.0, 1.0, 1.0};
        if (dims > 0) {
            const double spacing = std::pow(vol / (double)N, 1.0 / dims);   // mean distance of neighbouring centres
            double vol_ext = 1.0;
            for (int k = 0; k < 3; ++k)
                if (hi[k] > lo[k]) vol_ext *= hi[k] - lo[k] + spacing;
            const double edge = std::pow(vol_ext * target / (double)N, 1.0 / dims);
            for (int k = 0; k < 3; ++k) {
                if (!(hi[k] > lo[k])) continue;
                const double len = hi[k] - lo[k] + spacing;
                nbx[k] = (uint32_t)std::max<double>(1.0, std::min<double>(1024.0, std::floor(len / edge + 0.5)));
                org[k] = lo[k] - 0.5 * spacing;
                width[k] = len / nbx[k];
            }
        }
        const uint64_t P64 = (uint64_t)nbx[0] * nbx[1] * nbx[2];
        if (P64 > (1u << 20)) { why = "more than 2^20 patches"; return 0; }
        const uint32_t Pn = (uint32_t)P64;
        std::vector<uint32_t> count(Pn, 0);
        for (uint32_t c = 0; c < N; ++c) {
            uint32_t bx[3];
            for (int k = 0; k < 3; ++k) {
                const double t = (xyz[3 * (size_t)c + k] - org[k]) / width[k];
                bx[k] = (uint32_t)std::min<double>((double)nbx[k] - 1.0, std::max<double>(0.0, std::floor(t)));
            }
            pof[c] = (bx[0] * nbx[1] + bx[1]) * nbx[2] + bx[2];
            count[pof[c]]++;
        }
        const uint32_t mx = *std::max_element(count.begin(), count.end());
        if (mx > kMaxPatchCells) continue;
        poff.assign((size_t)Pn + 1, 0);
        for (uint32_t p = 0; p < Pn; ++p) poff[p + 1] = poff[p] + count[p];
        std::vector<uint32_t> cur(poff.begin(), poff.end() - 1);
        for (uint32_t c = 0; c < N; ++c) {   // ascending cell index inside a patch
            const uint32_t pos = cur[pof[c]]++;
            pcl[pos] = c;
            lidx[c] = (uint16_t)(pos - poff[pof[c]]);
        }
        max_cells = mx;
        for (int k = 0; k < 3; ++k) lattice[k] = nbx[k];
        return Pn;
    }
    why = "no patch lattice with <= 1024 cells per patch";
    return 0;
}

void Sweep::set_positions(const double *xyz) {
    if (!xyz) fail(SSW_E_INVALID, "null positions");
    have_patches = false;
    patch_note.clear();
    if (state) fail(SSW_E_INVALID, "cell positions must be set before the first all-cells schedule is compiled");
    // patch size: a direction shard (few local directions) is bound by the chain of dependent macro-tiles -> large
    // patches (8^3), few levels; with many directions the sweep is throughput-bound and small patches (5^3) keep
    // more macro-tiles resident per SM (measured on B200)
    const double target = (double)stream_env_u32("SSW_PATCH_CELLS", Dl <= 24 ? 512 : 125);
    std::vector<uint32_t> pof, poff, pcl;
    std::vector<uint16_t> lidx;
    uint32_t mx = 0;
    const uint32_t Pn = build_patch_lattice(xyz, N, target, pof, poff, pcl, lidx, mx, patch_note, patch_dims);
    if (!Pn) return;
    CUDA_CHECK(cudaSetDevice(device));
    patch_of.alloc(N); patch_of.upload(pof.data(), N, stream);
    patch_lidx.alloc(N); patch_lidx.upload(lidx.data(), N, stream);
    patch_off.alloc((size_t)Pn + 1); patch_off.upload(poff.data(), (size_t)Pn + 1, stream);
    patch_cells.alloc(N); patch_cells.upload(pcl.data(), N, stream);
    CUDA_CHECK(cudaStreamSynchronize(stream));
    n_patches = Pn;
    patch_max_cells = mx;
    have_patches = true;
}

// rotate_directions_system (src/sweep/direction/mod.rs:158-174): a new direction set of the same size.  The flux state
// follows: new direction i continues with the outgoing rates of the old direction it is best aligned with.  (The
// reference's `remap`, :190-205, intends the same kernel but ASSIGNS inside its double loop, `values[i] = old[j] *
// kernel[i][j]`, so that only the last old direction survives -- that bug is not reproduced.)  incoming_total_rate and
// periodic_source need no remap here: the gather form derives them from the neighbours' outgoing rates.
void Sweep::set_directions(const double *dirs_xyz) {
    if (!dirs_xyz) fail(SSW_E_INVALID, "null directions");
    if (P.world_size > 1) fail(SSW_E_INVALID, "ssw_set_directions is not available under direction sharding");
    bind();
    // the outgoing rates under the old directions, cell-major
    DevBuf<double> out_old;
    out_old.alloc((size_t)N * Dl);
    dir_state_kernel<<<cdiv(N, 256), 256, 0, stream>>>(grid_view(), state_view(), 1, Dl, out_old.p, nullptr);
    launched();
    std::vector<int32_t> map(D);
    for (int i = 0; i < D; ++i) {
        double best = -std::numeric_limits<double>::infinity();
        for (int j = 0; j < D; ++j) {
            const double dot = dirs_xyz[3 * i] * dirs_all[3 * j] + dirs_xyz[3 * i + 1] * dirs_all[3 * j + 1] +
                               dirs_xyz[3 * i + 2] * dirs_all[3 * j + 2];
            if (dot > best) { best = dot; map[i] = j; }
        }
    }
    CUDA_CHECK(cudaStreamSynchronize(stream));
    // everything compiled or cached for the old directions goes
    state = nullptr;
    for (auto &sc : sched) sc.reset(new Schedule());
    levels_version++;
    photon_valid = false;
    rotating = true;
    dirs_all.assign(dirs_xyz, dirs_xyz + 3 * (size_t)D);
    dirs_dev.upload(dirs_all.data(), 3 * (size_t)D, stream);
    DevBuf<int32_t> map_dev;
    map_dev.alloc(D);
    map_dev.upload(map.data(), D, stream);
    if (!q.p) q.alloc((size_t)N * Dl);
    remap_directions_kernel<<<cdiv(N, 256), 256, 0, stream>>>(grid_view(), out_old.p, map_dev.p, Dl, q.p);
    launched();
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaStreamSynchronize(stream));
}

void Sweep::refresh_histogram() {
    hist.zero(stream);
    histogram_kernel<<<cdiv(N, 256), 256, 0, stream>>>(level.p, N, hist.p);
    launched();
    unsigned long long h[33];
    CUDA_CHECK(cudaMemcpyAsync(h, hist.p, sizeof h, cudaMemcpyDeviceToHost, stream));
    CUDA_CHECK(cudaStreamSynchronize(stream));
    for (int l = 0; l < P.n_levels; ++l) bin_count[l] = h[l];
    levels_version++;
}

// ActiveList::enumerate_active (src/sweep/active_list.rs:55-66) as a compacted ascending list
void Sweep::build_active_list(Schedule &S, int cur) {
    iota_active_kernel<<<cdiv(N, 256), 256, 0, stream>>>(level.p, cur, N, flags.p);
    launched();
    S.act_list.ensure(S.n_act);
    cub::CountingInputIterator<uint32_t> iota(0);
    size_t bytes = 0;
    CUDA_CHECK(cub::DeviceSelect::Flagged(nullptr, bytes, iota, flags.p, S.act_list.p, n_selected.p, (int)N, stream));
    cub_temp.ensure(bytes);
    CUDA_CHECK(cub::DeviceSelect::Flagged(cub_temp.p, bytes, iota, flags.p, S.act_list.p, n_selected.p, (int)N, stream));
    launched(2);
    uint32_t got = 0;
    CUDA_CHECK(cudaMemcpyAsync(&got, n_selected.p, sizeof got, cudaMemcpyDeviceToHost, stream));
    CUDA_CHECK(cudaStreamSynchronize(stream));
    if (got != S.n_act) fail(SSW_E_CUDA, "active list has %u cells, histogram says %u", got, S.n_act);
    // the cells whose incoming rates a sweep over this set changes: Local neighbours of the set
    CUDA_CHECK(cudaMemsetAsync(flags.p, 0, N, stream));
    mark_touched_kernel<<<cdiv(S.n_act, 256), 256, 0, stream>>>(grid_view(), S.act_list.p, S.n_act, flags.p);
    S.touch_list.ensure(N);
    CUDA_CHECK(cub::DeviceSelect::Flagged(cub_temp.p, bytes, iota, flags.p, S.touch_list.p, n_selected.p, (int)N, stream));
    launched(3);
    CUDA_CHECK(cudaMemcpyAsync(&got, n_selected.p, sizeof got, cudaMemcpyDeviceToHost, stream));
    CUDA_CHECK(cudaStreamSynchronize(stream));
    S.n_touch = got;
}

void Sweep::gather_periodic(double *dst, const uint32_t *act, uint32_t n_act) {
    if (!n_periodic) return;
    dim3 grid(cdiv(act ? n_act : n_periodic, 256), Dl);
    periodic_gather_kernel<<<grid, 256, 0, stream>>>(grid_view(), pcells.p, n_periodic, act, n_act, pidx.p,
                                                     state_view(), dst);
    launched();
}

// Kahn peeling (+ solve) for the active set of `cur`: init_counts, get_initial_tasks, solve
// (src/sweep/mod.rs:346-398, 291-314).  The resulting level sets are stored in S.
void Sweep::build_schedule(Schedule &S, int cur, bool solve, int dl_base, int n_dl, int32_t *wl) {
    const uint64_t n_tasks = (uint64_t)S.n_act * n_dl;
    missing.ensure((size_t)N * Dl);
    queue_scratch.ensure(n_tasks);
    const uint32_t cap = (uint32_t)std::min<uint64_t>((uint64_t)S.n_act + 2, 1u << 24);
    level_off_scratch.ensure(cap);
    CUDA_CHECK(cudaMemsetAsync(ctl.p, 0, sizeof(QueueCtl), stream));
    const uint32_t *act = S.act_list.n && S.n_act != N ? S.act_list.p : nullptr;
    {
        dim3 grid(cdiv(S.n_act, 256), n_dl);
        init_counts_kernel<<<grid, 256, 0, stream>>>(grid_view(), level.p, cur, act, S.n_act,
                                                     missing.p, queue_scratch.p, ctl.p, wl, dl_base);
        launched();
        CUDA_CHECK(cudaGetLastError());
    }
    SweepArgs a = sweep_args(cur);
    if (!solve) a.solve = 0;
    uint32_t *qp = queue_scratch.p;
    QueueCtl *cp = ctl.p;
    uint32_t *lo = level_off_scratch.p;
    uint32_t lcap = cap;
    void *args[] = {&a, &qp, &cp, &lo, &lcap, &wl};
    // a grid barrier costs more the more blocks take part: small active sets get a small grid
    const unsigned build_blocks = std::max(1u, std::min<unsigned>((unsigned)coop_blocks_build, cdiv(n_tasks, 256)));
    CUDA_CHECK(cudaLaunchCooperativeKernel((const void *)sweep_build_kernel, dim3(build_blocks),
                                           dim3(256), args, 0, stream));
    launched();
    QueueCtl h;
    CUDA_CHECK(cudaMemcpyAsync(&h, ctl.p, sizeof h, cudaMemcpyDeviceToHost, stream));
    CUDA_CHECK(cudaStreamSynchronize(stream));
    if (h.overflow) fail(SSW_E_NOMEM, "more than %u wavefront levels", cap);
    if (h.solved != n_tasks) {
        if (P.check_deadlock && h.cnt[0] == 0 && h.n_levels == 0)
            fail(SSW_E_DEADLOCK, "deadlock: no initial task (level %d)", cur);   // deadlock_detection.rs:86-98
        fail(SSW_E_DEADLOCK, "dependency cycle among active cells: solved %u of %llu tasks at level %d",
             h.solved, (unsigned long long)n_tasks, cur);
    }
    S.n_tasks = n_tasks;
    S.n_levels = h.n_levels;
    S.level_off_host.resize(h.n_levels + 1);
    CUDA_CHECK(cudaMemcpyAsync(S.level_off_host.data(), level_off_scratch.p, sizeof(uint32_t) * (h.n_levels + 1),
                               cudaMemcpyDeviceToHost, stream));
    CUDA_CHECK(cudaStreamSynchronize(stream));
    S.max_level_tasks = 0;
    for (uint32_t l = 0; l < h.n_levels; ++l)
        S.max_level_tasks = std::max(S.max_level_tasks, S.level_off_host[l + 1] - S.level_off_host[l]);
}

// task records for replays of a cached partial schedule (kernels.cuh, MiniView)
void Sweep::build_mini(Schedule &S) {
    const uint32_t n = (uint32_t)S.n_tasks;
    S.mini_valid = false;
    if (n == 0) return;
    S.m_off.ensure((size_t)n + 1);
    CUDA_CHECK(cudaMemsetAsync(S.m_off.p + n, 0, sizeof(uint32_t), stream));
    queue_scratch.ensure((size_t)n + 1);   // entry counts (the level-sorted tasks live in S.tasks by now)
    CUDA_CHECK(cudaMemsetAsync(queue_scratch.p + n, 0, sizeof(uint32_t), stream));
    mini_count_kernel<<<cdiv(n, 256), 256, 0, stream>>>(grid_view(), S.tasks.p, n, queue_scratch.p);
    size_t bytes = 0;
    CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, queue_scratch.p, S.m_off.p, (int)n + 1, stream));
    cub_temp.ensure(bytes);
    CUDA_CHECK(cub::DeviceScan::ExclusiveSum(cub_temp.p, bytes, queue_scratch.p, S.m_off.p, (int)n + 1, stream));
    uint32_t n_entries = 0;
    CUDA_CHECK(cudaMemcpyAsync(&n_entries, S.m_off.p + n, sizeof n_entries, cudaMemcpyDeviceToHost, stream));
    CUDA_CHECK(cudaStreamSynchronize(stream));
    S.m_src.ensure(n_entries);
    S.m_w.ensure(n_entries);
    S.m_self.ensure(n);
    S.m_ttot.ensure(n);
    mini_fill_kernel<<<cdiv(n, 256), 256, 0, stream>>>(grid_view(), state_view(), S.tasks.p, n, S.m_off.p, S.m_src.p, S.m_w.p,
                                                       S.m_self.p, S.m_ttot.p);
    launched(4);
    CUDA_CHECK(cudaGetLastError());
    S.mini_valid = true;
    S.mini_slot_state = state != nullptr;
}

// ------------------------------------------------------------------------------------------
// direction sharding over peer-mapped memory (peer.cuh)
// ------------------------------------------------------------------------------------------
void Sweep::peer_attach(void *const *bases) {
    if (P.world_size <= 1) fail(SSW_E_INVALID, "peer attach needs world_size > 1");
    for (int r = 0; r < P.world_size; ++r) {
        if (r == P.rank) continue;
        if (!bases[r]) fail(SSW_E_INVALID, "no arena for rank %d", r);
        pt.base[r] = static_cast<unsigned char *>(bases[r]);
    }
    // The wait half of a synchronisation point.  One rank per device (production): a stream memory operation
    // (cuStreamWaitValue32) -- nothing of the host or of an SM is involved; a polling kernel where the driver has none.
    // Ranks that SHARE a device (tests): the host polls the flag words instead, because a stream blocked on a peer
    // would dead-lock with the implicit device-wide synchronisation of cudaFree / cudaMalloc on the peer's thread.
    // SSW_PEER_WAIT = 0 / 1 / 2 overrides.
    stream_wait_value32 = nullptr;
    peer_wait_mode = (int)stream_env_u32("SSW_PEER_WAIT", (P.flags & SSW_FLAG_SHARED_DEVICE) ? 1 : 0);
    if (peer_wait_mode == 0) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess && fn) {
            stream_wait_value32 = reinterpret_cast<StreamWaitValue32>(fn);
        } else {
            (void)cudaGetLastError();
            peer_wait_mode = 2;
        }
    }
    if (peer_wait_mode == 1 && !aux_stream) CUDA_CHECK(cudaStreamCreateWithFlags(&aux_stream, cudaStreamNonBlocking));
    peers = true;
}

// every rank signals every rank, then waits for everybody's signal: what this rank queued before the point is visible
// to all ranks behind it, and vice versa
void Sweep::peer_sync_point() {
    peer_signal_kernel<<<1, 32, 0, stream>>>(pt);
    launched();
    peer_wait();
}

// the wait half; the signal was given by peer_signal_kernel or by the tail of the producing kernel (peer_signal_tail)
void Sweep::peer_wait() {
    const size_t t = tic(T_ALLREDUCE);
    ++peer_epoch;
    const uint32_t target = peer_epoch * (uint32_t)P.world_size;
    if (peer_wait_mode == 0) {
        const unsigned long long addr = (unsigned long long)(uintptr_t)(arena.p + pt.L.flags);
        const int rc = stream_wait_value32(stream, addr, target, /*CU_STREAM_WAIT_VALUE_GEQ*/ 0u);
        if (rc != 0) fail(SSW_E_COMM, "cuStreamWaitValue32 failed (%d)", rc);
    } else if (peer_wait_mode == 1) {
        const auto t00 = std::chrono::steady_clock::now();
        CUDA_CHECK(cudaStreamSynchronize(stream));   // my signal is out
        if (!flags_host) CUDA_CHECK(cudaHostAlloc((void **)&flags_host, (size_t)kMaxPeers * kPeerFlagStride, cudaHostAllocDefault));
        unsigned char *flags_h = flags_host;
        const size_t flags_bytes = sizeof(uint32_t);
        const auto t0 = std::chrono::steady_clock::now();
        unsigned long polls = 0;
        for (;;) {
            ++polls;
            CUDA_CHECK(cudaMemcpyAsync(flags_h, arena.p + pt.L.flags, flags_bytes, cudaMemcpyDeviceToHost, aux_stream));
            CUDA_CHECK(cudaStreamSynchronize(aux_stream));
            uint32_t v;
            std::memcpy(&v, flags_h, sizeof v);
            if ((int32_t)(v - target) >= 0) break;
            if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(60))
                fail(SSW_E_COMM, "peer exchange: a rank did not reach synchronisation point %u within 60 s", peer_epoch);
            std::this_thread::yield();
        }
        if (stream_env_u32("SSW_PEER_DEBUG", 0)) {
            const auto t1 = std::chrono::steady_clock::now();
            fprintf(stderr, "[peer %d] point %u: stream sync %.3f ms, flag wait %.3f ms (%lu polls)\n", P.rank, peer_epoch,
                    std::chrono::duration<double, std::milli>(t0 - t00).count(), std::chrono::duration<double, std::milli>(t1 - t0).count(), polls);
        }
    } else {
        peer_wait_kernel<<<1, 32, 0, stream>>>(pt, target);
        launched();
    }
    CUDA_CHECK(cudaGetLastError());
    toc(t);
}

// in-place sum over the ranks, folded in rank order on every rank (deterministic): each rank exposes its buffer in
// its arena, everybody reads everybody's
void Sweep::peer_allreduce(double *buf, uint64_t n) {
    if (n > N) fail(SSW_E_INVALID, "peer all-reduce of more than n_cells values");
    CUDA_CHECK(cudaMemcpyAsync(arena.p + pt.L.scratch, buf, sizeof(double) * n, cudaMemcpyDeviceToDevice, stream));
    peer_sync_point();
    peer_sum_kernel<<<cdiv(n, 256), 256, 0, stream>>>(pt, pt.L.scratch, buf, n);
    launched();
    peer_sync_point();   // nobody overwrites its scratch before everybody has read it
}

void Sweep::peer_pull(uint64_t off) {
    peer_pull_kernel<<<cdiv(N, 256), 256, 0, stream>>>(pt, off);
    launched();
}

PeerChem Sweep::peer_chem() const {
    PeerChem pc{};
    pc.world = peers ? P.world_size : 0;
    if (!peers) return pc;
    pc.first = pt.first();
    pc.n_own = pt.n_own();
    pc.n_per = pt.n_per;
    pc.recv = pt.f64(P.rank, pt.L.recv);
    for (int r = 0; r < P.world_size; ++r) {
        pc.att[r] = pt.f64(r, pt.L.att);
        pc.counter[r] = reinterpret_cast<unsigned int *>(pt.base[r] + pt.L.flags);
    }
    pc.blocks_done = blocks_done.p;
    return pc;
}

PeerLevels Sweep::peer_levels() const {
    PeerLevels pl{};
    pl.world = peers ? P.world_size : 0;
    for (int r = 0; peers && r < P.world_size; ++r) pl.level[r] = pt.base[r] + pt.L.level;
    return pl;
}

void Sweep::maybe_allreduce(double *buf, uint64_t n) {
    if (P.world_size <= 1) return;  // one direction shard: no collective (north_star)
    if (peers) return peer_allreduce(buf, n);
    if (!allreduce) fail(SSW_E_COMM, "world_size > 1 but no all-reduce hook set (ssw_set_allreduce)");
    const size_t t = tic(T_ALLREDUCE);
    // stream-ordered: the hook enqueues the collective behind the work already queued on `stream`
    if (allreduce(allreduce_ctx, buf, n, (void *)stream) != 0) fail(SSW_E_COMM, "all-reduce hook failed");
    if (P.flags & SSW_FLAG_SHARED_DEVICE) bind();   // another handle of this process may have run inside the hook
    toc(t);
}

void Sweep::run_collective(int op, double *buf, uint64_t n_per_rank) {
    const size_t t = tic(T_ALLREDUCE);
    if (collective(collective_ctx, op, buf, n_per_rank, (void *)stream) != 0) fail(SSW_E_COMM, "collective hook failed (op %d)", op);
    if (P.flags & SSW_FLAG_SHARED_DEVICE) bind();
    toc(t);
}

// HydrogenOnly::update_abundances over the cells of one launch (src/sweep/mod.rs:559-572).  When the last step showed
// substepping (more attempts than cells) the launch order is sorted by each cell's last substep count.
void Sweep::launch_chemistry(const uint32_t *act, uint32_t n, const double *rate, uint32_t first_cell, const ChemParams &cp,
                             const PeerChem &pc) {
    if (n == 0) return;
    const uint32_t *order = nullptr;
    if (chem_balance && n >= 2048) {
        order_keys.ensure(n); order_keys_out.ensure(n); order_vals.ensure(n); order_vals_out.ensure(n);
        chem_order_keys_kernel<<<cdiv(n, 256), 256, 0, stream>>>(act, n, first_cell, last_attempts.p, order_keys.p, order_vals.p);
        size_t bytes = 0;
        CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, order_keys.p, order_keys_out.p, order_vals.p, order_vals_out.p,
                                                   (int)n, 0, 16, stream));
        cub_temp.ensure(bytes);
        CUDA_CHECK(cub::DeviceRadixSort::SortPairs(cub_temp.p, bytes, order_keys.p, order_keys_out.p, order_vals.p, order_vals_out.p,
                                                   (int)n, 0, 16, stream));
        launched(3);
        order = order_vals_out.p;
    }
    // one warp per block: a finished warp frees its registers at once (the kernel holds no shared memory and has no
    // block-level phase)
    chemistry_kernel<<<cdiv(n, kChemThreads), kChemThreads, 0, stream>>>(cell_view(), act, n, rate, cp, chem_stats.p, first_cell, pc, order,
                                                       last_attempts.p);
    launched();
}

// Sweep::single_sweep (src/sweep/mod.rs:274-289)
void Sweep::single_sweep(int cur) {
    const uint64_t n_act64 = count_at_least(cur);
    if (n_act64 == 0) return;
    const uint32_t n_act = (uint32_t)n_act64;
    const bool all = n_act == N;
    Schedule &S = *sched[all ? P.n_levels : cur];
    const bool cache_ok = !(P.flags & SSW_FLAG_NO_SCHEDULE_CACHE);
    const bool reuse = (all && S.compiled.valid) ||
                       (cache_ok && S.valid && S.n_act == n_act && (all || S.version == levels_version));
    const size_t t_sweep = tic(T_SWEEP, cur);

    const bool use_compiled = reuse && all && !rotating && !(P.flags & SSW_FLAG_NO_COMPILED_PATH) && compiled_supported();
    if (!reuse) {
        S.valid = false;
        S.mini_valid = false;
        if (state != &S.compiled) S.compiled.release();
        S.n_act = n_act;
        if (!all) {
            const size_t t_list = tic(T_SCHED);
            {
                SchedProbe pr(stream, "active list");
                build_active_list(S, cur);
            }
            toc(t_list);
        }
    }
    const uint32_t *act = all ? nullptr : S.act_list.p;
    // periodic_source as the tasks of this sweep will read it (lagged); the
    // compiled path reads the donors' previous outgoing rates directly (stream.cuh).  A partial
    // active set only refreshes the rows of its own cells.
    // A small active set whose level sets are cached runs as ONE launch (small.cuh): lagged periodic rows, replay,
    // photon_rate bookkeeping, new periodic rows, rate fold and -- under peer-mapped sharding, where it is the default --
    // the hand-over of the partial rates to their owners.  (On one GPU the chain of small kernels is as fast: measured.)
    // One block does all of it, so only while the set holds few (cell, direction) rows: measured with 64 active cells,
    // 10 and 21 local directions (8 and 4 ranks) are on par with the separate kernels, 42 (2 ranks) cost 0.74 ms per
    // step against 0.28 ms.
    bool fused_small = false;
    if (reuse && !all && !use_compiled && S.n_tasks > 0 && S.n_tasks <= (64u << 20) && S.max_level_tasks <= 2048 &&
        (uint64_t)n_act * Dl <= stream_env_u32("SSW_FUSED_SMALL_ROWS", 1536) && (uint64_t)S.n_touch * Dl <= 262144 &&
        (P.world_size == 1 || peers) && stream_env_u32("SSW_FUSED_SMALL", peers ? 1 : 0)) {
        if (!S.mini_valid || S.mini_slot_state != (state != nullptr)) {
            const size_t t_sched = tic(T_SCHED);
            build_mini(S);
            toc(t_sched);
        }
        fused_small = S.mini_valid;
    }
    if (!use_compiled && !fused_small) gather_periodic(per_lag.p, act, n_act);

    if (!reuse) {
        const size_t t_sched = tic(T_SCHED);
        const size_t t_k = tic(T_KERNEL, cur);
        {
            SchedProbe pr(stream, "init_counts + build");
            build_schedule(S, cur, /*solve=*/true, 0, Dl, nullptr);
            if (pr.on) fprintf(stderr, "[sched] level %d: %u active cells, %llu tasks, %u wavefront levels\n", cur, S.n_act,
                               (unsigned long long)S.n_tasks, S.n_levels);
        }
        toc(t_k);
        timings.sweep_kernel_launches += 1;
        timings.sweep_kernel_tasks += S.n_tasks;
        timings.kernel_level_tasks[cur] += S.n_tasks;
        timings.kernel_level_launches[cur] += 1;
        // keep the level sets: sort every level by (direction, cell) so replays walk memory in order
        S.tasks.ensure(S.n_tasks);
        S.level_off.ensure(S.n_levels + 1);
        CUDA_CHECK(cudaMemcpyAsync(S.level_off.p, level_off_scratch.p, sizeof(uint32_t) * (S.n_levels + 1),
                                   cudaMemcpyDeviceToDevice, stream));
        SchedProbe pr_keep(stream, "keep level sets");
        if (cache_ok) {
            // The level sets as the build left them (level by level, arbitrary order inside a level: the gather form does
            // not care).  Sorting every level by (direction, cell) makes replays walk memory in order; it pays for the
            // all-cells schedule (built once per grid), not for partial sets, which are rebuilt whenever a level changes
            // (measured on the ionization-front workload: the segmented sort was most of the rebuild).
            if (all || stream_env_u32("SSW_SORT_LEVELS", 0)) {
                size_t bytes = 0;
                CUDA_CHECK(cub::DeviceSegmentedSort::SortKeys(nullptr, bytes, queue_scratch.p, S.tasks.p,
                                                              (int64_t)S.n_tasks, (int64_t)S.n_levels,
                                                              S.level_off.p, S.level_off.p + 1, stream));
                cub_temp.ensure(bytes);
                CUDA_CHECK(cub::DeviceSegmentedSort::SortKeys(cub_temp.p, bytes, queue_scratch.p, S.tasks.p,
                                                              (int64_t)S.n_tasks, (int64_t)S.n_levels,
                                                              S.level_off.p, S.level_off.p + 1, stream));
                launched(3);
            } else {
                CUDA_CHECK(cudaMemcpyAsync(S.tasks.p, queue_scratch.p, sizeof(uint32_t) * S.n_tasks, cudaMemcpyDeviceToDevice, stream));
            }
            S.valid = true;
            S.version = levels_version;
        }
        stat[SSW_STAT_SCHEDULE_BUILDS]++;
        toc(t_sched);
    } else {
        if (use_compiled && !S.compiled.valid) {
            const size_t t_sched = tic(T_SCHED);
            try {
                bool patched = false;
                // Which compiled form?  The patch-ordered dataflow (no device-wide barriers, 46 instead of 382 dependent
                // steps at 128^3) whenever cell positions are known and the grid admits it; else the level-barrier
                // stream.  SSW_PATCH=0 switches the patch form off.
                const bool want_patch = stream_env_u32("SSW_PATCH", 1) != 0;
                if (!want_patch && have_patches) patch_note = "patch form switched off (SSW_PATCH=0)";
                if (have_patches && !(P.flags & SSW_FLAG_NO_PATCH_PATH) && want_patch) {
                    try {
                        compile_patch_schedule(S.compiled, grid_view(), patch_view(), dirs_all.data() + 3 * (size_t)d0,
                                               S.tasks.p, S.level_off.p, S.n_tasks, S.n_levels, Dl, pcells.p, n_periodic, q.p,
                                               &S.level_off_host, num_sms, stream, &stat[SSW_STAT_KERNEL_LAUNCHES]);
                        patched = true;
                        patch_note.clear();
                        stat[SSW_STAT_PATCH_MACRO_TILES] = S.compiled.n_mt;
                        stat[SSW_STAT_PATCH_LEVELS] = S.compiled.patch_levels;
                        stat[SSW_STAT_PATCH_PHASES] = S.compiled.n_phases;
                    } catch (const PatchUnsupported &e) {
                        patch_note = e.what();   // keep the level-barrier stream
                    }
                }
                if (!patched) {
                    // the walk form (one block per direction, recent rates in a shared-memory window, walk.cuh) unless
                    // the grid does not fit it; then the level-barrier stream
                    // Which one: the stream form crosses one device-wide barrier per wavefront level (~4.5 us each,
                    // profiles/), the walk form crosses none but keeps one direction on one SM.  Walk when the barriers
                    // would dominate the stream form's byte time (unstructured grids: thousands of levels) and there
                    // are enough directions to occupy the SMs; SSW_WALK = 0 / 1 decides by hand.
                    const double barrier_s = 4.5e-6 * (double)S.n_levels;
                    const double stream_s = 12.0 * 0.5 * (double)F * (double)Dl / 2.5e12;
                    bool want_walk = Dl >= 32 && barrier_s > 2.0 * stream_s;
                    // unstructured grids (many upwind faces per task, small levels): four interleaved direction groups
                    // hide the level barrier of the stream form better than two (measured, 128^3 Voronoi: 25.6 -> 17.5 ms)
                    const uint32_t groups = (double)F / (double)N > 10.0 ? 4u : 2u;
                    const uint32_t forced = stream_env_u32("SSW_WALK", 2);
                    if (forced < 2) want_walk = forced != 0;
                    bool walked = false;
                    try {
                        if (want_walk) {
                            compile_schedule(S.compiled, grid_view(), S.tasks.p, S.level_off.p, S.n_tasks, S.n_levels, Dl,
                                             pcells.p, n_periodic, pidx.p, q.p, num_sms, stream, &stat[SSW_STAT_KERNEL_LAUNCHES],
                                             /*allow_walk=*/true);
                            walked = true;
                            stat[SSW_STAT_WALK_WINDOW] = S.compiled.window;
                            stat[SSW_STAT_WALK_NEAR_PERMILLE] =
                                (uint64_t)(1000.0 * (double)S.compiled.n_near / (double)std::max<uint64_t>(1, S.compiled.n_near + S.compiled.n_far));
                        }
                    } catch (const WalkUnsupported &e) {
                        patch_note += std::string(patch_note.empty() ? "" : "; ") + e.what();
                    }
                    if (!walked)
                        compile_schedule(S.compiled, grid_view(), S.tasks.p, S.level_off.p, S.n_tasks, S.n_levels, Dl,
                                         pcells.p, n_periodic, pidx.p, q.p, num_sms, stream, &stat[SSW_STAT_KERNEL_LAUNCHES],
                                         /*allow_walk=*/false, groups);
                }
            } catch (const std::exception &e) {
                fail(SSW_E_CUDA, "%s", e.what());
            }
            state = &S.compiled;   // out_slot is now the flux state; the natural-layout copy goes
            q.release();
            toc(t_sched);
        }
        const size_t t_k = tic(T_KERNEL, cur);
        SweepArgs a = sweep_args(cur);
        if (use_compiled && S.compiled.valid) {
            try {
                if (S.compiled.solo) {
                    run_walk(S.compiled, att.p, src.p, (double)D, P.significant_rate_threshold_per_s, stream,
                             &stat[SSW_STAT_KERNEL_LAUNCHES]);
                } else {
                cellrec_kernel<<<cdiv(N, 256), 256, 0, stream>>>(att.p, src.p, (double)D, N, cellrec.p);
                launched();
                if (S.compiled.patch_mode)
                    run_patch(S.compiled, cellrec.p, P.significant_rate_threshold_per_s, stream,
                              &stat[SSW_STAT_KERNEL_LAUNCHES]);
                else
                    run_compiled(S.compiled, cellrec.p, P.significant_rate_threshold_per_s, stream,
                                 &stat[SSW_STAT_KERNEL_LAUNCHES]);
                }
            } catch (const std::exception &e) {
                fail(SSW_E_CUDA, "%s", e.what());
            }
        } else {
            const uint32_t *qp = S.tasks.p;
            const uint32_t *lo = S.level_off.p;
            uint32_t nl = S.n_levels;
            void *args[] = {&a, &qp, &lo, &nl};
            // partial schedules replay from stored task records; they are (re)built on first use and when the
            // flux state moved into slot order since
            const bool want_mini = !all && S.n_tasks <= (64u << 20);
            if (fused_small) {
                SmallSweepArgs fs;
                fs.a = a; fs.m = S.mini_view(); fs.queue = qp; fs.level_off = lo; fs.n_levels = nl;
                fs.act = act; fs.n_act = n_act; fs.per_lag = per_lag.p; fs.per_new = per_new.p;
                // photon_rate bookkeeping inside the block only while it is small; else its own (parallel) kernel below
                const bool photon_inside = (uint64_t)S.n_touch * Dl <= (uint64_t)kSmallPhotonScratch;
                fs.touch = photon_valid && S.n_touch && photon_inside ? S.touch_list.p : nullptr;
                fs.n_touch = S.n_touch; fs.photon = photon.p; fs.rate_act = rate_act.p; fs.n_local_dirs = Dl;
                fs.peers = P.world_size > 1 && peers ? 1 : 0;
                small_sweep_kernel<<<1, kMiniSmallThreads, 0, stream>>>(fs, pt);
                if (photon_valid && S.n_touch && !photon_inside) {
                    photon_patch_kernel<<<S.n_touch, kMaxDirs, 0, stream>>>(grid_view(), state_view(), S.touch_list.p, Dl, photon.p);
                    launched();
                }
            } else if (want_mini && (!S.mini_valid || S.mini_slot_state != (state != nullptr))) {
                SchedProbe pr(stream, "task records (mini)");
                build_mini(S);
            }
            if (fused_small) {
                // done above
            } else if (want_mini && S.mini_valid) {
                MiniView mv = S.mini_view();
                if (S.max_level_tasks <= 2048) {
                    mini_replay_small_kernel<<<1, kMiniSmallThreads, 0, stream>>>(a, mv, qp, lo, nl);
                } else {
                    void *margs[] = {&a, &mv, &qp, &lo, &nl};
                    const unsigned blocks = std::max(1u, std::min<unsigned>((unsigned)coop_blocks_mini, cdiv(S.max_level_tasks, 256)));
                    CUDA_CHECK(cudaLaunchCooperativeKernel((const void *)mini_replay_kernel, dim3(blocks), dim3(256), margs, 0, stream));
                }
            } else if (S.max_level_tasks <= 2048) {
                // every level fits one block: block barrier instead of the grid barrier
                sweep_replay_small_kernel<<<1, 512, 0, stream>>>(a, qp, lo, nl);
            } else {
                const unsigned replay_blocks =
                    std::max(1u, std::min<unsigned>((unsigned)coop_blocks_replay, cdiv(S.max_level_tasks, 256)));
                CUDA_CHECK(cudaLaunchCooperativeKernel((const void *)sweep_replay_kernel, dim3(replay_blocks),
                                                       dim3(256), args, 0, stream));
            }
            launched();
        }
        toc(t_k);
        timings.sweep_kernel_launches += 1;
        timings.sweep_kernel_tasks += S.n_tasks;
        timings.kernel_level_tasks[cur] += S.n_tasks;
        timings.kernel_level_launches[cur] += 1;
        stat[SSW_STAT_SCHEDULE_REPLAYS]++;
    }
    stat[SSW_STAT_TASKS_SOLVED] += S.n_tasks;
    stat[SSW_STAT_WAVEFRONT_LEVELS] = S.n_levels;
    toc(t_sweep);

    // update_chemistry (src/sweep/mod.rs:549-574)
    const size_t t_chem = tic(T_CHEM);
    if (use_compiled && S.compiled.valid) {
        // the compiled sweep left sum_d incoming and sum_d periodic_source in its group accumulators
        const Compiled &C = S.compiled;
        s_rate_finish_kernel<<<cdiv(N, 256), 256, 0, stream>>>(N, C.n_groups, C.n_groups_per, n_periodic, Dl, (double)D, C.acc_cell,
                                                               C.acc_per, pidx.p, src.p, rate_act.p, photon.p);
        photon_valid = true;
    } else if (fused_small) {
        // photon_rate bookkeeping, periodic rows and the rate fold ran inside small_sweep_kernel
    } else {
        if (all) {
            photon_valid = false;   // re-evaluated on demand (read_field)
        } else if (photon_valid && S.n_touch) {
            photon_patch_kernel<<<S.n_touch, kMaxDirs, 0, stream>>>(grid_view(), state_view(), S.touch_list.p, Dl, photon.p);
            launched();
        }
        gather_periodic(per_new.p, act, n_act);
        if (act && n_act <= 4096)
            rate_small_kernel<<<n_act, kMaxDirs, 0, stream>>>(act, N, Dl, incoming.p, pidx.p, per_new.p, n_periodic, rate_act.p);
        else
            rate_kernel<<<cdiv(n_act, 256), 256, 0, stream>>>(act, n_act, N, Dl, incoming.p, pidx.p, per_new.p,
                                                              n_periodic, rate_act.p);
    }
    launched();
    ChemParams cp;
    cp.max_timestep = P.max_timestep_s;
    cp.threshold = P.significant_rate_threshold_per_s;
    cp.scale_factor = P.scale_factor;
    cp.safety = P.chemistry_timestep_safety_factor;
    cp.prevent_cooling = P.prevent_cooling;
    if (P.world_size > 1 && peers) {
        // peer-mapped exchange: partial rates straight into the owners' receive buffers, chemistry on the owner, new
        // absorption factors straight into every rank's array (peer.cuh)
        if (!fused_small) {               // (the fused small sweep pushed its rates and signalled)
            peer_push_rates_kernel<<<cdiv(n_act, 256), 256, 0, stream>>>(pt, act, n_act, rate_act.p, blocks_done.p);
            launched();
        }
        peer_wait();                      // everybody's partial rates of my cells have arrived
        const uint32_t n_launch = all ? pt.n_own() : n_act;
        if (n_launch) {                   // the kernel's tail signals every rank
            launch_chemistry(act, n_launch, nullptr, all ? pt.first() : 0u, cp, peer_chem());
        } else {
            peer_signal_kernel<<<1, 32, 0, stream>>>(pt);
            launched();
        }
        peer_wait();                      // everybody's new absorption factors are in my array
    } else if (all && P.world_size > 1 && collective) {
        // all cells active on W ranks: reduce-scatter the partial rates, update the own slice of cells, all-gather
        // the results -- the chemistry is not replicated W times, and every rank ends with bit-identical state
        const uint32_t n_per = cells_per_rank();
        const uint32_t first = std::min<uint64_t>((uint64_t)P.rank * n_per, N);
        const uint32_t n_own = std::min<uint32_t>(n_per, N - first);
        chem_pack.ensure((size_t)P.world_size * kPackFields * n_per);
        run_collective(SSW_COLL_REDUCE_SCATTER, rate_act.p, n_per);
        double *chunk = chem_pack.p + (size_t)P.rank * kPackFields * n_per;
        if (n_own) {
            launch_chemistry(nullptr, n_own, rate_act.p + first, first, cp, PeerChem{});
            chem_pack_kernel<<<cdiv(n_own, 256), 256, 0, stream>>>(cell_view(), first, n_own, n_per, chunk);
            launched(2);
        }
        run_collective(SSW_COLL_ALL_GATHER, chem_pack.p, (uint64_t)kPackFields * n_per);
        chem_unpack_kernel<<<cdiv(N, 256), 256, 0, stream>>>(cell_view(), N, n_per, (uint32_t)P.rank, chem_pack.p);
        launched();
    } else {
        maybe_allreduce(rate_act.p, n_act);
        launch_chemistry(act, n_act, rate_act.p, 0u, cp, PeerChem{});
    }
    CUDA_CHECK(cudaGetLastError());
    toc(t_chem);
    stat[SSW_STAT_SINGLE_SWEEPS]++;
}

// Sweep::update_timestep_levels (src/sweep/mod.rs:576-589)
void Sweep::update_timestep_levels() {
    const size_t t = tic(T_LEVELS);
    hist.zero(stream);
    const uint32_t first = own_first(), n_own = own_count();
    if (n_own)
        levels_kernel<<<cdiv(n_own, 256), 256, 0, stream>>>(tau.p, level.p, first, n_own, P.n_levels, P.max_timestep_s,
                                                            P.timestep_safety_factor, lowest_allowed, hist.p, peer_levels());
    launched();
    unsigned long long h[33];
    if (peers) {
        // the owners pushed the new levels into every rank's array; the per-level counts travel the same way
        peer_hist_push_kernel<<<1, 64, 0, stream>>>(pt, hist.p, blocks_done.p);
        launched();
        peer_wait();
        std::vector<unsigned long long> all_h((size_t)P.world_size * kPeerHistWords);
        CUDA_CHECK(cudaMemcpyAsync(all_h.data(), arena.p + pt.L.hist, sizeof(unsigned long long) * all_h.size(),
                                   cudaMemcpyDeviceToHost, stream));
        toc(t);
        CUDA_CHECK(cudaStreamSynchronize(stream));
        for (int i = 0; i < 33; ++i) {
            h[i] = 0;
            for (int r = 0; r < P.world_size; ++r) h[i] += all_h[(size_t)r * kPeerHistWords + i];
        }
    } else {
    CUDA_CHECK(cudaMemcpyAsync(h, hist.p, sizeof h, cudaMemcpyDeviceToHost, stream));
    toc(t);
    CUDA_CHECK(cudaStreamSynchronize(stream));
    }
    for (int l = 0; l < P.n_levels; ++l) bin_count[l] = h[l];
    // level sets are rebuilt only when the active sets changed (north_star)
    if (h[32] != 0) levels_version++;
}

// Sweep::run_sweeps (src/sweep/mod.rs:258-272)
double Sweep::run_sweeps() {
    bind();
    const size_t t_step = tic(T_STEP);
    std::vector<uint64_t> counts(P.n_levels);
    for (int l = 0; l < P.n_levels; ++l) counts[l] = count_at_least(l);   // :240-245
    std::vector<int32_t> order(1u << (P.n_levels - lowest_allowed - 1));
    const int n = ssw_levels_in_sweep_order(P.n_levels, lowest_allowed, order.data(), (int)order.size());
    for (int i = 0; i < n; ++i)
        if (counts[order[i]] > 0) single_sweep(order[i]);
    const double elapsed = P.max_timestep_s * std::ldexp(1.0, -lowest_allowed);  // timestep_state.rs:77-79
    if (first_done && lowest_allowed > 0) lowest_allowed -= 1;                   // :37-48
    first_done = true;
    sim_time += elapsed;
    if (own_count()) {   // before the level update: its synchronisation point also publishes the ionization times
        ionization_time_kernel<<<cdiv(own_count(), 256), 256, 0, stream>>>(x.p, ion_time.p, own_first(), own_count(), sim_time);
        launched();
    }
    update_timestep_levels();
    // fold the chemistry statistics
    ChemStats cs;
    CUDA_CHECK(cudaMemcpyAsync(&cs, chem_stats.p, sizeof cs, cudaMemcpyDeviceToHost, stream));
    toc(t_step);
    resolve_timers();
    stat[SSW_STAT_CHEM_CELLS] = cs.cells;
    stat[SSW_STAT_CHEM_FAILURES] = cs.failures;
    stat[SSW_STAT_CHEM_ATTEMPTS] = cs.attempts;
    stat[SSW_STAT_CHEM_MAX_DEPTH] = cs.max_depth;
    // substepping in this step (more attempts than cell updates): balance the warps of the next step's launches
    {
        const uint64_t da = cs.attempts - chem_prev_attempts, dc = cs.cells - chem_prev_cells;
        chem_prev_attempts = cs.attempts;
        chem_prev_cells = cs.cells;
        const uint32_t forced = stream_env_u32("SSW_CHEM_BALANCE", 2);
        chem_balance = forced < 2 ? forced != 0 : (dc > 0 && (double)da > 1.25 * (double)dc);
    }
    return elapsed;
}

// sum_d get_rate(d) for every cell (Sweep::get_solver, src/sweep/mod.rs:616-620)
void Sweep::all_rates(double *dev_out) {
    dir_state_kernel<<<cdiv(N, 256), 256, 0, stream>>>(grid_view(), state_view(), 0, Dl, nullptr, cell_tmp.p);
    dir_state_kernel<<<cdiv(N, 256), 256, 0, stream>>>(grid_view(), state_view(), 2, Dl, nullptr, cell_tmp2.p);
    launched(2);
    const double frac = (double)Dl / (double)D;
    combine_rates_kernel<<<cdiv(N, 256), 256, 0, stream>>>(cell_tmp.p, cell_tmp2.p, src.p, frac, N, dev_out);
    launched();
    maybe_allreduce(dev_out, N);
}

void Sweep::read_field(int field, double *out, bool wait) {
    bind();
    const double *srcp = nullptr;
    // peer-mapped sharding: the chemistry state of a cell lives on its owner; the reading rank fetches the other
    // owners' slices (they are final once its own ssw_run_sweeps has returned, peer.cuh)
    auto owned = [&](const DevBuf<double> &buf, uint64_t off) {
        if (peers && out) peer_pull(off);
        return (const double *)buf.p;
    };
    switch (field) {
    case SSW_F_XHII: srcp = owned(x, pt.L.x); break;
    case SSW_F_TEMPERATURE: srcp = owned(T, pt.L.T); break;
    case SSW_F_TIMESTEP: srcp = owned(ts, pt.L.ts); break;
    case SSW_F_CHANGE_TIMESCALE: srcp = owned(tau, pt.L.tau); break;
    case SSW_F_PREVIOUS_RATE: srcp = owned(prev_rate, pt.L.prev); break;
    case SSW_F_DENSITY: srcp = rho.p; break;
    case SSW_F_SOURCE: srcp = src.p; break;
    case SSW_F_IONIZATION_TIME: srcp = owned(ion_time, pt.L.ion); break;
    case SSW_F_PHOTON_RATE:
        if (photon_valid && peers) {
            // every rank keeps the sum over ITS directions current in its arena: fold them in rank order
            peer_sync_point();
            peer_sum_kernel<<<cdiv(N, 256), 256, 0, stream>>>(pt, pt.L.photon, cell_tmp.p, N);
            launched();
            peer_sync_point();
            srcp = cell_tmp.p;
            break;
        }
        if (photon_valid) {
            // kept current by the sweeps themselves: the all-cells sweep leaves sum_d incoming in its
            // accumulators, partial sweeps patch the cells they touch
            if (P.world_size > 1) {
                CUDA_CHECK(cudaMemcpyAsync(cell_tmp.p, photon.p, sizeof(double) * N, cudaMemcpyDeviceToDevice, stream));
                maybe_allreduce(cell_tmp.p, N);
                srcp = cell_tmp.p;
            } else {
                srcp = photon.p;
            }
            break;
        }
        dir_state_kernel<<<cdiv(N, 256), 256, 0, stream>>>(grid_view(), state_view(), 0, Dl, nullptr, cell_tmp.p);
        launched();
        maybe_allreduce(cell_tmp.p, N);
        srcp = cell_tmp.p;
        break;
    case SSW_F_PHOTOIONIZATION_RATE:
    case SSW_F_HEATING_RATE:
    case SSW_F_RECOMBINATION_RATE:
    case SSW_F_COLLISIONAL_IONIZATION_RATE:
        if (peers) { peer_pull(pt.L.x); peer_pull(pt.L.T); }   // the outputs are evaluated for all cells on every rank
        all_rates(rate_act.p);
        chem_output_kernel<<<cdiv(N, 256), 256, 0, stream>>>(cell_view(), N, rate_act.p, P.scale_factor, field, cell_tmp.p);
        launched();
        srcp = cell_tmp.p;
        break;
    default: fail(SSW_E_INVALID, "unknown field %d", field);
    }
    CUDA_CHECK(cudaGetLastError());
    // out == NULL: a worker rank of a sharded job takes part in the field's collective but keeps no host copy
    if (out) CUDA_CHECK(cudaMemcpyAsync(out, srcp, sizeof(double) * N, cudaMemcpyDeviceToHost, stream));
    if (wait) CUDA_CHECK(cudaStreamSynchronize(stream));
}

}  // namespace ssw

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
struct ssw_handle {
    ssw::Sweep s;
};

#define SSW_TRY try {
#define SSW_CATCH                                        \
    }                                                    \
    catch (const ssw::Error &e) {                        \
        ssw::g_last_error = e.what();                    \
        return e.code;                                   \
    }                                                    \
    catch (const std::bad_alloc &) {                     \
        ssw::g_last_error = "host out of memory";        \
        return SSW_E_NOMEM;                              \
    }                                                    \
    catch (const std::exception &e) {                    \
        ssw::g_last_error = e.what();                    \
        return SSW_E_INVALID;                            \
    }                                                    \
    return SSW_OK;

#define REQUIRE_HANDLE(h) \
    if (!(h)) ssw::fail(SSW_E_INVALID, "null handle")

extern "C" {

const char *ssw_last_error(void) { return ssw::g_last_error.c_str(); }
int32_t ssw_abi_version(void) { return SSW_ABI_VERSION; }

int ssw_create(const ssw_params *params, const ssw_grid *grid, const double *density,
               const double *xhii, const double *temperature, const double *source,
               ssw_handle **out) {
    SSW_TRY
    if (!out) ssw::fail(SSW_E_INVALID, "null out pointer");
    *out = nullptr;
    std::unique_ptr<ssw_handle> h(new ssw_handle());
    h->s.create(params, grid, density, xhii, temperature, source);
    *out = h.release();
    SSW_CATCH
}

void ssw_destroy(ssw_handle *h) {
    if (!h) return;
    cudaSetDevice(h->s.device);
    if (h->s.stream) cudaStreamSynchronize(h->s.stream);
    delete h;
}

int ssw_set_allreduce(ssw_handle *h, ssw_allreduce_fn fn, void *ctx) {
    SSW_TRY
    REQUIRE_HANDLE(h);
    h->s.allreduce = fn;
    h->s.allreduce_ctx = ctx;
    SSW_CATCH
}

int ssw_set_directions(ssw_handle *h, const double *dirs_xyz) {
    SSW_TRY
    REQUIRE_HANDLE(h);
    h->s.set_directions(dirs_xyz);
    SSW_CATCH
}

int ssw_set_cell_positions(ssw_handle *h, const double *xyz) {
    SSW_TRY
    REQUIRE_HANDLE(h);
    h->s.set_positions(xyz);
    SSW_CATCH
}

const char *ssw_patch_note(ssw_handle *h) { return h ? h->s.patch_note.c_str() : ""; }

int ssw_set_collectives(ssw_handle *h, ssw_collective_fn fn, void *ctx) {
    SSW_TRY
    REQUIRE_HANDLE(h);
    h->s.collective = fn;
    h->s.collective_ctx = ctx;
    SSW_CATCH
}

int ssw_peer_arena(ssw_handle *h, void **base, uint64_t *bytes) {
    SSW_TRY
    REQUIRE_HANDLE(h);
    if (base) *base = h->s.arena.p;
    if (bytes) *bytes = h->s.pt.L.bytes;
    SSW_CATCH
}

int ssw_peer_export(ssw_handle *h, void *ipc_handle_out) {
    SSW_TRY
    REQUIRE_HANDLE(h);
    if (!ipc_handle_out) ssw::fail(SSW_E_INVALID, "null out pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == SSW_PEER_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
    CUDA_CHECK(cudaSetDevice(h->s.device));
    cudaIpcMemHandle_t m;
    CUDA_CHECK(cudaIpcGetMemHandle(&m, h->s.arena.p));
    std::memcpy(ipc_handle_out, &m, sizeof m);
    SSW_CATCH
}

int ssw_peer_attach_ipc(ssw_handle *h, const void *ipc_handles) {
    SSW_TRY
    REQUIRE_HANDLE(h);
    auto &s = h->s;
    if (!ipc_handles) ssw::fail(SSW_E_INVALID, "null handles");
    CUDA_CHECK(cudaSetDevice(s.device));
    std::vector<void *> bases(s.P.world_size, nullptr);
    for (int r = 0; r < s.P.world_size; ++r) {
        if (r == s.P.rank) { bases[r] = s.arena.p; continue; }
        cudaIpcMemHandle_t m;
        std::memcpy(&m, static_cast<const unsigned char *>(ipc_handles) + (size_t)r * SSW_PEER_HANDLE_BYTES, sizeof m);
        void *ptr = nullptr;
        CUDA_CHECK(cudaIpcOpenMemHandle(&ptr, m, cudaIpcMemLazyEnablePeerAccess));
        s.ipc_opened.push_back(ptr);
        bases[r] = ptr;
    }
    s.peer_attach(bases.data());
    SSW_CATCH
}

int ssw_peer_attach(ssw_handle *h, void *const *arena_bases) {
    SSW_TRY
    REQUIRE_HANDLE(h);
    auto &s = h->s;
    if (!arena_bases) ssw::fail(SSW_E_INVALID, "null bases");
    CUDA_CHECK(cudaSetDevice(s.device));
    // handles of one process on different devices: map the peers' memory
    for (int r = 0; r < s.P.world_size; ++r) {
        if (r == s.P.rank || !arena_bases[r]) continue;
        cudaPointerAttributes a;
        CUDA_CHECK(cudaPointerGetAttributes(&a, arena_bases[r]));
        if (a.device != s.device) {
            const cudaError_t e = cudaDeviceEnablePeerAccess(a.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CUDA_CHECK(e);
            (void)cudaGetLastError();
        }
    }
    s.peer_attach(arena_bases);
    SSW_CATCH
}

int ssw_run_sweeps(ssw_handle *h, double *time_elapsed_s) {
    SSW_TRY
    REQUIRE_HANDLE(h);
    const double e = h->s.run_sweeps();
    if (time_elapsed_s) *time_elapsed_s = e;
    SSW_CATCH
}

int ssw_set_inputs(ssw_handle *h, const double *density, const double *source) {
    SSW_TRY
    REQUIRE_HANDLE(h);
    auto &s = h->s;
    s.bind();
    if (density) s.rho.upload(density, s.N, s.stream);
    if (source) s.src.upload(source, s.N, s.stream);
    if (density) {
        // the absorption factor needs x: on the owner only under peer-mapped sharding, which then tells the others
        if (s.peers) ssw::peer_pull_kernel<<<ssw::cdiv(s.N, 256), 256, 0, s.stream>>>(s.pt, s.pt.L.x);
        ssw::attenuation_kernel<<<ssw::cdiv(s.N, 256), 256, 0, s.stream>>>(s.cell_view(), s.N);
        s.launched(2);
    }
    CUDA_CHECK(cudaStreamSynchronize(s.stream));
    SSW_CATCH
}

int ssw_read(ssw_handle *h, ssw_field field, double *out) {
    SSW_TRY
    REQUIRE_HANDLE(h);
    if (!out && h->s.P.world_size <= 1) ssw::fail(SSW_E_INVALID, "null out pointer");
    h->s.read_field((int)field, out);
    h->s.resolve_timers();
    SSW_CATCH
}

int ssw_read_begin(ssw_handle *h, ssw_field field, double *out) {
    SSW_TRY
    REQUIRE_HANDLE(h);
    if (!out && h->s.P.world_size <= 1) ssw::fail(SSW_E_INVALID, "null out pointer");
    h->s.read_field((int)field, out, /*wait=*/false);
    SSW_CATCH
}

int ssw_sync(ssw_handle *h) {
    SSW_TRY
    REQUIRE_HANDLE(h);
    CUDA_CHECK(cudaSetDevice(h->s.device));
    h->s.resolve_timers();   // synchronises the stream
    SSW_CATCH
}

int ssw_time_series_compute(ssw_handle *h, const double *mass, int32_t with_rates, ssw_time_series *out) {
    SSW_TRY
    REQUIRE_HANDLE(h);
    if (!out) ssw::fail(SSW_E_INVALID, "null out pointer");
    auto &s = h->s;
    s.bind();
    const uint32_t nb = ssw::cdiv(s.N, 256);
    auto &partial = s.series_partial;
    auto &sums = s.series_sums;
    auto &mass_dev = s.series_mass;
    partial.ensure((size_t)nb * ssw::kSeriesSums);   // allocated once, on the first call
    sums.ensure(ssw::kSeriesSums);
    if (mass) { mass_dev.ensure(s.N); mass_dev.upload(mass, s.N, s.stream); }
    if (s.peers) { s.peer_pull(s.pt.L.x); s.peer_pull(s.pt.L.T); }
    const double *gamma = nullptr;
    if (with_rates) {   // PhotoionizationRate of every cell, as sweep_optional_output_system computes it
        s.all_rates(s.rate_act.p);
        ssw::chem_output_kernel<<<nb, 256, 0, s.stream>>>(s.cell_view(), s.N, s.rate_act.p, s.P.scale_factor,
                                                         SSW_F_PHOTOIONIZATION_RATE, s.cell_tmp.p);
        s.launched();
        gamma = s.cell_tmp.p;
    }
    ssw::time_series_partial_kernel<<<nb, 256, 0, s.stream>>>(s.cell_view(), s.N, mass ? mass_dev.p : nullptr, gamma, partial.p);
    ssw::time_series_final_kernel<<<1, 256, 0, s.stream>>>(partial.p, nb, sums.p);
    s.launched(2);
    CUDA_CHECK(cudaGetLastError());
    double v[ssw::kSeriesSums];
    CUDA_CHECK(cudaMemcpyAsync(v, sums.p, sizeof v, cudaMemcpyDeviceToHost, s.stream));
    CUDA_CHECK(cudaStreamSynchronize(s.stream));
    const double nan = std::numeric_limits<double>::quiet_NaN();
    out->hydrogen_ionization_mass_average = v[0] / v[1];      // time_series.rs:81-83
    out->hydrogen_ionization_volume_average = v[2] / v[3];    // :91-97
    out->temperature_mass_average = v[4] / v[1];              // :106-112
    out->temperature_volume_average = v[5] / v[3];            // :120-125
    out->photoionization_rate_volume_average = with_rates ? v[6] / v[3] : nan;            // :134-139
    out->weighted_photoionization_rate_volume_average = with_rates ? v[7] / v[3] : nan;   // :148-153
    out->total_mass = v[1];
    out->total_volume = v[3];
    s.resolve_timers();
    SSW_CATCH
}

int ssw_read_levels(ssw_handle *h, uint8_t *out) {
    SSW_TRY
    REQUIRE_HANDLE(h);
    auto &s = h->s;
    CUDA_CHECK(cudaSetDevice(s.device));
    CUDA_CHECK(cudaMemcpyAsync(out, s.level.p, s.N, cudaMemcpyDeviceToHost, s.stream));
    CUDA_CHECK(cudaStreamSynchronize(s.stream));
    SSW_CATCH
}

int ssw_read_chem_attempts(ssw_handle *h, uint16_t *out) {
    SSW_TRY
    REQUIRE_HANDLE(h);
    if (!out) ssw::fail(SSW_E_INVALID, "null output");
    auto &s = h->s;
    CUDA_CHECK(cudaSetDevice(s.device));
    CUDA_CHECK(cudaMemcpyAsync(out, s.last_attempts.p, sizeof(uint16_t) * (size_t)s.N, cudaMemcpyDeviceToHost, s.stream));
    CUDA_CHECK(cudaStreamSynchronize(s.stream));
    SSW_CATCH
}

int ssw_level_counts(ssw_handle *h, uint64_t *out) {
    SSW_TRY
    REQUIRE_HANDLE(h);
    for (int l = 0; l < h->s.P.n_levels; ++l) out[l] = h->s.count_at_least(l);
    SSW_CATCH
}

int ssw_lowest_allowed_level(ssw_handle *h, int32_t *out) {
    SSW_TRY
    REQUIRE_HANDLE(h);
    *out = h->s.lowest_allowed;
    SSW_CATCH
}

int ssw_single_sweep(ssw_handle *h, int32_t level) {
    SSW_TRY
    REQUIRE_HANDLE(h);
    auto &s = h->s;
    if (level < 0 || level >= s.P.n_levels) ssw::fail(SSW_E_INVALID, "level out of range");
    s.bind();
    s.single_sweep(level);
    ssw::ChemStats cs;
    CUDA_CHECK(cudaMemcpyAsync(&cs, s.chem_stats.p, sizeof cs, cudaMemcpyDeviceToHost, s.stream));
    s.resolve_timers();
    s.stat[SSW_STAT_CHEM_CELLS] = cs.cells;
    s.stat[SSW_STAT_CHEM_FAILURES] = cs.failures;
    s.stat[SSW_STAT_CHEM_ATTEMPTS] = cs.attempts;
    s.stat[SSW_STAT_CHEM_MAX_DEPTH] = cs.max_depth;
    SSW_CATCH
}

int ssw_set_levels(ssw_handle *h, const uint8_t *levels) {
    SSW_TRY
    REQUIRE_HANDLE(h);
    auto &s = h->s;
    for (uint32_t c = 0; c < s.N; ++c)
        if (levels[c] >= s.P.n_levels) ssw::fail(SSW_E_INVALID, "level %d of cell %u out of range", (int)levels[c], c);
    s.bind();
    s.level.upload(levels, s.N, s.stream);
    s.refresh_histogram();
    SSW_CATCH
}

int ssw_set_change_timescale(ssw_handle *h, const double *tau) {
    SSW_TRY
    REQUIRE_HANDLE(h);
    auto &s = h->s;
    CUDA_CHECK(cudaSetDevice(s.device));
    s.tau.upload(tau, s.N, s.stream);
    CUDA_CHECK(cudaStreamSynchronize(s.stream));
    SSW_CATCH
}

int ssw_update_timestep_levels(ssw_handle *h) {
    SSW_TRY
    REQUIRE_HANDLE(h);
    h->s.bind();
    h->s.update_timestep_levels();
    h->s.resolve_timers();
    SSW_CATCH
}

int ssw_read_dir_state(ssw_handle *h, int32_t which, double *out) {
    SSW_TRY
    REQUIRE_HANDLE(h);
    auto &s = h->s;
    if (which < 0 || which > 2) ssw::fail(SSW_E_INVALID, "which must be 0, 1 or 2");
    s.bind();
    ssw::DevBuf<double> tmp;
    tmp.alloc((size_t)s.N * s.Dl);
    ssw::dir_state_kernel<<<ssw::cdiv(s.N, 256), 256, 0, s.stream>>>(s.grid_view(), s.state_view(), which, s.Dl, tmp.p, nullptr);
    s.launched();
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaMemcpyAsync(out, tmp.p, sizeof(double) * (size_t)s.N * s.Dl, cudaMemcpyDeviceToHost, s.stream));
    CUDA_CHECK(cudaStreamSynchronize(s.stream));
    SSW_CATCH
}

int ssw_read_wavefront_levels(ssw_handle *h, int32_t level, int32_t dir, int32_t *out) {
    SSW_TRY
    REQUIRE_HANDLE(h);
    auto &s = h->s;
    if (level < 0 || level >= s.P.n_levels) ssw::fail(SSW_E_INVALID, "level out of range");
    if (dir < s.d0 || dir >= s.d0 + s.Dl) ssw::fail(SSW_E_INVALID, "direction %d is not in this rank's shard [%d, %d)", dir, s.d0, s.d0 + s.Dl);
    s.bind();
    ssw::Schedule tmp;
    tmp.n_act = (uint32_t)s.count_at_least(level);
    s.wlevel.ensure(s.N);
    CUDA_CHECK(cudaMemsetAsync(s.wlevel.p, 0xff, sizeof(int32_t) * s.N, s.stream));
    if (tmp.n_act > 0) {
        if (tmp.n_act != s.N) s.build_active_list(tmp, level);
        s.build_schedule(tmp, level, /*solve=*/false, dir - s.d0, 1, s.wlevel.p);
    }
    CUDA_CHECK(cudaMemcpyAsync(out, s.wlevel.p, sizeof(int32_t) * s.N, cudaMemcpyDeviceToHost, s.stream));
    CUDA_CHECK(cudaStreamSynchronize(s.stream));
    SSW_CATCH
}

int ssw_get_stat(ssw_handle *h, ssw_stat which, uint64_t *out) {
    SSW_TRY
    REQUIRE_HANDLE(h);
    if ((int)which < 0 || (int)which >= 16) ssw::fail(SSW_E_INVALID, "unknown stat");
    *out = h->s.stat[which];
    SSW_CATCH
}

int ssw_get_timings(ssw_handle *h, ssw_timings *out) {
    SSW_TRY
    REQUIRE_HANDLE(h);
    *out = h->s.timings;
    SSW_CATCH
}

int ssw_reset_timings(ssw_handle *h) {
    SSW_TRY
    REQUIRE_HANDLE(h);
    std::memset(&h->s.timings, 0, sizeof(ssw_timings));
    SSW_CATCH
}

int ssw_set_timing_level(ssw_handle *h, int32_t level) {
    SSW_TRY
    REQUIRE_HANDLE(h);
    if (level < 0 || level > 1) ssw::fail(SSW_E_INVALID, "timing level %d (0 or 1)", level);
    h->s.timing_level = level;
    SSW_CATCH
}

int ssw_direction_shard(int32_t n_dirs, int32_t world_size, int32_t rank, int32_t *begin, int32_t *end) {
    if (world_size < 1 || rank < 0 || rank >= world_size || n_dirs < 0 || !begin || !end) return SSW_E_INVALID;
    *begin = (int32_t)((int64_t)n_dirs * rank / world_size);
    *end = (int32_t)((int64_t)n_dirs * (rank + 1) / world_size);
    return SSW_OK;
}

int32_t ssw_patch_lattice(const double *xyz, uint64_t n_cells, int32_t target_cells, uint32_t *patch_of) {
    try {
        if (!xyz || !patch_of || n_cells < 1 || n_cells > 0xfffffff0ull) return SSW_E_INVALID;
        std::vector<uint32_t> pof, poff, pcl;
        std::vector<uint16_t> lidx;
        uint32_t mx = 0;
        std::string why;
        uint32_t dims[3];
        const uint32_t Pn = ssw::build_patch_lattice(xyz, (uint32_t)n_cells, (double)target_cells, pof, poff, pcl, lidx, mx, why, dims);
        if (!Pn) { ssw::g_last_error = why; return 0; }
        std::copy(pof.begin(), pof.end(), patch_of);
        return (int32_t)Pn;
    } catch (const std::exception &e) {
        ssw::g_last_error = e.what();
        return SSW_E_INVALID;
    }
}

int32_t ssw_patch_levels(const uint32_t *upwind, int32_t n_groups, int32_t n_patches, uint32_t *level_out) {
    if (!upwind || !level_out || n_groups < 1 || n_patches < 1) return SSW_E_INVALID;
    std::vector<uint32_t> lvl, ndep;
    uint32_t max_level = 0;
    if (!ssw::level_macro_tiles(upwind, (uint32_t)n_groups, (uint32_t)n_patches, lvl, ndep, max_level)) return SSW_E_DEADLOCK;
    std::copy(lvl.begin(), lvl.end(), level_out);
    return (int32_t)max_level + 1;
}

int32_t ssw_direction_groups(const double *dirs_xyz, int32_t n_dirs, int32_t max_per_group, int32_t *group_of) {
    if (!dirs_xyz || !group_of || n_dirs < 1 || n_dirs > ssw::kMaxDirs || max_per_group < 1) return SSW_E_INVALID;
    std::vector<uint16_t> grp, rank;
    std::vector<uint32_t> kd;
    const uint32_t G = ssw::make_direction_groups(dirs_xyz, (uint32_t)n_dirs, (uint32_t)max_per_group, grp, rank, kd);
    for (int32_t d = 0; d < n_dirs; ++d) group_of[d] = grp[d];
    return (int32_t)G;
}

int32_t ssw_level_from_timesteps(int32_t max_num_levels, double max_timestep, double desired) {
    return ssw::level_rule(max_num_levels, max_timestep, desired);
}

int32_t ssw_levels_in_sweep_order(int32_t max_num_levels, int32_t lowest_allowed, int32_t *out, int32_t cap) {
    // TimestepState::iter_levels_in_sweep_order + lowest_active_from_iteration
    // (src/sweep/timestep_state.rs:22-27, 66-75)
    const int num = max_num_levels - lowest_allowed;
    if (num < 1 || num > 31) return 0;
    const uint32_t count = 1u << (num - 1);
    int n = 0;
    for (uint32_t i = 0; i < count; ++i) {
        int first_bit = num - 1;
        for (int b = 0; b < 32; ++b)
            if (i & (1u << b)) { first_bit = b; break; }
        if (n < cap) out[n] = lowest_allowed + (num - 1 - first_bit);
        ++n;
    }
    return n;
}

int ssw_chemistry_batch(int32_t device_id, uint64_t n, double *xhii, double *temperature,
                        const double *density, const double *volume, const double *length,
                        const double *rate, const double *timestep, double scale_factor,
                        double safety_factor, int32_t prevent_cooling, double *timescale_out,
                        int32_t *process_out, int32_t *depth_out, uint64_t *attempts_out) {
    SSW_TRY
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev < 1)
        ssw::fail(SSW_E_CUDA, "no CUDA device: this library has no CPU fallback");
    CUDA_CHECK(cudaSetDevice(device_id));
    ssw::DevBuf<double> dx, dT, drho, dvol, dlen, drate, ddt, dts;
    ssw::DevBuf<int> dproc, ddepth;
    ssw::DevBuf<unsigned long long> datt;
    cudaStream_t st = nullptr;
    dx.alloc(n); dT.alloc(n); drho.alloc(n); dvol.alloc(n); dlen.alloc(n); drate.alloc(n); ddt.alloc(n);
    dts.alloc(n); dproc.alloc(n); ddepth.alloc(n); datt.alloc(n);
    dx.upload(xhii, n, st); dT.upload(temperature, n, st); drho.upload(density, n, st);
    dvol.upload(volume, n, st); dlen.upload(length, n, st); drate.upload(rate, n, st); ddt.upload(timestep, n, st);
    ssw::chemistry_batch_kernel<<<ssw::cdiv(n, 128), 128, 0, st>>>(n, dx.p, dT.p, drho.p, dvol.p, dlen.p, drate.p, ddt.p,
                                                                   scale_factor, safety_factor, prevent_cooling,
                                                                   dts.p, dproc.p, ddepth.p, datt.p);
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaMemcpy(xhii, dx.p, n * sizeof(double), cudaMemcpyDeviceToHost));
    CUDA_CHECK(cudaMemcpy(temperature, dT.p, n * sizeof(double), cudaMemcpyDeviceToHost));
    if (timescale_out) CUDA_CHECK(cudaMemcpy(timescale_out, dts.p, n * sizeof(double), cudaMemcpyDeviceToHost));
    if (process_out) CUDA_CHECK(cudaMemcpy(process_out, dproc.p, n * sizeof(int), cudaMemcpyDeviceToHost));
    if (depth_out) CUDA_CHECK(cudaMemcpy(depth_out, ddepth.p, n * sizeof(int), cudaMemcpyDeviceToHost));
    if (attempts_out) CUDA_CHECK(cudaMemcpy(attempts_out, datt.p, n * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    SSW_CATCH
}

}  // extern "C"
