/*
 * subsweep_b200.h -- C ABI of libsubsweep_b200.so: a B200-native (sm_100a) replacement for
 * subsweep's per-timestep hot path, the directional upwind sweep over the Voronoi grid plus
 * the per-cell hydrogen ionization/temperature chemistry.
 *
 * This is the boundary a thin Rust FFI crate binds (INTEGRATION.md shows the crate): plain
 * pointers and sizes, no C++ or torch types.  Every entry point cites the reference interface
 * it replaces (paths relative to the reference repository root).
 *
 * Conventions
 *   - all floating point is f64 in SI base units, exactly like the reference's diman quantities
 *     (src/units/mod.rs); indices are 0-based; cell order = ParticleId.index order
 *     (src/sweep/active_list.rs:25-35).
 *   - host arrays passed in are borrowed for the duration of the call only.
 *   - every function returns 0 on success or a negative SSW_E_* code; ssw_last_error() gives
 *     the message (thread-local).  The handle is not thread-safe: one caller thread, like the
 *     reference's NonSend solver resource (src/sweep/mod.rs:139).
 *   - there is no CPU fallback: without a CUDA device of compute capability 10.x every call
 *     that needs the device fails with SSW_E_CUDA.
 */
#ifndef SUBSWEEP_B200_H
#define SUBSWEEP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSW_ABI_VERSION 1

enum {
    SSW_OK = 0,
    SSW_E_INVALID = -1,   /* bad argument / inconsistent grid                         */
    SSW_E_CUDA = -2,      /* CUDA runtime error, no device, wrong architecture        */
    SSW_E_DEADLOCK = -3,  /* dependency cycle among active Local faces: the reference */
                          /* would spin forever in Sweep::solve (src/sweep/mod.rs:291-300) */
    SSW_E_NOMEM = -4,
    SSW_E_COMM = -5       /* the all-reduce hook failed                               */
};

/* ParticleType of a face's neighbour, single-process subset of src/sweep/grid/cell.rs:16-23.
 * Remote / RemotePeriodic do not exist here: the MPI domain split is replaced by direction
 * sharding, every process holds the whole grid. */
enum { SSW_FACE_LOCAL = 0, SSW_FACE_BOUNDARY = 1, SSW_FACE_LOCAL_PERIODIC = 2 };

/* SweepParameters (src/sweep/parameters.rs:8-46) + HydrogenOnly (src/chemistry/hydrogen_only/
 * mod.rs:41-46) + what Sweep::new takes (src/sweep/mod.rs:194-232, call site :675-691). */
typedef struct ssw_params {
    int32_t n_dirs;                 /* D = sweep.directions (count)                              */
    const double *dirs_xyz;         /* D x 3; table values as in direction/healpix.rs (NOT         */
                                    /* re-normalised) or normalised explicit lists (mod.rs:97-109) */
    int32_t n_levels;               /* sweep.num_timestep_levels                                   */
    double max_timestep_s;          /* sweep.max_timestep                                          */
    double timestep_safety_factor;  /* sweep.timestep_safety_factor (default 0.1)                  */
    double chemistry_timestep_safety_factor; /* sweep.chemistry_timestep_safety_factor             */
    double significant_rate_threshold_per_s; /* sweep.significant_rate_threshold (default 0)       */
    int32_t prevent_cooling;        /* sweep.prevent_cooling (default true)                        */
    double scale_factor;            /* Cosmology::scale_factor() (src/sweep/mod.rs:687)            */
    int32_t check_deadlock;         /* sweep.check_deadlock                                        */
    int32_t device_id;              /* CUDA device ordinal of this process                         */
    int32_t rank, world_size;       /* direction sharding: this process sweeps the directions      */
                                    /* ssw_direction_shard(D, world_size, rank); 0,1 = all         */
    uint32_t flags;                 /* SSW_FLAG_*                                                  */
} ssw_params;

enum {
    SSW_FLAG_NO_SCHEDULE_CACHE = 1u << 0, /* rebuild wavefront level sets every single sweep      */
    SSW_FLAG_NO_COMPILED_PATH = 1u << 1,  /* replay cached level sets from the generic task list   */
    SSW_FLAG_NO_PATCH_PATH = 1u << 2,     /* keep the level-barrier stream even when cell positions */
                                          /* are known (ssw_set_cell_positions)                     */
    SSW_FLAG_SHARED_DEVICE = 1u << 3      /* several handles of ONE process take turns on one device */
                                          /* (tests that emulate ranks with threads): re-bind the    */
                                          /* per-process direction table after every hook call.  The */
                                          /* handles must never run concurrently.                    */
};

/* Flat (CSR) form of the per-particle `Cell` component (src/sweep/grid/cell.rs:92-133):
 * cell c owns faces [face_offsets[c], face_offsets[c+1]) in the order of Cell::neighbours. */
typedef struct ssw_grid {
    uint64_t n_cells;
    const uint64_t *face_offsets;  /* N+1                                                        */
    const double *face_area;       /* F, Face::area                                               */
    const double *face_normal;     /* F x 3, Face::normal (outward unit vector)                   */
    const int32_t *face_neighbour; /* F, neighbour's ParticleId.index; -1 for Boundary            */
    const uint8_t *face_kind;      /* F, SSW_FACE_*                                               */
    const double *cell_size;       /* N, Cell::size                                               */
    const double *cell_volume;     /* N, Cell::volume                                             */
} ssw_grid;

typedef struct ssw_handle ssw_handle;

/* Fields readable per cell.  The first block are the per-particle components the reference
 * writes back after every step (src/sweep/mod.rs:718-738, src/components.rs:14-83); the second
 * block are the optional chemistry outputs (src/sweep/chemistry_output.rs:25-55). */
typedef enum ssw_field {
    SSW_F_XHII = 0,               /* ionized_hydrogen_fraction                                   */
    SSW_F_TEMPERATURE = 1,        /* temperature [K]                                             */
    SSW_F_TIMESTEP = 2,           /* timestep: chemistry's recommended timescale [s]             */
    SSW_F_PHOTON_RATE = 3,        /* photon_rate = sum_d incoming_total_rate[d] [1/s]            */
    SSW_F_CHANGE_TIMESCALE = 4,   /* Site::change_timescale [s]                                  */
    SSW_F_PHOTOIONIZATION_RATE = 5,
    SSW_F_HEATING_RATE = 6,
    SSW_F_RECOMBINATION_RATE = 7,
    SSW_F_COLLISIONAL_IONIZATION_RATE = 8,
    SSW_F_PREVIOUS_RATE = 9,      /* Site::previous_incoming_total_rate                          */
    SSW_F_DENSITY = 10,
    SSW_F_SOURCE = 11,
    SSW_F_IONIZATION_TIME = 12    /* ionization_time (src/sweep/mod.rs:731-738); +inf = not yet  */
} ssw_field;

/* all-reduce hook for direction sharding: sum `n` doubles in place over all ranks.  `buf` is a
 * DEVICE pointer on params.device_id; `cuda_stream` is the cudaStream_t the library works on.
 * The call is stream-ordered: the values in `buf` are produced by work already queued on that
 * stream, and the hook must leave the reduced result ordered on the same stream (e.g.
 * ncclAllReduce(buf, buf, n, ncclDouble, ncclSum, comm, cuda_stream)); no host synchronisation
 * on either side.  Replaces the MPI flux messages of src/sweep/communicator.rs:59-95. */
typedef int (*ssw_allreduce_fn)(void *ctx, double *buf, uint64_t n, void *cuda_stream);

/* Optional second hook: reduce-scatter and all-gather, in place, with NCCL's in-place layout.  `buf` holds
 * world_size chunks of n_per_rank doubles (DEVICE pointer, stream-ordered like ssw_allreduce_fn):
 *   SSW_COLL_REDUCE_SCATTER  chunk `rank` of buf <- sum over ranks of their chunk `rank`
 *                            (ncclReduceScatter(buf, buf + rank * n, n, ncclDouble, ncclSum, ...))
 *   SSW_COLL_ALL_GATHER      every chunk r of buf <- chunk r of rank r's buf
 *                            (ncclAllGather(buf + rank * n, buf, n, ncclDouble, ...))
 * With it the chemistry after an all-cells sweep is sliced by cells instead of replicated on every rank
 * (reduce-scatter of the rates, update of the own slice, all-gather of x / T / timescales); without it the
 * all-reduce hook alone is used.  Sweeps over partial active sets always use the all-reduce. */
enum { SSW_COLL_REDUCE_SCATTER = 1, SSW_COLL_ALL_GATHER = 2 };
typedef int (*ssw_collective_fn)(void *ctx, int op, double *buf, uint64_t n_per_rank, void *cuda_stream);

/* -- life cycle ------------------------------------------------------------------------- */

/* Sweep::new (src/sweep/mod.rs:194-232) + init_sweep_system (:634-692).  Copies the grid and
 * the four per-cell components to the device; all cells start at level n_levels-1 (:206). */
int ssw_create(const ssw_params *params, const ssw_grid *grid, const double *density,
               const double *xhii, const double *temperature, const double *source,
               ssw_handle **out);
/* Drop (src/sweep/communicator.rs:116-125). */
void ssw_destroy(ssw_handle *h);
int ssw_set_allreduce(ssw_handle *h, ssw_allreduce_fn fn, void *ctx);
int ssw_set_collectives(ssw_handle *h, ssw_collective_fn fn, void *ctx);
/* Optional, once, before the first ssw_run_sweeps: the `Position` component of every cell
 * (src/components.rs, N x 3, cell order).  The solver itself never needs positions
 * (Sweep::new does not take them, src/sweep/mod.rs:194-232); they let the library group cells into
 * spatial patches and run the all-cells sweep as a dataflow of (patch, direction group) macro-tiles
 * without device-wide barriers.  Results do not depend on whether positions
 * were given beyond round-off in the per-cell rate sums.  Grids whose patches depend on each other
 * cyclically keep the level-barrier form; ssw_patch_note() says why (empty string: patch form in use
 * or not tried). */
int ssw_set_cell_positions(ssw_handle *h, const double *xyz /* N x 3 */);
/* rotate_directions_system (src/sweep/direction/mod.rs:158-174, `sweep.rotate_directions`): replace the direction set
 * by another one of the same size (the host rotates the bins; the Python mirror does it with the reference's
 * axis-angle construction).  New direction i continues with the outgoing rates of the old direction it is best
 * aligned with (largest dot product).  The reference's `remap` (:190-205) intends that kernel but assigns instead of
 * accumulating, which wipes all but the last direction's state; this library implements the intended remap and does NOT
 * reproduce the bug.  Level sets depend on the directions: after the first call nothing direction-dependent is
 * compiled or cached any more (every sweep peels its level sets while it solves).  Single rank only. */
int ssw_set_directions(ssw_handle *h, const double *dirs_xyz /* D x 3 */);
const char *ssw_patch_note(ssw_handle *h);

/* -- direction sharding without hooks: peer-mapped exchange over NVLink / NVSwitch ---------------------- */

/* The preferred way to run W ranks on one box (one ssw_handle per GPU).  Every handle owns an ARENA (one device
 * allocation) with its per-cell state and receive buffers; once the arenas of all ranks are attached the library moves
 * the per-cell rate partials, absorption factors and timestep levels itself, fused into its kernels as remote stores /
 * loads over NVLink with flag words for ordering -- no hook, no collective library, nothing of the host on the path
 * (DESIGN.md section 7).  This replaces SweepCommunicator (src/sweep/communicator.rs:59-95) on the single box.
 *
 *   several processes (one per GPU):  ssw_peer_export on every rank -> the host all-gathers the W handles of
 *       SSW_PEER_HANDLE_BYTES bytes each (MPI_Allgather in the reference's world) -> ssw_peer_attach_ipc;
 *   one process driving several handles:  ssw_peer_arena on every handle -> ssw_peer_attach with the W base pointers.
 *
 * Attach once, after ssw_create and before the first sweep, on every rank.  With peers attached every rank must make
 * the same sequence of ssw_run_sweeps / ssw_single_sweep / ssw_update_timestep_levels calls, and of ssw_read calls
 * for SSW_F_PHOTON_RATE and the optional chemistry outputs (they sum over all directions); the other fields may be
 * read by any single rank at any time between steps.  Cell state is held by the rank that owns the cell
 * (contiguous slices of ceil(N / W) cells); ssw_read assembles it on the reading rank. */
#define SSW_PEER_HANDLE_BYTES 64
int ssw_peer_arena(ssw_handle *h, void **base, uint64_t *bytes);
int ssw_peer_export(ssw_handle *h, void *ipc_handle_out /* SSW_PEER_HANDLE_BYTES */);
int ssw_peer_attach_ipc(ssw_handle *h, const void *ipc_handles /* world_size x SSW_PEER_HANDLE_BYTES, rank order */);
int ssw_peer_attach(ssw_handle *h, void *const *arena_bases /* world_size device pointers, rank order */);

/* -- the hot path ------------------------------------------------------------------------ */

/* Sweep::run_sweeps (src/sweep/mod.rs:258-272): all single sweeps of one full step in the
 * reference's level order, chemistry after each, then the timestep-level update.  Blocking.
 * *time_elapsed_s receives the value the reference returns (added to SimulationTime, :716-717). */
int ssw_run_sweeps(ssw_handle *h, double *time_elapsed_s);

/* Refresh per-cell inputs between steps (the `Source` / `Density` components,
 * src/sweep/mod.rs:637-644).  NULL leaves a field unchanged. */
int ssw_set_inputs(ssw_handle *h, const double *density, const double *source);

/* -- read-back (src/sweep/mod.rs:718-738, :612-632, :234-245) ---------------------------- */
/* In a direction-sharded job every rank holds the same cell state; the rank that owns the output (rank 0)
 * reads it back, the others may pass out = NULL: they then only take part in the collective a field needs
 * (PHOTON_RATE and the chemistry outputs sum over all directions) and copy nothing to the host. */
int ssw_read(ssw_handle *h, ssw_field field, double *out /* N, or NULL on worker ranks */);
/* The same without waiting: the copy into `out` (pinned host memory, or it degenerates to a blocking copy) is
 * queued on the library's stream; `out` is valid after the next ssw_sync / ssw_read / ssw_run_sweeps returns.
 * The write-back of run_sweep_system (five components, src/sweep/mod.rs:718-738) is five ssw_read_begin + one
 * ssw_sync. */
int ssw_read_begin(ssw_handle *h, ssw_field field, double *out);
int ssw_sync(ssw_handle *h);
int ssw_read_levels(ssw_handle *h, uint8_t *out /* N */);
/* Substep attempts of every cell's LAST chemistry update (perform_timestep_internal, hydrogen_only/mod.rs:394-441:
 * 1 = the first attempt was accepted; saturates at 65535; 0 = never updated).  Diagnostic: the substep histogram of an
 * ionization front; the library itself orders its chemistry launches by it.  Under
 * peer-mapped sharding a rank knows the counts of the cells it owns. */
int ssw_read_chem_attempts(ssw_handle *h, uint16_t *out /* N */);
int ssw_level_counts(ssw_handle *h, uint64_t *out /* n_levels, cumulative: #cells with level >= l */);
int ssw_lowest_allowed_level(ssw_handle *h, int32_t *out);

/* -- the step after the path: time series (src/sweep/time_series.rs:61-188) ------------------ */

/* compute_time_series_system (time_series.rs:61-155): mass- and volume-weighted averages over all
 * cells, reduced on the device from the resident state (no N-length read-back).  `mass` is the
 * per-particle Mass component (host array, N) or NULL for density * volume.  The two
 * photoionization-rate averages need the optional PhotoionizationRate output
 * (chemistry_output.rs:25-31) and are only evaluated when with_rates != 0 (else NaN). */
typedef struct ssw_time_series {
    double hydrogen_ionization_mass_average;
    double hydrogen_ionization_volume_average;
    double temperature_mass_average;                      /* K    */
    double temperature_volume_average;                    /* K    */
    double photoionization_rate_volume_average;           /* 1/s  */
    double weighted_photoionization_rate_volume_average;  /* 1/s  */
    double total_mass, total_volume;
} ssw_time_series;
int ssw_time_series_compute(ssw_handle *h, const double *mass /* N or NULL */, int32_t with_rates,
                            ssw_time_series *out);

/* -- pieces, exposed for parity tests and profiling --------------------------------------- */

/* Sweep::single_sweep (src/sweep/mod.rs:274-289) at `level`, including chemistry. */
int ssw_single_sweep(ssw_handle *h, int32_t level);
/* force ActiveList levels (src/sweep/active_list.rs:135-153) */
int ssw_set_levels(ssw_handle *h, const uint8_t *levels /* N */);
int ssw_set_change_timescale(ssw_handle *h, const double *tau /* N */);
/* Sweep::update_timestep_levels (src/sweep/mod.rs:576-589) with the current lowest allowed level */
int ssw_update_timestep_levels(ssw_handle *h);
/* per-direction state of this rank's directions, cell-major N x D_local like the reference's
 * Site vectors (src/sweep/site.rs:12-23): which = 0 incoming_total_rate, 1 outgoing_total_rate,
 * 2 periodic_source */
int ssw_read_dir_state(ssw_handle *h, int32_t which, double *out /* N x D_local */);
/* wavefront level of every cell for the active set of `level` and global direction `dir`
 * (0 = initial task; -1 = inactive); the level sets the kernels iterate over. */
int ssw_read_wavefront_levels(ssw_handle *h, int32_t level, int32_t dir, int32_t *out /* N */);

typedef enum ssw_stat {
    SSW_STAT_TASKS_SOLVED = 0,      /* cell-direction updates since create (this rank)          */
    SSW_STAT_SINGLE_SWEEPS = 1,
    SSW_STAT_CHEM_CELLS = 2,        /* chemistry cell updates                                   */
    SSW_STAT_CHEM_FAILURES = 3,     /* TimestepConvergenceFailed (hydrogen_only/mod.rs:431-440)  */
    SSW_STAT_SCHEDULE_BUILDS = 4,   /* wavefront level-set builds                               */
    SSW_STAT_SCHEDULE_REPLAYS = 5,
    SSW_STAT_KERNEL_LAUNCHES = 6,   /* kernels launched by this library since create            */
    SSW_STAT_WAVEFRONT_LEVELS = 7,  /* level count of the last single sweep                     */
    SSW_STAT_CHEM_ATTEMPTS = 8,     /* try_timestep_update calls                                 */
    SSW_STAT_CHEM_MAX_DEPTH = 9,
    SSW_STAT_PATCH_MACRO_TILES = 10, /* macro-tiles of the patch-ordered all-cells sweep (0: not in use) */
    SSW_STAT_PATCH_LEVELS = 11,      /* dependent macro-tile levels (vs SSW_STAT_WAVEFRONT_LEVELS)       */
    SSW_STAT_PATCH_PHASES = 12,      /* phases of the macro-tiles (1: the patch graph was acyclic)       */
    SSW_STAT_WALK_WINDOW = 13,       /* walk form of the all-cells sweep: slots of the shared-memory window (0: not in use) */
    SSW_STAT_WALK_NEAR_PERMILLE = 14 /* walk form: upwind entries read from the window, per thousand     */
} ssw_stat;
int ssw_get_stat(ssw_handle *h, ssw_stat which, uint64_t *out);

/* Device time (CUDA events, ms) accumulated per category since the last reset, named like the
 * reference's Performance timers (src/sweep/mod.rs:275,550,577). */
typedef struct ssw_timings {
    double sweep_ms;          /* sum of sweep_level_<L>                                         */
    double chemistry_ms;      /* "chemistry" (incl. rate reduction and all-reduce)              */
    double update_levels_ms;  /* "update levels"                                                */
    double schedule_ms;       /* level-set builds (part of sweep_level_<L> in the reference)    */
    double allreduce_ms;
    double sweep_kernel_ms;   /* the per-level sweep kernel(s) alone                            */
    uint64_t sweep_kernel_launches;
    uint64_t sweep_kernel_tasks;  /* cell-direction updates processed by those launches          */
    double sweep_level_ms[32];
    double step_ms;           /* whole ssw_run_sweeps calls, first to last CUDA event on the stream */
    uint64_t steps;
    double kernel_level_ms[32];      /* sweep kernel time per current timestep level               */
    uint64_t kernel_level_tasks[32]; /* cell-direction updates per current timestep level          */
    uint64_t kernel_level_launches[32];
} ssw_timings;
int ssw_get_timings(ssw_handle *h, ssw_timings *out);
int ssw_reset_timings(ssw_handle *h);
/* 0: time only whole steps and the sweep kernel of the all-cells sweeps (four CUDA event records per step: what a
 * production host wants); 1 (default, or SSW_TIMERS): every phase of every single sweep, the breakdown the reference's
 * `Performance` timers print (src/sweep/mod.rs:275-283, 550, 577).  The fine level records about ten events per single sweep. */
int ssw_set_timing_level(ssw_handle *h, int32_t level);

/* -- helpers -------------------------------------------------------------------------------- */

/* contiguous direction shard [begin, end) of rank `rank` out of `world_size` */
int ssw_direction_shard(int32_t n_dirs, int32_t world_size, int32_t rank, int32_t *begin,
                        int32_t *end);
/* host-side pieces of the patch-ordered sweep, exposed for tests; no device needed.
 * ssw_patch_lattice: patch index of every cell for boxes of about target_cells cells; returns the number of
 * patches (0: no lattice qualifies, < 0: error).  ssw_direction_groups: directions of one octant (sign pattern),
 * at most max_per_group per group; returns the number of groups. */
int32_t ssw_patch_lattice(const double *xyz /* N x 3 */, uint64_t n_cells, int32_t target_cells,
                          uint32_t *patch_of /* N */);
int32_t ssw_direction_groups(const double *dirs_xyz /* D x 3 */, int32_t n_dirs, int32_t max_per_group,
                             int32_t *group_of /* D */);
/* ssw_patch_levels: level of every macro-tile from the quotient graph; upwind[(g * P + p) * 64 ..] lists the upwind
 * patches of patch p for group g, terminated by 0xffffffff.  Returns the number of levels, SSW_E_DEADLOCK if a
 * group's graph has a cycle (the grid then keeps the level-barrier stream). */
int32_t ssw_patch_levels(const uint32_t *upwind, int32_t n_groups, int32_t n_patches, uint32_t *level_out /* G x P */);
/* TimestepLevel::from_max_timestep_and_desired_timestep (src/sweep/timestep_level.rs:27-36),
 * host-side scalar version of the device rule */
int32_t ssw_level_from_timesteps(int32_t max_num_levels, double max_timestep, double desired);
/* TimestepState::iter_levels_in_sweep_order (src/sweep/timestep_state.rs:22-27) */
int32_t ssw_levels_in_sweep_order(int32_t max_num_levels, int32_t lowest_allowed, int32_t *out,
                                  int32_t cap);
/* one chemistry update of n independent cells on the device, for parity tests of
 * HydrogenOnly::update_abundances (src/chemistry/hydrogen_only/mod.rs:90-119).  All arrays are
 * HOST arrays of length n; xhii / temperature are updated in place. */
int ssw_chemistry_batch(int32_t device_id, uint64_t n, double *xhii, double *temperature,
                        const double *density, const double *volume, const double *length,
                        const double *rate, const double *timestep, double scale_factor,
                        double safety_factor, int32_t prevent_cooling, double *timescale_out,
                        int32_t *process_out, int32_t *depth_out, uint64_t *attempts_out);

const char *ssw_last_error(void);
int32_t ssw_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SUBSWEEP_B200_H */
