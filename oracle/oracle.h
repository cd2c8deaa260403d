/*
 * oracle.h -- CPU restatement of subsweep's sweep + chemistry hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under subsweep_b200/ may include, link or
 * dlopen this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / the CPU baseline.
 *
 * PARITY UNPINNED: the reference (Rust nightly + MPI + HDF5) cannot be built in
 * this image and its own test-suite pins no fluxes, rates, xHII or T for this path
 * (SURVEY.md section 8c).  What the reference does pin -- the timestep-level rule,
 * the sweep order / warm-up, the derivative-consistency checks and the two
 * production-like chemistry inputs that must terminate -- is checked in
 * tests/test_oracle_*.py.  In lieu of reference vectors for the fluxes the restatement is
 * cross-checked by double entry: tests/test_oracle_double_entry.py types the sweep, the
 * chemistry substepper (recursively, as the reference writes it) and the whole of
 * run_sweeps incl. Rust's BinaryHeap a second time in Python, from the reference source,
 * and agrees with this file to 1e-10 ... 1e-12 (levels and attempt counts exactly).
 *
 * Every function cites the reference file:line it restates (paths relative to the
 * reference repository root).
 */
#ifndef SUBSWEEP_ORACLE_H
#define SUBSWEEP_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Face kinds of the flat grid (src/sweep/grid/cell.rs:16-23, single-rank subset). */
enum { ORC_FACE_LOCAL = 0, ORC_FACE_BOUNDARY = 1, ORC_FACE_LOCAL_PERIODIC = 2 };

/* How periodic_source is read by a task (DESIGN.md "periodic lag").
 * HEAP   : exactly the reference: read whatever has been scattered so far, which
 *          depends on the Rust BinaryHeap pop order (src/sweep/mod.rs:505-513, site.rs:53-56).
 * LAGGED : read the value periodic_source had when the single sweep started; this is
 *          what the reference does whenever the periodic donor is solved after its
 *          target (always, on fully active Cartesian grids) and is the order-independent
 *          definition the wavefront kernels implement. */
enum { ORC_PERIODIC_HEAP = 0, ORC_PERIODIC_LAGGED = 1 };

typedef struct {
    int32_t n_dirs;              /* D, total number of directions                      */
    const double *dirs_xyz;      /* D x 3, as the reference's tables (not re-normalised) */
    int32_t n_levels;            /* sweep.num_timestep_levels                          */
    double max_timestep;         /* s                                                  */
    double timestep_safety_factor;
    double chemistry_timestep_safety_factor;
    double significant_rate_threshold; /* 1/s                                          */
    int32_t prevent_cooling;
    double scale_factor;
    int32_t check_deadlock;
    int32_t periodic_mode;       /* ORC_PERIODIC_*                                     */
    int32_t dir_begin, dir_end;  /* this rank's direction shard [begin,end); 0,D = all  */
} orc_params;

typedef struct {
    uint64_t n_cells;
    const uint64_t *face_offsets;  /* N+1 */
    const double *face_area;       /* F   */
    const double *face_normal;     /* F x 3 */
    const int32_t *face_neighbour; /* F, local index, -1 = boundary */
    const uint8_t *face_kind;      /* F, ORC_FACE_* */
    const double *cell_size;       /* N */
    const double *cell_volume;     /* N */
} orc_grid;

typedef struct orc_sweep orc_sweep;

/* all-reduce hook used when the directions are sharded over ranks (DESIGN.md multi-GPU) */
typedef int (*orc_allreduce_fn)(void *ctx, double *buf, uint64_t n);

orc_sweep *orc_create(const orc_params *p, const orc_grid *g, const double *density,
                      const double *xhii, const double *temperature, const double *source);
void orc_destroy(orc_sweep *s);
void orc_set_allreduce(orc_sweep *s, orc_allreduce_fn fn, void *ctx);

/* Sweep::run_sweeps, src/sweep/mod.rs:258-272.  Returns elapsed time in s. */
double orc_run_sweeps(orc_sweep *s);

/* pieces, for tests */
void orc_single_sweep(orc_sweep *s, int level);          /* mod.rs:274-289 */
void orc_set_levels(orc_sweep *s, const uint8_t *levels);  /* test hook: force levels + rebuild bins */
void orc_set_change_timescale(orc_sweep *s, const double *tau);
void orc_update_timestep_levels(orc_sweep *s);           /* mod.rs:576-589 */
void orc_wavefront_levels(orc_sweep *s, int level, int dir, int32_t *out /*N, -1 inactive*/);

enum {
    ORC_F_XHII = 0, ORC_F_TEMPERATURE, ORC_F_TIMESTEP, ORC_F_PHOTON_RATE, ORC_F_CHANGE_TIMESCALE,
    ORC_F_PHOTOIONIZATION_RATE, ORC_F_HEATING_RATE, ORC_F_RECOMBINATION_RATE,
    ORC_F_COLLISIONAL_IONIZATION_RATE, ORC_F_PREVIOUS_RATE, ORC_F_DENSITY, ORC_F_SOURCE
};
int orc_read(orc_sweep *s, int field, double *out /*N*/);
void orc_read_levels(orc_sweep *s, uint8_t *out /*N*/);
void orc_level_counts(orc_sweep *s, uint64_t *out /*L, cumulative*/);
/* which: 0 incoming, 1 outgoing, 2 periodic_source; out is N x D_local, cell-major */
void orc_read_dir_state(orc_sweep *s, int which, double *out);
int orc_lowest_allowed_level(orc_sweep *s);
/* statistics of the last single sweep / all sweeps */
uint64_t orc_stat(orc_sweep *s, int which);
enum { ORC_STAT_TASKS_SOLVED = 0, ORC_STAT_NONLAGGED_PERIODIC_READS, ORC_STAT_CHEM_ATTEMPTS,
       ORC_STAT_CHEM_MAX_DEPTH, ORC_STAT_CHEM_FAILURES, ORC_STAT_SINGLE_SWEEPS, ORC_STAT_CHEM_CELLS };

/* ---- chemistry in isolation (src/chemistry/hydrogen_only/mod.rs) ---- */
typedef struct {
    double xhii, temperature, density, volume, length, rate, scale_factor;
    int32_t has_floor;
    double floor_temperature, floor_xhii;
} orc_solver;

typedef struct {
    double timescale;   /* returned Timescale.time */
    int32_t process;    /* 0 temperature, 1 ionization fraction */
    int32_t failed;     /* TimestepConvergenceFailed */
    uint64_t attempts;
    int32_t max_depth;
} orc_chem_result;

void orc_perform_timestep(orc_solver *s, double timestep, double safety, orc_chem_result *res);
/* rate fits; which = ORC_FIT_* ; returns SI value */
enum {
    ORC_FIT_ALPHA_B = 0, ORC_FIT_DALPHA_B, ORC_FIT_RECOMB_COOL, ORC_FIT_DRECOMB_COOL,
    ORC_FIT_COLL_ION, ORC_FIT_DCOLL_ION, ORC_FIT_COLL_ION_COOL, ORC_FIT_DCOLL_ION_COOL,
    ORC_FIT_COLL_EXC_COOL, ORC_FIT_DCOLL_EXC_COOL, ORC_FIT_BREMS, ORC_FIT_DBREMS,
    ORC_FIT_COMPTON, ORC_FIT_DCOMPTON, ORC_FIT_COOLING, ORC_FIT_DCOOLING
};
double orc_fit(const orc_solver *s, int which);
double orc_photoheating_rate(const orc_solver *s, double timestep);
double orc_photoionization_rate(const orc_solver *s, double timestep);

/* ---- scheduling helpers (src/sweep/timestep_level.rs, timestep_state.rs) ---- */
int orc_level_from_timesteps(int max_num_levels, double max_timestep, double desired_timestep);
/* fills out[] with the sweep order for the given state; returns count */
int orc_levels_in_sweep_order(int max_num_levels, int lowest_allowed, int *out, int cap);
/* Rust BinaryHeap<Task> pop order for a list of direction keys: fills order[] with
 * the indices (into keys) in pop order.  For testing Appendix B of SURVEY.md. */
void orc_heap_pop_order(const uint32_t *keys, uint32_t n, uint32_t *order);

/* unit constants (SI), src/units/mod.rs */
double orc_const(int which);
enum { ORC_C_PROTON_MASS = 0, ORC_C_BOLTZMANN, ORC_C_GAMMA, ORC_C_SIGMA, ORC_C_PHOTON_ENERGY,
       ORC_C_RYDBERG, ORC_C_YEAR, ORC_C_MEGAYEAR, ORC_C_PARSEC, ORC_C_KILOPARSEC };

/* ---- CPU baseline helper: the same algorithm, directions split over threads ----
 * Runs single sweeps like orc_run_sweeps but solves the direction shards on n_threads
 * POSIX threads (each with its own task heap), then does chemistry split by cells.
 * Results equal orc_run_sweeps in LAGGED mode (bitwise) -- see tests. */
double orc_run_sweeps_threads(orc_sweep *s, int n_threads);

#ifdef __cplusplus
}
#endif
#endif
