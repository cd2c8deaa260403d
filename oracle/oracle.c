/*
 * oracle.c -- CPU restatement of subsweep's directional sweep + hydrogen chemistry.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  PARITY UNPINNED (see oracle.h).
 *
 * Plain C, f64, single thread (plus one explicitly threaded baseline entry point).
 * Compile with -ffp-contract=off: the reference is Rust, which never contracts
 * a*b+c into an FMA, and the upwind/downwind classification depends on the sign of
 * a three-term dot product.
 *
 * Reference files restated here (relative to the reference repository root):
 *   src/sweep/mod.rs:194-610            Sweep<C>: run_sweeps, single_sweep, init_counts,
 *                                       get_initial_tasks, solve, solve_task, handle_*,
 *                                       update_chemistry, update_timestep_levels
 *   src/sweep/site.rs:12-56             Site state and get_rate
 *   src/sweep/active_list.rs:6-153      per-level bins, enumerate_active order
 *   src/sweep/timestep_level.rs:27-48   level rule
 *   src/sweep/timestep_state.rs:4-90    sweep order, warm-up
 *   src/sweep/task.rs:11-35             Task ordering (by direction only)
 *   src/sweep/grid/cell.rs:119-133      Face::points_upwind / points_downwind
 *   src/chemistry/mod.rs:56-82          Photons: make_positive, relative_change_to, below_threshold
 *   src/chemistry/timescale.rs:32-38    Timescale::min
 *   src/chemistry/hydrogen_only/mod.rs:32-461   HydrogenOnly, Solver, update
 *   src/sweep/chemistry_output.rs:15-55 optional outputs
 *   src/units/mod.rs:16-108             unit factors and constants
 * Third-party semantics restated (not vendored in the reference):
 *   Rust std BinaryHeap (from Vec / push / pop), f64::{min,max,clamp,powi}, `as usize`
 *   saturating float->int casts; glam DVec3::dot; diman Quantity = transparent f64.
 */
#include "oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* units and constants: src/units/mod.rs:16-108 (factors built in expression order) */
/* ------------------------------------------------------------------------- */
static const double U_METERS = 1.0;
static const double U_SECONDS = 1.0;
static const double U_KILOGRAMS = 1.0;
static const double U_KELVINS = 1.0;
static const double U_JOULES = 1.0;
#define U_CENTIMETERS (0.01 * U_METERS)
#define U_PARSEC (3.0857e16 * U_METERS)
#define U_KILOPARSEC (1000 * U_PARSEC)
#define U_YEARS (3.15576e7 * U_SECONDS)
#define U_MEGAYEARS (1e6 * U_YEARS)
#define U_ERGS (1e-7 * U_JOULES)
#define U_ELECTRON_VOLTS (1.602176634e-19 * U_JOULES)
#define U_CUBIC_METERS 1.0
#define U_CUBIC_CENTIMETERS (1e-6 * U_CUBIC_METERS)
#define U_CENTIMETERS_SQUARED (U_CENTIMETERS * U_CENTIMETERS)
#define U_CM3_PER_S (U_CUBIC_CENTIMETERS / U_SECONDS)
#define U_ERGS_CM3_PER_S (U_ERGS * U_CM3_PER_S)
#define U_ERGS_PER_S (U_ERGS / U_SECONDS)

static double BOLTZMANN_CONSTANT(void) { return 1.380649e-23 * U_JOULES / U_KELVINS; }
static double PROTON_MASS(void) { return 1.67262192369e-27 * U_KILOGRAMS; }
static double GAMMA(void) { return 5.0 / 3.0; }
static double SIGMA(void) { return 2.9580524545305314e-18 * U_CENTIMETERS_SQUARED; }
static double PHOTON_AVERAGE_ENERGY(void) { return 18.028356312818811 * U_ELECTRON_VOLTS; }
static double RYDBERG_CONSTANT(void) { return 13.65693 * U_ELECTRON_VOLTS; }

#define HYDROGEN_MASS_FRACTION 1.0          /* hydrogen_only/mod.rs:32 */
#define MAX_DEPTH 100                        /* hydrogen_only/mod.rs:34 */
#define XHII_EPSILON 1e-10                   /* hydrogen_only/mod.rs:38 */

double orc_const(int which) {
    switch (which) {
    case ORC_C_PROTON_MASS: return PROTON_MASS();
    case ORC_C_BOLTZMANN: return BOLTZMANN_CONSTANT();
    case ORC_C_GAMMA: return GAMMA();
    case ORC_C_SIGMA: return SIGMA();
    case ORC_C_PHOTON_ENERGY: return PHOTON_AVERAGE_ENERGY();
    case ORC_C_RYDBERG: return RYDBERG_CONSTANT();
    case ORC_C_YEAR: return U_YEARS;
    case ORC_C_MEGAYEAR: return U_MEGAYEARS;
    case ORC_C_PARSEC: return U_PARSEC;
    case ORC_C_KILOPARSEC: return U_KILOPARSEC;
    }
    return NAN;
}

/* Rust f64::min / f64::max: if one operand is NaN the other is returned. */
static double rs_min(double a, double b) { return fmin(a, b); }

/* Rust f64::powi -> llvm.powi -> compiler-rt __powidf2 (square-and-multiply). */
static double rs_powi(double a, int b) {
    const int recip = b < 0;
    double r = 1.0;
    while (1) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return recip ? 1.0 / r : r;
}

/* Rust `f as usize`: NaN -> 0, negative -> 0, >= 2^64 -> usize::MAX. */
static uint64_t rs_as_usize(double f) {
    if (isnan(f)) return 0;
    if (f <= 0.0) return 0;
    if (f >= 18446744073709551616.0) return UINT64_MAX;
    return (uint64_t)f;
}

/* ------------------------------------------------------------------------- */
/* chemistry: src/chemistry/hydrogen_only/mod.rs                              */
/* ------------------------------------------------------------------------- */

/* :139-159 */
static double hydrogen_number_density(const orc_solver *s) { return s->density / PROTON_MASS(); }
static double ionized_hydrogen_number_density(const orc_solver *s) {
    return hydrogen_number_density(s) * s->xhii;
}
static double neutral_hydrogen_number_density(const orc_solver *s) {
    return hydrogen_number_density(s) * (1.0 - s->xhii);
}
static double electron_number_density(const orc_solver *s) {
    return ionized_hydrogen_number_density(s);
}
static double mu(const orc_solver *s) { return 1.0 / (s->xhii + 1.0); }

/* :161-164 */
static double collision_fit_function(const orc_solver *s) {
    const double t = s->temperature / U_KELVINS;
    return sqrt(t) / (1.0 + sqrt(t / 1e5)) * exp(-157809.1 / t);
}
/* :166-173 */
static double collision_fit_function_derivative(const orc_solver *s) {
    const double const1 = 1.0 / 1e5;
    const double const2 = 157809.1;
    const double t = s->temperature / U_KELVINS;
    return (exp(-const2 / t) *
            (const1 * const2 * t + 0.5 * sqrt(const1 * t) * (2.0 * const2 + t))) /
           (sqrt(rs_powi(t, 3)) * sqrt(const1 * t) * rs_powi(sqrt(const1 * t) + 1.0, 2));
}
/* :175-180 */
static double case_b_recombination_rate(const orc_solver *s) {
    const double lambda = (315614.0 * U_KELVINS) / s->temperature;
    return (2.753e-14 * pow(lambda, 1.5) / pow(1.0 + pow(lambda / 2.74, 0.407), 2.242)) *
           U_CM3_PER_S;
}
/* :182-194 */
static double case_b_recombination_rate_derivative(const orc_solver *s) {
    const double lambda = (315614.0 * U_KELVINS) / s->temperature;
    const double dlambda_dt = -(315614.0 * U_KELVINS) / (s->temperature * s->temperature);
    const double c1 = 1.0 / 2.74;
    const double c2 = 0.407;
    const double c3 = 2.242;
    const double d = -sqrt(lambda) * pow(pow(c1 * lambda, c2) + 1.0, -c3 - 1.0) *
                     (c2 * c3 * pow(c1 * lambda, c2) - 1.5 * pow(c1 * lambda, c2) - 1.5);
    return ((2.753e-14 * d) * U_CM3_PER_S) * dlambda_dt;
}
/* :196-202 */
static double case_b_recombination_cooling_rate(const orc_solver *s) {
    const double lambda = (315614.0 * U_KELVINS) / s->temperature;
    return (3.435e-30 * (s->temperature / U_KELVINS) * pow(lambda, 1.97) /
            pow(1.0 + pow(lambda / 2.25, 0.376), 3.72)) *
           U_ERGS_CM3_PER_S;
}
/* :204-216 */
static double case_b_recombination_cooling_rate_derivative(const orc_solver *s) {
    const double c1 = 315614.0, c2 = 1.97, c3 = 0.376, c4 = 3.72, c5 = 2.25;
    const double t = s->temperature / U_KELVINS;
    const double derivative = pow(1.0 + pow(c1 / (c5 * t), c3), -1.0 - c4) *
                              (1.0 - 1.0 * c2 + (1.0 - 1.0 * c2 + c3 * c4) * pow(c1 / (c5 * t), c3)) *
                              pow(c1 / t, c2);
    return ((3.435e-30 * derivative) * U_ERGS_CM3_PER_S) / (1.0 * U_KELVINS);
}
/* :218-235 */
static double collisional_ionization_rate(const orc_solver *s) {
    return (5.85e-11 * collision_fit_function(s)) * U_CM3_PER_S;
}
static double collisional_ionization_rate_derivative(const orc_solver *s) {
    return ((5.85e-11 * collision_fit_function_derivative(s)) * U_CM3_PER_S) / (1.0 * U_KELVINS);
}
static double collisional_ionization_cooling_rate(const orc_solver *s) {
    return (1.27e-21 * collision_fit_function(s)) * U_ERGS_CM3_PER_S;
}
static double collisional_ionization_cooling_rate_derivative(const orc_solver *s) {
    return ((1.27e-21 * collision_fit_function_derivative(s)) * U_ERGS_CM3_PER_S) /
           (1.0 * U_KELVINS);
}
/* :237-253 */
static double collisional_excitation_cooling_rate(const orc_solver *s) {
    const double t = s->temperature / U_KELVINS;
    return (7.5e-19 / (1.0 + sqrt(t / 1e5)) * exp(-118348.0 / t)) * U_ERGS_CM3_PER_S;
}
static double collisional_excitation_cooling_rate_derivative(const orc_solver *s) {
    const double t = s->temperature / U_KELVINS;
    const double c1 = 7.5e-19, c2 = 118348.0, c3 = 1.0 / 1e5;
    return (((c1 * exp(-c2 / t) *
              (c2 * c3 * t - 0.5 * c3 * rs_powi(t, 2) + c2 * sqrt(c3 * t))) /
             (rs_powi(t, 2) * sqrt(c3 * t) * rs_powi(1.0 + sqrt(c3 * t), 2))) *
            U_ERGS_CM3_PER_S) /
           (1.0 * U_KELVINS);
}
/* :255-263 */
static double bremsstrahlung_cooling_rate(const orc_solver *s) {
    return (1.42e-27 * sqrt(s->temperature / U_KELVINS)) * U_ERGS_CM3_PER_S;
}
static double bremsstrahlung_cooling_rate_derivative(const orc_solver *s) {
    return ((1.42e-27 / (2.0 * sqrt(s->temperature / U_KELVINS))) * U_ERGS_CM3_PER_S) /
           (1.0 * U_KELVINS);
}
/* :265-273 */
static double compton_cooling_rate(const orc_solver *s) {
    const double x = 2.727 / s->scale_factor;
    return (1.017e-37 * rs_powi(x, 4) * ((s->temperature / U_KELVINS) - x)) * U_ERGS_PER_S;
}
static double compton_cooling_rate_derivative(const orc_solver *s) {
    const double x = 2.727 / s->scale_factor;
    return ((1.017e-37 * rs_powi(x, 4)) * U_ERGS_PER_S) / (1.0 * U_KELVINS);
}
/* :275-287 */
static double cooling_rate(const orc_solver *s) {
    const double ne = electron_number_density(s);
    const double nh_neutral = neutral_hydrogen_number_density(s);
    const double nh_ionized = ionized_hydrogen_number_density(s);
    const double collisional =
        (collisional_excitation_cooling_rate(s) + collisional_ionization_cooling_rate(s)) * ne *
        nh_neutral;
    const double recombination = case_b_recombination_cooling_rate(s) * ne * nh_ionized;
    const double bremsstrahlung = bremsstrahlung_cooling_rate(s) * ne * nh_ionized;
    const double compton = compton_cooling_rate(s) * ne;
    return collisional + recombination + bremsstrahlung + compton;
}
/* :289-302 */
static double cooling_rate_derivative(const orc_solver *s) {
    const double ne = electron_number_density(s);
    const double nh_neutral = neutral_hydrogen_number_density(s);
    const double nh_ionized = ionized_hydrogen_number_density(s);
    const double collisional = (collisional_excitation_cooling_rate_derivative(s) +
                                collisional_ionization_cooling_rate_derivative(s)) *
                               ne * nh_neutral;
    const double recombination = case_b_recombination_cooling_rate_derivative(s) * ne * nh_ionized;
    const double bremsstrahlung = bremsstrahlung_cooling_rate_derivative(s) * ne * nh_ionized;
    const double compton = compton_cooling_rate_derivative(s) * ne;
    return collisional + recombination + bremsstrahlung + compton;
}
/* :312-319 */
static double num_newly_ionized_hydrogen_atoms(const orc_solver *s, double timestep) {
    const double nhi = neutral_hydrogen_number_density(s);
    const double sigma = SIGMA();
    const double absorbed_fraction = 1.0 - exp(-nhi * sigma * s->length);
    const double num_photons = timestep * s->rate;
    return num_photons * absorbed_fraction;
}
/* :321-325 */
double orc_photoheating_rate(const orc_solver *s, double timestep) {
    const double n = num_newly_ionized_hydrogen_atoms(s, timestep);
    const double ionization_density = n / s->volume;
    return ionization_density * (PHOTON_AVERAGE_ENERGY() - RYDBERG_CONSTANT()) / timestep;
}
/* :327-332 */
double orc_photoionization_rate(const orc_solver *s, double timestep) {
    const double n = num_newly_ionized_hydrogen_atoms(s, timestep);
    const double fraction = n / (neutral_hydrogen_number_density(s) * s->volume);
    return fraction / timestep;
}
/* :304-310 */
static double temperature_change(const orc_solver *s, double timestep) {
    const double k = (GAMMA() - 1.0) * PROTON_MASS() / (s->density * BOLTZMANN_CONSTANT());
    const double lambda = orc_photoheating_rate(s, timestep) - cooling_rate(s);
    const double dlambdadt = -cooling_rate_derivative(s);
    const double m = mu(s);
    return k * m * lambda * timestep / (1.0 - k * m * dlambdadt * timestep);
}
/* :334-354 */
static double ionized_fraction_change(const orc_solver *s, double timestep) {
    const double nh = hydrogen_number_density(s);
    const double ne = electron_number_density(s);
    const double alpha = case_b_recombination_rate(s);
    const double dalpha = case_b_recombination_rate_derivative(s);
    const double beta = collisional_ionization_rate(s);
    const double dbeta = collisional_ionization_rate_derivative(s);
    const double photoionization_rate = orc_photoionization_rate(s, timestep);
    const double c = beta * ne + photoionization_rate;
    const double m = mu(s);
    const double d = alpha * ne;
    const double xhii = s->xhii;
    const double rhsc = ne * s->temperature * m * HYDROGEN_MASS_FRACTION * dbeta;
    const double dcdx = nh * beta - rhsc;
    const double rhsd = ne * s->temperature * m * HYDROGEN_MASS_FRACTION * dalpha;
    const double dddx = nh * alpha - rhsd;
    const double j = dcdx - (c + d) - xhii * (dcdx + dddx);
    return timestep * (c - xhii * (c + d)) / (1.0 - j * timestep);
}
/* :356-369 ; f64::clamp = `if x < min {min}; if x > max {max}` */
static void solver_clamp(orc_solver *s) {
    const double xfloor = s->has_floor ? s->floor_xhii : XHII_EPSILON;
    double x = s->xhii;
    if (x < xfloor) x = xfloor;
    if (x > 1.0 - XHII_EPSILON) x = 1.0 - XHII_EPSILON;
    s->xhii = x;
    if (s->has_floor) {
        if (s->temperature < s->floor_temperature) s->temperature = s->floor_temperature;
    }
}
/* :444-461 ; returns 0 on TimestepCriterionViolated */
static int update_value(double *value, double change, double max_allowed_change, double timestep,
                        double *recommendation) {
    const double relative_change = rs_min(fabs(change / *value), 1.0 / DBL_EPSILON);
    if (relative_change > max_allowed_change) return 0;
    *value += change;
    *recommendation = timestep * (max_allowed_change / relative_change);
    return 1;
}

typedef struct { double time; int process; } timescale_t;
enum { P_TEMPERATURE = 0, P_IONIZATION = 1, P_PHOTON_RATE = 2 };
/* timescale.rs:32-38 */
static timescale_t timescale_min(timescale_t a, timescale_t b) { return (a.time < b.time) ? a : b; }

/* :371-392 */
static int try_timestep_update(orc_solver *s, double timestep, double safety, timescale_t *out) {
    double t_rec, x_rec;
    const double dT = temperature_change(s, timestep);
    if (!update_value(&s->temperature, dT, safety, timestep, &t_rec)) return 0;
    const double dx = ionized_fraction_change(s, timestep);
    if (!update_value(&s->xhii, dx, safety, timestep, &x_rec)) return 0;
    solver_clamp(s);
    timescale_t a = {t_rec, P_TEMPERATURE}, b = {x_rec, P_IONIZATION};
    *out = timescale_min(a, b);
    return 1;
}

/* :394-424 ; returns 0 on TimestepConvergenceFailed */
static int perform_timestep_internal(orc_solver *s, double timestep, double safety, int depth,
                                     int max_depth, timescale_t *out, orc_chem_result *st) {
    solver_clamp(s);
    const double t0 = s->temperature, x0 = s->xhii;
    if (depth > max_depth) return 0;
    if (depth > st->max_depth) st->max_depth = depth;
    st->attempts++;
    if (try_timestep_update(s, timestep, safety, out)) return 1;
    s->temperature = t0;
    s->xhii = x0;
    timescale_t dummy;
    if (!perform_timestep_internal(s, timestep / 2.0, safety, depth + 1, max_depth, &dummy, st))
        return 0;
    return perform_timestep_internal(s, timestep / 2.0, safety, depth + 1, max_depth, out, st);
}

/* :426-441 */
void orc_perform_timestep(orc_solver *s, double timestep, double safety, orc_chem_result *res) {
    timescale_t ts;
    res->attempts = 0;
    res->max_depth = 0;
    res->failed = 0;
    if (!perform_timestep_internal(s, timestep, safety, 0, MAX_DEPTH, &ts, res)) {
        res->failed = 1;
        ts.time = timestep / 10.0;
        ts.process = P_TEMPERATURE;
    }
    res->timescale = ts.time;
    res->process = ts.process;
}

double orc_fit(const orc_solver *s, int which) {
    switch (which) {
    case ORC_FIT_ALPHA_B: return case_b_recombination_rate(s);
    case ORC_FIT_DALPHA_B: return case_b_recombination_rate_derivative(s);
    case ORC_FIT_RECOMB_COOL: return case_b_recombination_cooling_rate(s);
    case ORC_FIT_DRECOMB_COOL: return case_b_recombination_cooling_rate_derivative(s);
    case ORC_FIT_COLL_ION: return collisional_ionization_rate(s);
    case ORC_FIT_DCOLL_ION: return collisional_ionization_rate_derivative(s);
    case ORC_FIT_COLL_ION_COOL: return collisional_ionization_cooling_rate(s);
    case ORC_FIT_DCOLL_ION_COOL: return collisional_ionization_cooling_rate_derivative(s);
    case ORC_FIT_COLL_EXC_COOL: return collisional_excitation_cooling_rate(s);
    case ORC_FIT_DCOLL_EXC_COOL: return collisional_excitation_cooling_rate_derivative(s);
    case ORC_FIT_BREMS: return bremsstrahlung_cooling_rate(s);
    case ORC_FIT_DBREMS: return bremsstrahlung_cooling_rate_derivative(s);
    case ORC_FIT_COMPTON: return compton_cooling_rate(s);
    case ORC_FIT_DCOMPTON: return compton_cooling_rate_derivative(s);
    case ORC_FIT_COOLING: return cooling_rate(s);
    case ORC_FIT_DCOOLING: return cooling_rate_derivative(s);
    }
    return NAN;
}

/* ------------------------------------------------------------------------- */
/* scheduling: timestep_level.rs / timestep_state.rs                          */
/* ------------------------------------------------------------------------- */

/* timestep_level.rs:27-36 */
int orc_level_from_timesteps(int max_num_levels, double max_timestep, double desired_timestep) {
    const double ratio = max_timestep / desired_timestep;
    const uint64_t level = rs_as_usize(ceil(log2(ratio)));
    const uint64_t hi = (uint64_t)(max_num_levels - 1);
    return (int)(level > hi ? hi : level);
}

/* timestep_state.rs:81-89 */
static int lowest_set_bit(uint32_t v) {
    for (int b = 0; b < 32; b++)
        if (v & (1u << b)) return b;
    return -1;
}

/* timestep_state.rs:22-27, 70-75 */
int orc_levels_in_sweep_order(int max_num_levels, int lowest_allowed, int *out, int cap) {
    const int num = max_num_levels - lowest_allowed;
    const uint32_t count = 1u << (num - 1);
    int n = 0;
    for (uint32_t i = 0; i < count; i++) {
        int fb = lowest_set_bit(i);
        if (fb < 0) fb = num - 1;
        if (n < cap) out[n] = lowest_allowed + (num - 1 - fb);
        n++;
    }
    return n;
}

/* ------------------------------------------------------------------------- */
/* Rust std::collections::BinaryHeap<Task>; Task ordered by dir only (task.rs:25-35) */
/* ------------------------------------------------------------------------- */
typedef struct { uint32_t id; uint32_t dir; } task_t;
typedef struct { task_t *data; size_t len, cap; } heap_t;

static void heap_reserve(heap_t *h, size_t n) {
    if (n <= h->cap) return;
    size_t c = h->cap ? h->cap : 1024;
    while (c < n) c *= 2;
    h->data = (task_t *)realloc(h->data, c * sizeof(task_t));
    if (!h->data) { fprintf(stderr, "oracle: out of memory\n"); abort(); }
    h->cap = c;
}
/* sift_up(start, pos): move up while element > parent (strict) */
static void heap_sift_up(heap_t *h, size_t start, size_t pos) {
    const task_t elem = h->data[pos];
    while (pos > start) {
        const size_t parent = (pos - 1) / 2;
        if (elem.dir <= h->data[parent].dir) break;
        h->data[pos] = h->data[parent];
        pos = parent;
    }
    h->data[pos] = elem;
}
/* sift_down_range(pos, end) */
static void heap_sift_down_range(heap_t *h, size_t pos, size_t end) {
    const task_t elem = h->data[pos];
    size_t child = 2 * pos + 1;
    const size_t lim = end >= 2 ? end - 2 : 0; /* end.saturating_sub(2) */
    while (child <= lim) {
        if (h->data[child].dir <= h->data[child + 1].dir) child += 1;
        if (elem.dir >= h->data[child].dir) { h->data[pos] = elem; return; }
        h->data[pos] = h->data[child];
        pos = child;
        child = 2 * pos + 1;
    }
    if (child == end - 1 && elem.dir < h->data[child].dir) {
        h->data[pos] = h->data[child];
        pos = child;
    }
    h->data[pos] = elem;
}
/* BinaryHeap::from(Vec) -> rebuild */
static void heap_rebuild(heap_t *h) {
    size_t n = h->len / 2;
    while (n > 0) { n -= 1; heap_sift_down_range(h, n, h->len); }
}
static void heap_push(heap_t *h, task_t t) {
    heap_reserve(h, h->len + 1);
    const size_t old = h->len;
    h->data[h->len++] = t;
    heap_sift_up(h, 0, old);
}
/* pop: swap last into root, sift_down_to_bottom(0) then sift_up */
static int heap_pop(heap_t *h, task_t *out) {
    if (h->len == 0) return 0;
    task_t item = h->data[--h->len];
    if (h->len > 0) {
        const task_t root = h->data[0];
        h->data[0] = item;
        item = root;
        const size_t end = h->len;
        size_t pos = 0;
        const task_t elem = h->data[0];
        size_t child = 1;
        const size_t lim = end >= 2 ? end - 2 : 0;
        while (child <= lim) {
            if (h->data[child].dir <= h->data[child + 1].dir) child += 1;
            h->data[pos] = h->data[child];
            pos = child;
            child = 2 * pos + 1;
        }
        if (child == end - 1) { h->data[pos] = h->data[child]; pos = child; }
        h->data[pos] = elem;
        heap_sift_up(h, 0, pos);
    }
    *out = item;
    return 1;
}

void orc_heap_pop_order(const uint32_t *keys, uint32_t n, uint32_t *order) {
    heap_t h = {0};
    heap_reserve(&h, n);
    for (uint32_t i = 0; i < n; i++) { h.data[i].id = i; h.data[i].dir = keys[i]; }
    h.len = n;
    heap_rebuild(&h);
    task_t t;
    uint32_t k = 0;
    while (heap_pop(&h, &t)) order[k++] = t.id;
    free(h.data);
}

/* ------------------------------------------------------------------------- */
/* Sweep<HydrogenOnly>                                                        */
/* ------------------------------------------------------------------------- */
struct orc_sweep {
    orc_params p;
    double *dirs;         /* D x 3 */
    int D, Dl, d0;        /* total dirs, local dirs, first local dir */
    uint64_t N, F;
    uint64_t *face_offsets;
    double *face_area, *face_normal;
    int32_t *face_nb;
    uint8_t *face_kind;
    double *size, *volume;
    /* Site (site.rs:12-23), cell-major [c*Dl + dl] */
    double *in, *out, *per, *per_lag;
    uint32_t *miss;
    double *prev_rate, *x, *T, *ts, *rho, *tau, *src;
    uint8_t *level;       /* ActiveList.levels */
    uint32_t **bins;      /* ActiveList.bins */
    uint64_t *bin_len;
    /* TimestepState (timestep_state.rs:4-9) */
    int lowest_allowed;
    int first_done;
    int cur;              /* current_level */
    /* LAGGED/HEAP bookkeeping */
    uint32_t *solved_epoch; /* per (c,dl): epoch of the last sweep that solved it */
    uint32_t epoch;
    double *rate_buf;
    uint64_t stats[8];
    orc_allreduce_fn allreduce;
    void *allreduce_ctx;
};

static void *xcalloc(size_t n, size_t sz) {
    void *p = calloc(n ? n : 1, sz);
    if (!p) { fprintf(stderr, "oracle: out of memory (%zu x %zu)\n", n, sz); abort(); }
    return p;
}
static void *xdup(const void *src, size_t bytes) {
    void *p = malloc(bytes ? bytes : 1);
    if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); }
    memcpy(p, src, bytes);
    return p;
}

/* active_list.rs:143-153 */
static void update_bins(orc_sweep *s) {
    const int L = s->p.n_levels;
    for (int l = 0; l < L; l++) s->bin_len[l] = 0;
    for (uint64_t i = 0; i < s->N; i++) s->bin_len[s->level[i]]++;
    for (int l = 0; l < L; l++) {
        free(s->bins[l]);
        s->bins[l] = (uint32_t *)xcalloc(s->bin_len[l], sizeof(uint32_t));
        s->bin_len[l] = 0;
    }
    for (uint64_t i = 0; i < s->N; i++) {
        const int l = s->level[i];
        s->bins[l][s->bin_len[l]++] = (uint32_t)i;
    }
}

/* Sweep::new, mod.rs:194-232 ; Site::new, site.rs:26-43 ; ActiveList::new, active_list.rs:16-46 */
orc_sweep *orc_create(const orc_params *p, const orc_grid *g, const double *density,
                      const double *xhii, const double *temperature, const double *source) {
    orc_sweep *s = (orc_sweep *)xcalloc(1, sizeof(*s));
    s->p = *p;
    s->D = p->n_dirs;
    s->d0 = p->dir_begin;
    s->Dl = (p->dir_end > p->dir_begin) ? p->dir_end - p->dir_begin : p->n_dirs;
    if (!(p->dir_end > p->dir_begin)) s->d0 = 0;
    s->dirs = (double *)xdup(p->dirs_xyz, sizeof(double) * 3 * (size_t)s->D);
    s->p.dirs_xyz = s->dirs;
    s->N = g->n_cells;
    s->F = g->face_offsets[g->n_cells];
    s->face_offsets = (uint64_t *)xdup(g->face_offsets, sizeof(uint64_t) * (s->N + 1));
    s->face_area = (double *)xdup(g->face_area, sizeof(double) * s->F);
    s->face_normal = (double *)xdup(g->face_normal, sizeof(double) * 3 * s->F);
    s->face_nb = (int32_t *)xdup(g->face_neighbour, sizeof(int32_t) * s->F);
    s->face_kind = (uint8_t *)xdup(g->face_kind, s->F);
    s->size = (double *)xdup(g->cell_size, sizeof(double) * s->N);
    s->volume = (double *)xdup(g->cell_volume, sizeof(double) * s->N);
    const size_t ND = (size_t)s->N * (size_t)s->Dl;
    s->in = (double *)xcalloc(ND, sizeof(double));
    s->out = (double *)xcalloc(ND, sizeof(double));
    s->per = (double *)xcalloc(ND, sizeof(double));
    s->per_lag = (p->periodic_mode == ORC_PERIODIC_LAGGED) ? (double *)xcalloc(ND, sizeof(double)) : NULL;
    s->miss = (uint32_t *)xcalloc(ND, sizeof(uint32_t));
    s->solved_epoch = (uint32_t *)xcalloc(ND, sizeof(uint32_t));
    s->prev_rate = (double *)xcalloc(s->N, sizeof(double));
    s->x = (double *)xdup(xhii, sizeof(double) * s->N);
    s->T = (double *)xdup(temperature, sizeof(double) * s->N);
    s->ts = (double *)xcalloc(s->N, sizeof(double));
    s->rho = (double *)xdup(density, sizeof(double) * s->N);
    s->tau = (double *)xcalloc(s->N, sizeof(double));
    s->src = (double *)xdup(source, sizeof(double) * s->N);
    s->rate_buf = (double *)xcalloc(s->N, sizeof(double));
    s->level = (uint8_t *)xcalloc(s->N, 1);
    const int L = p->n_levels;
    memset(s->level, L - 1, s->N);            /* initial_level, mod.rs:206 */
    s->bins = (uint32_t **)xcalloc(L, sizeof(uint32_t *));
    s->bin_len = (uint64_t *)xcalloc(L, sizeof(uint64_t));
    update_bins(s);
    s->lowest_allowed = L - 1;                /* timestep_state.rs:17 */
    s->first_done = 0;
    s->cur = 0;
    return s;
}

void orc_destroy(orc_sweep *s) {
    if (!s) return;
    free(s->dirs); free(s->face_offsets); free(s->face_area); free(s->face_normal);
    free(s->face_nb); free(s->face_kind); free(s->size); free(s->volume);
    free(s->in); free(s->out); free(s->per); free(s->per_lag); free(s->miss);
    free(s->solved_epoch); free(s->prev_rate); free(s->x); free(s->T); free(s->ts);
    free(s->rho); free(s->tau); free(s->src); free(s->rate_buf); free(s->level);
    for (int l = 0; l < s->p.n_levels; l++) free(s->bins[l]);
    free(s->bins); free(s->bin_len);
    free(s);
}

void orc_set_allreduce(orc_sweep *s, orc_allreduce_fn fn, void *ctx) {
    s->allreduce = fn;
    s->allreduce_ctx = ctx;
}

/* glam DVec3::dot: (x*x + y*y) + z*z */
static inline double dot3(const double *a, const double *b) {
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}
/* timestep_level.rs:38-40 */
static inline int is_active(const orc_sweep *s, uint32_t c) { return s->level[c] >= s->cur; }

/* mod.rs:346-386 (single rank: only Local neighbours count) */
static void init_counts(orc_sweep *s, int dl_begin, int dl_end) {
    for (int l = s->cur; l < s->p.n_levels; l++) {
        for (uint64_t k = 0; k < s->bin_len[l]; k++) {
            const uint32_t c = s->bins[l][k];
            uint32_t *miss = s->miss + (size_t)c * s->Dl;
            for (int dl = dl_begin; dl < dl_end; dl++) {
                const double *dir = s->dirs + 3 * (size_t)(s->d0 + dl);
                uint32_t m = 0;
                for (uint64_t f = s->face_offsets[c]; f < s->face_offsets[c + 1]; f++) {
                    const double d = dot3(s->face_normal + 3 * f, dir);
                    if (!(d < 0.0) || s->face_kind[f] == ORC_FACE_BOUNDARY) continue;
                    if (!is_active(s, (uint32_t)s->face_nb[f])) continue;
                    if (s->face_kind[f] == ORC_FACE_LOCAL) m += 1;
                }
                miss[dl] = m;
            }
        }
    }
}

/* mod.rs:388-398 */
static void get_initial_tasks(orc_sweep *s, heap_t *h, int dl_begin, int dl_end) {
    h->len = 0;
    for (int dl = dl_begin; dl < dl_end; dl++) {
        for (int l = s->cur; l < s->p.n_levels; l++) {
            for (uint64_t k = 0; k < s->bin_len[l]; k++) {
                const uint32_t c = s->bins[l][k];
                if (s->miss[(size_t)c * s->Dl + dl] == 0) {
                    heap_reserve(h, h->len + 1);
                    h->data[h->len].id = c;
                    h->data[h->len].dir = (uint32_t)dl;
                    h->len++;
                }
            }
        }
    }
    heap_rebuild(h);
}

/* mod.rs:412-421 + site.rs:49-56 + hydrogen_only/mod.rs:72-88 */
static double get_outgoing_rate(orc_sweep *s, uint32_t c, int dl, uint64_t *nonlagged) {
    const size_t i = (size_t)c * s->Dl + dl;
    if (s->in[i] < 0.0) s->in[i] = 0.0;                 /* make_positive, chemistry/mod.rs:61-65 */
    const double source = s->src[c] / (double)s->D;     /* site.rs:49-51 */
    const double per = s->per_lag ? s->per_lag[i] : s->per[i];
    const double incoming = s->in[i] + source + per;    /* site.rs:53-56 */
    (void)nonlagged;
    const double nhi = s->rho[c] / PROTON_MASS() * (1.0 - s->x[c]);
    const double sigma = SIGMA();
    if (incoming < s->p.significant_rate_threshold) return 0.0;
    const double non_absorbed = exp(-nhi * sigma * s->size[c]);
    return incoming * non_absorbed;
}

/* mod.rs:423-485, 487-513 */
static void solve_task(orc_sweep *s, heap_t *h, task_t t, uint64_t *nonlagged) {
    const uint32_t c = t.id;
    const int dl = (int)t.dir;
    const size_t i = (size_t)c * s->Dl + dl;
    const double *dir = s->dirs + 3 * (size_t)(s->d0 + dl);
    /* statistic only: would the reference have read a periodic contribution that was
       scattered earlier in this same sweep?  (DESIGN.md "periodic lag") */
    for (uint64_t f = s->face_offsets[c]; f < s->face_offsets[c + 1]; f++) {
        if (s->face_kind[f] != ORC_FACE_LOCAL_PERIODIC) continue;
        if (!(dot3(s->face_normal + 3 * f, dir) < 0.0)) continue;
        const uint32_t nb = (uint32_t)s->face_nb[f];
        if (s->solved_epoch[(size_t)nb * s->Dl + dl] == s->epoch) (*nonlagged)++;
    }
    const double outgoing = get_outgoing_rate(s, c, dl, nonlagged);
    const double correction = outgoing - s->out[i];
    s->out[i] = outgoing;
    s->solved_epoch[i] = s->epoch;
    double total_effective_area = 0.0;                   /* iterator .sum() from 0 */
    for (uint64_t f = s->face_offsets[c]; f < s->face_offsets[c + 1]; f++) {
        const double d = dot3(s->face_normal + 3 * f, dir);
        if (d > 0.0) total_effective_area += s->face_area[f] * d;
    }
    for (uint64_t f = s->face_offsets[c]; f < s->face_offsets[c + 1]; f++) {
        const double d = dot3(s->face_normal + 3 * f, dir);
        if (!(d > 0.0)) continue;
        const double effective_area = s->face_area[f] * d;
        const double share = correction * (effective_area / total_effective_area);
        const uint8_t kind = s->face_kind[f];
        if (kind == ORC_FACE_LOCAL) {
            const uint32_t nb = (uint32_t)s->face_nb[f];
            const size_t j = (size_t)nb * s->Dl + dl;
            s->in[j] += share;
            if (is_active(s, nb)) {
                if (--s->miss[j] == 0) {
                    task_t nt = {nb, (uint32_t)dl};
                    heap_push(h, nt);
                }
            }
        } else if (kind == ORC_FACE_LOCAL_PERIODIC) {
            const uint32_t nb = (uint32_t)s->face_nb[f];
            s->per[(size_t)nb * s->Dl + dl] += share;
        }
    }
}

static uint64_t count_active(const orc_sweep *s) {
    uint64_t n = 0;
    for (int l = s->cur; l < s->p.n_levels; l++) n += s->bin_len[l];
    return n;
}

/* init_counts + get_initial_tasks + solve for local directions [dl_begin, dl_end) */
static void sweep_directions(orc_sweep *s, int dl_begin, int dl_end, uint64_t *solved,
                             uint64_t *nonlagged) {
    heap_t h = {0};
    init_counts(s, dl_begin, dl_end);
    get_initial_tasks(s, &h, dl_begin, dl_end);
    const uint64_t expect = count_active(s) * (uint64_t)(dl_end - dl_begin);
    if (s->p.check_deadlock && expect > 0 && h.len == 0) {
        fprintf(stderr, "oracle: deadlock: no initial task\n");   /* deadlock_detection.rs:86-98 */
        abort();
    }
    task_t t;
    uint64_t n = 0;
    while (heap_pop(&h, &t)) { solve_task(s, &h, t, nonlagged); n++; }
    if (n != expect) {
        fprintf(stderr, "oracle: dependency cycle: solved %llu of %llu tasks\n",
                (unsigned long long)n, (unsigned long long)expect);
        abort();   /* the reference would spin forever in solve(), mod.rs:291-300 */
    }
    *solved += n;
    free(h.data);
}

/* hydrogen_only/mod.rs:90-119 */
static timescale_t update_abundances(orc_sweep *s, uint32_t c, double rate, double timestep,
                                     orc_chem_result *res) {
    orc_solver sol;
    sol.xhii = s->x[c];
    sol.temperature = s->T[c];
    sol.density = s->rho[c];
    sol.volume = s->volume[c];
    sol.length = s->size[c];
    sol.rate = rate;
    sol.scale_factor = s->p.scale_factor;
    sol.has_floor = s->p.prevent_cooling ? 1 : 0;
    sol.floor_temperature = s->T[c];
    sol.floor_xhii = s->x[c];
    orc_perform_timestep(&sol, timestep, s->p.chemistry_timestep_safety_factor, res);
    s->T[c] = sol.temperature;
    s->x[c] = sol.xhii;
    s->ts[c] = res->timescale;
    timescale_t t = {res->timescale, res->process};
    return t;
}

/* Σ_d get_rate(d) over the local directions, site.rs:53-56 ; mod.rs:554-558 */
static double rate_sum(const orc_sweep *s, uint32_t c) {
    const double source = s->src[c] / (double)s->D;
    const double *in = s->in + (size_t)c * s->Dl, *per = s->per + (size_t)c * s->Dl;
    double rate = 0.0;
    for (int dl = 0; dl < s->Dl; dl++) rate += in[dl] + source + per[dl];
    return rate;
}

/* mod.rs:549-574 for one cell, given the (all-direction) rate */
static void chemistry_cell(orc_sweep *s, uint32_t c, double rate, uint64_t *stats) {
    const double timestep = s->p.max_timestep * rs_powi(0.5, s->level[c]);  /* timestep_level.rs:42-48 */
    double relative_change;
    if (fabs(rate) < fabs(s->p.significant_rate_threshold)) {    /* chemistry/mod.rs:79-81 */
        relative_change = 0.0;
    } else {                                                     /* chemistry/mod.rs:73-77 */
        relative_change = fabs(rs_min(fabs(fabs(rate - s->prev_rate[c]) / rate), 1.0 / DBL_EPSILON));
    }
    s->prev_rate[c] = rate;
    timescale_t rate_ts = {timestep / relative_change, P_PHOTON_RATE};
    orc_chem_result res;
    const timescale_t chem_ts = update_abundances(s, c, rate, timestep, &res);
    const timescale_t change = timescale_min(rate_ts, chem_ts);
    s->tau[c] = change.time;
    stats[ORC_STAT_CHEM_ATTEMPTS] += res.attempts;
    if ((uint64_t)res.max_depth > stats[ORC_STAT_CHEM_MAX_DEPTH]) stats[ORC_STAT_CHEM_MAX_DEPTH] = res.max_depth;
    stats[ORC_STAT_CHEM_FAILURES] += res.failed;
    stats[ORC_STAT_CHEM_CELLS] += 1;
}

/* gathers the rates of the active cells (and all-reduces them when sharded) */
static void compute_rates(orc_sweep *s) {
    for (int l = s->cur; l < s->p.n_levels; l++)
        for (uint64_t k = 0; k < s->bin_len[l]; k++) {
            const uint32_t c = s->bins[l][k];
            s->rate_buf[c] = rate_sum(s, c);
        }
    if (s->allreduce && s->Dl != s->D) {
        /* inactive entries are whatever they were; zero them so the sum is well defined */
        for (uint64_t c = 0; c < s->N; c++)
            if (!is_active(s, (uint32_t)c)) s->rate_buf[c] = 0.0;
        if (s->allreduce(s->allreduce_ctx, s->rate_buf, s->N) != 0) {
            fprintf(stderr, "oracle: allreduce failed\n");
            abort();
        }
    }
}

/* mod.rs:549-574 */
static void update_chemistry(orc_sweep *s) {
    compute_rates(s);
    for (int l = s->cur; l < s->p.n_levels; l++)
        for (uint64_t k = 0; k < s->bin_len[l]; k++) {
            const uint32_t c = s->bins[l][k];
            chemistry_cell(s, c, s->rate_buf[c], s->stats);
        }
}

/* mod.rs:274-289 */
void orc_single_sweep(orc_sweep *s, int level) {
    s->cur = level;
    s->epoch++;
    if (s->per_lag) memcpy(s->per_lag, s->per, sizeof(double) * (size_t)s->N * s->Dl);
    sweep_directions(s, 0, s->Dl, &s->stats[ORC_STAT_TASKS_SOLVED],
                     &s->stats[ORC_STAT_NONLAGGED_PERIODIC_READS]);
    update_chemistry(s);
    s->stats[ORC_STAT_SINGLE_SWEEPS]++;
}

/* mod.rs:240-245 : counts[l] = #cells with level >= l */
void orc_level_counts(orc_sweep *s, uint64_t *out) {
    const int L = s->p.n_levels;
    uint64_t acc = 0;
    for (int l = L - 1; l >= 0; l--) { acc += s->bin_len[l]; out[l] = acc; }
}

/* mod.rs:576-589 + timestep_state.rs:54-64 */
void orc_update_timestep_levels(orc_sweep *s) {
    for (uint64_t c = 0; c < s->N; c++) {
        const double desired = s->p.timestep_safety_factor * s->tau[c];
        int lv = orc_level_from_timesteps(s->p.n_levels, s->p.max_timestep, desired);
        if (lv < s->lowest_allowed) lv = s->lowest_allowed;
        s->level[c] = (uint8_t)lv;
    }
    update_bins(s);
}

/* timestep_state.rs:37-48 */
static void advance_allowed_levels(orc_sweep *s) {
    if (s->first_done && s->lowest_allowed > 0) s->lowest_allowed -= 1;
    if (!s->first_done) s->first_done = 1;
}

/* mod.rs:258-272 */
double orc_run_sweeps(orc_sweep *s) {
    uint64_t counts[64];
    int order[1 << 16];
    orc_level_counts(s, counts);
    const int n = orc_levels_in_sweep_order(s->p.n_levels, s->lowest_allowed, order, 1 << 16);
    for (int i = 0; i < n; i++)
        if (counts[order[i]] > 0) orc_single_sweep(s, order[i]);
    const double elapsed = s->p.max_timestep * rs_powi(0.5, s->lowest_allowed); /* timestep_state.rs:77-79 */
    advance_allowed_levels(s);
    orc_update_timestep_levels(s);
    return elapsed;
}

/* ---- threaded baseline: same algorithm, direction shards on threads ---- */
typedef struct { orc_sweep *s; int a, b; uint64_t solved, nonlagged; } dir_job;
static void *dir_worker(void *arg) {
    dir_job *j = (dir_job *)arg;
    sweep_directions(j->s, j->a, j->b, &j->solved, &j->nonlagged);
    return NULL;
}
typedef struct { orc_sweep *s; uint32_t *cells; uint64_t a, b; uint64_t stats[8]; } chem_job;
static void *chem_worker(void *arg) {
    chem_job *j = (chem_job *)arg;
    for (uint64_t k = j->a; k < j->b; k++) chemistry_cell(j->s, j->cells[k], j->s->rate_buf[j->cells[k]], j->stats);
    return NULL;
}

static void single_sweep_threads(orc_sweep *s, int level, int nt) {
    s->cur = level;
    s->epoch++;
    if (s->per_lag) memcpy(s->per_lag, s->per, sizeof(double) * (size_t)s->N * s->Dl);
    if (nt > s->Dl) nt = s->Dl;
    pthread_t *th = (pthread_t *)xcalloc(nt, sizeof(pthread_t));
    dir_job *jobs = (dir_job *)xcalloc(nt, sizeof(dir_job));
    for (int t = 0; t < nt; t++) {
        jobs[t].s = s;
        jobs[t].a = (int)((long)s->Dl * t / nt);
        jobs[t].b = (int)((long)s->Dl * (t + 1) / nt);
        pthread_create(&th[t], NULL, dir_worker, &jobs[t]);
    }
    for (int t = 0; t < nt; t++) {
        pthread_join(th[t], NULL);
        s->stats[ORC_STAT_TASKS_SOLVED] += jobs[t].solved;
        s->stats[ORC_STAT_NONLAGGED_PERIODIC_READS] += jobs[t].nonlagged;
    }
    compute_rates(s);
    const uint64_t na = count_active(s);
    uint32_t *cells = (uint32_t *)xcalloc(na, sizeof(uint32_t));
    uint64_t k = 0;
    for (int l = s->cur; l < s->p.n_levels; l++) {
        memcpy(cells + k, s->bins[l], s->bin_len[l] * sizeof(uint32_t));
        k += s->bin_len[l];
    }
    chem_job *cj = (chem_job *)xcalloc(nt, sizeof(chem_job));
    /* interleave cells over threads in blocks of 64: chemistry cost per cell is very uneven */
    for (int t = 0; t < nt; t++) {
        cj[t].s = s; cj[t].cells = cells;
        cj[t].a = na * t / nt; cj[t].b = na * (t + 1) / nt;
        pthread_create(&th[t], NULL, chem_worker, &cj[t]);
    }
    for (int t = 0; t < nt; t++) {
        pthread_join(th[t], NULL);
        s->stats[ORC_STAT_CHEM_ATTEMPTS] += cj[t].stats[ORC_STAT_CHEM_ATTEMPTS];
        if (cj[t].stats[ORC_STAT_CHEM_MAX_DEPTH] > s->stats[ORC_STAT_CHEM_MAX_DEPTH])
            s->stats[ORC_STAT_CHEM_MAX_DEPTH] = cj[t].stats[ORC_STAT_CHEM_MAX_DEPTH];
        s->stats[ORC_STAT_CHEM_FAILURES] += cj[t].stats[ORC_STAT_CHEM_FAILURES];
        s->stats[ORC_STAT_CHEM_CELLS] += cj[t].stats[ORC_STAT_CHEM_CELLS];
    }
    free(cj); free(cells); free(jobs); free(th);
    s->stats[ORC_STAT_SINGLE_SWEEPS]++;
}

double orc_run_sweeps_threads(orc_sweep *s, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    uint64_t counts[64];
    int order[1 << 16];
    orc_level_counts(s, counts);
    const int n = orc_levels_in_sweep_order(s->p.n_levels, s->lowest_allowed, order, 1 << 16);
    for (int i = 0; i < n; i++)
        if (counts[order[i]] > 0) single_sweep_threads(s, order[i], n_threads);
    const double elapsed = s->p.max_timestep * rs_powi(0.5, s->lowest_allowed);
    advance_allowed_levels(s);
    orc_update_timestep_levels(s);
    return elapsed;
}

/* ---- test hooks and read-back ---- */
void orc_set_levels(orc_sweep *s, const uint8_t *levels) {
    memcpy(s->level, levels, s->N);
    update_bins(s);
}
void orc_set_change_timescale(orc_sweep *s, const double *tau) {
    memcpy(s->tau, tau, sizeof(double) * s->N);
}
void orc_read_levels(orc_sweep *s, uint8_t *out) { memcpy(out, s->level, s->N); }
int orc_lowest_allowed_level(orc_sweep *s) { return s->lowest_allowed; }
uint64_t orc_stat(orc_sweep *s, int which) { return s->stats[which]; }

void orc_read_dir_state(orc_sweep *s, int which, double *out) {
    const double *src = which == 0 ? s->in : which == 1 ? s->out : s->per;
    memcpy(out, src, sizeof(double) * (size_t)s->N * s->Dl);
}

/* wavefront level of (c, dir) for the active set of `level`:
 * 0 for initial tasks, else 1 + max over active Local upwind neighbours. */
void orc_wavefront_levels(orc_sweep *s, int level, int dir, int32_t *out) {
    const int saved = s->cur;
    s->cur = level;
    const double *d = s->dirs + 3 * (size_t)dir;
    uint32_t *miss = (uint32_t *)xcalloc(s->N, sizeof(uint32_t));
    uint32_t *queue = (uint32_t *)xcalloc(s->N, sizeof(uint32_t));
    uint64_t head = 0, tail = 0;
    for (uint64_t c = 0; c < s->N; c++) {
        out[c] = -1;
        if (!is_active(s, (uint32_t)c)) continue;
        uint32_t m = 0;
        for (uint64_t f = s->face_offsets[c]; f < s->face_offsets[c + 1]; f++) {
            if (!(dot3(s->face_normal + 3 * f, d) < 0.0)) continue;
            if (s->face_kind[f] != ORC_FACE_LOCAL) continue;
            if (is_active(s, (uint32_t)s->face_nb[f])) m++;
        }
        miss[c] = m;
        if (m == 0) { out[c] = 0; queue[tail++] = (uint32_t)c; }
    }
    while (head < tail) {
        const uint32_t c = queue[head++];
        for (uint64_t f = s->face_offsets[c]; f < s->face_offsets[c + 1]; f++) {
            if (!(dot3(s->face_normal + 3 * f, d) > 0.0)) continue;
            if (s->face_kind[f] != ORC_FACE_LOCAL) continue;
            const uint32_t nb = (uint32_t)s->face_nb[f];
            if (!is_active(s, nb)) continue;
            if (out[nb] < out[c] + 1) out[nb] = out[c] + 1;
            if (--miss[nb] == 0) queue[tail++] = nb;
        }
    }
    free(miss); free(queue);
    s->cur = saved;
}

/* Sweep::get_solver, mod.rs:612-632 (floor: None) */
static void get_solver(orc_sweep *s, uint32_t c, double rate, orc_solver *sol) {
    sol->xhii = s->x[c];
    sol->temperature = s->T[c];
    sol->density = s->rho[c];
    sol->volume = s->volume[c];
    sol->length = s->size[c];
    sol->rate = rate;
    sol->scale_factor = s->p.scale_factor;
    sol->has_floor = 0;
    sol->floor_temperature = 0.0;
    sol->floor_xhii = 0.0;
}

int orc_read(orc_sweep *s, int field, double *out) {
    const double one_year = 1.0 * U_YEARS;               /* chemistry_output.rs:15-19 */
    switch (field) {
    case ORC_F_XHII: memcpy(out, s->x, sizeof(double) * s->N); return 0;
    case ORC_F_TEMPERATURE: memcpy(out, s->T, sizeof(double) * s->N); return 0;
    case ORC_F_TIMESTEP: memcpy(out, s->ts, sizeof(double) * s->N); return 0;
    case ORC_F_CHANGE_TIMESCALE: memcpy(out, s->tau, sizeof(double) * s->N); return 0;
    case ORC_F_PREVIOUS_RATE: memcpy(out, s->prev_rate, sizeof(double) * s->N); return 0;
    case ORC_F_DENSITY: memcpy(out, s->rho, sizeof(double) * s->N); return 0;
    case ORC_F_SOURCE: memcpy(out, s->src, sizeof(double) * s->N); return 0;
    case ORC_F_PHOTON_RATE:                              /* mod.rs:727-730 */
        for (uint64_t c = 0; c < s->N; c++) {
            double acc = 0.0;
            for (int dl = 0; dl < s->Dl; dl++) acc += s->in[(size_t)c * s->Dl + dl];
            out[c] = acc;
        }
        if (s->allreduce && s->Dl != s->D) return s->allreduce(s->allreduce_ctx, out, s->N);
        return 0;
    case ORC_F_PHOTOIONIZATION_RATE:
    case ORC_F_HEATING_RATE:
    case ORC_F_RECOMBINATION_RATE:
    case ORC_F_COLLISIONAL_IONIZATION_RATE: {
        double *rates = (double *)xcalloc(s->N, sizeof(double));
        for (uint64_t c = 0; c < s->N; c++) rates[c] = rate_sum(s, (uint32_t)c);
        if (s->allreduce && s->Dl != s->D) s->allreduce(s->allreduce_ctx, rates, s->N);
        for (uint64_t c = 0; c < s->N; c++) {
            orc_solver sol;
            get_solver(s, (uint32_t)c, rates[c], &sol);
            if (field == ORC_F_PHOTOIONIZATION_RATE)                      /* chemistry_output.rs:25-29 */
                out[c] = orc_photoionization_rate(&sol, one_year);
            else if (field == ORC_F_HEATING_RATE)                         /* :31-35 */
                out[c] = orc_photoheating_rate(&sol, one_year) - cooling_rate(&sol);
            else if (field == ORC_F_RECOMBINATION_RATE)                   /* :37-45 */
                out[c] = case_b_recombination_rate(&sol) * electron_number_density(&sol) * sol.xhii;
            else                                                          /* :47-55 */
                out[c] = collisional_ionization_rate(&sol) * electron_number_density(&sol) *
                         (1.0 - sol.xhii);
        }
        free(rates);
        return 0;
    }
    }
    return -1;
}
