// chemistry.cuh -- device-side hydrogen ionization / temperature update, one cell per thread.
//
// Replaces HydrogenOnly / Solver of the reference (src/chemistry/hydrogen_only/mod.rs:68-461).
// All quantities are f64 in SI base units like the reference's diman quantities
// (src/units/mod.rs).  The recursive binary substepping of perform_timestep_internal
// (:394-424) is executed iteratively with a 128-bit path word instead of a call stack.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>

namespace ssw {

// ---- unit factors and constants, src/units/mod.rs:16-108 (built in expression order) ----
namespace units {
constexpr double centimeters = 0.01;
constexpr double years = 3.15576e7;
constexpr double ergs = 1e-7;
constexpr double electron_volts = 1.602176634e-19;
constexpr double cubic_centimeters = 1e-6;
constexpr double centimeters_squared = centimeters * centimeters;
constexpr double cm3_per_s = cubic_centimeters / 1.0;
constexpr double ergs_cm3_per_s = ergs * cm3_per_s;
constexpr double ergs_per_s = ergs / 1.0;
constexpr double BOLTZMANN_CONSTANT = 1.380649e-23;
constexpr double PROTON_MASS = 1.67262192369e-27;
constexpr double GAMMA = 5.0 / 3.0;
constexpr double SIGMA = 2.9580524545305314e-18 * centimeters_squared;  // number weighted cross section
constexpr double PHOTON_AVERAGE_ENERGY = 18.028356312818811 * electron_volts;
constexpr double RYDBERG_CONSTANT = 13.65693 * electron_volts;
}  // namespace units

constexpr int kMaxDepth = 100;           // hydrogen_only/mod.rs:34
constexpr double kXhiiEpsilon = 1e-10;   // hydrogen_only/mod.rs:38
constexpr double kInvEps = 1.0 / DBL_EPSILON;

enum Process : int { kTemperature = 0, kIonizationFraction = 1, kPhotonRate = 2 };

// Non-absorbed fraction over one cell: HydrogenOnly::get_outgoing_rate, hydrogen_only/mod.rs:78-86.
__host__ __device__ inline double non_absorbed_fraction(double density, double xhii, double size) {
    const double nhi = density / units::PROTON_MASS * (1.0 - xhii);
    return exp(-nhi * units::SIGMA * size);
}

struct Solver {  // hydrogen_only/mod.rs:126-135
    double xhii, temperature, density, volume, length, rate, scale_factor;
    bool has_floor;
    double floor_temperature, floor_xhii;

    __host__ __device__ double nh() const { return density / units::PROTON_MASS; }            // :139-141
    __host__ __device__ double nh_ionized() const { return nh() * xhii; }                      // :143-145
    __host__ __device__ double nh_neutral() const { return nh() * (1.0 - xhii); }              // :147-149
    __host__ __device__ double ne() const { return nh_ionized(); }                             // :151-154
    __host__ __device__ double mu() const { return 1.0 / (xhii + 1.0); }                       // :156-159

    __host__ __device__ double collision_fit() const {                                         // :161-164
        const double t = temperature;
        return sqrt(t) / (1.0 + sqrt(t / 1e5)) * exp(-157809.1 / t);
    }
    __host__ __device__ double collision_fit_derivative() const {                              // :166-173
        const double c1 = 1.0 / 1e5, c2 = 157809.1, t = temperature;
        const double s = sqrt(c1 * t);
        return (exp(-c2 / t) * (c1 * c2 * t + 0.5 * s * (2.0 * c2 + t))) /
               (sqrt(t * (t * t)) * s * ((s + 1.0) * (s + 1.0)));
    }
    __host__ __device__ double alpha_b() const {                                               // :175-180
        const double lambda = 315614.0 / temperature;
        return (2.753e-14 * pow(lambda, 1.5) / pow(1.0 + pow(lambda / 2.74, 0.407), 2.242)) *
               units::cm3_per_s;
    }
    __host__ __device__ double alpha_b_derivative() const {                                    // :182-194
        const double lambda = 315614.0 / temperature;
        const double dlambda_dt = -315614.0 / (temperature * temperature);
        const double c1 = 1.0 / 2.74, c2 = 0.407, c3 = 2.242;
        const double p = pow(c1 * lambda, c2);
        const double d = -sqrt(lambda) * pow(p + 1.0, -c3 - 1.0) * (c2 * c3 * p - 1.5 * p - 1.5);
        return ((2.753e-14 * d) * units::cm3_per_s) * dlambda_dt;
    }
    __host__ __device__ double recombination_cooling() const {                                 // :196-202
        const double lambda = 315614.0 / temperature;
        return (3.435e-30 * temperature * pow(lambda, 1.97) /
                pow(1.0 + pow(lambda / 2.25, 0.376), 3.72)) *
               units::ergs_cm3_per_s;
    }
    __host__ __device__ double recombination_cooling_derivative() const {                      // :204-216
        const double c1 = 315614.0, c2 = 1.97, c3 = 0.376, c4 = 3.72, c5 = 2.25;
        const double t = temperature;
        const double p = pow(c1 / (c5 * t), c3);
        const double derivative = pow(1.0 + p, -1.0 - c4) *
                                  (1.0 - 1.0 * c2 + (1.0 - 1.0 * c2 + c3 * c4) * p) * pow(c1 / t, c2);
        return (3.435e-30 * derivative) * units::ergs_cm3_per_s;
    }
    __host__ __device__ double beta() const { return (5.85e-11 * collision_fit()) * units::cm3_per_s; }  // :218-220
    __host__ __device__ double beta_derivative() const {                                       // :222-225
        return (5.85e-11 * collision_fit_derivative()) * units::cm3_per_s;
    }
    __host__ __device__ double collisional_ionization_cooling() const {                        // :227-229
        return (1.27e-21 * collision_fit()) * units::ergs_cm3_per_s;
    }
    __host__ __device__ double collisional_ionization_cooling_derivative() const {             // :231-235
        return (1.27e-21 * collision_fit_derivative()) * units::ergs_cm3_per_s;
    }
    __host__ __device__ double collisional_excitation_cooling() const {                        // :237-242
        const double t = temperature;
        return (7.5e-19 / (1.0 + sqrt(t / 1e5)) * exp(-118348.0 / t)) * units::ergs_cm3_per_s;
    }
    __host__ __device__ double collisional_excitation_cooling_derivative() const {             // :244-253
        const double t = temperature, c1 = 7.5e-19, c2 = 118348.0, c3 = 1.0 / 1e5;
        const double s = sqrt(c3 * t);
        return ((c1 * exp(-c2 / t) * (c2 * c3 * t - 0.5 * c3 * (t * t) + c2 * s)) /
                ((t * t) * s * ((1.0 + s) * (1.0 + s)))) *
               units::ergs_cm3_per_s;
    }
    __host__ __device__ double bremsstrahlung_cooling() const {                                // :255-257
        return (1.42e-27 * sqrt(temperature)) * units::ergs_cm3_per_s;
    }
    __host__ __device__ double bremsstrahlung_cooling_derivative() const {                     // :259-263
        return (1.42e-27 / (2.0 * sqrt(temperature))) * units::ergs_cm3_per_s;
    }
    __host__ __device__ double compton_x4() const {
        const double x = 2.727 / scale_factor;
        const double x2 = x * x;
        return x2 * x2;  // powi(x, 4) by squaring
    }
    __host__ __device__ double compton_cooling() const {                                       // :265-268
        const double x = 2.727 / scale_factor;
        return (1.017e-37 * compton_x4() * (temperature - x)) * units::ergs_per_s;
    }
    __host__ __device__ double compton_cooling_derivative() const {                            // :270-273
        return (1.017e-37 * compton_x4()) * units::ergs_per_s;
    }
    __host__ __device__ double cooling_rate() const {                                          // :275-287
        const double n_e = ne(), n0 = nh_neutral(), n1 = nh_ionized();
        const double collisional =
            (collisional_excitation_cooling() + collisional_ionization_cooling()) * n_e * n0;
        const double recombination = recombination_cooling() * n_e * n1;
        const double bremsstrahlung = bremsstrahlung_cooling() * n_e * n1;
        const double compton = compton_cooling() * n_e;
        return collisional + recombination + bremsstrahlung + compton;
    }
    __host__ __device__ double cooling_rate_derivative() const {                               // :289-302
        const double n_e = ne(), n0 = nh_neutral(), n1 = nh_ionized();
        const double collisional = (collisional_excitation_cooling_derivative() +
                                    collisional_ionization_cooling_derivative()) *
                                   n_e * n0;
        const double recombination = recombination_cooling_derivative() * n_e * n1;
        const double bremsstrahlung = bremsstrahlung_cooling_derivative() * n_e * n1;
        const double compton = compton_cooling_derivative() * n_e;
        return collisional + recombination + bremsstrahlung + compton;
    }
    __host__ __device__ double num_newly_ionized(double timestep) const {                      // :312-319
        const double absorbed_fraction = 1.0 - exp(-nh_neutral() * units::SIGMA * length);
        const double num_photons = timestep * rate;
        return num_photons * absorbed_fraction;
    }
    __host__ __device__ double photoheating_rate(double timestep) const {                      // :321-325
        const double ionization_density = num_newly_ionized(timestep) / volume;
        return ionization_density * (units::PHOTON_AVERAGE_ENERGY - units::RYDBERG_CONSTANT) / timestep;
    }
    __host__ __device__ double photoionization_rate(double timestep) const {                   // :327-332
        const double fraction = num_newly_ionized(timestep) / (nh_neutral() * volume);
        return fraction / timestep;
    }
    __host__ __device__ double temperature_change(double timestep) const {                     // :304-310
        const double k = (units::GAMMA - 1.0) * units::PROTON_MASS / (density * units::BOLTZMANN_CONSTANT);
        const double lambda = photoheating_rate(timestep) - cooling_rate();
        const double dlambdadt = -cooling_rate_derivative();
        const double m = mu();
        return k * m * lambda * timestep / (1.0 - k * m * dlambdadt * timestep);
    }
    __host__ __device__ double ionized_fraction_change(double timestep) const {                // :334-354
        const double n_h = nh(), n_e = ne();
        const double alpha = alpha_b(), dalpha = alpha_b_derivative();
        const double b = beta(), dbeta = beta_derivative();
        const double c = b * n_e + photoionization_rate(timestep);
        const double m = mu();
        const double d = alpha * n_e;
        const double rhsc = n_e * temperature * m * 1.0 * dbeta;
        const double dcdx = n_h * b - rhsc;
        const double rhsd = n_e * temperature * m * 1.0 * dalpha;
        const double dddx = n_h * alpha - rhsd;
        const double j = dcdx - (c + d) - xhii * (dcdx + dddx);
        return timestep * (c - xhii * (c + d)) / (1.0 - j * timestep);
    }
    __host__ __device__ void clamp() {                                                         // :356-369
        const double lo = has_floor ? floor_xhii : kXhiiEpsilon;
        double x = xhii;
        if (x < lo) x = lo;
        if (x > 1.0 - kXhiiEpsilon) x = 1.0 - kXhiiEpsilon;
        xhii = x;
        if (has_floor && temperature < floor_temperature) temperature = floor_temperature;
    }
};

// `update`, hydrogen_only/mod.rs:444-461.  fmin ignores a NaN operand like Rust's f64::min.
__host__ __device__ inline bool update_value(double &value, double change, double max_allowed,
                                             double timestep, double &recommendation) {
    const double relative_change = fmin(fabs(change / value), kInvEps);
    if (relative_change > max_allowed) return false;
    value += change;
    recommendation = timestep * (max_allowed / relative_change);
    return true;
}

struct ChemResult {
    double timescale;
    int process;
    int failed;
    int max_depth;
    unsigned long long attempts;
};

// try_timestep_update, hydrogen_only/mod.rs:371-392
__host__ __device__ inline bool try_timestep_update(Solver &s, double timestep, double safety,
                                                    double &timescale, int &process) {
    double t_rec, x_rec;
    const double dT = s.temperature_change(timestep);
    if (!update_value(s.temperature, dT, safety, timestep, t_rec)) return false;
    const double dx = s.ionized_fraction_change(timestep);
    if (!update_value(s.xhii, dx, safety, timestep, x_rec)) return false;
    s.clamp();
    if (t_rec < x_rec) {  // Timescale::min, chemistry/timescale.rs:32-38
        timescale = t_rec;
        process = kTemperature;
    } else {
        timescale = x_rec;
        process = kIonizationFraction;
    }
    return true;
}

// perform_timestep + perform_timestep_internal, hydrogen_only/mod.rs:394-441, iteratively.
// path bit k = "the first half at depth k is done, the second half is running".
__host__ __device__ inline ChemResult perform_timestep(Solver &s, double timestep, double safety) {
    ChemResult r;
    r.failed = 0;
    r.max_depth = 0;
    r.attempts = 0;
    r.timescale = 0.0;
    r.process = kTemperature;
    unsigned long long path_lo = 0, path_hi = 0;  // bits 0..63, 64..127
    int depth = 0;
    double h = timestep;
    while (true) {
        s.clamp();
        const double t0 = s.temperature, x0 = s.xhii;
        if (depth > kMaxDepth) {  // TimestepConvergenceFailed
            r.failed = 1;
            r.timescale = timestep / 10.0;
            r.process = kTemperature;
            return r;
        }
        if (depth > r.max_depth) r.max_depth = depth;
        r.attempts++;
        if (!try_timestep_update(s, h, safety, r.timescale, r.process)) {
            s.temperature = t0;
            s.xhii = x0;
            depth += 1;
            h = h / 2.0;
            if (depth < 64) path_lo &= ~(1ull << depth);
            else path_hi &= ~(1ull << (depth - 64));
            continue;
        }
        // success: pop every frame whose second half just finished
        while (depth > 0) {
            const bool second = depth < 64 ? ((path_lo >> depth) & 1ull) : ((path_hi >> (depth - 64)) & 1ull);
            if (!second) break;
            depth -= 1;
            h = h * 2.0;
        }
        if (depth == 0) return r;
        if (depth < 64) path_lo |= (1ull << depth);
        else path_hi |= (1ull << (depth - 64));
    }
}

}  // namespace ssw
