/* c_abi_smoke.c -- a plain C program against include/subsweep_b200.h: the header itself (not its ctypes twin) is
 * compiled and linked with libsubsweep_b200.so, and drives create -> run -> read on a 4^3 periodic Cartesian grid
 * built the way src/sweep/grid/cartesian.rs builds it (faces in the order -x,+x,-y,+y,-z,+z; area h^2, size h,
 * volume h^3).  Built by __graft_entry__.build(), run by tests/test_gpu_c_abi.py under -m gpu.
 * Prints "ok <mean xHII> <tasks>" and exits 0, or a message and a non-zero code. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "subsweep_b200.h"

#define N1 4
#define NC (N1 * N1 * N1)

static int idx(int i, int j, int k) { return ((i + N1) % N1 * N1 + (j + N1) % N1) * N1 + (k + N1) % N1; }

int main(void) {
    if (ssw_abi_version() != SSW_ABI_VERSION) { fprintf(stderr, "ABI version mismatch\n"); return 2; }
    static uint64_t off[NC + 1];
    static double area[6 * NC], normal[18 * NC], size[NC], volume[NC], rho[NC], x[NC], T[NC], src[NC], out[NC];
    static int32_t nb[6 * NC];
    static uint8_t kind[6 * NC];
    const double h = 3.0857e19 * 5.0;   /* 5 kpc */
    const int step[6][3] = {{-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1}};
    for (int i = 0; i < N1; ++i)
        for (int j = 0; j < N1; ++j)
            for (int k = 0; k < N1; ++k) {
                const int c = idx(i, j, k);
                off[c] = 6u * (uint64_t)c;
                for (int f = 0; f < 6; ++f) {
                    const int ii = i + step[f][0], jj = j + step[f][1], kk = k + step[f][2];
                    const int wraps = ii < 0 || ii >= N1 || jj < 0 || jj >= N1 || kk < 0 || kk >= N1;
                    area[6 * c + f] = h * h;
                    for (int a = 0; a < 3; ++a) normal[18 * c + 3 * f + a] = step[f][a];
                    nb[6 * c + f] = idx(ii, jj, kk);
                    kind[6 * c + f] = wraps ? SSW_FACE_LOCAL_PERIODIC : SSW_FACE_LOCAL;
                }
                size[c] = h; volume[c] = h * h * h;
                rho[c] = 1e-4 * 1e6 * 1.67262192369e-27; x[c] = 1e-10; T[c] = 100.0; src[c] = 0.0;
            }
    off[NC] = 6u * NC;
    src[idx(1, 2, 3)] = 1e51;
    const double dirs[6 * 3] = {1, 0, 0, -1, 0, 0, 0, 1, 0, 0, -1, 0, 0.6, 0.48, 0.64, -0.6, -0.48, -0.64};
    ssw_params p = {0};
    p.n_dirs = 6; p.dirs_xyz = dirs; p.n_levels = 2; p.max_timestep_s = 3.15576e12; p.timestep_safety_factor = 0.1;
    p.chemistry_timestep_safety_factor = 0.1; p.significant_rate_threshold_per_s = 1e-5; p.prevent_cooling = 1;
    p.scale_factor = 1.0; p.world_size = 1;
    ssw_grid g = {NC, off, area, normal, nb, kind, size, volume};
    ssw_handle *hnd = NULL;
    if (ssw_create(&p, &g, rho, x, T, src, &hnd) != SSW_OK) { fprintf(stderr, "ssw_create: %s\n", ssw_last_error()); return 3; }
    double elapsed = 0.0, total = 0.0;
    for (int s = 0; s < 3; ++s) {
        if (ssw_run_sweeps(hnd, &elapsed) != SSW_OK) { fprintf(stderr, "ssw_run_sweeps: %s\n", ssw_last_error()); return 4; }
        total += elapsed;
    }
    if (ssw_read(hnd, SSW_F_XHII, out) != SSW_OK) { fprintf(stderr, "ssw_read: %s\n", ssw_last_error()); return 5; }
    double mean = 0.0;
    for (int c = 0; c < NC; ++c) {
        if (!(out[c] >= 1e-10 && out[c] <= 1.0)) { fprintf(stderr, "xHII out of range in cell %d: %g\n", c, out[c]); return 6; }
        mean += out[c] / NC;
    }
    uint64_t counts[2], tasks = 0;
    if (ssw_level_counts(hnd, counts) != SSW_OK || counts[0] != NC) { fprintf(stderr, "level counts\n"); return 7; }
    ssw_get_stat(hnd, SSW_STAT_TASKS_SOLVED, &tasks);
    if (!(out[idx(1, 2, 3)] > 1e-10) || tasks < 3u * NC * 6u || fabs(total - 2.0 * p.max_timestep_s) > 1.0   /* 1/2 + 1/2 + 1: two levels unlock in the first calls */) {
        fprintf(stderr, "unexpected result: x_src %g tasks %llu elapsed %g\n", out[idx(1, 2, 3)], (unsigned long long)tasks, total);
        return 8;
    }
    ssw_destroy(hnd);
    printf("ok %.17g %llu\n", mean, (unsigned long long)tasks);
    return 0;
}
