// small.cuh -- a single sweep over a SMALL active set in one launch.
//
// After the first step most cells of a reionisation box sit at timestep level 0; the higher levels hold a few dozen
// cells around the sources, and a step sweeps them 2^(L-1) - 1 times (src/sweep/timestep_state.rs:22-27).  As separate
// kernels (lagged periodic rows, replay of the cached level sets, photon_rate bookkeeping, new periodic rows, rate fold,
// and under direction sharding the exchange of the partial rates) such a sweep is a chain of six launches that each do
// microseconds of work: at 8 GPUs the seven sub-level sweeps of a step cost more than the all-cells chemistry.  Here one
// thread block does all of it, phase by phase with block barriers, and under peer-mapped sharding (peer.cuh) also pushes
// the partial rates to their owners and signals them.  Per task and per cell the arithmetic is that of the separate
// kernels (same device functions, same summation orders): results are bit-identical.
#pragma once

namespace ssw {

constexpr int kSmallPhotonScratch = 6144;   // (touched cell, direction) terms the block can hold in shared memory

struct SmallSweepArgs {
    SweepArgs a;
    MiniView m;
    const uint32_t *queue, *level_off;
    uint32_t n_levels;
    const uint32_t *act;
    uint32_t n_act;
    double *per_lag, *per_new;      // [dl][periodic row]
    const uint32_t *touch;          // Local neighbours of the active cells (photon_rate bookkeeping), or null
    uint32_t n_touch;
    double *photon;
    double *rate_act;               // out: [k] partial rate of active cell k over this rank's directions
    int n_local_dirs;
    int peers;                      // push the rates to the owners and signal (peer.cuh)
};

// sum over the LocalPeriodic upwind faces of (cell, dl): handle_local_periodic_neighbour in gather form
__device__ __forceinline__ double small_periodic_row(const GridView &g, const StateView &st, uint32_t c, int dl) {
    const double dx = g.dirs[3 * dl], dy = g.dirs[3 * dl + 1], dz = g.dirs[3 * dl + 2];
    double acc = 0.0;
    for (uint32_t f = g.face_off[c]; f < g.face_off[c + 1]; ++f) {
        if (g.face_kind[f] != 2) continue;
        const double d = dot_dir(ld_geo(g.face_geo + f), dx, dy, dz);
        if (d < 0.0) acc += st.load_q(dl, (uint32_t)g.face_nb[f]) * (g.face_rev[f] * (-d));
    }
    return acc;
}

// incoming_total_rate of (cell, dl) from the neighbours' current outgoing rates (photon_patch_kernel's term)
__device__ __forceinline__ double small_incoming(const GridView &g, const StateView &st, uint32_t c, int dl) {
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
    return acc;
}

__global__ void __launch_bounds__(kMiniSmallThreads)
small_sweep_kernel(SmallSweepArgs s, PeerTable pt) {
    __shared__ double s_term[kSmallPhotonScratch];
    const GridView &g = s.a.g;
    const StateView &st = s.a.st;
    const uint32_t T = blockDim.x, tid = threadIdx.x;
    const int Dl = s.n_local_dirs;
    const uint32_t n_rows = s.n_act * (uint32_t)Dl;
    // (a) periodic_source as the tasks of this sweep read it: the rows of the active cells, lagged
    if (s.a.n_periodic) {
        for (uint32_t i = tid; i < n_rows; i += T) {
            const uint32_t k = i / Dl;
            const int dl = (int)(i - k * Dl);
            const uint32_t c = s.act[k];
            const int32_t p = s.a.pidx[c];
            if (p >= 0) s.per_lag[(size_t)dl * s.a.n_periodic + p] = small_periodic_row(g, st, c, dl);
        }
        __syncthreads();
    }
    // (b) the sweep: cached level sets, block barrier between wavefront levels
    mini_replay_small_body(s.a, s.m, s.queue, s.level_off, s.n_levels);
    __syncthreads();
    // (c) photon_rate of the cells whose incoming rates this sweep changed: terms in parallel, fold in direction order
    if (s.touch) {   // the host passes the list only while n_touch * Dl fits the scratch
        const uint32_t n_terms = s.n_touch * (uint32_t)Dl;
        for (uint32_t i = tid; i < n_terms; i += T) {
            const uint32_t j = i / Dl;
            s_term[i] = small_incoming(g, st, s.touch[j], (int)(i - j * Dl));
        }
        __syncthreads();
        for (uint32_t j = tid; j < s.n_touch; j += T) {
            double total = 0.0;
            for (int dl = 0; dl < Dl; ++dl) total += s_term[j * Dl + dl];
            s.photon[s.touch[j]] = total;
        }
    }
    // (d) periodic_source of this sweep's rates (they enter the rate, site.rs:53-56)
    if (s.a.n_periodic) {
        for (uint32_t i = tid; i < n_rows; i += T) {
            const uint32_t k = i / Dl;
            const int dl = (int)(i - k * Dl);
            const uint32_t c = s.act[k];
            const int32_t p = s.a.pidx[c];
            if (p >= 0) s.per_new[(size_t)dl * s.a.n_periodic + p] = small_periodic_row(g, st, c, dl);
        }
    }
    __syncthreads();
    // (e) rate fold over this rank's directions, in direction order (rate_kernel), and (f) hand-over to the owners
    for (uint32_t k = tid; k < s.n_act; k += T) {
        const uint32_t c = s.act[k];
        const int32_t p = s.a.pidx[c];
        double rate = 0.0;
        for (int dl = 0; dl < Dl; ++dl) {
            const double per = p >= 0 ? s.per_new[(size_t)dl * s.a.n_periodic + p] : 0.0;
            rate += __ldcg(s.a.incoming + (size_t)dl * g.n_cells + c) + per;
        }
        s.rate_act[k] = rate;
        if (s.peers) {
            const uint32_t owner = c / pt.n_per;
            pt.f64((int)owner, pt.L.recv)[(uint64_t)pt.rank * pt.n_per + (c - owner * pt.n_per)] = rate;
        }
    }
    if (s.peers) {
        __syncthreads();
        if (tid == 0) {
            __threadfence_system();
            for (int p = 0; p < pt.world; ++p) peer_arrive(pt, p);
        }
    }
}

}  // namespace ssw
