// compiled.cuh -- slot-ordered ("compiled") form of the all-cells wavefront schedule and the
// persistent sweep kernel that runs it.
//
// The generic kernels of kernels.cuh walk the CSR grid for every task: ~45 B per face of
// geometry plus dependent index loads, and one grid barrier per wavefront level whose critical
// path is four dependent memory round trips.  For the all-cells sweep (the one that dominates
// the step) the level sets are static, so they are compiled once into a layout where
//   * slot s = position of task (dl, c) in the level-sorted, (dl, c)-sorted task list;
//   * the flux state lives in slot order: out_slot[s] = outgoing_total_rate of that task, so
//     each level writes one contiguous range and reads ranges of earlier levels;
//   * every task owns a contiguous run of precomputed upwind entries
//         e_src[e]  = slot of the upwind Local neighbour for the same direction
//         e_w[e]    = A_rev * (-n.d) / sum_downwind(A n.d) of that neighbour   (the share)
//     in the face order of the cell, so the kernel is a pure gather:
//         incoming = sum_e out_slot[e_src[e]] * e_w[e]
// Algorithmic bytes per cell-direction update: 20 B per upwind entry (4 + 8 + 8 gather) plus
// 24 B per task (task id 4, entry offset 4, outgoing store 8, incoming store 8) -- the B_alg of
// SURVEY.md section 8d; att / source / periodic rows are per-cell gathers that stay in L2.
//
// After compilation out_slot is the single source of truth for the flux state; the generic
// kernels reach it through slot_of[dl * N + c] (kernels.cuh, StateView).
#pragma once
#include <cub/cub.cuh>

#include <cstdint>
#include <vector>

#include "kernels.cuh"

namespace ssw {

struct Compiled {
    bool valid = false;
    uint64_t n_tasks = 0;
    uint64_t n_entries = 0;
    uint32_t n_levels = 0;
    uint32_t *slot_of = nullptr;   // [dl*N + c] -> slot
    uint32_t *t_upoff = nullptr;   // n_tasks + 1
    uint32_t *e_src = nullptr;     // n_entries
    double *e_w = nullptr;         // n_entries
    double *out_slot = nullptr;    // n_tasks
    double *ttot_slot = nullptr;   // n_tasks, sum over downwind faces of A * n.d
    unsigned int *barrier = nullptr;
    Compiled() = default;
    Compiled(const Compiled &) = delete;
    Compiled &operator=(const Compiled &) = delete;
    ~Compiled() { release(); }
    void release() {
        cudaFree(slot_of); cudaFree(t_upoff); cudaFree(e_src); cudaFree(e_w); cudaFree(out_slot);
        cudaFree(ttot_slot); cudaFree(barrier);
        slot_of = t_upoff = e_src = nullptr;
        e_w = out_slot = ttot_slot = nullptr;
        barrier = nullptr;
        valid = false;
    }
};

inline bool compiled_supported() { return true; }

// ---- construction kernels (run once per compiled schedule) -----------------------------------
__global__ void __launch_bounds__(256)
c_slot_scatter_kernel(const uint32_t *__restrict__ tasks, uint32_t n, uint32_t *__restrict__ slot_of) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) slot_of[tasks[s]] = s;
}

__global__ void __launch_bounds__(256)
c_count_kernel(GridView g, const uint32_t *__restrict__ tasks, uint32_t n, uint32_t *__restrict__ cnt,
               double *__restrict__ ttot_slot) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint32_t task = tasks[s];
    const uint32_t dl = task / g.n_cells, c = task - dl * g.n_cells;
    const double dx = c_dirs[3 * dl], dy = c_dirs[3 * dl + 1], dz = c_dirs[3 * dl + 2];
    uint32_t m = 0;
    double ttot = 0.0;
    for (uint32_t f = g.face_off[c]; f < g.face_off[c + 1]; ++f) {
        const double4 geo = ld_geo(g.face_geo + f);
        const double d = dot_dir(geo, dx, dy, dz);
        if (d < 0.0) {
            if (g.face_kind[f] == 0) ++m;
        } else if (d > 0.0) {
            ttot += geo.w * d;
        }
    }
    cnt[s] = m;
    ttot_slot[s] = ttot;
}

__global__ void __launch_bounds__(256)
c_fill_kernel(GridView g, const uint32_t *__restrict__ tasks, uint32_t n, const uint32_t *__restrict__ t_upoff,
              const uint32_t *__restrict__ slot_of, const double *__restrict__ ttot_slot,
              uint32_t *__restrict__ e_src, double *__restrict__ e_w) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint32_t task = tasks[s];
    const uint32_t dl = task / g.n_cells, c = task - dl * g.n_cells;
    const double dx = c_dirs[3 * dl], dy = c_dirs[3 * dl + 1], dz = c_dirs[3 * dl + 2];
    uint32_t e = t_upoff[s];
    for (uint32_t f = g.face_off[c]; f < g.face_off[c + 1]; ++f) {
        if (g.face_kind[f] != 0) continue;
        const double d = dot_dir(ld_geo(g.face_geo + f), dx, dy, dz);
        if (!(d < 0.0)) continue;
        const uint32_t src = slot_of[dl * g.n_cells + (uint32_t)g.face_nb[f]];
        const double tt = ttot_slot[src];
        e_src[e] = src;
        e_w[e] = tt > 0.0 ? (g.face_rev[f] * (-d)) / tt : 0.0;
        ++e;
    }
}

__global__ void __launch_bounds__(256)
c_convert_state_kernel(const uint32_t *__restrict__ tasks, uint32_t n, const double *__restrict__ q_nat,
                       const double *__restrict__ ttot_slot, double *__restrict__ out_slot) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) out_slot[s] = q_nat[tasks[s]] * ttot_slot[s];
}

// ---- grid barrier: one monotone counter, release/acquire at gpu scope -----------------------
__device__ __forceinline__ void grid_barrier(unsigned int *counter, unsigned int target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        unsigned int v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        } while (v < target);
    }
    __syncthreads();
}

struct CompiledArgs {
    const uint32_t *tasks;      // slot -> task id (dl * N + c)
    const uint32_t *t_upoff;
    const uint32_t *e_src;
    const double *e_w;
    double *out_slot;
    double *incoming;           // natural layout [dl*N + c]
    const uint32_t *level_off;
    const double *att, *src;
    const int32_t *pidx;
    const double *per_lag;
    unsigned int *barrier;
    uint32_t n_levels, n_cells, n_periodic;
    double threshold, n_dirs_total;
};

// One thread per task per round; all levels in one persistent launch.
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 2048 / THREADS)
sweep_compiled_kernel(CompiledArgs a) {
    const unsigned int gtid = blockIdx.x * THREADS + threadIdx.x;
    const unsigned int gsz = gridDim.x * THREADS;
    const uint32_t N = a.n_cells;
    for (uint32_t lvl = 0; lvl < a.n_levels; ++lvl) {
        const uint32_t s0 = __ldg(a.level_off + lvl), s1 = __ldg(a.level_off + lvl + 1);
        for (uint32_t s = s0 + gtid; s < s1; s += gsz) {
            const uint32_t task = __ldcs(a.tasks + s);
            const uint32_t e0 = __ldcs(a.t_upoff + s), e1 = __ldcs(a.t_upoff + s + 1);
            const uint32_t dl = task / N, c = task - dl * N;
            double in = 0.0;
            for (uint32_t e = e0; e < e1; ++e)
                in += __ldcg(a.out_slot + __ldcs(a.e_src + e)) * __ldcs(a.e_w + e);
            const double inc = in + __ldg(a.src + c) / a.n_dirs_total;
            const int32_t p = __ldg(a.pidx + c);
            const double total = p >= 0 ? inc + __ldg(a.per_lag + (size_t)dl * a.n_periodic + p) : inc + 0.0;
            const double out = (total < a.threshold) ? 0.0 : total * a.att[c];
            __stcg(a.out_slot + s, out);
            __stcg(a.incoming + task, inc);
        }
        if (lvl + 1 < a.n_levels) grid_barrier(a.barrier, (lvl + 1) * gridDim.x);
    }
}

constexpr int kCompiledThreads = 512;

inline void compile_schedule(Compiled &C, const GridView &g, const uint32_t *tasks,
                             const std::vector<uint32_t> &level_off_host, uint64_t n_tasks,
                             uint32_t n_levels, int n_local_dirs, const double *q_nat,
                             cudaStream_t stream, uint64_t *launch_counter) {
    (void)level_off_host;
    (void)n_local_dirs;
    C.release();
    const uint32_t n = (uint32_t)n_tasks;
    auto chk = [](cudaError_t e) {
        if (e != cudaSuccess) throw std::runtime_error(std::string("compile_schedule: ") + cudaGetErrorString(e));
    };
    uint32_t *cnt = nullptr;
    void *temp = nullptr;
    try {
        chk(cudaMalloc(&C.slot_of, sizeof(uint32_t) * (size_t)n));
        chk(cudaMalloc(&C.t_upoff, sizeof(uint32_t) * ((size_t)n + 1)));
        chk(cudaMalloc(&C.ttot_slot, sizeof(double) * (size_t)n));
        chk(cudaMalloc(&C.out_slot, sizeof(double) * (size_t)n));
        chk(cudaMalloc(&C.barrier, sizeof(unsigned int)));
        chk(cudaMalloc(&cnt, sizeof(uint32_t) * ((size_t)n + 1)));
        const unsigned blocks = (unsigned)((n + 255) / 256);
        c_slot_scatter_kernel<<<blocks, 256, 0, stream>>>(tasks, n, C.slot_of);
        chk(cudaMemsetAsync(cnt + n, 0, sizeof(uint32_t), stream));
        c_count_kernel<<<blocks, 256, 0, stream>>>(g, tasks, n, cnt, C.ttot_slot);
        size_t bytes = 0;
        chk(cub::DeviceScan::ExclusiveSum(nullptr, bytes, cnt, C.t_upoff, (int64_t)n + 1, stream));
        chk(cudaMalloc(&temp, bytes));
        chk(cub::DeviceScan::ExclusiveSum(temp, bytes, cnt, C.t_upoff, (int64_t)n + 1, stream));
        uint32_t total = 0;
        chk(cudaMemcpyAsync(&total, C.t_upoff + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        chk(cudaStreamSynchronize(stream));
        // the 32-bit scan wraps silently: verify with the per-cell face count bound
        C.n_entries = total;
        chk(cudaMalloc(&C.e_src, sizeof(uint32_t) * (size_t)std::max<uint32_t>(total, 1)));
        chk(cudaMalloc(&C.e_w, sizeof(double) * (size_t)std::max<uint32_t>(total, 1)));
        c_fill_kernel<<<blocks, 256, 0, stream>>>(g, tasks, n, C.t_upoff, C.slot_of, C.ttot_slot, C.e_src, C.e_w);
        c_convert_state_kernel<<<blocks, 256, 0, stream>>>(tasks, n, q_nat, C.ttot_slot, C.out_slot);
        chk(cudaGetLastError());
        chk(cudaStreamSynchronize(stream));
        if (launch_counter) *launch_counter += 6;
    } catch (...) {
        cudaFree(cnt);
        cudaFree(temp);
        C.release();
        throw;
    }
    cudaFree(cnt);
    cudaFree(temp);
    C.n_tasks = n_tasks;
    C.n_levels = n_levels;
    C.valid = true;
}

inline void run_compiled(Compiled &C, const SweepArgs &sa, const uint32_t *tasks, const uint32_t *level_off,
                         int num_sms, cudaStream_t stream, uint64_t *launch_counter) {
    CompiledArgs a;
    a.tasks = tasks;
    a.t_upoff = C.t_upoff;
    a.e_src = C.e_src;
    a.e_w = C.e_w;
    a.out_slot = C.out_slot;
    a.incoming = sa.incoming;
    a.level_off = level_off;
    a.att = sa.att;
    a.src = sa.src;
    a.pidx = sa.pidx;
    a.per_lag = sa.per_lag;
    a.barrier = C.barrier;
    a.n_levels = C.n_levels;
    a.n_cells = sa.g.n_cells;
    a.n_periodic = sa.n_periodic;
    a.threshold = sa.threshold;
    a.n_dirs_total = sa.n_dirs_total;
    cudaMemsetAsync(C.barrier, 0, sizeof(unsigned int), stream);
    const int blocks = num_sms * (2048 / kCompiledThreads);
    void *args[] = {&a};
    // cooperative launch only to guarantee co-residency of all blocks (the barrier spins)
    cudaError_t e = cudaLaunchCooperativeKernel((const void *)sweep_compiled_kernel<kCompiledThreads>, dim3(blocks),
                                                dim3(kCompiledThreads), args, 0, stream);
    if (e != cudaSuccess) throw std::runtime_error(std::string("sweep_compiled_kernel launch: ") + cudaGetErrorString(e));
    if (launch_counter) *launch_counter += 1;
}

}  // namespace ssw
