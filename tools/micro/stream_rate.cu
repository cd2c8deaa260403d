// Micro-benchmark: how fast can ONE thread block per SM stream a private, sequential byte range from HBM into shared memory?
//   mode 0: cp.async.bulk (TMA 1-D bulk copy) ring, one producer lane           mode 1: cp.async 16 B (LDGSTS) by a loader warp group
//   mode 2: plain 16-byte loads into registers by all threads (no shared memory)
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o stream_rate stream_rate.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__global__ void __launch_bounds__(256, 1) stream_kernel(const unsigned char *data, size_t bytes_per_block, uint32_t chunk, uint32_t stages, int mode,
                                                      int split, unsigned long long *sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)stages * chunk);
    uint64_t *empty = full + 16;
    const unsigned char *mine = data + (size_t)blockIdx.x * bytes_per_block;
    const uint32_t n_chunks = (uint32_t)(bytes_per_block / chunk);
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned long long acc = 0;
    if (mode == 2) {
        const uint4 *p = reinterpret_cast<const uint4 *>(mine);
        const size_t n16 = bytes_per_block / 16;
        for (size_t i = tid; i < n16; i += 256 * 4) {
            uint4 v0 = p[i], v1 = i + 256 < n16 ? p[i + 256] : v0, v2 = i + 512 < n16 ? p[i + 512] : v0, v3 = i + 768 < n16 ? p[i + 768] : v0;
            acc += v0.x ^ v1.y ^ v2.z ^ v3.w;
        }
        if (acc == 0x1234567) sink[0] = acc;
        return;
    }
    if (tid == 0) {
        for (uint32_t s = 0; s < stages; ++s) { mbar_init(smem_u32(full + s), mode == 0 ? 1 : 32 * 2); mbar_init(smem_u32(empty + s), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (mode == 0) {
        if (warp == 7) {
            if (lane == 0) {
                uint32_t stage = 0, use = 0;
                for (uint32_t k = 0; k < n_chunks; ++k) {
                    if (use > 0) mbar_wait(smem_u32(empty + stage), (use - 1) & 1);
                    mbar_expect_tx(smem_u32(full + stage), chunk);
                    const uint32_t part = chunk / split;
                    for (int q = 0; q < split; ++q)
                        tma_bulk_load(smem_u32(smem + (size_t)stage * chunk + q * part), mine + (size_t)k * chunk + q * part, part, smem_u32(full + stage));
                    if (++stage == stages) { stage = 0; ++use; }
                }
            }
            return;
        }
    } else {
        if (warp >= 6) {   // two loader warps, 16-byte cp.async
            uint32_t stage = 0, use = 0;
            for (uint32_t k = 0; k < n_chunks; ++k) {
                if (use > 0) mbar_wait(smem_u32(empty + stage), (use - 1) & 1);
                const uint32_t dst = smem_u32(smem + (size_t)stage * chunk);
                const unsigned char *src = mine + (size_t)k * chunk;
                for (uint32_t o = ((warp - 6) * 32 + lane) * 16; o < chunk; o += 64 * 16)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + o), "l"(src + o) : "memory");
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(full + stage)) : "memory");
                if (++stage == stages) { stage = 0; ++use; }
            }
            return;
        }
    }
    // consumers: wait, touch one word, release
    const uint32_t n_cons = mode == 0 ? 224 : 192;
    uint32_t stage = 0, parity = 0;
    for (uint32_t k = 0; k < n_chunks; ++k) {
        mbar_wait(smem_u32(full + stage), parity);
        acc += smem[(size_t)stage * chunk + tid * 16];
        asm volatile("bar.sync 1, %0;" ::"r"(n_cons) : "memory");
        if (tid == 0) mbar_arrive(smem_u32(empty + stage));
        if (++stage == stages) { stage = 0; parity ^= 1; }
    }
    if (acc == 0x1234567) sink[0] = acc;
}

int main(int argc, char **argv) {
    const size_t total = (size_t)16 << 30;   // 16 GiB
    unsigned char *data; unsigned long long *sink;
    cudaMalloc(&data, total); cudaMemset(data, 1, total); cudaMalloc(&sink, 8);
    cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int grids[] = {42, 84, 148};
    struct Cfg { int mode; uint32_t chunk, stages; int split; } cfgs[] = {
        {0, 8192, 8, 1}, {0, 21504, 4, 1}, {0, 21504, 8, 1}, {0, 21504, 8, 4}, {0, 65536, 3, 1}, {0, 65536, 3, 8},
        {1, 8192, 8, 1}, {1, 21504, 8, 1}, {1, 65536, 3, 1}, {2, 21504, 1, 1}};
    for (int g : grids)
        for (auto c : cfgs) {
            size_t per_block = ((size_t)(1536u << 20) / g) / c.chunk * c.chunk;   // ~1.5 GiB per launch in total... enough to leave L2
            per_block = per_block / (c.chunk) * c.chunk;
            const size_t smem = (size_t)c.stages * c.chunk + 256;
            float best = 1e9f;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(a);
                stream_kernel<<<g, 256, c.mode == 2 ? 0 : smem>>>(data + (size_t)rep * ((size_t)4 << 30), per_block, c.chunk, c.stages, c.mode, c.split, sink);
                cudaEventRecord(b); cudaEventSynchronize(b);
                float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
            }
            cudaError_t e = cudaGetLastError();
            const double gbs = (double)per_block * g / (best * 1e-3) / 1e9;
            printf("grid %3d mode %d chunk %6u stages %u split %d : %8.3f ms  %8.1f GB/s total  %6.1f GB/s per SM  %s\n", g, c.mode, c.chunk, c.stages, c.split,
                   best, gbs, gbs / g, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    return 0;
}
