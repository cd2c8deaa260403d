// compiled.cuh -- slot-ordered ("compiled") form of a cached wavefront schedule.
// Placeholder until the general path is validated on hardware; see DESIGN.md section 5.
#pragma once
#include <cstdint>
#include <vector>

#include "kernels.cuh"

namespace ssw {

struct Compiled {
    bool valid = false;
    void release() { valid = false; }
};

inline bool compiled_supported() { return false; }

inline void compile_schedule(Compiled &, const GridView &, const uint32_t *, const std::vector<uint32_t> &,
                             uint64_t, uint32_t, int, const int32_t *, cudaStream_t, uint64_t *) {}

inline void run_compiled(Compiled &, const SweepArgs &, int, cudaStream_t, uint64_t *) {}

}  // namespace ssw
