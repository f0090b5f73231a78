// cge_inst.cu -- instantiates the tile kernels for five exponents per translation unit
// (compiled eight times with -DCGE_PART=0..7 so the instantiations build in parallel).
#include "cge_kernels.cuh"
#include "cge_ring.cuh"

#ifndef CGE_PART
#error "compile with -DCGE_PART=0..7"
#endif

namespace cge {

#define CGE_CASE(MM)                                                         \
    case MM:                                                                 \
        switch (kind) {                                                      \
            case 0: k_sweep<MM, false><<<grid, NTHREADS, 0, stream>>>(a); break;  \
            case 1: k_sweep<MM, true><<<grid, NTHREADS, 0, stream>>>(a); break;   \
            case 2: k_bsweep<MM, false><<<grid, NTHREADS, 0, stream>>>(a); break; \
            case 4: launch_bfp<MM>(grid, stream, a); break;                   \
            default: k_bsweep<MM, true><<<grid, NTHREADS, 0, stream>>>(a); break; \
        }                                                                    \
        break;

#define CGE_CAT_(a, b) a##b
#define CGE_CAT(a, b) CGE_CAT_(a, b)

void CGE_CAT(launch_tiles_part, CGE_PART)(int m, int kind, int grid, cudaStream_t stream,
                                          const SweepArgs &a) {
    switch (m) {
        CGE_CASE(CGE_PART * 5 + 1)
        CGE_CASE(CGE_PART * 5 + 2)
        CGE_CASE(CGE_PART * 5 + 3)
        CGE_CASE(CGE_PART * 5 + 4)
        CGE_CASE(CGE_PART * 5 + 5)
        default: break;
    }
}

#define CGE_FP_CASE(MM)                                                       \
    case MM:                                                                  \
        return directed ? (const void *)k_fixed_point<MM, true>               \
                        : (const void *)k_fixed_point<MM, false>;

const void *CGE_CAT(fp_kernel_part, CGE_PART)(int m, int directed) {
    switch (m) {
        CGE_FP_CASE(CGE_PART * 5 + 1)
        CGE_FP_CASE(CGE_PART * 5 + 2)
        CGE_FP_CASE(CGE_PART * 5 + 3)
        CGE_FP_CASE(CGE_PART * 5 + 4)
        CGE_FP_CASE(CGE_PART * 5 + 5)
        default: return nullptr;
    }
}

#define CGE_RING_CASE(MM)                                                     \
    case MM:                                                                  \
        return directed ? (const void *)k_fixed_point_ring<MM, true>          \
                        : (const void *)k_fixed_point_ring<MM, false>;

const void *CGE_CAT(fp_ring_kernel_part, CGE_PART)(int m, int directed) {
    switch (m) {
        CGE_RING_CASE(CGE_PART * 5 + 1)
        CGE_RING_CASE(CGE_PART * 5 + 2)
        CGE_RING_CASE(CGE_PART * 5 + 3)
        CGE_RING_CASE(CGE_PART * 5 + 4)
        CGE_RING_CASE(CGE_PART * 5 + 5)
        default: return nullptr;
    }
}

}  // namespace cge
