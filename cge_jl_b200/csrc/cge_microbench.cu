// cge_microbench.cu -- what the FP64 datapath of this device can do, measured: the denominators
// and design evidence of the recompute regime (north_star (a): "the row-norm/dot form may use FP64
// DMMA only if ncu shows the dot products dominate").  Nothing here is on the scoring path.
//
//   out[0]  DFMA only                      TFLOP/s (2 flop per FMA), 8 independent chains per thread
//   out[1]  DMMA m8n8k4 only               TFLOP/s (512 flop per warp instruction)
//   out[2]  DMMA m16n8k16 only             TFLOP/s (4096 flop per warp instruction)
//   out[3]  mixed: DMMA m8n8k4 share       TFLOP/s  } one DMMA + 2 DFMA per thread and step, to see
//   out[4]  mixed: DFMA share              TFLOP/s  } whether the two share an execution pipe
//   out[5]  mixed m16n8k16 + 8 DFMA: DMMA share
//   out[6]  mixed m16n8k16 + 8 DFMA: DFMA share
//   out[7]  DMMA m16n8k8 only              TFLOP/s
#include "cge_kernels.cuh"

namespace cge {

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
        "{%0,%1,%2,%3};"
        : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
        : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, "
        "{%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
        : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
        : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
          "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// MODE 0: DFMA; 1: m8n8k4; 2: m16n8k16; 3: m8n8k4 + NF DFMA per DMMA; 4: m16n8k16 + NF DFMA; 5: m16n8k8
template <int MODE, int NF>
__global__ void __launch_bounds__(256) k_pipe(double *out, int iters, double x) {
    const double t = threadIdx.x * 1e-3;
    double f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = t + i;
    double c2[8][2], c4[4][4], a8[8], b4[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        c2[i][0] = c2[i][1] = t;
        a8[i] = x + 1e-6 * i;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        b4[i] = x - 1e-6 * i;
#pragma unroll
        for (int j = 0; j < 4; ++j) c4[i][j] = t;
    }
    for (int it = 0; it < iters; ++it) {
        if constexpr (MODE == 0) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = fma(f[i], x, 1e-9);
        } else if constexpr (MODE == 1 || MODE == 3) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                dmma884(c2[i], a8[i], b4[i & 3]);
#pragma unroll
                for (int r = 0; r < NF; ++r) f[(i * NF + r) & 7] = fma(f[(i * NF + r) & 7], x, 1e-9);
            }
        } else if constexpr (MODE == 2 || MODE == 4) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                dmma16816(c4[i], a8, b4);
#pragma unroll
                for (int r = 0; r < NF; ++r) f[r & 7] = fma(f[r & 7], x, 1e-9);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double a4[4] = {a8[0], a8[1], a8[2], a8[3]};
                const double b2[2] = {b4[0], b4[1]};
                dmma1688(c4[i], a4, b2);
            }
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += f[i] + c2[i][0] + c2[i][1];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += c4[i][j];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int NF>
static float time_pipe(int blocks, int iters, double *buf, cudaStream_t st) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_pipe<MODE, NF><<<blocks, 256, 0, st>>>(buf, 64, 0.999999);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0, st);
        k_pipe<MODE, NF><<<blocks, 256, 0, st>>>(buf, iters, 0.999999);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return best;
}

int measure_fp64_pipes(int sm_count, cudaStream_t st, double *out) {
    const int blocks = sm_count * 8, iters = 1 << 13;
    double *buf = nullptr;
    if (cudaMalloc(&buf, (size_t)blocks * 256 * 8) != cudaSuccess) return -1;
    const double warps = (double)blocks * 8.0, thr = (double)blocks * 256.0, n = (double)iters;
    float ms;
    ms = time_pipe<0, 0>(blocks, iters, buf, st);
    out[0] = 2.0 * 32.0 * n * thr / (ms * 1e-3) / 1e12;
    ms = time_pipe<1, 0>(blocks, iters, buf, st);
    out[1] = 512.0 * 8.0 * n * warps / (ms * 1e-3) / 1e12;
    ms = time_pipe<2, 0>(blocks, iters, buf, st);
    out[2] = 4096.0 * 4.0 * n * warps / (ms * 1e-3) / 1e12;
    ms = time_pipe<3, 2>(blocks, iters, buf, st);
    out[3] = 512.0 * 8.0 * n * warps / (ms * 1e-3) / 1e12;
    out[4] = 2.0 * 16.0 * n * thr / (ms * 1e-3) / 1e12;
    ms = time_pipe<4, 8>(blocks, iters, buf, st);
    out[5] = 4096.0 * 4.0 * n * warps / (ms * 1e-3) / 1e12;
    out[6] = 2.0 * 32.0 * n * thr / (ms * 1e-3) / 1e12;
    ms = time_pipe<5, 0>(blocks, iters, buf, st);
    out[7] = 2048.0 * 4.0 * n * warps / (ms * 1e-3) / 1e12;
    cudaFree(buf);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace cge
