// cge_host.cu -- C ABI (include/cge_b200.h), host driver of the alpha loop and the small
// kernels around the pair-matrix sweeps (distance tiles, normalisation, finalize, local score).
//
// The control flow mirrors /root/reference/src/divergence.jl:27-257 (wGCL) and :282-561
// (wGCL_directed); every step cites the lines it replaces.  Nothing here computes on the CPU
// except O(m) / O(k^2) bookkeeping (C vector, degrees, star check, Jensen-Shannon over <= k^2
// bins, early-stopping counters) -- the reference's "negligible" rows A3, A14-A16 of SURVEY.md
// section 8(a).
#include "../../include/cge_b200.h"
#include "cge_kernels.cuh"
#include "cge_landmarks.cuh"
#include "cge_rc.cuh"
#include "cge_ring.cuh"

#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <numeric>
#include <string>
#include <vector>

namespace cge {

static thread_local std::string g_err;

static int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
void set_last_error(const std::string &msg) { g_err = msg; }  // for cge_ingest.cpp

#define CUDA_TRY(expr)                                                                     \
    do {                                                                                   \
        cudaError_t e__ = (expr);                                                          \
        if (e__ != cudaSuccess)                                                            \
            return fail(e__ == cudaErrorMemoryAllocation ? CGE_B200_ERR_OOM                \
                                                         : CGE_B200_ERR_CUDA,              \
                        std::string(#expr) + ": " + cudaGetErrorString(e__));              \
    } while (0)

void launch_tiles(int m, int kind, int grid, cudaStream_t stream, const SweepArgs &a) {
    switch ((m - 1) / 5) {
        case 0: launch_tiles_part0(m, kind, grid, stream, a); break;
        case 1: launch_tiles_part1(m, kind, grid, stream, a); break;
        case 2: launch_tiles_part2(m, kind, grid, stream, a); break;
        case 3: launch_tiles_part3(m, kind, grid, stream, a); break;
        case 4: launch_tiles_part4(m, kind, grid, stream, a); break;
        case 5: launch_tiles_part5(m, kind, grid, stream, a); break;
        case 6: launch_tiles_part6(m, kind, grid, stream, a); break;
        default: launch_tiles_part7(m, kind, grid, stream, a); break;
    }
}

const void *fp_kernel(int m, int directed) {
    switch ((m - 1) / 5) {
        case 0: return fp_kernel_part0(m, directed);
        case 1: return fp_kernel_part1(m, directed);
        case 2: return fp_kernel_part2(m, directed);
        case 3: return fp_kernel_part3(m, directed);
        case 4: return fp_kernel_part4(m, directed);
        case 5: return fp_kernel_part5(m, directed);
        case 6: return fp_kernel_part6(m, directed);
        default: return fp_kernel_part7(m, directed);
    }
}

const void *fp_ring_kernel(int m, int directed) {
    switch ((m - 1) / 5) {
        case 0: return fp_ring_kernel_part0(m, directed);
        case 1: return fp_ring_kernel_part1(m, directed);
        case 2: return fp_ring_kernel_part2(m, directed);
        case 3: return fp_ring_kernel_part3(m, directed);
        case 4: return fp_ring_kernel_part4(m, directed);
        case 5: return fp_ring_kernel_part5(m, directed);
        case 6: return fp_ring_kernel_part6(m, directed);
        default: return fp_ring_kernel_part7(m, directed);
    }
}
size_t fp_ring_smem_bytes(int directed) {
    return directed ? ring_smem_bytes<true>() : ring_smem_bytes<false>();
}
int fp_ring_threads() { return RING_THREADS; }

// ---------------------------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------------------------
constexpr int DK = 16;  // embedding dimensions staged per step; dp is padded to a multiple

__device__ __forceinline__ void atomic_min_max_nonneg(unsigned long long *lohi, double mn,
                                                      double mx) {
    // non-negative doubles order like their bit patterns
    atomicMin(lohi, (unsigned long long)__double_as_longlong(mn));
    atomicMax(lohi + 1, (unsigned long long)__double_as_longlong(mx));
}

// Distance tiles: D_ij = sqrt(sum_c (x_ic - x_jc)^2) in difference form (auxilary.jl:14-20),
// D_ii = distances[i] (divergence.jl:85-86; zeros for the full graph, :106-111), extrema over all
// entries (divergence.jl:92 / :113).  STORE writes the raw tile (pads = -1), otherwise only the
// extrema are produced (landmark mode needs nothing else of the full graph).
template <bool STORE>
__global__ void __launch_bounds__(NTHREADS)
k_build_dist(const double *__restrict__ emb, int dp, const double *__restrict__ diag, int n,
             const int2 *__restrict__ tile_ij, long long tile_begin, long long tile_end,
             double *__restrict__ q, unsigned long long *lohi,
             const int *__restrict__ tile_list = nullptr) {
    __shared__ double sA[TILE][DK + 1];
    __shared__ double sB[TILE][DK + 1];
    __shared__ double s_mn[NWARPS], s_mx[NWARPS];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    double lmin = INFINITY, lmax = 0.0;
    for (long long idx = tile_begin + blockIdx.x; idx < tile_end; idx += gridDim.x) {
        // with a tile list (diameter verification) [tile_begin, tile_end) indexes the list
        const long long t = tile_list ? (long long)tile_list[idx] : idx;
        const int2 ij = tile_ij[t];
        const int bi = ij.x, bj = ij.y;
        double acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;
        for (int k0 = 0; k0 < dp; k0 += DK) {
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int e = tid + NTHREADS * i, r = e >> 4, c = e & 15;
                sA[r][c] = emb[(size_t)(bi * TILE + r) * dp + k0 + c];
                sB[r][c] = emb[(size_t)(bj * TILE + r) * dp + k0 + c];
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < DK; ++kk) {
                double av[8], bv[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) av[i] = sA[ty + 16 * i][kk];
#pragma unroll
                for (int j = 0; j < 8; ++j) bv[j] = sB[tx + 16 * j][kk];
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const double df = av[i] - bv[j];
                        acc[i][j] = fma(df, df, acc[i][j]);
                    }
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int gi = bi * TILE + ty + 16 * i;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int gj = bj * TILE + tx + 16 * j;
                double v;
                if (gi >= n || gj >= n) {
                    v = -1.0;
                } else {
                    v = gi == gj ? (diag ? diag[gi] : 0.0) : sqrt(acc[i][j]);
                    lmin = fmin(lmin, v);
                    lmax = fmax(lmax, v);
                }
                if (STORE)
                    q[(size_t)(t - tile_begin) * TILE_ELEMS + (size_t)(ty + 16 * i) * TILE + tx +
                      16 * j] = v;
            }
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        lmin = fmin(lmin, __shfl_xor_sync(FULL, lmin, off));
        lmax = fmax(lmax, __shfl_xor_sync(FULL, lmax, off));
    }
    if ((tid & 31) == 0) {
        s_mn[tid >> 5] = lmin;
        s_mx[tid >> 5] = lmax;
    }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < NWARPS; ++w) {
            lmin = fmin(lmin, s_mn[w]);
            lmax = fmax(lmax, s_mx[w]);
        }
        if (lmin <= lmax) atomic_min_max_nonneg(lohi, lmin, lmax);
    }
}

// D -> q = ((1 - (D - lo)/(hi - lo)))^(1/4)  (divergence.jl:93 and the alpha-independent part of
// :146); pads -> 0 so that they never contribute.
__global__ void k_transform(double *__restrict__ q, size_t count,
                            const unsigned long long *__restrict__ lohi) {
    const double lo = __longlong_as_double((long long)lohi[0]);
    const double hi = __longlong_as_double((long long)lohi[1]);
    const double range = hi - lo;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (size_t)gridDim.x * blockDim.x) {
        const double x = q[i];
        q[i] = x < 0.0 ? 0.0 : sqrt(sqrt(1.0 - (x - lo) / range));
    }
}

__global__ void k_qdiag(const double *__restrict__ dist, int n,
                        const unsigned long long *__restrict__ lohi, double *__restrict__ qdiag) {
    const double lo = __longlong_as_double((long long)lohi[0]);
    const double hi = __longlong_as_double((long long)lohi[1]);
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < n) qdiag[v] = sqrt(sqrt(1.0 - (dist[v] - lo) / (hi - lo)));
}

// q for the sampled pairs of the local score (divergence.jl:187-189,196-198 landmark mode with
// the full-graph extrema, :205,210 exact mode).  Same summation order as k_build_dist.
__global__ void k_sample_q(const double *__restrict__ emb, int dp, const int *__restrict__ ia,
                           const int *__restrict__ ib, const double *__restrict__ diag,
                           const unsigned long long *__restrict__ lohi, int full_graph,
                           long long count, double *__restrict__ out) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= count) return;
    const double lo = full_graph ? 0.0 : __longlong_as_double((long long)lohi[0]);
    const double hi = __longlong_as_double((long long)lohi[1]);
    const int i = ia[s], j = ib[s];
    double dv;
    if (i == j) {
        dv = diag ? diag[i] : 0.0;
    } else {
        const double *a = emb + (size_t)i * dp, *b = emb + (size_t)j * dp;
        double acc = 0.0;
        for (int c = 0; c < dp; ++c) {
            const double df = a[c] - b[c];
            acc = fma(df, df, acc);
        }
        dv = sqrt(acc);
    }
    out[s] = sqrt(sqrt(1.0 - (dv - lo) / (hi - lo)));
}

__device__ __forceinline__ void block_max_to_slot(double e, unsigned long long *slot) {
    __shared__ double s_max[32];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) e = fmax(e, __shfl_xor_sync(FULL, e, off));
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = e;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) e = fmax(e, s_max[w]);
        atomicMax(slot, (unsigned long long)__double_as_longlong(e));
    }
}

// sum of the per-block partial slots in fixed order: a CTA takes 32 vertices, warp w adds the
// slots b = w, w+8, ... (coalesced 256-byte rows), warp 0 adds the eight sub-sums
__global__ void __launch_bounds__(NTHREADS)
k_reduce_part(const SweepArgs a, const double *__restrict__ part, double *__restrict__ sraw) {
    __shared__ double s_red[NWARPS * 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int v = blockIdx.x * 32 + lane;
    double p = 0.0;
    if (v < a.n) {
        int b_lo, b_hi;
        part_range(a, v, b_lo, b_hi);
        for (int b = b_lo + w; b < b_hi; b += NWARPS) p += __ldcg(part + (size_t)b * a.np + v);
    }
    s_red[w * 32 + lane] = p;
    __syncthreads();
    if (w == 0 && v < a.n) {
        double s = 0.0;
#pragma unroll
        for (int w2 = 0; w2 < NWARPS; ++w2) s += s_red[w2 * 32 + lane];
        sraw[v] = s;
    }
}

// divergence.jl:160-166: S_i = T_i * sum, T_i += eps*T_i*(w_i/S_i - 1), diff = max|w_i - S_i|
__global__ void k_update_u(const double *__restrict__ sraw, double *__restrict__ T,
                           const double *__restrict__ w, int n, double eps,
                           double *__restrict__ S, unsigned long long *slot,
                           unsigned long long *next_slot) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    double e = 0.0;
    if (v < n) {
        const double t = T[v];
        const double s = t * sraw[v];
        const double move = eps * t * (w[v] / s - 1.0);
        T[v] = t + move;
        S[v] = s;
        e = fabs(w[v] - s);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *next_slot = 0ull;
    block_max_to_slot(e, slot);
}

// divergence.jl:451-461 with the doubled diagonal of :442-447
__global__ void k_update_d(const double *__restrict__ sraw_in, const double *__restrict__ sraw_out,
                           double *__restrict__ Tin, double *__restrict__ Tout,
                           const double *__restrict__ deg_in, const double *__restrict__ deg_out,
                           const double *__restrict__ qdiag, int m, int n, double eps,
                           double *__restrict__ Sin, double *__restrict__ Sout,
                           unsigned long long *slot, unsigned long long *next_slot) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    double e = 0.0;
    if (v < n) {
        const double ti = Tin[v], to = Tout[v];
        const double gd = powm_rt(qdiag[v], m);
        const double sin = ti * (sraw_in[v] + to * gd);
        const double sout = to * (sraw_out[v] + ti * gd);
        Sin[v] = sin;
        Sout[v] = sout;
        const double di = deg_in[v], dout = deg_out[v];
        if (di > 0.0) {
            Tin[v] = ti + eps * ti * (di / sin - 1.0);
            e = fmax(e, fabs(di - sin));
        }
        if (dout > 0.0) {
            Tout[v] = to + eps * to * (dout / sout - 1.0);
            e = fmax(e, fabs(dout - sout));
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *next_slot = 0ull;
    block_max_to_slot(e, slot);
}

// Local score, divergence.jl:178-213 / 478-517: pos/neg = (f_a*f_b) * q^m, then
// sum((pos > neg) * w) and sum(w).  One block, fixed reduction order.
struct SampleSide {
    const int *a, *b;        // scored-graph (sorted) vertex of each endpoint
    const double *w0a, *wla; // landmark mode: init_vweights[i], vweights[v_to_l[i]] (else NULL)
    const double *w0b, *wlb;
    const double *q;         // (1 - D)^(1/4) of the sampled pair
};

__device__ __forceinline__ double sample_value(const SampleSide &s, long long i,
                                               const double *Ta, const double *Tb, int m) {
    double fa = __ldcg(Ta + s.a[i]), fb = __ldcg(Tb + s.b[i]);
    if (s.w0a) {  // adj_T = T[l]*w0/w_l (divergence.jl:187-188)
        fa = fa * s.w0a[i] / s.wla[i];
        fb = fb * s.w0b[i] / s.wlb[i];
    }
    return fa * fb * powm_rt(s.q[i], m);
}

constexpr int AUC_THREADS = 256, AUC_MAX_BLOCKS = 64;
// block b handles samples [b*chunk, (b+1)*chunk) and writes its (sum of winning weights, sum of
// weights) to out[2b], out[2b+1]; the host adds the per-block pairs in block order.
__global__ void __launch_bounds__(AUC_THREADS)
k_auc(SampleSide pos, SampleSide neg, const double *__restrict__ wts, long long offset,
      int K, int chunk, const double *Ta, const double *Tb, int m, double *out) {
    __shared__ double s_num[AUC_THREADS], s_den[AUC_THREADS];
    double num = 0.0, den = 0.0;
    const int s0 = blockIdx.x * chunk, s1 = min(K, s0 + chunk);
    for (int s = s0 + threadIdx.x; s < s1; s += AUC_THREADS) {
        const long long i = offset + s;
        const double pv = sample_value(pos, i, Ta, Tb, m);
        const double nv = sample_value(neg, i, Ta, Tb, m);
        const double w = wts[i];
        num += pv > nv ? w : 0.0;
        den += w;
    }
    s_num[threadIdx.x] = num;
    s_den[threadIdx.x] = den;
    __syncthreads();
    for (int h = AUC_THREADS >> 1; h > 0; h >>= 1) {
        if ((int)threadIdx.x < h) {
            s_num[threadIdx.x] += s_num[threadIdx.x + h];
            s_den[threadIdx.x] += s_den[threadIdx.x + h];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[2 * blockIdx.x] = s_num[0];
        out[2 * blockIdx.x + 1] = s_den[0];
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(CGE_B200_ERR_OOM, std::string("cudaMalloc(") + std::to_string(bytes) +
                                              " bytes): " + cudaGetErrorString(e));
        }
        cap = bytes;
        return 0;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T *as() const {
        return reinterpret_cast<T *>(p);
    }
};

// NCCL is bound at run time so that single-GPU use has no NCCL dependency
struct Id128 {  // ncclUniqueId (128 bytes, passed by value)
    char b[128];
};
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, Id128, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static int nccl_load() {
    if (g_nccl.lib) return 0;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        g_nccl.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) return fail(CGE_B200_ERR_NCCL, std::string("dlopen libnccl: ") + dlerror());
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(g_nccl.lib, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(g_nccl.lib, "ncclCommInitRank");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(g_nccl.lib, "ncclAllReduce");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(g_nccl.lib, "ncclCommDestroy");
    g_nccl.GetErrorString =
        (decltype(g_nccl.GetErrorString))dlsym(g_nccl.lib, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy)
        return fail(CGE_B200_ERR_NCCL, "libnccl lacks a required symbol");
    return 0;
}
// nccl.h enum values (stable across NCCL 2.x): ncclFloat64 = 8, ncclUint64 = 5;
// ncclSum = 0, ncclMax = 2, ncclMin = 3
constexpr int kNcclF64 = 8, kNcclU64 = 5, kNcclSum = 0, kNcclMax = 2, kNcclMin = 3;

static inline int64_t tile_index(int64_t nb, int64_t bi, int64_t bj) {
    return bi * nb - bi * (bi - 1) / 2 + (bj - bi);
}

static void shard_range(int64_t n_tiles, int rank, int n_ranks, int64_t *b, int64_t *e) {
    // contiguous, near-equal tile counts (every tile costs the same 128 KB read)
    *b = n_tiles * rank / n_ranks;
    *e = n_tiles * (rank + 1) / n_ranks;
}

// Jensen-Shannon with the +1 prior (auxilary.jl:34-52) over the bins listed in `bins`
static double js_bins(const std::vector<double> &C, const std::vector<double> &B,
                      const std::vector<int64_t> &bins, const std::vector<uint8_t> &is_int,
                      int use_mask, int internal) {
    double sp1 = 0.0, sp2 = 0.0;
    int64_t cnt = 0;
    for (size_t t = 0; t < bins.size(); ++t) {
        if (use_mask && (is_int[t] != 0) != (internal != 0)) continue;
        sp1 += C[bins[t]];
        sp2 += B[bins[t]];
        ++cnt;
    }
    sp1 += (double)cnt;
    sp2 += (double)cnt;
    double f = 0.0;
    for (size_t t = 0; t < bins.size(); ++t) {
        if (use_mask && (is_int[t] != 0) != (internal != 0)) continue;
        const double p = (C[bins[t]] + 1.0) / sp1, q = (B[bins[t]] + 1.0) / sp2;
        const double mm = (p + q) / 2.0;
        f += p * std::log(p / mm) + q * std::log(q / mm);
    }
    return f / 2.0;
}

}  // namespace cge

using namespace cge;

// Ranks that live in ONE process (cge_b200_score_multi: one host thread per GPU).  They exchange
// the per-pass sums exactly like separate processes do (peer stores inside the kernel) but reach
// the peers' buffers through plain peer access, and the two tiny host-visible reductions (distance
// extrema, B matrix) go through this structure instead of NCCL.
struct LocalGroup {
    int n = 0;
    std::mutex mu;
    std::condition_variable cv;
    int arrived = 0;
    unsigned long gen = 0;
    std::atomic<int> failed{0};
    std::vector<unsigned long long> lohi;  // [n][2]
    std::vector<double> B;                 // [n][k*k]
    // returns false when some rank has failed (everybody then gives up instead of waiting)
    bool barrier() {
        std::unique_lock<std::mutex> lk(mu);
        const unsigned long g = gen;
        if (++arrived == n) {
            arrived = 0;
            ++gen;
            cv.notify_all();
        } else {
            cv.wait(lk, [&] { return gen != g || failed.load() != 0; });
        }
        return failed.load() == 0;
    }
    void fail() {
        failed.store(1);
        std::lock_guard<std::mutex> lk(mu);
        cv.notify_all();
    }
};

struct cge_b200_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    int sm_count = 148;
    bool uploaded = false;
    // problem
    bool directed = false, split = false, landmark = false, star = false;
    int64_t n = 0, np = 0, nb = 0, d = 0, dp = 0, k = 0, K = 0, n_sets = 1;
    int64_t n_full = 0, npf = 0, nbf = 0;
    int max_alphas = CGE_B200_N_ALPHA, driver = CGE_B200_DRIVER_HOSTLOOP;
    int regime = CGE_B200_REGIME_STORED;  // regime in use after upload()
    int64_t n_tiles = 0, tile_begin = 0, tile_end = 0, n_tiles_full = 0;
    // multi-rank
    int rank = 0, n_ranks = 1;
    void *nccl_comm = nullptr;
    // NVLink peer exchange (optional)
    void *xbuf = nullptr;           // own exchange allocation
    void *xpeer[8] = {nullptr};     // every rank's allocation as mapped here
    int64_t xcap = 0;               // vertex capacity of a slot
    bool p2p_ready = false;
    unsigned pass_total = 0;        // fixed-point passes executed with the exchange so far
    LocalGroup *group = nullptr;    // set when the ranks are threads of this process
    // host copies
    std::vector<int64_t> perm;           // sorted position -> caller's 0-based vertex
    std::vector<double> C;               // k*k observed community mass (divergence.jl:55-63/337-345)
    std::vector<int64_t> bins;           // bins of C/B that enter JS, in the reference's order
    std::vector<uint8_t> bin_internal;   // vect_I (divergence.jl:66-71 / 348-351)
    // device
    DevBuf q, tile_ij, tile_ij_full, emb, emb_full, dist, w, w2, T0a, T0b, Ta, Tb, Sa, Sb, sraw_a,
        sraw_b, partA, partB, comm, B, qdiag, lohi, slots, auc_out, fpres;
    // recompute regime: super-tile table, operand image (centred for the row-norm / dot form), row norms
    DevBuf st_ij, opT, nrm, rc_mean, st_pre, qst;
    std::vector<int64_t> st_pre_host;    // tiles before super-tile s in the global sequence
    int64_t st_store_end = 0;            // super-tiles [st_begin, st_store_end) keep their q tiles in HBM
    bool store_auto = false;             // regime chosen by AUTO: keep as much of the matrix as fits
    bool rc_dot = false;
    int regime_reported = CGE_B200_REGIME_STORED;
    int sb = 1;                          // tiles per super-block side
    int64_t nsb = 0, n_st = 0, st_begin = 0, st_end = 0;
    int srow_begin = 0, srow_end = -1;
    // tensor-core diameter filter (landmark mode, large original graphs)
    DevBuf diam_strips, diam_packed, diam_norms, diam_tilemax, diam_list, diam_ctr, diam_mean;
    int diam_n_strips = 0;
    bool diam_ok = false;
    DevBuf s_pda, s_pdb, s_nda, s_ndb, s_pa, s_pb, s_na, s_nb, s_pw, s_pw0a, s_pwla, s_pw0b, s_pwlb, s_nw0a, s_nwla, s_nw0b,
        s_nwlb, s_pq, s_nq;
    float ms_upload = 0.f;
    int64_t launches = 0;
    void *pinned = nullptr;           // page-locked staging for the per-alpha results
    size_t pinned_cap = 0;
    int ensure_pinned(size_t bytes) {
        if (bytes <= pinned_cap) return 0;
        if (pinned) cudaFreeHost(pinned);
        pinned = nullptr;
        pinned_cap = 0;
        if (cudaMallocHost(&pinned, bytes) != cudaSuccess) {
            cudaGetLastError();
            return fail(CGE_B200_ERR_OOM, "cudaMallocHost failed");
        }
        pinned_cap = bytes;
        return 0;
    }
    std::vector<cudaEvent_t> evpool;  // pairs of events around every sweep launch
    size_t ev_used = 0;
    std::vector<uint8_t> ev_is_b;
    cudaEvent_t next_event() {
        if (ev_used == evpool.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            evpool.push_back(e);
        }
        return evpool[ev_used++];
    }
};

namespace cge {

static int upload_vec(DevBuf &buf, const void *src, size_t bytes, cudaStream_t st) {
    if (int rc = buf.ensure(std::max<size_t>(bytes, 16))) return rc;
    if (bytes) CUDA_TRY(cudaMemcpyAsync(buf.p, src, bytes, cudaMemcpyHostToDevice, st));
    return 0;
}

static std::vector<int2> make_tile_table(int64_t nb) {
    std::vector<int2> t;
    t.reserve((size_t)(nb * (nb + 1) / 2));
    for (int bi = 0; bi < nb; ++bi)
        for (int bj = bi; bj < nb; ++bj) t.push_back(make_int2(bi, bj));
    return t;
}

static int nccl_check(int rc, const char *what);

// max of a small host integer over the ranks (the in-process group, or NCCL on a device word)
static int agree_max_across_ranks(cge_b200_handle *h, int *value) {
    if (h->group) {
        LocalGroup &G = *h->group;
        G.lohi[2 * h->rank] = (unsigned long long)*value;
        if (!G.barrier()) return fail(CGE_B200_ERR_STATE, "another rank failed");
        unsigned long long mx = 0;
        for (int r = 0; r < G.n; ++r) mx = std::max(mx, G.lohi[2 * r]);
        if (!G.barrier()) return fail(CGE_B200_ERR_STATE, "another rank failed");
        *value = (int)mx;
        return 0;
    }
    if (!h->nccl_comm) return fail(CGE_B200_ERR_STATE, "multi-rank handle without a communicator");
    if (int rc = h->slots.ensure(128)) return rc;
    unsigned long long v = (unsigned long long)*value;
    CUDA_TRY(cudaMemcpyAsync(h->slots.p, &v, 8, cudaMemcpyHostToDevice, h->stream));
    if (int rc = nccl_check(g_nccl.AllReduce(h->slots.p, h->slots.p, 1, kNcclU64, kNcclMax,
                                             h->nccl_comm, h->stream),
                            "ncclAllReduce(regime)"))
        return rc;
    CUDA_TRY(cudaMemcpyAsync(&v, h->slots.p, 8, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    *value = (int)v;
    return 0;
}

static int do_upload(cge_b200_handle *h, const cge_b200_problem *p) {
    auto t0 = std::chrono::steady_clock::now();
    const bool trace = getenv("CGE_B200_PHASES") && atoi(getenv("CGE_B200_PHASES"));
    auto lap = [&](const char *what) {
        if (trace)
            fprintf(stderr, "[cge_b200 upload] %-28s %.3f ms\n", what,
                    std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0)
                        .count());
    };
    if (!p || p->struct_size != (int32_t)sizeof(cge_b200_problem))
        return fail(CGE_B200_ERR_ARG, "cge_b200_problem.struct_size mismatch");
    if (p->index_base != 0 && p->index_base != 1) return fail(CGE_B200_ERR_ARG, "index_base");
    if (p->m <= 0 || !p->edge_src || !p->edge_dst || !p->eweights || !p->comm || !p->embed ||
        !p->distances || !p->vweights || p->d <= 0)
        return fail(CGE_B200_ERR_ARG, "missing input array");
    const int64_t base = p->index_base;
    h->uploaded = false;
    h->directed = p->directed != 0;
    h->split = p->split != 0;
    h->landmark = p->n_full > 0;
    h->max_alphas = p->max_alphas > 0 ? std::min<int>(p->max_alphas, CGE_B200_N_ALPHA)
                                      : CGE_B200_N_ALPHA;
    h->driver = p->driver;
    // no_vertices = maximum(edges) (divergence.jl:41 / :294)
    int64_t n = 0;
    for (int64_t e = 0; e < p->m; ++e) {
        n = std::max(n, std::max(p->edge_src[e], p->edge_dst[e]) - base + 1);
        if (p->edge_src[e] < base || p->edge_dst[e] < base)
            return fail(CGE_B200_ERR_ARG, "edge endpoint below index_base");
    }
    if (p->n_comm != n)
        return fail(CGE_B200_ERR_ASSERT_COMM, "No. communities not matching no. vertices");
    if (p->n_distances != n)
        return fail(CGE_B200_ERR_ASSERT_DIST,
                    "Distances vector length is not equal to no. vertices");
    if (p->embed_rows < n) return fail(CGE_B200_ERR_ARG, "embedding has fewer rows than vertices");
    if (n >= (int64_t)1 << 30) return fail(CGE_B200_ERR_ARG, "too many vertices");
    h->n = n;
    h->d = p->d;
    h->dp = (p->d + DK - 1) / DK * DK;
    h->nb = (n + TILE - 1) / TILE;
    h->np = h->nb * TILE;
    h->K = p->n_samples;
    h->n_sets = p->n_samples > 0 ? std::max<int64_t>(p->n_sets, 1) : 1;
    // one set for every alpha, or one set per evaluated alpha (the run indexes set m-1 for alpha m)
    if (h->n_sets != 1 && h->n_sets < h->max_alphas)
        return fail(CGE_B200_ERR_ARG, "n_sets must be 1 or at least the number of alpha values evaluated");
    h->n_full = p->n_full;
    // communities, 0-based; n_parts = maximum(comm) (divergence.jl:51)
    int64_t k = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (p->comm[i] < base) return fail(CGE_B200_ERR_ARG, "community id below index_base");
        k = std::max(k, p->comm[i] - base + 1);
    }
    h->k = k;
    // sort vertices by community (stable): row/column community of a tile changes rarely
    h->perm.resize((size_t)n);
    std::iota(h->perm.begin(), h->perm.end(), 0);
    std::stable_sort(h->perm.begin(), h->perm.end(),
                     [&](int64_t a, int64_t b) { return p->comm[a] < p->comm[b]; });
    std::vector<int64_t> inv((size_t)n);
    for (int64_t s = 0; s < n; ++s) inv[(size_t)h->perm[(size_t)s]] = s;

    // C vector (divergence.jl:55-63 undirected: bin (min c, max c); :337-345 directed: (c_src, c_dst))
    h->C.assign((size_t)(k * k), 0.0);
    std::vector<double> deg_in, deg_out;
    if (h->directed) {
        deg_in.assign((size_t)n, 0.0);
        deg_out.assign((size_t)n, 0.0);
    }
    std::vector<int64_t> star((size_t)(h->directed ? n : 0), 0);
    for (int64_t e = 0; e < p->m; ++e) {
        const int64_t u = p->edge_src[e] - base, v = p->edge_dst[e] - base;
        const int64_t cu = p->comm[u] - base, cv = p->comm[v] - base;
        if (h->directed) {
            h->C[(size_t)(cu * k + cv)] += p->eweights[e];
            deg_out[(size_t)u] += p->eweights[e];  // divergence.jl:311-319
            deg_in[(size_t)v] += p->eweights[e];
            star[(size_t)u] += 1;
            star[(size_t)v] += 1;
        } else {
            h->C[(size_t)(std::min(cu, cv) * k + std::max(cu, cv))] += p->eweights[e];
        }
    }
    h->star = false;
    if (h->directed) {  // divergence.jl:322-334
        bool has_nm1 = false, has_2nm1 = false;
        int64_t sum = 0, cnt2 = 0;
        for (int64_t i = 0; i < n; ++i) {
            has_nm1 |= star[(size_t)i] == n - 1;
            has_2nm1 |= star[(size_t)i] == 2 * (n - 1);
            cnt2 += star[(size_t)i] == 2;
            sum += star[(size_t)i];
        }
        if (has_nm1 && sum == 2 * (n - 1)) h->star = true;
        else if (has_2nm1 && cnt2 == n - 1) h->star = true;
    }
    h->bins.clear();
    h->bin_internal.clear();
    if (h->directed) {
        for (int64_t a = 0; a < k; ++a)
            for (int64_t b = 0; b < k; ++b) {
                h->bins.push_back(a * k + b);
                h->bin_internal.push_back(a == b);
            }
    } else {
        for (int64_t a = 0; a < k; ++a)
            for (int64_t b = a; b < k; ++b) {
                h->bins.push_back(a * k + b);
                h->bin_internal.push_back(a == b);
            }
    }
    if (h->star) {
        h->uploaded = true;
        return 0;
    }

    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const int64_t np = h->np, dp = h->dp;
    lap("validate, sort, C, degrees");
    // sorted + padded per-vertex arrays
    std::vector<double> emb((size_t)(np * dp), 0.0), dist((size_t)np, 0.0), w((size_t)np, 1.0),
        w2, Ta((size_t)np, 0.0), Tb;
    std::vector<int> comm((size_t)np, -1);
    for (int64_t s = 0; s < n; ++s) {
        const int64_t v = h->perm[(size_t)s];
        for (int64_t c = 0; c < p->d; ++c)
            emb[(size_t)(s * dp + c)] = p->embed[v * p->embed_row_stride + c * p->embed_col_stride];
        dist[(size_t)s] = p->distances[v];
        comm[(size_t)s] = (int)(p->comm[v] - base);
    }
    if (h->directed) {  // Tin/Tout = 1, 0 where the degree is 0 (divergence.jl:399-402)
        w.assign((size_t)np, 0.0);
        w2.assign((size_t)np, 0.0);
        Tb.assign((size_t)np, 0.0);
        for (int64_t s = 0; s < n; ++s) {
            const int64_t v = h->perm[(size_t)s];
            w[(size_t)s] = deg_in[(size_t)v];
            w2[(size_t)s] = deg_out[(size_t)v];
            Ta[(size_t)s] = deg_in[(size_t)v] == 0.0 ? 0.0 : 1.0;
            Tb[(size_t)s] = deg_out[(size_t)v] == 0.0 ? 0.0 : 1.0;
        }
    } else {  // T = ones (divergence.jl:118)
        for (int64_t s = 0; s < n; ++s) {
            w[(size_t)s] = p->vweights[h->perm[(size_t)s]];
            Ta[(size_t)s] = 1.0;
        }
    }
    int rc;
    if ((rc = upload_vec(h->emb, emb.data(), emb.size() * 8, st))) return rc;
    if ((rc = upload_vec(h->dist, dist.data(), dist.size() * 8, st))) return rc;
    if ((rc = upload_vec(h->w, w.data(), w.size() * 8, st))) return rc;
    if ((rc = upload_vec(h->T0a, Ta.data(), Ta.size() * 8, st))) return rc;
    if ((rc = upload_vec(h->comm, comm.data(), comm.size() * 4, st))) return rc;
    if (h->directed) {
        if ((rc = upload_vec(h->w2, w2.data(), w2.size() * 8, st))) return rc;
        if ((rc = upload_vec(h->T0b, Tb.data(), Tb.size() * 8, st))) return rc;
    }
    lap("per-vertex arrays uploaded");
    // this rank's share of the tile sequence, regime
    h->n_tiles = h->nb * (h->nb + 1) / 2;
    shard_range(h->n_tiles, h->rank, h->n_ranks, &h->tile_begin, &h->tile_end);
    const size_t local_tiles = (size_t)(h->tile_end - h->tile_begin);
    const size_t q_bytes = std::max<size_t>(local_tiles, 1) * TILE_ELEMS * 8;
    const int want = p->regime;
    if (want < CGE_B200_REGIME_AUTO || want > CGE_B200_REGIME_RECOMPUTE_DIFF)
        return fail(CGE_B200_ERR_ARG, "unknown regime");
    h->regime = want >= CGE_B200_REGIME_RECOMPUTE ? CGE_B200_REGIME_RECOMPUTE : want;
    if (h->regime == CGE_B200_REGIME_AUTO && h->q.cap >= q_bytes) {
        h->regime = CGE_B200_REGIME_STORED;  // the handle already holds a large enough matrix
    } else if (h->regime == CGE_B200_REGIME_AUTO) {
        // stored when the tiles fit next to everything else (2 GB + 5 % head-room), else recompute
        // (cudaMemGetInfo costs milliseconds: only asked when the matrix has to be (re)allocated)
        size_t free_b = 0, total_b = 0;
        CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
        const size_t avail = free_b + h->q.cap + h->partA.cap + h->partB.cap;
        const size_t other = (size_t)h->nb * (size_t)np * 8 * (h->directed ? 2 : 1) + ((size_t)2 << 30);
        h->regime = q_bytes + other + total_b / 20 <= avail ? CGE_B200_REGIME_STORED
                                                            : CGE_B200_REGIME_RECOMPUTE;
    }
    // every rank must run the same regime (a rank that alone falls back to recomputing would leave
    // its peers polling the exchange records of a kernel that never starts): the ranks agree on
    // the larger code, i.e. recompute as soon as one of them cannot store its share
    if (h->n_ranks > 1 && want == CGE_B200_REGIME_AUTO)
        if ((rc = agree_max_across_ranks(h, &h->regime))) return rc;
    if (h->regime == CGE_B200_REGIME_STORED) {
        if ((rc = h->q.ensure(q_bytes))) return rc;
    } else {
        h->q.release();
    }
    const bool stored = h->regime == CGE_B200_REGIME_STORED;
    // Distances of the recompute regime: d^2 = n_i + n_j - 2 x_i.x_j on the centred embedding
    // (cge_recompute.cu) unless the difference form of the reference is asked for (regime
    // RECOMPUTE_DIFF, or CGE_B200_RC_FORM=diff for hosts that cannot set the field).
    const char *rc_form = getenv("CGE_B200_RC_FORM");
    h->rc_dot = !stored && want != CGE_B200_REGIME_RECOMPUTE_DIFF &&
                !(rc_form && std::strcmp(rc_form, "diff") == 0);
    h->regime_reported = stored ? CGE_B200_REGIME_STORED
                         : !h->rc_dot ? CGE_B200_REGIME_RECOMPUTE_DIFF
                         : want == CGE_B200_REGIME_RECOMPUTE_DOT ? CGE_B200_REGIME_RECOMPUTE_DOT
                                                                 : CGE_B200_REGIME_RECOMPUTE;
    if (stored) {  // the tile table of the stored sweeps and of k_build_dist
        std::vector<int2> tij = make_tile_table(h->nb);
        if ((rc = upload_vec(h->tile_ij, tij.data(), tij.size() * sizeof(int2), st))) return rc;
        CUDA_TRY(cudaStreamSynchronize(st));  // tij is a local
    }
    size_t part_rows = (size_t)h->nb;
    if (!stored) {
        // ---- super-tiles (cge_rc.cuh): the smallest super-block that keeps the partial slots
        // under 2 GiB, at most RC_MAX_SB; CGE_B200_RC_SB overrides (tests) ----
        int sb = 1;
        while (sb < RC_MAX_SB && (size_t)((h->nb + sb - 1) / sb) * (size_t)np * 8 * (h->directed ? 2 : 1) >
                                     ((size_t)2 << 30))
            sb *= 2;
        if (const char *e = getenv("CGE_B200_RC_SB")) {
            const int v = atoi(e);
            if (v == 1 || v == 2 || v == 4 || v == 8) sb = v;
        }
        h->sb = sb;
        h->nsb = (h->nb + sb - 1) / sb;
        h->n_st = h->nsb * (h->nsb + 1) / 2;
        std::vector<int2> stij;
        stij.reserve((size_t)h->n_st);
        std::vector<int64_t> pre((size_t)h->n_st + 1, 0);  // tiles before super-tile s
        auto side = [&](int64_t I) { return std::min<int64_t>(sb, h->nb - I * sb); };
        for (int64_t I = 0; I < h->nsb; ++I)
            for (int64_t J = I; J < h->nsb; ++J) {
                const int64_t r = side(I), c = side(J);
                pre[stij.size() + 1] = pre[stij.size()] + (I == J ? r * (r + 1) / 2 : r * c);
                stij.push_back(make_int2((int)I, (int)J));
            }
        // contiguous shares of near-equal tile counts
        auto cut = [&](int r) {
            const int64_t target = pre[(size_t)h->n_st] * r / h->n_ranks;
            return (int64_t)(std::lower_bound(pre.begin(), pre.end(), target) - pre.begin());
        };
        h->st_begin = h->rank == 0 ? 0 : cut(h->rank);
        h->st_end = h->rank == h->n_ranks - 1 ? h->n_st : cut(h->rank + 1);
        h->srow_begin = h->st_end > h->st_begin ? stij[(size_t)h->st_begin].x : 0;
        h->srow_end = h->st_end > h->st_begin ? stij[(size_t)h->st_end - 1].x : -1;
        if ((rc = upload_vec(h->st_ij, stij.data(), stij.size() * sizeof(int2), st))) return rc;
        if ((rc = upload_vec(h->st_pre, pre.data(), pre.size() * 8, st))) return rc;
        h->st_pre_host = pre;
        h->st_store_end = h->st_begin;
        h->store_auto = want == CGE_B200_REGIME_AUTO;
        part_rows = (size_t)h->nsb;
        // ---- operand image and row norms ----
        std::vector<double> mean((size_t)dp, 0.0);
        if (h->rc_dot) {
            for (int64_t r = 0; r < n; ++r)
                for (int64_t c = 0; c < p->d; ++c) mean[(size_t)c] += emb[(size_t)(r * dp + c)];
            for (int64_t c = 0; c < p->d; ++c) mean[(size_t)c] /= (double)n;
        }
        if ((rc = upload_vec(h->rc_mean, mean.data(), mean.size() * 8, st))) return rc;
        if ((rc = h->opT.ensure((size_t)h->nb * (size_t)(dp / RC_DK) * RC_CHUNK * 8))) return rc;
        if ((rc = h->nrm.ensure((size_t)np * 8))) return rc;
        launch_rc_pack(h->emb.as<double>(), h->rc_mean.as<double>(), (int)n, (int)np, (int)dp,
                       h->opT.as<double>(), h->rc_dot ? h->nrm.as<double>() : nullptr, st);
        CUDA_TRY(cudaStreamSynchronize(st));  // stij, mean are locals
        CUDA_TRY(cudaGetLastError());
        if (trace)
            fprintf(stderr, "[cge_b200] recompute regime: %s form, super-block %d (%lld super-tiles, %lld..%lld here)\n",
                    h->rc_dot ? "row-norm/dot" : "difference", sb, (long long)h->n_st,
                    (long long)h->st_begin, (long long)h->st_end);
    }
    const size_t part_bytes = part_rows * (size_t)np * 8;
    if ((rc = h->partA.ensure(part_bytes))) return rc;
    if (h->directed && (rc = h->partB.ensure(part_bytes))) return rc;
    for (DevBuf *b : {&h->Ta, &h->Tb, &h->Sa, &h->Sb, &h->sraw_a, &h->sraw_b, &h->qdiag})
        if ((rc = b->ensure((size_t)np * 8))) return rc;
    if ((rc = h->B.ensure(std::max<size_t>((size_t)(k * k) * 8, 16)))) return rc;
    if ((rc = h->lohi.ensure(64))) return rc;
    if ((rc = h->slots.ensure(128))) return rc;
    if ((rc = h->auc_out.ensure(2 * AUC_MAX_BLOCKS * 8))) return rc;
    if ((rc = h->ensure_pinned(64 + 2 * AUC_MAX_BLOCKS * 8 + (size_t)(k * k) * 8))) return rc;
    if ((rc = h->fpres.ensure(64))) return rc;

    lap("tile table, buffers");
    // landmark mode: original graph arrays for the local score
    if (h->landmark && h->K > 0) {
        if (!p->init_vweights || !p->v_to_l || !p->init_embed)
            return fail(CGE_B200_ERR_ARG, "landmark mode needs init_vweights, v_to_l, init_embed");
        h->nbf = (h->n_full + TILE - 1) / TILE;
        h->npf = h->nbf * TILE;
        h->n_tiles_full = h->nbf * (h->nbf + 1) / 2;
        std::vector<double> ef((size_t)(h->npf * dp), 0.0), mean((size_t)dp, 0.0);
        for (int64_t v = 0; v < h->n_full; ++v)
            for (int64_t c = 0; c < p->d; ++c) {
                const double x = p->init_embed[v * p->init_row_stride + c * p->init_col_stride];
                ef[(size_t)(v * dp + c)] = x;
                mean[(size_t)c] += x;
            }
        for (int64_t c = 0; c < p->d; ++c) mean[(size_t)c] /= (double)h->n_full;
        if ((rc = upload_vec(h->diam_mean, mean.data(), mean.size() * 8, st))) return rc;
        if ((rc = upload_vec(h->emb_full, ef.data(), ef.size() * 8, st))) return rc;
        std::vector<int2> tij = make_tile_table(h->nbf);
        if ((rc = upload_vec(h->tile_ij_full, tij.data(), tij.size() * sizeof(int2), st)))
            return rc;
        // the diameter of a large original graph goes through the tensor-core filter
        // (cge_diameter.cu): runs of <= 64 tiles of one tile row are the work units
        int64_t diam_min = 8192;
        if (const char *e = getenv("CGE_B200_DIAM_MIN")) diam_min = atoll(e);
        h->diam_ok = h->n_full >= diam_min && dp <= 128 && h->n_tiles_full < ((int64_t)1 << 31);
        if (h->diam_ok) {
            std::vector<int4> strips;
            for (int64_t bi = 0; bi < h->nbf; ++bi)
                for (int64_t bj = bi; bj < h->nbf; bj += 64)
                    strips.push_back(make_int4((int)bi, (int)bj, (int)std::min<int64_t>(64, h->nbf - bj),
                                               (int)tile_index(h->nbf, bi, bj)));
            h->diam_n_strips = (int)strips.size();
            if ((rc = upload_vec(h->diam_strips, strips.data(), strips.size() * sizeof(int4), st)))
                return rc;
            const size_t opb = (size_t)(dp / 16) * 4096;
            if ((rc = h->diam_packed.ensure((size_t)h->nbf * 2 * opb))) return rc;
            if ((rc = h->diam_norms.ensure((size_t)h->npf * 4))) return rc;
            if ((rc = h->diam_tilemax.ensure((size_t)h->n_tiles_full * 4))) return rc;
            if ((rc = h->diam_list.ensure((size_t)(1 << 20) * 4))) return rc;
            if ((rc = h->diam_ctr.ensure(64))) return rc;
            CUDA_TRY(cudaStreamSynchronize(st));  // strips is a local
        }
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    // sampled pairs: (da,db) = vertices whose embedding rows give the distance (original graph in
    // landmark mode), (ta,tb) = sorted scored-graph vertices whose T enters (divergence.jl:187-189)
    if (h->K > 0) {
        if (!p->pos_i || !p->pos_j || !p->pos_w || !p->neg_i || !p->neg_j)
            return fail(CGE_B200_ERR_ARG, "n_samples > 0 but sample arrays missing");
        const int64_t S = h->K * h->n_sets;
        const int64_t lim = h->landmark ? h->n_full : n;
        auto side = [&](const int64_t *si, const int64_t *sj, DevBuf &d_da, DevBuf &d_db,
                        DevBuf &d_ta, DevBuf &d_tb, DevBuf &w0a, DevBuf &wla, DevBuf &w0b,
                        DevBuf &wlb) -> int {
            std::vector<int> da((size_t)S), db((size_t)S), ta((size_t)S), tb2((size_t)S);
            std::vector<double> v0a, vla, v0b, vlb;
            if (h->landmark) {
                v0a.resize((size_t)S); vla.resize((size_t)S);
                v0b.resize((size_t)S); vlb.resize((size_t)S);
            }
            for (int64_t s = 0; s < S; ++s) {
                int64_t i = si[s] - base, j = sj[s] - base;
                if (i < 0 || j < 0 || i >= lim || j >= lim)
                    return fail(CGE_B200_ERR_ARG, "sampled vertex id out of range");
                // undirected pairs are addressed as (min,max) (divergence.jl:133; idx needs i<=j)
                if (!h->directed && i > j) std::swap(i, j);
                if (h->landmark) {
                    const int64_t li = p->v_to_l[i] - base, lj = p->v_to_l[j] - base;
                    if (li < 0 || lj < 0 || li >= n || lj >= n)
                        return fail(CGE_B200_ERR_ARG, "v_to_l out of range");
                    da[(size_t)s] = (int)i;
                    db[(size_t)s] = (int)j;
                    ta[(size_t)s] = (int)inv[(size_t)li];
                    tb2[(size_t)s] = (int)inv[(size_t)lj];
                    v0a[(size_t)s] = p->init_vweights[i];
                    vla[(size_t)s] = p->vweights[li];
                    v0b[(size_t)s] = p->init_vweights[j];
                    vlb[(size_t)s] = p->vweights[lj];
                } else {
                    da[(size_t)s] = ta[(size_t)s] = (int)inv[(size_t)i];
                    db[(size_t)s] = tb2[(size_t)s] = (int)inv[(size_t)j];
                }
            }
            int r;
            if ((r = upload_vec(d_da, da.data(), da.size() * 4, st))) return r;
            if ((r = upload_vec(d_db, db.data(), db.size() * 4, st))) return r;
            if ((r = upload_vec(d_ta, ta.data(), ta.size() * 4, st))) return r;
            if ((r = upload_vec(d_tb, tb2.data(), tb2.size() * 4, st))) return r;
            if (h->landmark) {
                if ((r = upload_vec(w0a, v0a.data(), v0a.size() * 8, st))) return r;
                if ((r = upload_vec(wla, vla.data(), vla.size() * 8, st))) return r;
                if ((r = upload_vec(w0b, v0b.data(), v0b.size() * 8, st))) return r;
                if ((r = upload_vec(wlb, vlb.data(), vlb.size() * 8, st))) return r;
            }
            CUDA_TRY(cudaStreamSynchronize(st));  // the vectors are locals
            return 0;
        };
        if ((rc = side(p->pos_i, p->pos_j, h->s_pda, h->s_pdb, h->s_pa, h->s_pb, h->s_pw0a,
                       h->s_pwla, h->s_pw0b, h->s_pwlb)))
            return rc;
        if ((rc = side(p->neg_i, p->neg_j, h->s_nda, h->s_ndb, h->s_na, h->s_nb, h->s_nw0a,
                       h->s_nwla, h->s_nw0b, h->s_nwlb)))
            return rc;
        if ((rc = upload_vec(h->s_pw, p->pos_w, (size_t)S * 8, st))) return rc;
        if ((rc = h->s_pq.ensure((size_t)S * 8))) return rc;
        if ((rc = h->s_nq.ensure((size_t)S * 8))) return rc;
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    if (!stored) {
        // ---- "store what fits": with everything else allocated, the leading super-tiles of this rank
        // keep their q tiles in what is left of HBM (minus 4 GB + 3 % head-room) and are read in every
        // pass instead of being recomputed.  Only when the regime was left to AUTO (an explicit
        // recompute request recomputes everything), or with CGE_B200_STORE_MB = the budget in MB ----
        size_t budget = 0;
        if (const char *e = getenv("CGE_B200_STORE_MB")) {
            budget = (size_t)std::max(0.0, atof(e)) << 20;
        } else if (h->store_auto) {
            size_t free_b = 0, total_b = 0;
            CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
            const size_t avail = free_b + h->qst.cap, reserve = ((size_t)4 << 30) + total_b / 32;
            budget = avail > reserve ? avail - reserve : 0;
        }
        const int64_t tiles_fit = (int64_t)(budget / ((size_t)TILE_ELEMS * 8));
        const std::vector<int64_t> &pre = h->st_pre_host;
        int64_t e = h->st_begin;
        while (e < h->st_end && pre[(size_t)e + 1] - pre[(size_t)h->st_begin] <= tiles_fit) ++e;
        h->st_store_end = e;
        const size_t bytes = (size_t)(pre[(size_t)e] - pre[(size_t)h->st_begin]) * TILE_ELEMS * 8;
        if (bytes) {
            if ((rc = h->qst.ensure(bytes))) return rc;
        } else {
            h->qst.release();
        }
        if (trace)
            fprintf(stderr, "[cge_b200] store what fits: %lld of %lld super-tiles of this rank (%.2f GB)\n",
                    (long long)(e - h->st_begin), (long long)(h->st_end - h->st_begin), bytes / 1e9);
    }
    h->uploaded = true;
    lap("samples");
    h->ms_upload =
        std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return 0;
}

static int nccl_check(int rc, const char *what) {
    if (rc == 0) return 0;
    return fail(CGE_B200_ERR_NCCL, std::string(what) + ": " +
                                       (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"));
}

// Distance extrema over all ranks (divergence.jl:92): the bit patterns of non-negative doubles
// order like the values, so an integer min / max does it -- through the in-process group or NCCL.
static int reduce_extrema_across_ranks(cge_b200_handle *h, unsigned long long *lohi,
                                       cudaStream_t st) {
    if (h->n_ranks > 1 && h->group) {
        LocalGroup &G = *h->group;
        unsigned long long mine[2];
        CUDA_TRY(cudaMemcpyAsync(mine, lohi, 16, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        G.lohi[2 * h->rank] = mine[0];
        G.lohi[2 * h->rank + 1] = mine[1];
        if (!G.barrier()) return fail(CGE_B200_ERR_STATE, "another rank failed");
        for (int r = 0; r < G.n; ++r) {  // bit patterns of non-negative doubles order like the values
            mine[0] = std::min(mine[0], G.lohi[2 * r]);
            mine[1] = std::max(mine[1], G.lohi[2 * r + 1]);
        }
        if (!G.barrier()) return fail(CGE_B200_ERR_STATE, "another rank failed");
        CUDA_TRY(cudaMemcpyAsync(lohi, mine, 16, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaStreamSynchronize(st));  // `mine` is a local
    } else if (h->n_ranks > 1) {
        if (int rc = nccl_check(g_nccl.AllReduce(lohi, lohi, 1, kNcclU64, kNcclMin, h->nccl_comm, st),
                                "ncclAllReduce(min)"))
            return rc;
        if (int rc = nccl_check(
                g_nccl.AllReduce(lohi + 1, lohi + 1, 1, kNcclU64, kNcclMax, h->nccl_comm, st),
                "ncclAllReduce(max)"))
            return rc;
    }
    return 0;
}

// Landmark mode: extrema of the ORIGINAL graph's distances (divergence.jl:104-115 / 386-397); only
// the maximum matters (the minimum is the zero diagonal).  Large graphs go through the tensor-core
// filter of cge_diameter.cu and an FP64 check of its candidate tiles, the rest (and any case the
// filter cannot prune) through the all-FP64 pass.  Writes the bit pattern into lohi_full[0..1].
static int full_graph_extrema(cge_b200_handle *h, cge_b200_stats &S, unsigned long long *lohi_full,
                              int dp, cudaStream_t st) {
    const int gridf = (int)std::max<int64_t>(
        1, std::min<int64_t>(h->n_tiles_full, (int64_t)2 * h->sm_count));
    bool exact_all = true;
    S.diam_candidate_tiles = -1;
    if (h->diam_ok) {
        // tensor-core filter, then FP64 verification of the candidate tiles only
        unsigned *ctr = h->diam_ctr.as<unsigned>();  // [0] strip counter [1] gmax [2] rmax [3] count
        CUDA_TRY(cudaMemsetAsync(ctr, 0, 64, st));
        launch_pack_bf16(h->emb_full.as<double>(), h->diam_mean.as<double>(), dp,
                         (int)h->n_full, (int)h->d, (int)h->nbf,
                         h->diam_packed.as<unsigned char>(), h->diam_norms.as<float>(),
                         ctr + 2, st);
        DiamArgs da;
        da.packed = h->diam_packed.as<unsigned char>();
        da.norms = h->diam_norms.as<float>();
        da.nb = (int)h->nbf;
        da.ksteps = dp / 16;
        da.strips = h->diam_strips.as<int4>();
        da.n_strips = h->diam_n_strips;
        da.strip_counter = ctr;
        da.tile_max = h->diam_tilemax.as<float>();
        da.gmax_bits = ctr + 1;
        CUDA_TRY(launch_diameter_filter(da, std::min(h->diam_n_strips, h->sm_count), st));
        float rel = 1e-3f;
        if (const char *e = getenv("CGE_B200_DIAM_REL")) rel = (float)atof(e);
        const int cap = 1 << 20;
        launch_select_candidates(da.tile_max, h->n_tiles_full, ctr + 1, ctr + 2, rel,
                                 h->diam_list.as<int>(), cap, (int *)(ctr + 3), st);
        h->launches += 3;
        unsigned hc[4];
        CUDA_TRY(cudaMemcpyAsync(hc, ctr, 16, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        const int64_t cand = (int)hc[3];
        S.diam_candidate_tiles = (int32_t)cand;
        if (cand > 0 && cand <= cap && cand * 8 <= h->n_tiles_full) {
            const int gridc = (int)std::max<int64_t>(1, std::min<int64_t>(cand, 2 * h->sm_count));
            k_build_dist<false><<<gridc, NTHREADS, 0, st>>>(
                h->emb_full.as<double>(), dp, nullptr, (int)h->n_full,
                h->tile_ij_full.as<int2>(), 0, cand, nullptr, lohi_full, h->diam_list.as<int>());
            ++h->launches;
            exact_all = false;
        }
    }
    if (exact_all) {
        k_build_dist<false><<<gridf, NTHREADS, 0, st>>>(
            h->emb_full.as<double>(), dp, nullptr, (int)h->n_full, h->tile_ij_full.as<int2>(),
            0, h->n_tiles_full, nullptr, lohi_full);
        ++h->launches;
    }
    return 0;
}

static int do_run(cge_b200_handle *h, double *out, int32_t *out_len, cge_b200_stats *stats) {
    if (!h->uploaded) return fail(CGE_B200_ERR_STATE, "run() before upload()");
    auto wall0 = std::chrono::steady_clock::now();
    const double inf = std::numeric_limits<double>::infinity();
    cge_b200_stats st_local;
    cge_b200_stats &S = stats ? *stats : st_local;
    std::memset(&S, 0, sizeof(S));
    S.struct_size = (int32_t)sizeof(cge_b200_stats);
    for (int a = 0; a < CGE_B200_N_ALPHA; ++a) S.div[a] = S.auc[a] = NAN;
    S.diam_candidate_tiles = -1;
    S.n = h->n;
    S.n_pairs = h->n * (h->n + 1) / 2;
    S.n_ranks = h->n_ranks;
    // the persistent kernel cannot call NCCL between passes: multi-rank runs without the peer
    // exchange use the host loop.  Ranks that are threads of this process have no NCCL communicator
    // at all (their small reductions go through the LocalGroup), so they always run persistent.
    const bool can_p2p = h->n_ranks > 1 && h->p2p_ready && h->np <= h->xcap;
    if (h->group && !can_p2p)
        return fail(CGE_B200_ERR_STATE, "in-process multi-GPU needs the peer exchange");
    const int driver =
        h->n_ranks > 1 ? ((can_p2p && (h->driver != CGE_B200_DRIVER_HOSTLOOP || h->group))
                              ? CGE_B200_DRIVER_PERSISTENT
                              : CGE_B200_DRIVER_HOSTLOOP)
        : h->driver == CGE_B200_DRIVER_AUTO
            ? CGE_B200_DRIVER_PERSISTENT  // measured: 65.3 us/pass vs 72.2 for the TMA ring (r01)
            : h->driver;
    S.driver = driver;
    S.regime = h->regime_reported;
    const bool stored = h->regime == CGE_B200_REGIME_STORED;
    // small problems (fewer tiles than resident CTAs): one run-time-exponent kernel for the whole
    // alpha grid instead of one instantiation per alpha (first-use load time, see powm_any)
    bool small = stored && (h->tile_end - h->tile_begin) < 2 * (int64_t)h->sm_count;
    if (const char *e = getenv("CGE_B200_RT_EXPONENT")) small = stored && atoi(e) != 0;
    S.ms_upload = h->ms_upload;
    if (h->star) {  // divergence.jl:332-334
        out[0] = -1.0;
        for (int i = 1; i < 6; ++i) out[i] = 0.0;
        *out_len = 6;
        return 0;
    }
    *out_len = 7;
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    h->launches = 0;
    h->ev_used = 0;
    h->ev_is_b.clear();
    const int n = (int)h->n, np = (int)h->np, nb = (int)h->nb, dp = (int)h->dp, k = (int)h->k;
    const long long tb = h->tile_begin, te = h->tile_end;
    // work units of this rank: tiles (stored regime) or super-tiles (recompute regime, one CTA per SM)
    const int local_tiles = stored ? (int)(te - tb) : (int)(h->st_end - h->st_begin);
    const int grid = std::max(1, std::min(local_tiles, (stored ? 2 : 1) * h->sm_count));
    const int build_tiles = (int)(te - tb), build_grid = std::max(1, std::min(build_tiles, 2 * h->sm_count));
    S.n_tiles = (int)h->n_tiles;
    S.grid = grid;
    S.matrix_bytes = stored ? (int64_t)build_tiles * TILE_ELEMS * 8
                            : (h->st_store_end > h->st_begin
                                   ? (h->st_pre_host[(size_t)h->st_store_end] - h->st_pre_host[(size_t)h->st_begin]) *
                                         (int64_t)TILE_ELEMS * 8
                                   : 0);
    // three phase markers come from the handle's event pool (no create/destroy per run, nothing
    // to leak on the error paths below)
    cudaEvent_t ev0 = h->next_event(), ev1 = h->next_event(), ev2 = h->next_event();
    CUDA_TRY(cudaEventRecord(ev0, st));

    // ---- distances, extrema, q (divergence.jl:79-93 / 359-375) ----
    unsigned long long *lohi = h->lohi.as<unsigned long long>();  // [0..1] scored graph, [2..3] full
    {
        const double pinf = inf;
        unsigned long long init[4];
        std::memcpy(&init[0], &pinf, 8);
        init[1] = 0ull;
        init[2] = init[0];
        init[3] = 0ull;
        CUDA_TRY(cudaMemcpyAsync(lohi, init, sizeof(init), cudaMemcpyHostToDevice, st));
    }
    RcArgs A = {};  // the stored-regime kernels take its SweepArgs base
    A.tile_ij = h->tile_ij.as<int2>();
    A.tile_begin = tb;
    A.tile_end = te;
    A.nb = nb; A.np = np; A.n = n; A.k = k;
    A.emb = h->emb.as<double>();
    A.nrm = h->rc_dot ? h->nrm.as<double>() : nullptr;
    A.diag = h->dist.as<double>();
    A.lohi = lohi;
    A.dp = dp;
    A.m = 1;
    A.st_ij = h->st_ij.as<int2>();
    A.st_begin = h->st_begin;
    A.st_end = h->st_end;
    A.sb = h->sb;
    A.nsb = (int)h->nsb;
    A.srow_begin = h->srow_begin;
    A.srow_end = h->srow_end;
    A.nchunk = dp / RC_DK;
    A.opT = h->opT.as<double>();
    A.qst = h->qst.as<double>();
    A.st_pre = h->st_pre.as<long long>();
    A.st_store_end = stored ? 0 : h->st_store_end;
    if (stored) {
        if (build_tiles > 0) {
            k_build_dist<true><<<build_grid, NTHREADS, 0, st>>>(h->emb.as<double>(), dp,
                                                                h->dist.as<double>(), n,
                                                                h->tile_ij.as<int2>(), tb, te,
                                                                h->q.as<double>(), lohi);
            ++h->launches;
        }
    } else if (local_tiles > 0) {  // recompute regime: only the extrema are needed up front, in
        launch_extrema_rc(grid, st, A, lohi, h->rc_dot);  // the arithmetic the passes will use
        ++h->launches;
    }
    if (h->n_ranks > 1)
        if (int rc = reduce_extrema_across_ranks(h, lohi, st)) return rc;
    if (!stored && h->st_store_end > h->st_begin) {  // store what fits: q of the leading super-tiles
        RcArgs Bld = A;
        Bld.st_end = h->st_store_end;
        Bld.st_store_end = h->st_begin;
        Bld.m = 1;
        launch_store_rc(std::max(1, std::min((int)(h->st_store_end - h->st_begin), h->sm_count)), st,
                        Bld, h->rc_dot);
        ++h->launches;
    }
    if (build_tiles > 0 && stored) {
        k_transform<<<4 * h->sm_count, 256, 0, st>>>(h->q.as<double>(),
                                                     (size_t)build_tiles * TILE_ELEMS, lohi);
        ++h->launches;
    }
    k_qdiag<<<(n + 255) / 256, 256, 0, st>>>(h->dist.as<double>(), n, lohi, h->qdiag.as<double>());
    ++h->launches;
    // ---- landmark mode: extrema of the full graph (divergence.jl:104-115 / 386-397) ----
    const long long SK = h->K * h->n_sets;
    if (h->K > 0) {
        if (h->landmark)
            if (int rc = full_graph_extrema(h, S, lohi + 2, dp, st)) return rc;  // lohi[2..3]
        const double *e = h->landmark ? h->emb_full.as<double>() : h->emb.as<double>();
        const double *dg = h->landmark ? nullptr : h->dist.as<double>();
        const unsigned long long *lh = h->landmark ? lohi + 2 : lohi;
        const int blocks = (int)((SK + 255) / 256);
        if (!stored && !h->landmark) {  // the sampled pairs get the bits the passes use
            const double *nr = h->rc_dot ? h->nrm.as<double>() : nullptr;
            launch_sample_q_dot(h->opT.as<double>(), dp / RC_DK, nr, e, dp, h->s_pda.as<int>(),
                                h->s_pdb.as<int>(), dg, lh, SK, h->s_pq.as<double>(), st);
            launch_sample_q_dot(h->opT.as<double>(), dp / RC_DK, nr, e, dp, h->s_nda.as<int>(),
                                h->s_ndb.as<int>(), dg, lh, SK, h->s_nq.as<double>(), st);
        } else {
            k_sample_q<<<blocks, 256, 0, st>>>(e, dp, h->s_pda.as<int>(), h->s_pdb.as<int>(), dg, lh,
                                               h->landmark, SK, h->s_pq.as<double>());
            k_sample_q<<<blocks, 256, 0, st>>>(e, dp, h->s_nda.as<int>(), h->s_ndb.as<int>(), dg, lh,
                                               h->landmark, SK, h->s_nq.as<double>());
        }
        h->launches += 2;
    }
    // ---- T (divergence.jl:118 / 399-402), partial slots ----
    CUDA_TRY(cudaMemcpyAsync(h->Ta.p, h->T0a.p, (size_t)np * 8, cudaMemcpyDeviceToDevice, st));
    if (h->directed)
        CUDA_TRY(cudaMemcpyAsync(h->Tb.p, h->T0b.p, (size_t)np * 8, cudaMemcpyDeviceToDevice, st));
    const size_t part_bytes = (size_t)(stored ? nb : (int)h->nsb) * np * 8;
    CUDA_TRY(cudaMemsetAsync(h->partA.p, 0, part_bytes, st));
    if (h->directed) CUDA_TRY(cudaMemsetAsync(h->partB.p, 0, part_bytes, st));
    CUDA_TRY(cudaMemsetAsync(h->slots.p, 0, 64, st));
    CUDA_TRY(cudaEventRecord(ev1, st));

    A.q = h->q.as<double>();
    {   // tile rows covered by [tb, te): row bi starts at tile bi*nb - bi(bi-1)/2
        auto row_of = [&](long long t) {
            int bi = 0;
            while (bi + 1 < nb && tile_index(nb, bi + 1, bi + 1) <= t) ++bi;
            return bi;
        };
        A.row_begin = te > tb ? row_of(tb) : 0;
        A.row_end = te > tb ? row_of(te - 1) : -1;
    }
    A.Ta = h->Ta.as<double>();
    A.Tb = h->Tb.as<double>();
    A.partA = h->partA.as<double>();
    A.partB = h->partB.as<double>();
    A.comm = h->comm.as<int>();
    A.B = h->B.as<double>();
    A.Tw_a = h->Ta.as<double>();
    A.Tw_b = h->Tb.as<double>();
    A.w_a = h->w.as<double>();
    A.w_b = h->w2.as<double>();
    A.qdiag = h->qdiag.as<double>();
    A.S_a = h->Sa.as<double>();
    A.S_b = h->Sb.as<double>();
    A.slots = h->slots.as<unsigned long long>();
    A.delta = 0.001;
    A.max_iter = 200000;
    {
        // L2-resident share of the matrix (MB); CGE_B200_L2_MB overrides the default
        double mb = 80.0;
        if (const char *e = getenv("CGE_B200_L2_MB")) mb = atof(e);
        A.resident_tiles = (long long)(mb * 1e6 / (TILE_ELEMS * 8.0));
        if (mb < 0) A.resident_tiles = -1;  // plain evict_normal loads
    }
    A.out_iters = reinterpret_cast<int *>(h->fpres.as<char>());
    A.out_diff = reinterpret_cast<double *>(h->fpres.as<char>() + 8);
    A.rank = h->rank;
    A.n_ranks = (driver == CGE_B200_DRIVER_PERSISTENT && h->n_ranks > 1) ? h->n_ranks : 1;
    A.xcap = h->xcap;
    for (int r = 0; r < 8; ++r) A.xbuf_peer[r] = reinterpret_cast<uint4 *>(h->xpeer[r]);
    A.pass_base = h->pass_total;
    A.phase_ns = nullptr;
    const bool phases = getenv("CGE_B200_PHASES") && atoi(getenv("CGE_B200_PHASES"));
    if (phases) {  // diagnostic: per-phase time of block 0 inside the persistent kernel
        A.phase_ns = h->slots.as<unsigned long long>() + 8;  // bytes 64..127 of the slots buffer
        CUDA_TRY(cudaMemsetAsync(A.phase_ns, 0, 64, st));
    }

    SampleSide sp, sn;
    if (h->K > 0) {
        sp.a = h->s_pa.as<int>(); sp.b = h->s_pb.as<int>();
        sn.a = h->s_na.as<int>(); sn.b = h->s_nb.as<int>();
        sp.q = h->s_pq.as<double>(); sn.q = h->s_nq.as<double>();
        if (h->landmark) {
            sp.w0a = h->s_pw0a.as<double>(); sp.wla = h->s_pwla.as<double>();
            sp.w0b = h->s_pw0b.as<double>(); sp.wlb = h->s_pwlb.as<double>();
            sn.w0a = h->s_nw0a.as<double>(); sn.wla = h->s_nwla.as<double>();
            sn.w0b = h->s_nw0b.as<double>(); sn.wlb = h->s_nwlb.as<double>();
        } else {
            sp.w0a = sp.wla = sp.w0b = sp.wlb = nullptr;
            sn.w0a = sn.wla = sn.w0b = sn.wlb = nullptr;
        }
    }

    // ---- alpha loop (divergence.jl:139-254 / 423-558) ----
    const double delta = 0.001;
    int alpha_div_counter = 5, alpha_auc_counter = 5;
    bool skip_div = false, skip_auc = h->K <= 0;
    double best_div = inf, best_div_ext = inf, best_div_int = inf, best_auc_err = inf,
           best_auc = inf, best_alpha = -1.0, best_alpha_auc = -1.0;
    unsigned long long *slots = h->slots.as<unsigned long long>();
    std::vector<double> Bh((size_t)k * k);
    const int ublocks = (n + 255) / 256;
    long long sweep_no = 0;
    // Deferred B sweep: the global score of alpha m needs one more read of the matrix with the
    // converged T, and so does the first fixed-point pass of alpha m+1 (T is warm-started), so when
    // alpha m+1 is certain to run, B of alpha m is computed by k_bfp<m+1> together with that pass
    // and its score is booked one iteration late -- same numbers, one matrix read less per alpha.
    // Each k_bfp<m> is one more kernel for CUDA to load on first use (~10 ms per instantiation in a fresh
    // process), more than the sweeps it saves on a small problem, so by default the sweep is deferred only
    // when a rank's share of the matrix is at least 1 GiB (a rank-independent test: every rank must take the
    // same path).  CGE_B200_FUSE_B=1 / 0 forces it on / off.
    bool can_fuse = stored && !small && !h->directed && driver == CGE_B200_DRIVER_PERSISTENT;
    if (const char *e = getenv("CGE_B200_FUSE_B"))
        can_fuse = can_fuse && atoi(e) != 0;
    else
        can_fuse = can_fuse && h->n_tiles * (int64_t)TILE_ELEMS * 8 / std::max(h->n_ranks, 1) >= ((int64_t)1 << 30);
    int pending_b = 0;  // exponent whose B rides on the next alpha's first pass (0: none)
    char *pin = static_cast<char *>(h->pinned);
    double *pin_auc = reinterpret_cast<double *>(pin + 64);
    double *pin_B = reinterpret_cast<double *>(pin + 64 + 2 * AUC_MAX_BLOCKS * 8);
    // queue the all-reduce (one rank per process) and the read-back of B behind its kernel
    auto fetch_B = [&]() -> int {
        if (h->n_ranks > 1 && !h->group)
            if (int rc = nccl_check(g_nccl.AllReduce(h->B.p, h->B.p, (size_t)k * k, kNcclF64,
                                                     kNcclSum, h->nccl_comm, st),
                                    "ncclAllReduce(B)"))
                return rc;
        CUDA_TRY(cudaMemcpyAsync(pin_B, h->B.p, (size_t)k * k * 8, cudaMemcpyDeviceToHost, st));
        return 0;
    };
    // after the stream is synchronized: B of exponent mb is in pin_B -> JS score, :226-252 / :530-556
    auto book_div = [&](int mb) -> int {
        if (h->n_ranks > 1 && h->group) {  // sum the ranks' B on the host, in rank order
            LocalGroup &G = *h->group;
            const size_t kk = (size_t)k * k;
            std::memcpy(G.B.data() + (size_t)h->rank * kk, pin_B, kk * 8);
            if (!G.barrier()) return fail(CGE_B200_ERR_STATE, "another rank failed");
            for (size_t i = 0; i < kk; ++i) {
                double acc = 0.0;
                for (int r = 0; r < G.n; ++r) acc += G.B[(size_t)r * kk + i];
                pin_B[i] = acc;
            }
            if (!G.barrier()) return fail(CGE_B200_ERR_STATE, "another rank failed");
        }
        std::memcpy(Bh.data(), pin_B, (size_t)k * k * 8);
        double f, div_int = 0.0, div_ext = 0.0;
        if (!h->split) {
            f = js_bins(h->C, Bh, h->bins, h->bin_internal, 0, 1);
        } else {
            div_int = js_bins(h->C, Bh, h->bins, h->bin_internal, 1, 1);
            div_ext = js_bins(h->C, Bh, h->bins, h->bin_internal, 1, 0);
            f = (div_int + div_ext) / 2.0;
        }
        S.div[mb - 1] = f;
        if (f < best_div) {  // :242-251
            best_div = f;
            best_alpha = 0.25 * mb;
            best_div_ext = !h->split ? 0.0 : div_ext;
            best_div_int = !h->split ? 0.0 : div_int;
            alpha_div_counter = 5;
        } else {
            alpha_div_counter -= 1;
            skip_div = alpha_div_counter == 0;
        }
        return 0;
    };
    for (int m = 1; m <= h->max_alphas; ++m) {
        const double alpha = 0.25 * m;
        A.m = m;
        double diff = 1.0, eps = h->directed ? 0.9 : 0.25;  // :150,:34 / :434-435
        int it = 0;
        const bool fused = pending_b > 0;  // == m - 1
        A.skip_first_tiles = 0;
        if (fused) {
            CUDA_TRY(cudaMemsetAsync(h->B.p, 0, (size_t)k * k * 8, st));
            if (local_tiles > 0) {
                cudaEventRecord(h->next_event(), st);
                launch_tiles(m, 4, grid, st, A);  // B of m-1 + pass 1 of m
                cudaEventRecord(h->next_event(), st);
                h->ev_is_b.push_back(2);  // carries a fixed-point pass: booked with the sweeps (and ms_fused)
                ++h->launches;
            }
            ++S.b_sweeps;
            ++S.b_fused;
            if (int rc = fetch_B()) return rc;
            A.skip_first_tiles = 1;
        }
        if (driver == CGE_B200_DRIVER_PERSISTENT || driver == CGE_B200_DRIVER_RING) {
            // one cooperative launch runs every pass of this alpha
            const bool ring = stored && driver == CGE_B200_DRIVER_RING;
            const void *fn = !stored ? fp_kernel_rc(h->directed, h->rc_dot)
                             : ring  ? fp_ring_kernel(m, h->directed)
                             : small ? fp_kernel_rt(h->directed)
                                     : fp_kernel(m, h->directed);
            const int threads = ring ? fp_ring_threads() : NTHREADS;
            const size_t smem = ring ? fp_ring_smem_bytes(h->directed) : (!stored ? rc_smem_bytes() : 0);
            if (smem > 0)
                CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)smem));
            int bps = 0;
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, fn, threads, smem));
            if (bps < 1) return fail(CGE_B200_ERR_CUDA, "fixed-point kernel does not fit on an SM");
            const int want = stored ? std::max(local_tiles, (n + 31) / 32)
                                    : std::max(local_tiles, 1);  // recompute: never more CTAs than units
            const int cgrid = std::max(1, std::min(want, bps * h->sm_count));
            A.eps0 = eps;
            A.pass_base = h->pass_total;
            CUDA_TRY(cudaMemsetAsync(h->slots.p, 0, 64, st));
            void *kargs[] = {(void *)&A};
            cudaEventRecord(h->next_event(), st);
            CUDA_TRY(cudaLaunchCooperativeKernel(fn, dim3(cgrid), dim3(threads), kargs, smem, st));
            cudaEventRecord(h->next_event(), st);
            h->ev_is_b.push_back(0);
            ++h->launches;
            S.grid = cgrid;
            // its result (pass count, residual) is fetched together with the scores below
            CUDA_TRY(cudaMemcpyAsync(h->pinned, h->fpres.p, 16, cudaMemcpyDeviceToHost, st));
        }
        while (driver == CGE_B200_DRIVER_HOSTLOOP && diff > delta) {  // :151 / :436
            if (local_tiles > 0) {
                cudaEventRecord(h->next_event(), st);
                if (!stored) launch_tiles_rc(h->directed ? 1 : 0, grid, st, A, h->rc_dot);
                else if (small) launch_tiles_rt(h->directed ? 1 : 0, grid, st, A);
                else launch_tiles(m, h->directed ? 1 : 0, grid, st, A);
                cudaEventRecord(h->next_event(), st);
                h->ev_is_b.push_back(0);
                ++h->launches;
            }
            if (stored) k_reduce_part<<<(n + 31) / 32, NTHREADS, 0, st>>>(A, A.partA, h->sraw_a.as<double>());
            else launch_reduce_part_rc(A, A.partA, h->sraw_a.as<double>(), st);
            ++h->launches;
            if (h->directed) {
                if (stored) k_reduce_part<<<(n + 31) / 32, NTHREADS, 0, st>>>(A, A.partB, h->sraw_b.as<double>());
                else launch_reduce_part_rc(A, A.partB, h->sraw_b.as<double>(), st);
                ++h->launches;
            }
            if (h->n_ranks > 1) {
                if (int rc = nccl_check(g_nccl.AllReduce(h->sraw_a.p, h->sraw_a.p, (size_t)n,
                                                         kNcclF64, kNcclSum, h->nccl_comm, st),
                                        "ncclAllReduce(S)"))
                    return rc;
                if (h->directed)
                    if (int rc = nccl_check(g_nccl.AllReduce(h->sraw_b.p, h->sraw_b.p, (size_t)n,
                                                             kNcclF64, kNcclSum, h->nccl_comm, st),
                                            "ncclAllReduce(Sout)"))
                        return rc;
            }
            unsigned long long *slot = slots + (sweep_no & 1), *next = slots + ((sweep_no + 1) & 1);
            if (h->directed)
                k_update_d<<<ublocks, 256, 0, st>>>(
                    h->sraw_a.as<double>(), h->sraw_b.as<double>(), h->Ta.as<double>(),
                    h->Tb.as<double>(), h->w.as<double>(), h->w2.as<double>(),
                    h->qdiag.as<double>(), m, n, eps, h->Sa.as<double>(), h->Sb.as<double>(), slot,
                    next);
            else
                k_update_u<<<ublocks, 256, 0, st>>>(h->sraw_a.as<double>(), h->Ta.as<double>(),
                                                    h->w.as<double>(), n, eps, h->Sa.as<double>(),
                                                    slot, next);
            ++h->launches;
            unsigned long long bits = 0;
            CUDA_TRY(cudaMemcpyAsync(&bits, slot, 8, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            double f;
            std::memcpy(&f, &bits, 8);
            if (h->directed && f > diff) eps *= 0.99;  // :462-464
            diff = f;
            ++it;
            ++sweep_no;
            ++S.fp_sweeps;
            if (it >= 200000)
                return fail(CGE_B200_ERR_STATE, "fixed point did not converge in 200000 passes");
        }
        // ---- local score (divergence.jl:178-224 / 478-528) and global score (:226-252 /
        // :530-556): both kernels are queued behind the fixed point, one sync per alpha ----
        const bool do_auc = !skip_auc;
        int auc_blocks = 0;
        if (do_auc) {
            const long long off = (h->n_sets > 1 ? (long long)(m - 1) : 0) * h->K;
            // undirected: T_a*T_b; directed: Tout of the source, Tin of the target (:488-490,:507)
            const double *fa = h->directed ? h->Tb.as<double>() : h->Ta.as<double>();
            const double *fb = h->Ta.as<double>();
            auc_blocks = (int)std::min<int64_t>(AUC_MAX_BLOCKS, (h->K + AUC_THREADS - 1) / AUC_THREADS);
            const int chunk = (int)((h->K + auc_blocks - 1) / auc_blocks);
            k_auc<<<auc_blocks, AUC_THREADS, 0, st>>>(sp, sn, h->s_pw.as<double>(), off, (int)h->K,
                                                      chunk, fa, fb, m, h->auc_out.as<double>());
            ++h->launches;
            CUDA_TRY(cudaMemcpyAsync(pin_auc, h->auc_out.p, (size_t)auc_blocks * 16,
                                     cudaMemcpyDeviceToHost, st));
        }
        bool synced = false;
        if (fused) {  // the previous alpha's global score, before this alpha's scores (:226-252)
            CUDA_TRY(cudaStreamSynchronize(st));
            synced = true;
            if (int rc = book_div(m - 1)) return rc;
            pending_b = 0;
        }
        const bool do_div = !skip_div;
        // alpha m+1 runs for certain when neither score can reach its patience limit at alpha m
        // (:215-223, :242-253: a counter of c before alpha m is >= c-1 after it)
        const bool defer = do_div && can_fuse && m < h->max_alphas &&
                           (alpha_div_counter >= 2 || (do_auc && alpha_auc_counter >= 2));
        if (do_div && !defer) {
            CUDA_TRY(cudaMemsetAsync(h->B.p, 0, (size_t)k * k * 8, st));
            if (local_tiles > 0) {
                cudaEventRecord(h->next_event(), st);
                if (!stored) launch_tiles_rc(h->directed ? 3 : 2, grid, st, A, h->rc_dot);
                else if (small) launch_tiles_rt(h->directed ? 3 : 2, grid, st, A);
                else launch_tiles(m, h->directed ? 3 : 2, grid, st, A);
                cudaEventRecord(h->next_event(), st);
                h->ev_is_b.push_back(1);
                ++h->launches;
            }
            ++S.b_sweeps;
            if (int rc = fetch_B()) return rc;
            synced = false;
        }
        if (!synced) CUDA_TRY(cudaStreamSynchronize(st));
        if (driver != CGE_B200_DRIVER_HOSTLOOP) {
            std::memcpy(&it, pin, 4);
            std::memcpy(&diff, pin + 8, 8);
            S.fp_sweeps += it;
            if (A.n_ranks > 1) h->pass_total += (unsigned)it;
            if (it >= A.max_iter)
                return fail(CGE_B200_ERR_STATE, "fixed point did not converge in 200000 passes");
        }
        S.iters[m - 1] = it;
        S.n_alpha_run = m;
        if (do_auc) {
            double num = 0.0, den = 0.0;
            for (int b = 0; b < auc_blocks; ++b) {
                num += pin_auc[2 * b];
                den += pin_auc[2 * b + 1];
            }
            const double auc = 1.0 - num / den;  // :213
            S.auc[m - 1] = auc;
            if (auc < best_auc) {  // :215-223
                best_auc = auc;
                best_auc_err = 1.96 * std::sqrt(auc * (1.0 - auc) / (double)h->K);
                best_alpha_auc = alpha;
                alpha_auc_counter = 5;
            } else {
                alpha_auc_counter -= 1;
                skip_auc = alpha_auc_counter == 0;
            }
        }
        if (do_div && !defer) {
            if (int rc = book_div(m)) return rc;
        } else if (defer) {
            pending_b = m;
        }
        if (skip_div && skip_auc) break;  // :253
    }
    CUDA_TRY(cudaEventRecord(ev2, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventElapsedTime(&S.ms_build, ev0, ev1));
    CUDA_TRY(cudaEventElapsedTime(&S.ms_solve, ev1, ev2));
    for (size_t i = 0; i < h->ev_is_b.size(); ++i) {
        float ms = 0.f;  // pool entries 0..2 are the phase markers, sweep pairs follow
        if (cudaEventElapsedTime(&ms, h->evpool[3 + 2 * i], h->evpool[4 + 2 * i]) == cudaSuccess)
            (h->ev_is_b[i] == 1 ? S.ms_bsweeps : S.ms_sweeps) += ms;
        if (h->ev_is_b[i] == 2) S.ms_fused += ms;
    }
    {
        unsigned long long lh[4];
        CUDA_TRY(cudaMemcpy(lh, lohi, sizeof(lh), cudaMemcpyDeviceToHost));
        std::memcpy(&S.lo, &lh[0], 8);
        std::memcpy(&S.hi, &lh[1], 8);
        std::memcpy(&S.hi_full, &lh[3], 8);
    }
    if (phases) {
        unsigned long long ph[8];
        CUDA_TRY(cudaMemcpy(ph, A.phase_ns, 64, cudaMemcpyDeviceToHost));
        const double np_ = (double)std::max<int64_t>(S.fp_sweeps, 1) * 1e3;
        fprintf(stderr,
                "[cge_b200 rank %d] us per pass (block 0): tiles %.2f | barrier A %.2f | reduce%s %.2f | "
                "peer wait+sum+update %.2f | barrier C %.2f\n",
                h->rank, ph[0] / np_, ph[1] / np_, A.n_ranks > 1 ? "+peer stores" : "+update",
                ph[2] / np_, ph[5] / np_, ph[6] / np_);
    }
    out[0] = best_alpha; out[1] = best_div; out[2] = best_div_ext; out[3] = best_div_int;
    out[4] = best_alpha_auc; out[5] = best_auc; out[6] = best_auc_err;  // :256
    S.launches = h->launches;
    S.ms_total =
        std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - wall0).count();
    return 0;
}

}  // namespace cge

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

void cge_b200_version(int *major, int *minor, int *patch) {
    if (major) *major = CGE_B200_VERSION_MAJOR;
    if (minor) *minor = CGE_B200_VERSION_MINOR;
    if (patch) *patch = CGE_B200_VERSION_PATCH;
}

int cge_b200_device_count(void) {
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return c;
}

const char *cge_b200_last_error(void) { return g_err.c_str(); }

int cge_b200_create(int device, cge_b200_handle **out) {
    if (!out) return fail(CGE_B200_ERR_ARG, "out is NULL");
    *out = nullptr;
    int c = cge_b200_device_count();
    if (c <= 0) return fail(CGE_B200_ERR_CUDA, "no CUDA device is usable (there is no CPU fallback)");
    if (device < 0 || device >= c) return fail(CGE_B200_ERR_ARG, "device index out of range");
    CUDA_TRY(cudaSetDevice(device));
    cge_b200_handle *h = new cge_b200_handle();
    h->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete h;
        return fail(CGE_B200_ERR_CUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(e));
    }
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
    *out = h;
    return 0;
}

void cge_b200_destroy(cge_b200_handle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->nccl_comm);
    for (int r = 0; r < 8; ++r)
        if (h->xpeer[r] && h->xpeer[r] != h->xbuf) cudaIpcCloseMemHandle(h->xpeer[r]);
    if (h->xbuf) cudaFree(h->xbuf);
    for (DevBuf *b :
         {&h->q, &h->tile_ij, &h->tile_ij_full, &h->emb, &h->emb_full, &h->dist, &h->w, &h->w2,
          &h->T0a, &h->T0b, &h->Ta, &h->Tb, &h->Sa, &h->Sb, &h->sraw_a, &h->sraw_b, &h->partA,
          &h->partB, &h->comm, &h->B, &h->qdiag, &h->lohi, &h->slots, &h->auc_out, &h->fpres, &h->st_ij, &h->opT, &h->nrm, &h->rc_mean, &h->st_pre, &h->qst, &h->diam_strips, &h->diam_packed, &h->diam_norms, &h->diam_tilemax,
          &h->diam_list, &h->diam_ctr, &h->diam_mean, &h->s_pda, &h->s_pdb, &h->s_nda, &h->s_ndb, &h->s_pa,
          &h->s_pb, &h->s_na, &h->s_nb, &h->s_pw, &h->s_pw0a, &h->s_pwla, &h->s_pw0b, &h->s_pwlb,
          &h->s_nw0a, &h->s_nwla, &h->s_nw0b, &h->s_nwlb, &h->s_pq, &h->s_nq})
        b->release();
    for (cudaEvent_t e : h->evpool) cudaEventDestroy(e);
    if (h->pinned) cudaFreeHost(h->pinned);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int cge_b200_upload(cge_b200_handle *h, const cge_b200_problem *p) {
    if (!h) return fail(CGE_B200_ERR_ARG, "handle is NULL");
    return do_upload(h, p);
}

int cge_b200_run(cge_b200_handle *h, double *out, int32_t *out_len, cge_b200_stats *stats) {
    if (!h || !out || !out_len) return fail(CGE_B200_ERR_ARG, "NULL argument");
    if (stats && stats->struct_size != 0 && stats->struct_size != (int32_t)sizeof(cge_b200_stats))
        return fail(CGE_B200_ERR_ARG, "cge_b200_stats.struct_size mismatch");
    return do_run(h, out, out_len, stats);
}

int cge_b200_score_multi(const cge_b200_problem *p, int n_gpus, double *out, int32_t *out_len,
                         cge_b200_stats *stats) {
    if (!p || !out || !out_len) return fail(CGE_B200_ERR_ARG, "NULL argument");
    if (n_gpus < 2 || n_gpus > 8) return fail(CGE_B200_ERR_ARG, "n_gpus must be 2..8");
    if (cge_b200_device_count() < n_gpus) return fail(CGE_B200_ERR_CUDA, "not enough CUDA devices");
    if (p->n_full > 0)
        return fail(CGE_B200_ERR_ARG, "landmark-mode runs stay on one GPU (replicas only)");
    LocalGroup G;
    G.n = n_gpus;
    G.lohi.assign((size_t)2 * n_gpus, 0ull);
    std::vector<cge_b200_handle *> hs((size_t)n_gpus, nullptr);
    std::vector<int> rcs((size_t)n_gpus, 0);
    std::vector<std::string> errs((size_t)n_gpus);
    std::vector<double> outs((size_t)n_gpus * 7, 0.0);
    std::vector<int32_t> lens((size_t)n_gpus, 7);
    std::vector<cge_b200_stats> sts((size_t)n_gpus);
    // n = maximum(edges) decides the exchange capacity and the size of the B scratch
    int64_t n = 0, k = 0;
    for (int64_t e = 0; e < p->m; ++e)
        n = std::max(n, std::max(p->edge_src[e], p->edge_dst[e]) - p->index_base + 1);
    for (int64_t i = 0; i < std::min(n, p->n_comm); ++i) k = std::max(k, p->comm[i] - p->index_base + 1);
    if (n <= 0 || k <= 0) return fail(CGE_B200_ERR_ARG, "empty problem");
    G.B.assign((size_t)n_gpus * (size_t)(k * k), 0.0);
    auto worker = [&](int r) {
        auto bail = [&](int rc) {
            rcs[(size_t)r] = rc;
            errs[(size_t)r] = g_err;
            G.fail();
        };
        int rc = cge_b200_create(r, &hs[(size_t)r]);
        if (rc) return bail(rc);
        cge_b200_handle *h = hs[(size_t)r];
        h->rank = r;
        h->n_ranks = n_gpus;
        h->group = &G;
        // exchange buffer: same layout as cge_b200_p2p_export, reached by plain peer access
        h->xcap = (n + TILE - 1) / TILE * TILE;
        const size_t bytes = (size_t)2 * n_gpus * 2 * (size_t)h->xcap * 16;
        if (cudaMalloc(&h->xbuf, bytes) != cudaSuccess || cudaMemset(h->xbuf, 0, bytes) != cudaSuccess ||
            cudaDeviceSynchronize() != cudaSuccess) {
            cudaGetLastError();
            fail(CGE_B200_ERR_OOM, "exchange buffer allocation failed");
            return bail(CGE_B200_ERR_OOM);
        }
        for (int q = 0; q < n_gpus; ++q)
            if (q != r) {
                cudaError_t e = cudaDeviceEnablePeerAccess(q, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                    fail(CGE_B200_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
                    return bail(CGE_B200_ERR_CUDA);
                }
                cudaGetLastError();
            }
        if (!G.barrier()) return;
        for (int q = 0; q < n_gpus; ++q) h->xpeer[q] = hs[(size_t)q]->xbuf;
        h->p2p_ready = true;
        rc = cge_b200_upload(h, p);
        if (rc) return bail(rc);
        if (!G.barrier()) return;  // nobody launches a kernel that waits for a rank that gave up
        rc = cge_b200_run(h, &outs[(size_t)r * 7], &lens[(size_t)r], &sts[(size_t)r]);
        if (rc) return bail(rc);
    };
    std::vector<std::thread> th;
    for (int r = 0; r < n_gpus; ++r) th.emplace_back(worker, r);
    for (auto &t : th) t.join();
    int rc = 0;
    for (int r = 0; r < n_gpus && !rc; ++r)
        if (rcs[(size_t)r]) {
            rc = rcs[(size_t)r];
            g_err = "rank " + std::to_string(r) + ": " + errs[(size_t)r];
        }
    if (!rc && G.failed.load()) rc = fail(CGE_B200_ERR_STATE, "a rank failed");
    for (int r = 0; r < n_gpus; ++r)
        if (hs[(size_t)r]) {
            for (int q = 0; q < 8; ++q) hs[(size_t)r]->xpeer[q] = nullptr;  // not IPC mappings
            cge_b200_destroy(hs[(size_t)r]);
        }
    if (rc) return rc;
    std::memcpy(out, outs.data(), 7 * sizeof(double));
    *out_len = lens[0];
    if (stats) *stats = sts[0];
    return 0;
}

int cge_b200_score(const cge_b200_problem *p, double *out, int32_t *out_len,
                   cge_b200_stats *stats) {
    if (!p || !out || !out_len) return fail(CGE_B200_ERR_ARG, "NULL argument");
    // CGE_B200_GPUS=N shards an exact-mode call over N GPUs of this box without any change to
    // the caller (the Julia wrapper keeps calling cge_b200_score)
    if (const char *e = getenv("CGE_B200_GPUS")) {
        const int g = atoi(e);
        if (g > 1 && p->n_full == 0) return cge_b200_score_multi(p, g, out, out_len, stats);
    }
    cge_b200_handle *h = nullptr;
    int rc = cge_b200_create(0, &h);
    if (rc) return rc;
    rc = cge_b200_upload(h, p);
    if (!rc) rc = cge_b200_run(h, out, out_len, stats);
    cge_b200_destroy(h);
    return rc;
}

int cge_b200_comm_id_size(void) { return (int)sizeof(Id128); }

int cge_b200_comm_unique_id(void *id_bytes) {
    if (!id_bytes) return fail(CGE_B200_ERR_ARG, "id_bytes is NULL");
    if (int rc = nccl_load()) return rc;
    return nccl_check(g_nccl.GetUniqueId(id_bytes), "ncclGetUniqueId");
}

int cge_b200_comm_init(cge_b200_handle *h, const void *id_bytes, int rank, int n_ranks) {
    if (!h || !id_bytes || n_ranks < 1 || rank < 0 || rank >= n_ranks)
        return fail(CGE_B200_ERR_ARG, "bad comm_init argument");
    if (n_ranks == 1) {
        h->rank = 0;
        h->n_ranks = 1;
        return 0;
    }
    if (int rc = nccl_load()) return rc;
    CUDA_TRY(cudaSetDevice(h->device));
    Id128 id;
    std::memcpy(&id, id_bytes, sizeof(id));
    if (int rc = nccl_check(g_nccl.CommInitRank(&h->nccl_comm, n_ranks, id, rank),
                            "ncclCommInitRank"))
        return rc;
    h->rank = rank;
    h->n_ranks = n_ranks;
    h->uploaded = false;
    return 0;
}

int cge_b200_p2p_handle_size(void) { return (int)sizeof(cudaIpcMemHandle_t); }

int cge_b200_p2p_export(cge_b200_handle *h, int64_t max_vertices, void *handle_out) {
    if (!h || !handle_out || max_vertices <= 0) return fail(CGE_B200_ERR_ARG, "bad p2p_export argument");
    if (h->n_ranks < 2 || h->n_ranks > 8)
        return fail(CGE_B200_ERR_STATE, "p2p exchange needs comm_init with 2..8 ranks first");
    CUDA_TRY(cudaSetDevice(h->device));
    if (h->xbuf) {
        cudaFree(h->xbuf);
        h->xbuf = nullptr;
    }
    h->p2p_ready = false;
    h->xcap = (max_vertices + TILE - 1) / TILE * TILE;
    // [2 parities][n_ranks writers][2][xcap] records of 16 bytes (value halves + pass number)
    const size_t bytes = (size_t)2 * h->n_ranks * 2 * (size_t)h->xcap * 16;
    CUDA_TRY(cudaMalloc(&h->xbuf, bytes));
    CUDA_TRY(cudaMemset(h->xbuf, 0, bytes));
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t *>(handle_out), h->xbuf));
    return 0;
}

int cge_b200_p2p_import(cge_b200_handle *h, const void *all_handles) {
    if (!h || !all_handles) return fail(CGE_B200_ERR_ARG, "bad p2p_import argument");
    if (!h->xbuf) return fail(CGE_B200_ERR_STATE, "p2p_import before p2p_export");
    CUDA_TRY(cudaSetDevice(h->device));
    const cudaIpcMemHandle_t *hs = static_cast<const cudaIpcMemHandle_t *>(all_handles);
    for (int r = 0; r < h->n_ranks; ++r) {
        if (r == h->rank) {
            h->xpeer[r] = h->xbuf;
        } else {
            void *ptr = nullptr;
            CUDA_TRY(cudaIpcOpenMemHandle(&ptr, hs[r], cudaIpcMemLazyEnablePeerAccess));
            h->xpeer[r] = ptr;
        }
    }
    // a fresh exchange: pass numbers restart at 1, so no record of an earlier one may survive.
    // Peers write into this buffer only from inside run(), which every rank enters after its own
    // import (the caller's collective upload / barrier comes in between).
    CUDA_TRY(cudaMemset(h->xbuf, 0, (size_t)2 * h->n_ranks * 2 * (size_t)h->xcap * 16));
    CUDA_TRY(cudaDeviceSynchronize());
    h->pass_total = 0;
    h->p2p_ready = true;
    return 0;
}

int cge_b200_measure_fp64_peak(cge_b200_handle *h, double *tflops) {
    if (!h || !tflops) return fail(CGE_B200_ERR_ARG, "NULL argument");
    CUDA_TRY(cudaSetDevice(h->device));
    *tflops = measure_fp64_peak_tflops(h->sm_count, h->stream);
    if (*tflops <= 0.0) return fail(CGE_B200_ERR_CUDA, "FP64 peak measurement failed");
    return 0;
}

int cge_b200_measure_fp64_pipes(cge_b200_handle *h, double out[8]) {
    if (!h || !out) return fail(CGE_B200_ERR_ARG, "NULL argument");
    CUDA_TRY(cudaSetDevice(h->device));
    if (measure_fp64_pipes(h->sm_count, h->stream, out) != 0)
        return fail(CGE_B200_ERR_CUDA, "FP64 pipe measurement failed");
    return 0;
}

int cge_b200_sample_non_edges(cge_b200_handle *h, int64_t n, int64_t m, const int64_t *edge_src,
                              const int64_t *edge_dst, int32_t index_base, int32_t directed,
                              int64_t n_samples, int64_t n_sets, uint64_t seed, int64_t *out_i,
                              int64_t *out_j, double *draws_per_sample) {
    if (!h || n < 2 || n >= ((int64_t)1 << 31) || m < 0 || (m > 0 && (!edge_src || !edge_dst)) ||
        (index_base != 0 && index_base != 1) || n_samples <= 0 || n_sets <= 0 || !out_i || !out_j)
        return fail(CGE_B200_ERR_ARG, "bad sample_non_edges argument");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const int64_t total = n_samples * n_sets;
    uint64_t slots = 1024;
    while (slots < 2 * (uint64_t)m) slots <<= 1;  // load factor <= 0.5
    DevBuf src, dst, table, counts, oi, oj;
    auto release = [&]() {
        for (DevBuf *b : {&src, &dst, &table, &counts, &oi, &oj}) b->release();
    };
    int rc = 0;
    if (!rc) rc = src.ensure((size_t)std::max<int64_t>(m, 1) * 8);
    if (!rc) rc = dst.ensure((size_t)std::max<int64_t>(m, 1) * 8);
    if (!rc) rc = table.ensure((size_t)slots * 8);
    if (!rc) rc = counts.ensure(32);
    if (!rc) rc = oi.ensure((size_t)total * 8);
    if (!rc) rc = oj.ensure((size_t)total * 8);
    unsigned long long hc[4] = {0, 0, 0, 0};
    cudaError_t e = cudaSuccess;
    auto step = [&](cudaError_t r) {
        if (e == cudaSuccess) e = r;
    };
    if (!rc) {
        if (m > 0) {
            step(cudaMemcpyAsync(src.p, edge_src, (size_t)m * 8, cudaMemcpyHostToDevice, st));
            step(cudaMemcpyAsync(dst.p, edge_dst, (size_t)m * 8, cudaMemcpyHostToDevice, st));
        }
        step(cudaMemsetAsync(table.p, 0xFF, (size_t)slots * 8, st));
        step(cudaMemsetAsync(counts.p, 0, 32, st));
        if (m > 0 && e == cudaSuccess)
            launch_edge_set_insert(src.as<long long>(), dst.as<long long>(), m, n, index_base,
                                   directed, table.as<unsigned long long>(), slots - 1,
                                   counts.as<unsigned long long>(), 4 * h->sm_count, st);
        step(cudaMemcpyAsync(hc, counts.p, 32, cudaMemcpyDeviceToHost, st));
        step(cudaStreamSynchronize(st));
        step(cudaGetLastError());
    }
    const uint64_t candidates = directed ? (uint64_t)n * (uint64_t)(n - 1)
                                         : (uint64_t)n * (uint64_t)(n - 1) / 2;
    if (!rc && e == cudaSuccess) {
        if (hc[1] > 0)
            rc = fail(CGE_B200_ERR_ARG, std::to_string(hc[1]) + " edges with an endpoint outside 1..n");
        else if (hc[0] >= candidates)  // sample() of an empty collection (divergence.jl:137 leaves NE empty)
            rc = fail(CGE_B200_ERR_ARG, "collection must be non-empty: the graph has no non-edges");
    }
    if (!rc && e == cudaSuccess) {
        launch_sample_non_edges(n, directed, index_base, table.as<unsigned long long>(), slots - 1,
                                seed, total, oi.as<long long>(), oj.as<long long>(),
                                counts.as<unsigned long long>(), 4 * h->sm_count, st);
        step(cudaMemcpyAsync(out_i, oi.p, (size_t)total * 8, cudaMemcpyDeviceToHost, st));
        step(cudaMemcpyAsync(out_j, oj.p, (size_t)total * 8, cudaMemcpyDeviceToHost, st));
        step(cudaMemcpyAsync(hc, counts.p, 32, cudaMemcpyDeviceToHost, st));
        step(cudaStreamSynchronize(st));
        step(cudaGetLastError());
        if (e == cudaSuccess && hc[2] > 0)
            rc = fail(CGE_B200_ERR_STATE, std::to_string(hc[2]) +
                                              " samples found no non-edge in 2^22 draws each");
        if (draws_per_sample) *draws_per_sample = (double)hc[3] / (double)total;
    }
    release();
    if (rc) return rc;
    CUDA_TRY(e);
    return 0;
}

int cge_b200_landmarks_select(cge_b200_handle *h, int64_t n, int64_t d, const double *embed,
                              int64_t embed_row_stride, int64_t embed_col_stride, const double *vweights,
                              int64_t n_clusters, const int64_t *cluster_ptr, const int64_t *cluster_members,
                              int32_t index_base, int64_t land, int64_t forced, int32_t rule,
                              cge_b200_eigvec_fn eig, void *eig_user, int64_t *out_group, int64_t *out_cuts) {
    if (!h || n <= 0 || d <= 0 || n >= ((int64_t)1 << 31) || d > 4096 || !embed || !vweights ||
        n_clusters <= 0 || !cluster_ptr || !cluster_members || !out_group || land <= 0 || forced <= 0 ||
        (index_base != 0 && index_base != 1) ||
        (rule != CGE_B200_RULE_RSS && rule != CGE_B200_RULE_SIZE && rule != CGE_B200_RULE_DIAMETER))
        return fail(CGE_B200_ERR_ARG, "bad landmarks_select argument");
    if (cluster_ptr[0] != 0 || cluster_ptr[n_clusters] != n)
        return fail(CGE_B200_ERR_ARG, "the clusters must cover every vertex exactly once");
    std::vector<int> members((size_t)n);
    std::vector<char> seen((size_t)n, 0);
    for (int64_t c = 0; c < n_clusters; ++c)
        if (cluster_ptr[c + 1] < cluster_ptr[c])
            return fail(CGE_B200_ERR_ARG, "cluster_ptr must be non-decreasing");
    for (int64_t i = 0; i < n; ++i) {
        const int64_t v = cluster_members[i] - index_base;
        if (v < 0 || v >= n || seen[(size_t)v])
            return fail(CGE_B200_ERR_ARG, "the clusters must cover every vertex exactly once");
        seen[(size_t)v] = 1;
        members[(size_t)i] = (int)v;
    }
    CUDA_TRY(cudaSetDevice(h->device));
    // row-major copy of the embedding (the caller's matrix may be column-major)
    std::vector<double> x;
    const double *xr = embed;
    if (!(embed_col_stride == 1 && embed_row_stride == d)) {
        x.resize((size_t)(n * d));
        for (int64_t i = 0; i < n; ++i)
            for (int64_t j = 0; j < d; ++j)
                x[(size_t)(i * d + j)] = embed[i * embed_row_stride + j * embed_col_stride];
        xr = x.data();
    }
    std::string msg;
    long long cuts = 0;
    static_assert(sizeof(long long) == sizeof(int64_t), "int64_t is long long here");
    const int rc = landmarks_select_device(h->device, h->stream, n, (int)d, xr, vweights, n_clusters,
                                           reinterpret_cast<const long long *>(cluster_ptr), members.data(),
                                           land, forced, rule, reinterpret_cast<SelectEigFn>(eig), eig_user,
                                           reinterpret_cast<long long *>(out_group), &cuts, msg);
    if (out_cuts) *out_cuts = cuts;
    if (rc == -1) return fail(CGE_B200_ERR_CUDA, msg);
    if (rc == -2) return fail(CGE_B200_ERR_STATE, msg);
    return 0;
}

int cge_b200_unique_rows(cge_b200_handle *h, int64_t n, int64_t d, const double *embed,
                         int64_t embed_row_stride, int64_t embed_col_stride, int64_t *out_count) {
    if (!h || n <= 0 || d <= 0 || n >= ((int64_t)1 << 31) || !embed || !out_count)
        return fail(CGE_B200_ERR_ARG, "bad unique_rows argument");
    CUDA_TRY(cudaSetDevice(h->device));
    std::vector<double> x;
    const double *xr = embed;
    if (!(embed_col_stride == 1 && embed_row_stride == d)) {
        x.resize((size_t)(n * d));
        for (int64_t i = 0; i < n; ++i)
            for (int64_t j = 0; j < d; ++j)
                x[(size_t)(i * d + j)] = embed[i * embed_row_stride + j * embed_col_stride];
        xr = x.data();
    }
    std::string msg;
    long long c = 0;
    if (count_unique_rows_device(h->stream, n, (int)d, xr, &c, msg) != 0) return fail(CGE_B200_ERR_CUDA, msg);
    *out_count = c;
    return 0;
}

int cge_b200_sym_top_eigvec(const double *a, int64_t d, double *v_out, double *lambda) {
    if (!a || !v_out || d <= 0 || d > 4096) return fail(CGE_B200_ERR_ARG, "bad sym_top_eigvec argument");
    sym_top_eigvec(a, (int)d, v_out, lambda);
    return 0;
}

int cge_b200_landmarks_aggregate(cge_b200_handle *h, int64_t n, int64_t d, int64_t n_landmarks,
                                 const int64_t *landmark, int32_t index_base, const double *vweights,
                                 const int64_t *comm, const double *embed, int64_t embed_row_stride,
                                 int64_t embed_col_stride, int64_t m, const int64_t *edge_src,
                                 const int64_t *edge_dst, const double *eweights, int32_t directed,
                                 double *out_embed, double *out_lweight, double *out_dii,
                                 int64_t *out_cluster, int64_t *out_edge_src, int64_t *out_edge_dst,
                                 double *out_eweights, int64_t edge_cap, int64_t *out_n_edges) {
    if (!h || n <= 0 || d <= 0 || n_landmarks <= 0 || n >= ((int64_t)1 << 31) ||
        n_landmarks >= ((int64_t)1 << 31) || m < 0 || m >= ((int64_t)1 << 31) || !landmark ||
        !vweights || !embed || !out_embed || !out_lweight || !out_dii ||
        (index_base != 0 && index_base != 1) || (out_cluster && !comm) ||
        (m > 0 && (!edge_src || !edge_dst || !eweights || !out_edge_src || !out_edge_dst ||
                   !out_eweights || !out_n_edges)))
        return fail(CGE_B200_ERR_ARG, "bad landmarks_aggregate argument");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const int64_t N = n_landmarks;
    // row-major copy of the embedding (the caller's matrix may be column-major)
    std::vector<double> x((size_t)(n * d));
    for (int64_t i = 0; i < n; ++i)
        for (int64_t j = 0; j < d; ++j)
            x[(size_t)(i * d + j)] = embed[i * embed_row_stride + j * embed_col_stride];
    std::vector<int64_t> zero_comm;
    if (!comm) {
        zero_comm.assign((size_t)n, 0);
        comm = zero_comm.data();
    }
    DevBuf d_lm, d_vw, d_comm, d_x, d_src, d_dst, d_ew, d_embed, d_lw, d_dii, d_cl, d_oa, d_ob, d_ow;
    auto release = [&]() {
        for (DevBuf *b : {&d_lm, &d_vw, &d_comm, &d_x, &d_src, &d_dst, &d_ew, &d_embed, &d_lw, &d_dii,
                          &d_cl, &d_oa, &d_ob, &d_ow})
            b->release();
    };
    const size_t mm = (size_t)std::max<int64_t>(m, 1);
    int rc = 0;
    if (!rc) rc = upload_vec(d_lm, landmark, (size_t)n * 8, st);
    if (!rc) rc = upload_vec(d_vw, vweights, (size_t)n * 8, st);
    if (!rc) rc = upload_vec(d_comm, comm, (size_t)n * 8, st);
    if (!rc) rc = upload_vec(d_x, x.data(), x.size() * 8, st);
    if (!rc) rc = upload_vec(d_src, edge_src, (size_t)m * 8, st);
    if (!rc) rc = upload_vec(d_dst, edge_dst, (size_t)m * 8, st);
    if (!rc) rc = upload_vec(d_ew, eweights, (size_t)m * 8, st);
    if (!rc) rc = d_embed.ensure((size_t)(N * d) * 8);
    if (!rc) rc = d_lw.ensure((size_t)N * 8);
    if (!rc) rc = d_dii.ensure((size_t)N * 8);
    if (!rc) rc = d_cl.ensure((size_t)N * 8);
    if (!rc) rc = d_oa.ensure(mm * 8);
    if (!rc) rc = d_ob.ensure(mm * 8);
    if (!rc) rc = d_ow.ensure(mm * 8);
    int cells = 0, bad = 0;
    cudaError_t e = cudaSuccess;
    if (!rc)
        e = landmarks_aggregate_device((int)n, (int)d, (int)N, index_base, d_lm.as<long long>(),
                                       d_vw.as<double>(), d_comm.as<long long>(), d_x.as<double>(), m,
                                       d_src.as<long long>(), d_dst.as<long long>(), d_ew.as<double>(),
                                       directed, d_embed.as<double>(), d_lw.as<double>(),
                                       d_dii.as<double>(), d_cl.as<long long>(), d_oa.as<long long>(),
                                       d_ob.as<long long>(), d_ow.as<double>(), &cells, &bad, st);
    std::vector<long long> oa((size_t)cells), ob((size_t)cells), cl((size_t)N);
    std::vector<double> ow((size_t)cells);
    auto step = [&](cudaError_t r) {
        if (e == cudaSuccess) e = r;
    };
    if (!rc && e == cudaSuccess && bad == 0) {
        step(cudaMemcpyAsync(out_embed, d_embed.p, (size_t)(N * d) * 8, cudaMemcpyDeviceToHost, st));
        step(cudaMemcpyAsync(out_lweight, d_lw.p, (size_t)N * 8, cudaMemcpyDeviceToHost, st));
        step(cudaMemcpyAsync(out_dii, d_dii.p, (size_t)N * 8, cudaMemcpyDeviceToHost, st));
        step(cudaMemcpyAsync(cl.data(), d_cl.p, (size_t)N * 8, cudaMemcpyDeviceToHost, st));
        if (cells > 0) {
            step(cudaMemcpyAsync(oa.data(), d_oa.p, (size_t)cells * 8, cudaMemcpyDeviceToHost, st));
            step(cudaMemcpyAsync(ob.data(), d_ob.p, (size_t)cells * 8, cudaMemcpyDeviceToHost, st));
            step(cudaMemcpyAsync(ow.data(), d_ow.p, (size_t)cells * 8, cudaMemcpyDeviceToHost, st));
        }
        step(cudaStreamSynchronize(st));
    }
    release();
    if (rc) return rc;
    CUDA_TRY(e);
    if (bad > 0)
        return fail(CGE_B200_ERR_ARG, std::to_string(bad) + " landmark ids or edge endpoints out of range");
    if (out_cluster)
        for (int64_t L = 0; L < N; ++L) out_cluster[L] = cl[(size_t)L];
    int64_t w = 0;
    for (int c = 0; c < cells; ++c) {  // landmark_edges[landmark_edges[:,3] .> 0, :] (landmarks.jl:461)
        if (!(ow[(size_t)c] > 0.0)) continue;
        if (w >= edge_cap) return fail(CGE_B200_ERR_ARG, "edge_cap too small");
        out_edge_src[w] = oa[(size_t)c] + index_base;
        out_edge_dst[w] = ob[(size_t)c] + index_base;
        out_eweights[w] = ow[(size_t)c];
        ++w;
    }
    if (out_n_edges) *out_n_edges = w;
    return 0;
}

int cge_b200_selftest_math(cge_b200_handle *h, int64_t n_samples, uint64_t seed,
                           int64_t *sqrt_mismatches, int64_t *div_mismatches) {
    if (!h || n_samples <= 0 || !sqrt_mismatches || !div_mismatches)
        return fail(CGE_B200_ERR_ARG, "bad selftest_math argument");
    CUDA_TRY(cudaSetDevice(h->device));
    unsigned long long *dev = nullptr, host[2] = {0, 0};
    CUDA_TRY(cudaMalloc(&dev, 16));
    cudaError_t e = cudaMemsetAsync(dev, 0, 16, h->stream);
    if (e == cudaSuccess) {
        launch_selftest_math(n_samples, seed, dev, 8 * h->sm_count, h->stream);
        e = cudaMemcpyAsync(host, dev, 16, cudaMemcpyDeviceToHost, h->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(dev);
    CUDA_TRY(e);
    *sqrt_mismatches = (int64_t)host[0];
    *div_mismatches = (int64_t)host[1];
    return 0;
}

int cge_b200_shard_plan(int64_t n, int rank, int n_ranks, int64_t *n_tiles, int64_t *tile_begin,
                        int64_t *tile_end) {
    if (n <= 0 || n_ranks < 1 || rank < 0 || rank >= n_ranks || !n_tiles || !tile_begin ||
        !tile_end)
        return fail(CGE_B200_ERR_ARG, "bad shard_plan argument");
    const int64_t nb = (n + TILE - 1) / TILE;
    *n_tiles = nb * (nb + 1) / 2;
    shard_range(*n_tiles, rank, n_ranks, tile_begin, tile_end);
    return 0;
}

int cge_b200_debug_read(cge_b200_handle *h, int what, double *buf, int64_t n_elems) {
    if (!h || !buf) return fail(CGE_B200_ERR_ARG, "NULL argument");
    if (!h->uploaded || h->star) return fail(CGE_B200_ERR_STATE, "nothing uploaded");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    const int64_t n = h->n;
    if (what == 0) {
        if (n_elems < n * n) return fail(CGE_B200_ERR_ARG, "buffer too small");
        if (h->n_ranks != 1) return fail(CGE_B200_ERR_STATE, "dense read needs all tiles local");
        if (h->regime != CGE_B200_REGIME_STORED)
            return fail(CGE_B200_ERR_STATE, "no stored matrix in the recompute regime");
        std::vector<double> tile((size_t)TILE_ELEMS);
        for (int64_t bi = 0; bi < h->nb; ++bi)
            for (int64_t bj = bi; bj < h->nb; ++bj) {
                const int64_t t = tile_index(h->nb, bi, bj);
                CUDA_TRY(cudaMemcpy(tile.data(), h->q.as<double>() + (size_t)t * TILE_ELEMS,
                                    (size_t)TILE_ELEMS * 8, cudaMemcpyDeviceToHost));
                for (int r = 0; r < TILE; ++r)
                    for (int c = 0; c < TILE; ++c) {
                        const int64_t si = bi * TILE + r, sj = bj * TILE + c;
                        if (si >= n || sj >= n) continue;
                        const int64_t vi = h->perm[(size_t)si], vj = h->perm[(size_t)sj];
                        buf[vi * n + vj] = tile[(size_t)r * TILE + c];
                        if (bi != bj) buf[vj * n + vi] = tile[(size_t)r * TILE + c];
                    }
            }
        return 0;
    }
    const DevBuf *src = what == 1 ? &h->Ta : what == 2 ? &h->Tb : what == 3 ? &h->Sa
                        : what == 4 ? &h->Sb : nullptr;
    if (!src || !src->p) return fail(CGE_B200_ERR_ARG, "unknown or unavailable probe");
    if (n_elems < n) return fail(CGE_B200_ERR_ARG, "buffer too small");
    std::vector<double> tmp((size_t)n);
    CUDA_TRY(cudaMemcpy(tmp.data(), src->p, (size_t)n * 8, cudaMemcpyDeviceToHost));
    for (int64_t s = 0; s < n; ++s) buf[h->perm[(size_t)s]] = tmp[(size_t)s];
    return 0;
}

}  // extern "C"
