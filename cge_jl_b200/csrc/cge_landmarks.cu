// cge_landmarks.cu -- SURVEY.md 8(f) F2: the aggregation half of landmarks()
// (/root/reference/src/landmarks.jl:387-463) on the device.  Landmark SELECTION (runsplit, the PCA
// bisection of :155-345) stays on the host; given its vertex -> landmark assignment this file
// produces every landmark-mode input of the scorer:
//   embed[L][j]  = sum_{i in L} vweights[i] * x[i][j] / lweight[L]           (:387-404)
//   lweight[L]   = sum_{i in L} vweights[i]
//   dii[L]       = sqrt( sum_{i in L} sum_j (embed[L][j] - x[i][j])^2 / lweight[L] )   (:407-423; the
//                  numerator is unweighted, as in the reference)
//   cluster[L]   = comm of the landmark's last member                         (:426-430)
//   wedges       = edge weights summed per (landmark, landmark) cell, cells in idx order, > 0 only (:433-463)
//
// Bit-identical to the reference's loops: every sum runs in the reference's order.  Vertices are
// bucketed by landmark with a STABLE radix sort (members in ascending vertex order, which is the
// order the Julia loops meet them), each (landmark, dimension) and each landmark is summed by one
// thread over its members in that order, products and sums are rounded separately (__dmul_rn /
// __dadd_rn: no FMA contraction, as in Julia), the inner sum over the dimensions of d_ii starts at 0
// and runs over ascending j.  Edges are keyed by their cell, stably sorted, and each cell is summed
// by one thread in edge order.
#include <cub/cub.cuh>

#include "cge_landmarks.cuh"

namespace cge {

__global__ void k_lm_keys(const long long *__restrict__ lm, int base, int n, int *__restrict__ key,
                          int *__restrict__ val, int *bad, int N) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long L = lm[i] - base;
    if (L < 0 || L >= N) atomicAdd(bad, 1);
    key[i] = (int)L;
    val[i] = i;
}

// first member position of every landmark in the sorted list (keys ascending); start[N] = n
__global__ void k_lm_starts(const int *__restrict__ key, int n, int N, int *__restrict__ start) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p > n) return;
    const int cur = p < n ? key[p] : N, prev = p > 0 ? key[p - 1] : -1;
    for (int L = prev + 1; L <= cur; ++L) start[L] = p;  // empty landmarks get an empty range
}

// one thread per (landmark, dimension): weighted sum in member order, then the division (:390-404)
__global__ void k_lm_centroid(const int *__restrict__ start, const int *__restrict__ member,
                              const double *__restrict__ vw, const double *__restrict__ x, int d, int N,
                              double *__restrict__ embed, double *__restrict__ lweight) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)N * d) return;
    const int L = (int)(t / d), j = (int)(t % d);
    double acc = 0.0, w = 0.0;
    for (int p = start[L]; p < start[L + 1]; ++p) {
        const int i = member[p];
        const double wi = vw[i];
        w = __dadd_rn(w, wi);
        acc = __dadd_rn(acc, __dmul_rn(wi, x[(size_t)i * d + j]));
    }
    embed[(size_t)L * d + j] = acc / w;
    if (j == 0) lweight[L] = w;
}

// one thread per vertex: its squared distance to the centroid, dimensions in ascending order (:409-415)
__global__ void k_lm_sqdev(const long long *__restrict__ lm, int base, const double *__restrict__ x,
                           const double *__restrict__ embed, int d, int n, int N,
                           double *__restrict__ dev) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long L = lm[i] - base;
    if (L < 0 || L >= N) {  // reported by k_lm_keys; the call fails afterwards
        dev[i] = 0.0;
        return;
    }
    const double *c = embed + (size_t)L * d, *xi = x + (size_t)i * d;
    double dist = 0.0;
    for (int j = 0; j < d; ++j) {
        const double t = __dadd_rn(c[j], -xi[j]);
        dist = __dadd_rn(dist, __dmul_rn(t, t));
    }
    dev[i] = dist;
}

// one thread per landmark: d_ii and the community of its last member (:416-430)
__global__ void k_lm_dii(const int *__restrict__ start, const int *__restrict__ member,
                         const double *__restrict__ dev, const double *__restrict__ lweight,
                         const long long *__restrict__ comm, int N, double *__restrict__ dii,
                         long long *__restrict__ cluster) {
    const int L = blockIdx.x * blockDim.x + threadIdx.x;
    if (L >= N) return;
    double acc = 0.0;
    for (int p = start[L]; p < start[L + 1]; ++p) acc = __dadd_rn(acc, dev[member[p]]);
    const double w = lweight[L];
    dii[L] = w > 0.0 ? sqrt(acc / w) : acc;
    cluster[L] = start[L + 1] > start[L] ? comm[member[start[L + 1] - 1]] : 0;
}

// edge -> cell key a*N + b (undirected: a <= b), value = edge index (:436-437 / :450-451)
__global__ void k_lm_edge_keys(const long long *__restrict__ src, const long long *__restrict__ dst,
                               const long long *__restrict__ lm, int base, long long m, int n, int N,
                               int directed, unsigned long long *__restrict__ key,
                               int *__restrict__ val, int *bad) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= m) return;
    const long long u = src[e] - base, v = dst[e] - base;
    if (u < 0 || v < 0 || u >= n || v >= n) {
        atomicAdd(bad, 1);
        key[e] = 0;
        val[e] = (int)e;
        return;
    }
    long long a = lm[u] - base, b = lm[v] - base;
    if (a < 0 || b < 0 || a >= N || b >= N) a = b = 0;  // reported by k_lm_keys
    if (!directed && a > b) {
        const long long t = a;
        a = b;
        b = t;
    }
    key[e] = (unsigned long long)(a * N + b);
    val[e] = (int)e;
}

__global__ void k_lm_heads(const unsigned long long *__restrict__ key, long long m, int *__restrict__ head) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= m) return;
    head[e] = (e == 0 || key[e] != key[e - 1]) ? 1 : 0;
}

// position p is the head of cell number seg[p] (exclusive scan of the head flags): record where the cell starts
__global__ void k_lm_cell_starts(const int *__restrict__ head, const int *__restrict__ seg, long long m,
                                 int *__restrict__ cell_start) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= m) return;
    if (head[e]) cell_start[seg[e]] = (int)e;
}

// one thread per cell: the edge weights of the cell in edge order (the sort is stable) (:437 / :451)
__global__ void k_lm_cells(const unsigned long long *__restrict__ key, const int *__restrict__ val,
                           const int *__restrict__ cell_start, int n_cells, long long m,
                           const double *__restrict__ ew, int N, long long *__restrict__ oa,
                           long long *__restrict__ ob, double *__restrict__ ow) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    const long long p0 = cell_start[c], p1 = c + 1 < n_cells ? cell_start[c + 1] : m;
    double acc = 0.0;
    for (long long p = p0; p < p1; ++p) acc = __dadd_rn(acc, ew[val[p]]);
    const unsigned long long k = key[p0];
    oa[c] = (long long)(k / (unsigned long long)N);
    ob[c] = (long long)(k % (unsigned long long)N);
    ow[c] = acc;
}

#define LM_TRY(expr)                      \
    do {                                  \
        cudaError_t e__ = (expr);         \
        if (e__ != cudaSuccess) return e__; \
    } while (0)

// Device pipeline; all pointers are device pointers.  Returns the number of cells in *n_cells and the
// count of out-of-range ids in *n_bad.
cudaError_t landmarks_aggregate_device(int n, int d, int N, int base, const long long *lm,
                                       const double *vw, const long long *comm, const double *x,
                                       long long m, const long long *src, const long long *dst,
                                       const double *ew, int directed, double *embed, double *lweight,
                                       double *dii, long long *cluster, long long *oa, long long *ob,
                                       double *ow, int *n_cells, int *n_bad, cudaStream_t st) {
    int *key = nullptr, *val = nullptr, *key2 = nullptr, *val2 = nullptr, *start = nullptr, *bad = nullptr;
    double *dev = nullptr;
    unsigned long long *ek = nullptr, *ek2 = nullptr;
    int *ev = nullptr, *ev2 = nullptr, *head = nullptr, *seg = nullptr, *cstart = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0, need = 0;
    auto cleanup = [&]() {
        for (void *p : {(void *)key, (void *)val, (void *)key2, (void *)val2, (void *)start, (void *)bad,
                        (void *)dev, (void *)ek, (void *)ek2, (void *)ev, (void *)ev2, (void *)head,
                        (void *)seg, (void *)cstart, tmp})
            if (p) cudaFree(p);
    };
    auto run = [&]() -> cudaError_t {
        const size_t mm = (size_t)(m > 0 ? m : 1);
        LM_TRY(cudaMalloc(&key, (size_t)n * 4));
        LM_TRY(cudaMalloc(&val, (size_t)n * 4));
        LM_TRY(cudaMalloc(&key2, (size_t)n * 4));
        LM_TRY(cudaMalloc(&val2, (size_t)n * 4));
        LM_TRY(cudaMalloc(&start, (size_t)(N + 1) * 4));
        LM_TRY(cudaMalloc(&bad, 8));
        LM_TRY(cudaMalloc(&dev, (size_t)n * 8));
        LM_TRY(cudaMalloc(&ek, mm * 8));
        LM_TRY(cudaMalloc(&ek2, mm * 8));
        LM_TRY(cudaMalloc(&ev, mm * 4));
        LM_TRY(cudaMalloc(&ev2, mm * 4));
        LM_TRY(cudaMalloc(&head, mm * 4));
        LM_TRY(cudaMalloc(&seg, mm * 4));
        LM_TRY(cudaMalloc(&cstart, mm * 4));
        // scratch of the three CUB calls
        LM_TRY(cub::DeviceRadixSort::SortPairs(nullptr, need, key, key2, val, val2, n, 0, 32, st));
        tmp_bytes = need;
        LM_TRY(cub::DeviceRadixSort::SortPairs(nullptr, need, ek, ek2, ev, ev2, (int)mm, 0, 64, st));
        tmp_bytes = need > tmp_bytes ? need : tmp_bytes;
        LM_TRY(cub::DeviceScan::ExclusiveSum(nullptr, need, head, seg, (int)mm, st));
        tmp_bytes = need > tmp_bytes ? need : tmp_bytes;
        LM_TRY(cudaMalloc(&tmp, tmp_bytes));
        LM_TRY(cudaMemsetAsync(bad, 0, 8, st));
        const int T = 256;
        // members of every landmark in ascending vertex order
        k_lm_keys<<<(n + T - 1) / T, T, 0, st>>>(lm, base, n, key, val, bad, N);
        need = tmp_bytes;
        LM_TRY(cub::DeviceRadixSort::SortPairs(tmp, need, key, key2, val, val2, n, 0, 32, st));
        k_lm_starts<<<(n + 1 + T - 1) / T, T, 0, st>>>(key2, n, N, start);
        k_lm_centroid<<<(unsigned)(((long long)N * d + T - 1) / T), T, 0, st>>>(start, val2, vw, x, d, N, embed,
                                                                              lweight);
        k_lm_sqdev<<<(n + T - 1) / T, T, 0, st>>>(lm, base, x, embed, d, n, N, dev);
        k_lm_dii<<<(N + T - 1) / T, T, 0, st>>>(start, val2, dev, lweight, comm, N, dii, cluster);
        int cells = 0;
        if (m > 0) {
            k_lm_edge_keys<<<(unsigned)((m + T - 1) / T), T, 0, st>>>(src, dst, lm, base, m, n, N, directed, ek,
                                                                     ev, bad + 1);
            need = tmp_bytes;
            LM_TRY(cub::DeviceRadixSort::SortPairs(tmp, need, ek, ek2, ev, ev2, (int)m, 0, 64, st));
            k_lm_heads<<<(unsigned)((m + T - 1) / T), T, 0, st>>>(ek2, m, head);
            need = tmp_bytes;
            LM_TRY(cub::DeviceScan::ExclusiveSum(tmp, need, head, seg, (int)m, st));
            k_lm_cell_starts<<<(unsigned)((m + T - 1) / T), T, 0, st>>>(head, seg, m, cstart);
            int last_seg = 0, last_head = 0;
            LM_TRY(cudaMemcpyAsync(&last_seg, seg + (m - 1), 4, cudaMemcpyDeviceToHost, st));
            LM_TRY(cudaMemcpyAsync(&last_head, head + (m - 1), 4, cudaMemcpyDeviceToHost, st));
            LM_TRY(cudaStreamSynchronize(st));
            cells = last_seg + last_head;
            k_lm_cells<<<(cells + T - 1) / T, T, 0, st>>>(ek2, ev2, cstart, cells, m, ew, N, oa, ob, ow);
        }
        int hb[2] = {0, 0};
        LM_TRY(cudaMemcpyAsync(hb, bad, 8, cudaMemcpyDeviceToHost, st));
        LM_TRY(cudaStreamSynchronize(st));
        LM_TRY(cudaGetLastError());
        *n_cells = cells;
        *n_bad = hb[0] + hb[1];
        return cudaSuccess;
    };
    const cudaError_t rc = run();
    cleanup();
    return rc;
}

}  // namespace cge
