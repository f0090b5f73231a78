// cge_select.cu -- SURVEY.md 8(f) F4: landmark SELECTION on the device
// (/root/reference/src/landmarks.jl: runsplit :279-345, the split rules split_cluster_rss :155-210,
// split_cluster_size :218-238, split_cluster_diameter :247-267, the heap :12-46, WSSE :50-66).
//
// What the reference does: a priority queue of vertex clusters keyed by -RSS; the cluster with the
// largest RSS is popped and cut in two along its first principal component (weighted PCA of its
// embedding rows) by the chosen rule, until `land` clusters exist (first every initial cluster is
// cut into at most `forced` pieces on a local queue).  The queue is sequential by construction and
// tiny (one entry per landmark); all the work is in the O(s d^2) and O(s d) passes over the rows of
// the cluster being cut, which at 10^6 vertices x 128 dimensions is ~1 GB per level of the bisection.
//
// Here the embedding, the weights and the vertex order live in HBM for the whole selection:
//   * every cluster is a contiguous SEGMENT of one index array, kept in the reference's member
//     order (cuts are stable partitions, so ties are broken as the reference breaks them);
//   * per cut, kernels compute the weighted moments (mean, RSS), the d x d weighted covariance, the
//     projection z on the principal axis, a stable radix sort of z (cub), the moments of the
//     ranges of z the rss rule asks for, and the stable regrouping of the segment;
//   * the host keeps what is O(d^2) or O(log s) per cut: the queue, the principal axis of the d x d
//     covariance (Householder tridiagonalisation + bisection + inverse iteration), the medians and
//     the rule's comparisons of O(d) moment vectors.
// All sums have a fixed order (per-block partials added in block order): bit-reproducible run to run.
//
// Parity.  The reference's cut depends on LAPACK's eigenvector including its SIGN, which LAPACK leaves
// to its algorithm (the sign decides which child is "low" and on which side the median element of an
// odd cluster falls).  So the d x d eigenproblem can be handed back to the HOST's LAPACK through a
// callback -- Julia's eigvecs in the Julia binding, numpy.linalg.eigh in the Python mirror: the very
// routine the reference (or the mirror) calls, on a covariance that agrees with theirs to rounding.
// Without a callback the built-in solver below is used and the sign is fixed (largest-magnitude
// component positive).  Tests hold the vertex -> landmark assignment equal, label by label, to the
// host mirror's for both choices.
#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "cge_landmarks.cuh"

namespace cge {

void set_last_error(const std::string &msg);

namespace {

constexpr int SEL_TX = 32, SEL_TY = 8;  // moments / projection blocks: 32 dimension lanes x 8 row lanes
constexpr int COV_THREADS = 256, COV_ROWS = 32;
constexpr int MAX_GROUPS = 96;

struct Range {
    long long off;  // first element (position in the segment's order, or rank in its sorted order)
    long long len;
};

// ---- moments of up to two ranges: out[range][block][1 + 2 d] = (ws, s[d], ss[d]) partials ----------
// rows are addressed as vertex = idx[seg_off + (via ? via[seg_off + r] : r)]
__global__ void k_sel_moments(const double *__restrict__ x, const double *__restrict__ w,
                              const int *__restrict__ idx, const int *__restrict__ via, long long seg_off,
                              Range r0, Range r1, int d, double *__restrict__ out) {
    const Range rg = blockIdx.y == 0 ? r0 : r1;
    const int nblk = gridDim.x;
    const long long per = (rg.len + nblk - 1) / nblk;
    const long long lo = rg.off + (long long)blockIdx.x * per;
    const long long hi = min(rg.off + rg.len, lo + per);
    extern __shared__ __align__(16) double sm[];  // [SEL_TY][1 + 2 d]
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int stride = 1 + 2 * d;
    double *mine = sm + (size_t)ty * stride;
    for (int j = tx; j < stride; j += SEL_TX) mine[j] = 0.0;
    __syncwarp();
    // a warp (= one ty) takes rows lo + ty, lo + ty + 8, ...; lane tx the dimensions tx, tx + 32, ...
    for (int j0 = 0; j0 < d; j0 += SEL_TX * 4) {  // four dimensions per lane and sweep over the rows
        double s[4] = {0, 0, 0, 0}, ss[4] = {0, 0, 0, 0}, ws = 0.0;
        for (long long r = lo + ty; r < hi; r += SEL_TY) {
            const int v = idx[seg_off + (via ? via[seg_off + r] : r)];
            const double wv = w[v];
            ws += wv;
            const double *row = x + (size_t)v * d;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int j = j0 + tx + k * SEL_TX;
                if (j < d) {
                    const double xv = row[j], wx = wv * xv;
                    s[k] += wx;
                    ss[k] = fma(wx, xv, ss[k]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int j = j0 + tx + k * SEL_TX;
            if (j < d) {
                mine[1 + j] = s[k];
                mine[1 + d + j] = ss[k];
            }
        }
        if (j0 == 0 && tx == 0) mine[0] = ws;
    }
    __syncthreads();
    double *dst = out + ((size_t)blockIdx.y * nblk + blockIdx.x) * stride;
    for (int j = ty * SEL_TX + tx; j < stride; j += SEL_TX * SEL_TY) {
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < SEL_TY; ++q) acc += sm[(size_t)q * stride + j];
        dst[j] = acc;
    }
}

// sums the per-block partials in block order: fin[range][comp]
__global__ void k_sel_sum_blocks(const double *__restrict__ part, int nblk, int comps, int nranges,
                                 double *__restrict__ fin) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= comps * nranges) return;
    const int rg = t / comps, c = t % comps;
    double acc = 0.0;
    for (int b = 0; b < nblk; ++b) acc += part[((size_t)rg * nblk + b) * comps + c];
    fin[t] = acc;
}

// ---- weighted covariance of a segment: C = sum_r w_r (x_r - mu)(x_r - mu)^T, upper 4x4 tiles -------
__global__ void __launch_bounds__(COV_THREADS) k_sel_cov(const double *__restrict__ x,
                                                         const double *__restrict__ w,
                                                         const int *__restrict__ idx, long long seg_off,
                                                         long long len, const double *__restrict__ mu,
                                                         int d, int dp, double *__restrict__ part) {
    extern __shared__ __align__(16) double sm[];  // [COV_ROWS][dp]: y = (x - mu) sqrt(w), zero beyond d
    const int nblk = gridDim.x;
    const long long per = (len + nblk - 1) / nblk;
    const long long lo = (long long)blockIdx.x * per, hi = min(len, lo + per);
    const int nt = dp / 4, ntile = nt * nt;
    double *mine = part + (size_t)blockIdx.x * dp * dp;
    for (int e = threadIdx.x; e < dp * dp; e += COV_THREADS) mine[e] = 0.0;
    for (long long c0 = lo; c0 < hi; c0 += COV_ROWS) {
        const int rows = (int)min((long long)COV_ROWS, hi - c0);
        __syncthreads();
        for (int e = threadIdx.x; e < COV_ROWS * dp; e += COV_THREADS) {
            const int r = e / dp, j = e % dp;
            double y = 0.0;
            if (r < rows && j < d) {
                const int v = idx[seg_off + c0 + r];
                y = (x[(size_t)v * d + j] - mu[j]) * sqrt(w[v]);
            }
            sm[e] = y;
        }
        __syncthreads();
        for (int t = threadIdx.x; t < ntile; t += COV_THREADS) {
            const int ti = t / nt, tj = t % nt;
            if (tj < ti) continue;
            double acc[4][4] = {};
            for (int r = 0; r < rows; ++r) {
                const double *yr = sm + (size_t)r * dp;
                const double2 a01 = *reinterpret_cast<const double2 *>(yr + 4 * ti);
                const double2 a23 = *reinterpret_cast<const double2 *>(yr + 4 * ti + 2);
                const double2 b01 = *reinterpret_cast<const double2 *>(yr + 4 * tj);
                const double2 b23 = *reinterpret_cast<const double2 *>(yr + 4 * tj + 2);
                const double a[4] = {a01.x, a01.y, a23.x, a23.y}, b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
                for (int p = 0; p < 4; ++p)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[p][q] = fma(a[p], b[q], acc[p][q]);
            }
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int q = 0; q < 4; ++q) mine[(size_t)(4 * ti + p) * dp + 4 * tj + q] += acc[p][q];
        }
    }
}

// ---- projection: z_r = sum_j ((x_rj - mu_j) sqrt(w_r)) v_j; also the local position as payload ----
__global__ void k_sel_project(const double *__restrict__ x, const double *__restrict__ w,
                              const int *__restrict__ idx, long long seg_off, long long len,
                              const double *__restrict__ mu, const double *__restrict__ v, int d,
                              double *__restrict__ z, int *__restrict__ pos) {
    const long long r = (long long)blockIdx.x * SEL_TY + threadIdx.y;
    if (r >= len) return;
    const int vx = idx[seg_off + r];
    const double sw = sqrt(w[vx]);
    const double *row = x + (size_t)vx * d;
    double acc = 0.0;
    for (int j = threadIdx.x; j < d; j += SEL_TX) acc = fma((row[j] - mu[j]) * sw, v[j], acc);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (threadIdx.x == 0) {
        z[seg_off + r] = acc;
        pos[seg_off + r] = (int)r;
    }
}

// rank[perm[r]] = r  (perm: local positions in ascending z)
__global__ void k_sel_invert(const int *__restrict__ perm, long long seg_off, long long len,
                             int *__restrict__ rank) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < len) rank[seg_off + perm[seg_off + r]] = (int)r;
}

struct Groups {
    int n;
    int end[MAX_GROUPS];     // group g holds the sorted ranks [end[g-1], end[g])
    int layout[MAX_GROUPS];  // its place in the new order of the segment
};
// key of the element at local position p: the layout place of the group its rank falls in
__global__ void k_sel_group_keys(const int *__restrict__ rank, long long seg_off, long long len,
                                 Groups g, unsigned *__restrict__ key) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= len) return;
    const int r = rank[seg_off + p];
    int q = 0;
    while (q < g.n - 1 && r >= g.end[q]) ++q;
    key[seg_off + p] = (unsigned)g.layout[q];
}
__global__ void k_sel_side_keys(const unsigned char *__restrict__ side, long long seg_off, long long len,
                                unsigned *__restrict__ key) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < len) key[seg_off + p] = side[p];
}

// ---- number of distinct embedding rows (landmarks.jl:369: size(unique(embedding, dims=1), 1)) ----------
// one warp per row: 64-bit hash of the row's bit patterns (isequal semantics, as Julia's unique)
__global__ void k_row_hash(const double *__restrict__ x, long long n, int d, unsigned long long *__restrict__ h,
                           int *__restrict__ id) {
    const long long r = (long long)blockIdx.x * SEL_TY + threadIdx.y;
    if (r >= n) return;
    const unsigned long long *row = reinterpret_cast<const unsigned long long *>(x) + (size_t)r * d;
    unsigned long long acc = 0x9E3779B97F4A7C15ull * (unsigned long long)(threadIdx.x + 1);
    for (int j = threadIdx.x; j < d; j += SEL_TX) {
        unsigned long long v = row[j] + 0x9E3779B97F4A7C15ull * (unsigned long long)(j + 1);
        v = (v ^ (v >> 30)) * 0xBF58476D1CE4E5B9ull;  // splitmix64 finaliser
        v = (v ^ (v >> 27)) * 0x94D049BB133111EBull;
        acc += v ^ (v >> 31);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (threadIdx.x == 0) {
        h[r] = acc;
        id[r] = (int)r;
    }
}
// rows sorted by hash: position p starts a new distinct row unless its hash AND its content equal those of
// p - 1 (one warp per position compares the rows only when the hashes agree)
__global__ void k_row_distinct(const double *__restrict__ x, long long n, int d,
                               const unsigned long long *__restrict__ h, const int *__restrict__ id,
                               unsigned long long *__restrict__ count) {
    const long long p = (long long)blockIdx.x * SEL_TY + threadIdx.y;
    if (p >= n) return;
    bool fresh = p == 0 || h[p] != h[p - 1];
    if (!fresh) {
        const unsigned long long *a = reinterpret_cast<const unsigned long long *>(x) + (size_t)id[p] * d;
        const unsigned long long *b = reinterpret_cast<const unsigned long long *>(x) + (size_t)id[p - 1] * d;
        bool diff = false;
        for (int j = threadIdx.x; j < d; j += SEL_TX) diff |= a[j] != b[j];
        fresh = __any_sync(0xffffffffu, diff);
    }
    if (fresh && threadIdx.x == 0) atomicAdd(count, 1ull);
}


// ---------------------------------------------------------------------------------------------------
// host: the reference's queue (landmarks.jl:12-46), entries are segments
// ---------------------------------------------------------------------------------------------------
struct Entry {
    long long off, len;
    double value;
    std::vector<double> mom;  // (ws, s[d], ss[d]) of the segment; empty for singletons
};
void heap_put(std::vector<Entry> &pq, Entry e) {  // landmark_put!
    pq.push_back(Entry());
    size_t i = pq.size();
    while (i / 2 >= 1) {
        const size_t j = i / 2;
        if (e.value < pq[j - 1].value) {
            pq[i - 1] = std::move(pq[j - 1]);
            i = j;
        } else {
            break;
        }
    }
    pq[i - 1] = std::move(e);
}
Entry heap_pop(std::vector<Entry> &pq) {  // landmark_pop!
    Entry x = std::move(pq[0]);
    Entry y = std::move(pq.back());
    pq.pop_back();
    if (!pq.empty()) {
        size_t i = 1;
        const size_t len = pq.size();
        while (2 * i <= len) {
            const size_t l = 2 * i, r = l + 1;
            const size_t j = (r > len || pq[l - 1].value < pq[r - 1].value) ? l : r;
            if (pq[j - 1].value < y.value) {
                pq[i - 1] = std::move(pq[j - 1]);
                i = j;
            } else {
                break;
            }
        }
        pq[i - 1] = std::move(y);
    }
    return x;
}

// sum over the dimensions of ss - s^2 / ws (wsse, landmarks.jl:63; total_rss :269)
double total_of(const double *m, int d) {
    double t = 0.0;
    for (int j = 0; j < d; ++j) t += m[1 + d + j] - m[1 + j] * m[1 + j] / m[0];
    return t;
}
void add_to(std::vector<double> &a, const double *b) {
    for (size_t i = 0; i < a.size(); ++i) a[i] += b[i];
}

struct Sel {
    int d, dp, device;
    long long n;
    cudaStream_t st;
    // device
    double *x = nullptr, *w = nullptr, *z = nullptr, *zs = nullptr, *mu = nullptr, *vec = nullptr,
           *part = nullptr, *fin = nullptr, *covp = nullptr;
    int *idx = nullptr, *idx2 = nullptr, *pos = nullptr, *perm = nullptr, *rank = nullptr;
    unsigned *key = nullptr, *key2 = nullptr;
    unsigned char *side = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    int sm_count = 148, max_mblk = 0, max_cblk = 0;
    // host
    // pinned, so that the many small device -> host reads of a cut are real asynchronous copies
    double *h_fin = nullptr, *h_cov = nullptr, *h_z = nullptr;  // [2 (1+2d)], [dp dp], [n] (z or sorted z)
    cudaError_t err = cudaSuccess;
    std::string msg;
    long long n_cuts = 0;

    bool ok(cudaError_t e) {
        if (e != cudaSuccess && err == cudaSuccess) err = e;
        return err == cudaSuccess;
    }
    template <typename T>
    bool alloc(T *&p, size_t count) {
        return ok(cudaMalloc(reinterpret_cast<void **>(&p), std::max<size_t>(count, 1) * sizeof(T)));
    }
    ~Sel() {
        for (void *p : {(void *)x, (void *)w, (void *)z, (void *)zs, (void *)mu, (void *)vec, (void *)part,
                        (void *)fin, (void *)covp, (void *)idx, (void *)idx2, (void *)pos, (void *)perm,
                        (void *)rank, (void *)key, (void *)key2, (void *)side, tmp})
            if (p) cudaFree(p);
        for (void *p : {(void *)h_fin, (void *)h_cov, (void *)h_z})
            if (p) cudaFreeHost(p);
    }

    int mom_blocks(long long len) const {
        return (int)std::max<long long>(1, std::min<long long>(max_mblk, (len + 255) / 256));
    }
    // moments of one or two ranges of segment seg_off into h_fin[range * (1 + 2d) ...]
    bool moments(long long seg_off, const int *via, Range r0, Range r1, int nranges) {
        const int comps = 1 + 2 * d;
        const int nblk = mom_blocks(std::max(r0.len, nranges > 1 ? r1.len : 0));
        const size_t smem = (size_t)SEL_TY * comps * 8;
        k_sel_moments<<<dim3(nblk, nranges), dim3(SEL_TX, SEL_TY), smem, st>>>(x, w, idx, via, seg_off, r0,
                                                                               r1, d, part);
        k_sel_sum_blocks<<<(comps * nranges + 127) / 128, 128, 0, st>>>(part, nblk, comps, nranges, fin);
        if (!ok(cudaMemcpyAsync(h_fin, fin, (size_t)nranges * comps * 8, cudaMemcpyDeviceToHost, st)))
            return false;
        return ok(cudaStreamSynchronize(st));
    }
};

}  // namespace

// ---------------------------------------------------------------------------------------------------
// runsplit (landmarks.jl:279-345) with the cuts on the device
// rule: 0 = split_cluster_rss, 2 = split_cluster_size, 3 = split_cluster_diameter
// clusters: CSR over 0-based vertex ids, already in the order of sort(initial_clusters)
// returns 0, or a negative code: -1 CUDA (see msg), -2 the reference's ErrorException / assertion
// ---------------------------------------------------------------------------------------------------
int landmarks_select_device(int device, cudaStream_t st, long long n, int d, const double *x_rowmajor,
                            const double *vweights, long long n_clusters, const long long *cl_ptr,
                            const int *cl_members, long long land, long long forced, int rule,
                            SelectEigFn eig, void *eig_user, long long *out_group, long long *out_cuts,
                            std::string &msg) {
    Sel S;
    S.d = d;
    S.dp = (d + 3) / 4 * 4;
    S.n = n;
    S.st = st;
    S.device = device;
    cudaDeviceGetAttribute(&S.sm_count, cudaDevAttrMultiProcessorCount, device);
    S.max_mblk = 4 * S.sm_count;
    S.max_cblk = S.sm_count;
    const int comps = 1 + 2 * d;
    const size_t mom_smem = (size_t)SEL_TY * comps * 8, cov_smem = (size_t)COV_ROWS * S.dp * 8;
    if (mom_smem > 200 * 1024 || cov_smem > 200 * 1024) {
        msg = "embedding dimension too large for the selection kernels";
        return -1;
    }
    if (mom_smem > 48 * 1024)
        S.ok(cudaFuncSetAttribute(k_sel_moments, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mom_smem));
    if (cov_smem > 48 * 1024)
        S.ok(cudaFuncSetAttribute(k_sel_cov, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cov_smem));
    S.alloc(S.x, (size_t)n * d);
    S.alloc(S.w, (size_t)n);
    S.alloc(S.z, (size_t)n);
    S.alloc(S.zs, (size_t)n);
    S.alloc(S.mu, (size_t)d);
    S.alloc(S.vec, (size_t)d);
    S.alloc(S.part, (size_t)2 * S.max_mblk * comps);
    S.alloc(S.fin, (size_t)2 * comps);
    S.alloc(S.covp, (size_t)(S.max_cblk + 1) * S.dp * S.dp);
    S.alloc(S.idx, (size_t)n);
    S.alloc(S.idx2, (size_t)n);
    S.alloc(S.pos, (size_t)n);
    S.alloc(S.perm, (size_t)n);
    S.alloc(S.rank, (size_t)n);
    S.alloc(S.key, (size_t)n);
    S.alloc(S.key2, (size_t)n);
    S.alloc(S.side, (size_t)n);
    S.ok(cudaMallocHost(reinterpret_cast<void **>(&S.h_fin), (size_t)2 * comps * 8));
    S.ok(cudaMallocHost(reinterpret_cast<void **>(&S.h_cov), (size_t)S.dp * S.dp * 8));
    S.ok(cudaMallocHost(reinterpret_cast<void **>(&S.h_z), (size_t)n * 8));
    {
        size_t b1 = 0, b2 = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, b1, S.z, S.zs, S.pos, S.perm, (int)n, 0, 64, st);
        cub::DeviceRadixSort::SortPairs(nullptr, b2, S.key, S.key2, S.idx, S.idx2, (int)n, 0, 8, st);
        S.tmp_bytes = std::max(b1, b2);
        S.ok(cudaMalloc(&S.tmp, std::max<size_t>(S.tmp_bytes, 1)));
    }
    auto cuda_fail = [&]() {
        msg = std::string("landmark selection: ") + cudaGetErrorString(S.err);
        cudaGetLastError();
        return -1;
    };
    if (S.err != cudaSuccess) return cuda_fail();
    S.ok(cudaMemcpyAsync(S.x, x_rowmajor, (size_t)n * d * 8, cudaMemcpyHostToDevice, st));
    S.ok(cudaMemcpyAsync(S.w, vweights, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    // the index array: clusters one after the other
    S.ok(cudaMemcpyAsync(S.idx, cl_members, (size_t)cl_ptr[n_clusters] * 4, cudaMemcpyHostToDevice, st));
    if (S.err != cudaSuccess) return cuda_fail();
    const double eps = 2.220446049250313e-16;  // eps()

    // moments + value of a fresh segment
    auto make_entry = [&](long long off, long long len, Entry &e) -> bool {
        e.off = off;
        e.len = len;
        if (len == 1) {
            e.value = eps;  // "make sure 1-length cluster is in tail of the queue"
            return true;
        }
        if (!S.moments(off, nullptr, Range{0, len}, Range{0, 0}, 1)) return false;
        e.mom.assign(S.h_fin, S.h_fin + comps);
        e.value = -total_of(e.mom.data(), d);
        return true;
    };

    // cut segment e by the rule; returns the length of the low child (the segment is regrouped so that
    // low = [off, off + cut), high = the rest, both in the reference's member order)
    auto cut_segment = [&](const Entry &e, long long &cut) -> int {
        const long long off = e.off, len = e.len;
        ++S.n_cuts;
        if (len == 2) {  // :157-159
            cut = 1;
            return 0;
        }
        // mean, covariance, principal axis
        std::vector<double> mu(d);
        for (int j = 0; j < d; ++j) mu[j] = e.mom[1 + j] / e.mom[0];
        S.ok(cudaMemcpyAsync(S.mu, mu.data(), (size_t)d * 8, cudaMemcpyHostToDevice, st));
        const int cblk = (int)std::max<long long>(1, std::min<long long>(S.max_cblk, (len + 4 * COV_ROWS - 1) / (4 * COV_ROWS)));
        k_sel_cov<<<cblk, COV_THREADS, cov_smem, st>>>(S.x, S.w, S.idx, off, len, S.mu, d, S.dp, S.covp);
        double *cfin = S.covp + (size_t)S.max_cblk * S.dp * S.dp;
        k_sel_sum_blocks<<<(S.dp * S.dp + 127) / 128, 128, 0, st>>>(S.covp, cblk, S.dp * S.dp, 1, cfin);
        S.ok(cudaMemcpyAsync(S.h_cov, cfin, (size_t)S.dp * S.dp * 8, cudaMemcpyDeviceToHost, st));
        if (!S.ok(cudaStreamSynchronize(st))) return -1;
        std::vector<double> C((size_t)d * d), v(d);
        for (int i = 0; i < d; ++i)
            for (int j = i; j < d; ++j)
                C[(size_t)i * d + j] = C[(size_t)j * d + i] = S.h_cov[(size_t)i * S.dp + j];
        if (eig) {  // the host's LAPACK (what the reference's eigvecs calls): its vector, its sign
            if (eig(C.data(), (long long)d, v.data(), eig_user) != 0) {
                msg = "the eigenvector callback failed";
                return -2;
            }
        } else {
            sym_top_eigvec(C.data(), d, v.data(), nullptr);
        }
        S.ok(cudaMemcpyAsync(S.vec, v.data(), (size_t)d * 8, cudaMemcpyHostToDevice, st));
        k_sel_project<<<(unsigned)((len + SEL_TY - 1) / SEL_TY), dim3(SEL_TX, SEL_TY), 0, st>>>(
            S.x, S.w, S.idx, off, len, S.mu, S.vec, d, S.z, S.pos);
        const unsigned lblocks = (unsigned)((len + 255) / 256);
        if (rule == 0) {
            // ---- split_cluster_rss (:155-210) ----
            size_t tb = S.tmp_bytes;
            S.ok(cub::DeviceRadixSort::SortPairs(S.tmp, tb, S.z + off, S.zs + off, S.pos + off, S.perm + off,
                                                 (int)len, 0, 64, st));
            k_sel_invert<<<lblocks, 256, 0, st>>>(S.perm, off, len, S.rank);
            S.ok(cudaMemcpyAsync(S.h_z, S.zs + off, (size_t)len * 8, cudaMemcpyDeviceToHost, st));
            if (!S.ok(cudaStreamSynchronize(st))) return -1;
            const double *zs = S.h_z;
            if (!(zs[0] < zs[len - 1])) {
                msg = "Trying to split homogenous cluster";  // :165-167
                return -2;
            }
            // l1 starts as [argmin], l2 as [argmax]: FIRST occurrences (:163-164).  The sort is stable, so
            // rank 0 is the first minimum; the first maximum is the first rank holding the top value.
            long long rb = len - 1;
            while (rb > 0 && zs[rb - 1] == zs[len - 1]) --rb;
            // gray = every other element; with a tied maximum (rb < len-1) the ranks above rb stay gray
            // and are >= every median, so they behave as part of the upper end of gray: handled by
            // treating gray as the rank set [1, len) \ {rb}.  Ranges below exclude rb explicitly.
            const bool tied_top = rb != len - 1;
            if (tied_top) {  // rare (duplicated rows at the extreme): move rb to the end in rank space
                // rotate ranks rb..len-1 on the host copy and on the device permutation
                std::vector<int> pm((size_t)(len - rb));
                S.ok(cudaMemcpyAsync(pm.data(), S.perm + off + rb, pm.size() * 4, cudaMemcpyDeviceToHost, st));
                if (!S.ok(cudaStreamSynchronize(st))) return -1;
                std::rotate(pm.begin(), pm.begin() + 1, pm.end());
                S.ok(cudaMemcpyAsync(S.perm + off + rb, pm.data(), pm.size() * 4, cudaMemcpyHostToDevice, st));
                k_sel_invert<<<lblocks, 256, 0, st>>>(S.perm, off, len, S.rank);
                // zs values in that range are all equal: nothing to rotate
            }
            if (!S.moments(off, S.perm, Range{0, 1}, Range{len - 1, 1}, 2)) return -1;
            std::vector<double> rss_low(S.h_fin, S.h_fin + comps);
            std::vector<double> rss_high(S.h_fin + comps, S.h_fin + 2 * comps);
            long long lo = 1, hi = len - 1;  // gray = ranks [lo, hi)
            // accepted ranges in acceptance order
            std::vector<Range> acc_low, acc_high;
            auto median = [&](long long a, long long b) {
                const long long c = b - a;
                return (c & 1) ? zs[a + c / 2] : 0.5 * (zs[a + c / 2 - 1] + zs[a + c / 2]);
            };
            if (hi > lo) {
                double med = median(0, len);  // :173 median(z), the whole cluster
                while (true) {
                    const long long p = std::lower_bound(zs + lo, zs + hi, med) - zs;  // first z >= med
                    if (!S.moments(off, S.perm, Range{lo, p - lo}, Range{p, hi - p}, 2)) return -1;
                    std::vector<double> low_tmp = rss_low, high_tmp = rss_high;
                    add_to(low_tmp, S.h_fin);
                    add_to(high_tmp, S.h_fin + comps);
                    if (total_of(low_tmp.data(), d) < total_of(high_tmp.data(), d)) {
                        if (p == lo) break;
                        rss_low.swap(low_tmp);
                        acc_low.push_back(Range{lo, p - lo});
                        lo = p;
                    } else {
                        if (p == hi) break;
                        rss_high.swap(high_tmp);
                        acc_high.push_back(Range{p, hi - p});
                        hi = p;
                    }
                    if (lo == hi) break;
                    med = median(lo, hi);
                }
            }
            bool gray_low = false;
            if (hi > lo) {  // :199-207
                if (!S.moments(off, S.perm, Range{lo, hi - lo}, Range{0, 0}, 1)) return -1;
                std::vector<double> low_tmp = rss_low, high_tmp = rss_high;
                add_to(low_tmp, S.h_fin);
                add_to(high_tmp, S.h_fin);
                gray_low = std::max(total_of(low_tmp.data(), d), total_of(rss_high.data(), d)) <
                           std::max(total_of(rss_low.data(), d), total_of(high_tmp.data(), d));
            }
            // new order: l1 = [a], accepted low ranges in order, (gray); l2 = [b], accepted high ranges in
            // order, (gray); members of a range in their previous order
            std::vector<std::pair<Range, int>> lay;  // (rank range, layout place)
            int place = 0;
            lay.push_back({Range{0, 1}, place++});
            for (const Range &r : acc_low) lay.push_back({r, place++});
            if (hi > lo && gray_low) lay.push_back({Range{lo, hi - lo}, place++});
            lay.push_back({Range{len - 1, 1}, place++});
            for (const Range &r : acc_high) lay.push_back({r, place++});
            if (hi > lo && !gray_low) lay.push_back({Range{lo, hi - lo}, place++});
            cut = gray_low ? hi : lo;
            std::sort(lay.begin(), lay.end(),
                      [](const std::pair<Range, int> &a, const std::pair<Range, int> &b) { return a.first.off < b.first.off; });
            if ((int)lay.size() > MAX_GROUPS) {
                msg = "landmark selection: too many groups in one cut";
                return -1;
            }
            Groups g;
            g.n = (int)lay.size();
            for (int i = 0; i < g.n; ++i) {
                g.end[i] = (int)(lay[i].first.off + lay[i].first.len);
                g.layout[i] = lay[i].second;
            }
            k_sel_group_keys<<<lblocks, 256, 0, st>>>(S.rank, off, len, g, S.key);
        } else {
            // ---- split_cluster_size (:218-238) / split_cluster_diameter (:247-267): threshold on z, ties
            // alternate by the running lengths in member order ----
            S.ok(cudaMemcpyAsync(S.h_z, S.z + off, (size_t)len * 8, cudaMemcpyDeviceToHost, st));
            if (!S.ok(cudaStreamSynchronize(st))) return -1;
            const double *z = S.h_z;
            double thr;
            if (rule == 2) {
                std::vector<double> t(z, z + len);
                const size_t h = (size_t)len / 2;
                std::nth_element(t.begin(), t.begin() + h, t.end());
                thr = t[h];
                if (len % 2 == 0) thr = 0.5 * (*std::max_element(t.begin(), t.begin() + h) + thr);
            } else {
                const auto mm = std::minmax_element(z, z + len);
                thr = (*mm.first + *mm.second) / 2.0;
            }
            std::vector<unsigned char> side((size_t)len);
            long long nl = 0, nh = 0;
            for (long long i = 0; i < len; ++i) {
                bool low;
                if (z[i] == thr) low = nl < nh;
                else low = z[i] < thr;
                side[(size_t)i] = low ? 0 : 1;
                (low ? nl : nh)++;
            }
            if (nl == 0 || nh == 0) {
                msg = "Unexpected empty cluster generated";  // :309, :317
                return -2;
            }
            cut = nl;
            S.ok(cudaMemcpyAsync(S.side, side.data(), (size_t)len, cudaMemcpyHostToDevice, st));
            k_sel_side_keys<<<lblocks, 256, 0, st>>>(S.side, off, len, S.key);
        }
        // stable regrouping of the segment by the keys
        size_t tb = S.tmp_bytes;
        S.ok(cub::DeviceRadixSort::SortPairs(S.tmp, tb, S.key + off, S.key2 + off, S.idx + off, S.idx2 + off,
                                             (int)len, 0, 8, st));
        S.ok(cudaMemcpyAsync(S.idx + off, S.idx2 + off, (size_t)len * 4, cudaMemcpyDeviceToDevice, st));
        if (!S.ok(cudaGetLastError())) return -1;
        return 0;
    };

    auto split_and_put = [&](std::vector<Entry> &pq, const Entry &e) -> int {
        if (e.len < 2) {
            msg = "AssertionError: size(m, 1) > 1";  // :156
            return -2;
        }
        long long cut = 0;
        if (int rc = cut_segment(e, cut)) return rc;
        if (cut <= 0 || cut >= e.len) {
            msg = "Unexpected empty cluster generated";
            return -2;
        }
        // both children's moments (-> their queue values and, later, their means) in one launch
        Entry lowc, highc;
        lowc.off = e.off; lowc.len = cut;
        highc.off = e.off + cut; highc.len = e.len - cut;
        lowc.value = highc.value = eps;  // singletons go to the tail of the queue (:311, :319)
        if (lowc.len > 1 || highc.len > 1) {
            if (!S.moments(e.off, nullptr, Range{0, lowc.len}, Range{cut, highc.len}, 2)) return -1;
            if (lowc.len > 1) {
                lowc.mom.assign(S.h_fin, S.h_fin + comps);
                lowc.value = -total_of(lowc.mom.data(), d);
            }
            if (highc.len > 1) {
                highc.mom.assign(S.h_fin + comps, S.h_fin + 2 * comps);
                highc.value = -total_of(highc.mom.data(), d);
            }
        }
        heap_put(pq, std::move(lowc));
        heap_put(pq, std::move(highc));
        return 0;
    };

    std::vector<Entry> pq;
    for (long long c = 0; c < n_clusters; ++c) {  // :281-303
        const long long off = cl_ptr[c], len = cl_ptr[c + 1] - cl_ptr[c];
        if (len <= forced) {
            for (long long j = 0; j < len; ++j) {
                Entry e;
                e.off = off + j;
                e.len = 1;
                e.value = eps;
                heap_put(pq, std::move(e));
            }
        } else {
            std::vector<Entry> local;
            Entry e;
            if (!make_entry(off, len, e)) return cuda_fail();
            heap_put(local, std::move(e));
            while ((long long)local.size() < forced) {
                Entry top = heap_pop(local);
                if (int rc = split_and_put(local, top)) return rc == -1 && S.err != cudaSuccess ? cuda_fail() : rc;
            }
            while (!local.empty()) {
                Entry top = heap_pop(local);
                heap_put(pq, std::move(top));
            }
        }
    }
    while ((long long)pq.size() < land) {  // :305-323
        if (pq.empty()) {
            msg = "no cluster to split";
            return -2;
        }
        Entry top = heap_pop(pq);
        if (int rc = split_and_put(pq, top)) return rc == -1 && S.err != cudaSuccess ? cuda_fail() : rc;
    }
    // group ids = position in the queue's array (:325-331)
    std::vector<int> h_idx((size_t)n);
    S.ok(cudaMemcpyAsync(h_idx.data(), S.idx, (size_t)cl_ptr[n_clusters] * 4, cudaMemcpyDeviceToHost, st));
    if (!S.ok(cudaStreamSynchronize(st))) return cuda_fail();
    for (long long i = 0; i < n; ++i) out_group[i] = -1;
    for (size_t g = 0; g < pq.size(); ++g)
        for (long long r = 0; r < pq[g].len; ++r) out_group[h_idx[(size_t)(pq[g].off + r)]] = (long long)g;
    if (out_cuts) *out_cuts = S.n_cuts;
    return 0;
}


// Number of distinct rows of the row-major n x d matrix (host pointer).  Rows are hashed (64 bits), sorted by
// hash, and neighbours with equal hashes are compared in full, so the count is exact unless two DIFFERENT
// rows collide in 64 bits and interleave with duplicates of themselves.  Returns 0 or -1 (CUDA failure).
int count_unique_rows_device(cudaStream_t st, long long n, int d, const double *x_rowmajor,
                             long long *out_count, std::string &msg) {
    double *x = nullptr;
    unsigned long long *h = nullptr, *h2 = nullptr, *cnt = nullptr;
    int *id = nullptr, *id2 = nullptr;
    void *tmp = nullptr;
    cudaError_t err = cudaSuccess;
    auto ok = [&](cudaError_t e) {
        if (e != cudaSuccess && err == cudaSuccess) err = e;
        return err == cudaSuccess;
    };
    size_t tb = 0;
    ok(cudaMalloc(&x, (size_t)n * d * 8));
    ok(cudaMalloc(&h, (size_t)n * 8));
    ok(cudaMalloc(&h2, (size_t)n * 8));
    ok(cudaMalloc(&id, (size_t)n * 4));
    ok(cudaMalloc(&id2, (size_t)n * 4));
    ok(cudaMalloc(&cnt, 8));
    cub::DeviceRadixSort::SortPairs(nullptr, tb, h, h2, id, id2, (int)n, 0, 64, st);
    ok(cudaMalloc(&tmp, std::max<size_t>(tb, 1)));
    unsigned long long c = 0;
    if (err == cudaSuccess) {
        ok(cudaMemcpyAsync(x, x_rowmajor, (size_t)n * d * 8, cudaMemcpyHostToDevice, st));
        ok(cudaMemsetAsync(cnt, 0, 8, st));
        const unsigned blocks = (unsigned)((n + SEL_TY - 1) / SEL_TY);
        k_row_hash<<<blocks, dim3(SEL_TX, SEL_TY), 0, st>>>(x, n, d, h, id);
        ok(cub::DeviceRadixSort::SortPairs(tmp, tb, h, h2, id, id2, (int)n, 0, 64, st));
        k_row_distinct<<<blocks, dim3(SEL_TX, SEL_TY), 0, st>>>(x, n, d, h2, id2, cnt);
        ok(cudaMemcpyAsync(&c, cnt, 8, cudaMemcpyDeviceToHost, st));
        ok(cudaStreamSynchronize(st));
        ok(cudaGetLastError());
    }
    for (void *p : {(void *)x, (void *)h, (void *)h2, (void *)id, (void *)id2, (void *)cnt, tmp})
        if (p) cudaFree(p);
    if (err != cudaSuccess) {
        msg = std::string("unique rows: ") + cudaGetErrorString(err);
        cudaGetLastError();
        return -1;
    }
    *out_count = (long long)c;
    return 0;
}

}  // namespace cge
