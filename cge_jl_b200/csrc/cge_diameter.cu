// cge_diameter.cu -- exact diameter (max pairwise distance) of the full graph for the landmark-mode
// local score: hi of divergence.jl:113 is max_ij dist(i,j) over ALL n_full^2/2 pairs, the only
// thing (besides 2K sampled distances) that is needed of full_graph_D.
//
// The FP64 difference-form kernel (k_build_dist<false>) costs 2d FP64 instructions per pair:
// 14 s for 1M vertices, d = 128.  This file gets the SAME value (bit for bit) in two steps:
//   1. filter on the 5th-generation tensor cores: d2_ij ~ n_i + n_j - 2 x_i.x_j, the Gram tile
//      x_i.x_j by tcgen05.mma (kind::f16, BF16 operands split hi + lo: hi.hi + hi.lo + lo.hi,
//      FP32 accumulator in TMEM), operands brought in by cp.async.bulk (TMA) from a pre-packed
//      copy of the embedding that already has the UMMA K-major core-matrix layout; the epilogue
//      reads the accumulator with tcgen05.ld and keeps one number per 128x128 tile: its largest
//      approximate squared distance;
//   2. every tile whose approximate maximum is within 2E of the global approximate maximum -- E a
//      rigorous bound on the filter's error -- is recomputed in FP64 by k_build_dist<false>; the
//      true argmax tile is always among them, so the result is the exact FP64 maximum.
// Error bound: |x - hi - lo| <= 2^-18 |x| per coordinate, hence the three-term Gram entry is within
// (3*2^-18 + accumulation) |x_i||x_j| of the exact dot product; norms are exact FP64 rounded to
// FP32.  E = rel * max_i n_i with rel = 1e-3 leaves a > 10x margin (cge_diameter_rel overrides).
#include "cge_kernels.cuh"
#include "cge_ring.cuh"  // mbarrier + bulk-copy helpers

#include <cuda_bf16.h>

namespace cge {

constexpr int DM_THREADS = 160;  // 4 epilogue warps + the producer warp
constexpr float DM_PAD_NORM = -1e30f;

// ---- packing: FP64 rows -> BF16 hi/lo in the canonical K-major no-swizzle UMMA layout ----------
// core matrix = 8 rows x 16 bytes (8 BF16 along K), stored as 128 contiguous bytes; within one
// operand part: offset(kc, rg, r) = ((kc*16 + rg)*8 + r)*16 with kc = k/8, rg = row/8, r = row%8.
// The rows are centred first (x - mean; distances are translation invariant): the filter's error
// bound scales with the largest squared norm, which must not be inflated by a common offset.
__global__ void k_pack_bf16(const double *__restrict__ emb, const double *__restrict__ mean, int dp,
                            int n, int d_true, long long n_rows, unsigned char *__restrict__ packed,
                            float *__restrict__ norms, unsigned *rmax_bits) {
    const int ksteps = dp / 16, kchunks = dp / 8;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long row = gid / kchunks;
    const int kc = (int)(gid % kchunks);
    if (row >= n_rows) return;  // n_rows = nb*128 (rows >= n are zero pads)
    const int b = (int)(row / TILE), r = (int)(row % TILE);
    const size_t opb = (size_t)ksteps * 4096;
    unsigned char *blk = packed + (size_t)b * 2 * opb;
    const size_t off = ((size_t)(kc * 16 + r / 8) * 8 + (r % 8)) * 16;
    __nv_bfloat16 hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int c = kc * 8 + e;
        const double x = (row < n && c < d_true) ? emb[(size_t)row * dp + c] - mean[c] : 0.0;
        hi[e] = __float2bfloat16_rn((float)x);
        lo[e] = __float2bfloat16_rn((float)(x - (double)__bfloat162float(hi[e])));
    }
    *reinterpret_cast<uint4 *>(blk + off) = *reinterpret_cast<const uint4 *>(hi);
    *reinterpret_cast<uint4 *>(blk + opb + off) = *reinterpret_cast<const uint4 *>(lo);
    if (kc == 0) {
        float nf = DM_PAD_NORM;
        if (row < n) {
            double s = 0.0;
            for (int c = 0; c < d_true; ++c) {
                const double x = emb[(size_t)row * dp + c] - mean[c];
                s = fma(x, x, s);
            }
            nf = (float)s;
            atomicMax(rmax_bits, __float_as_uint(nf));
        }
        norms[row] = nf;
    }
}

// ---- tcgen05 helpers (forms as in the CUTLASS sm100 headers) -----------------------------------
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes,
                                              uint32_t sbo_bytes) {
    // start address [0,14) >>4 | LBO [16,30) >>4 | SBO [32,46) >>4 | version 1 at [46,48) |
    // layout type SWIZZLE_NONE (0) at [61,64)
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
          "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// One CTA per SM, warp-specialised: warps 0-3 (128 threads = the 128 TMEM lanes) are the epilogue,
// lane 0 of warp 4 is the producer -- it claims strips of tiles of one tile row from a counter, brings
// the operand blocks in by TMA and issues the MMAs.  TWO accumulators in TMEM (2 x 128 columns): the 24
// MMAs of tile i+1 run on the tensor pipe while the epilogue warps read tile i's accumulator back
// (tcgen05.ld, one row per thread) and take its maximum.  Hand-offs are mbarriers:
//   b_full[2]   TMA -> producer      column block of tile i has landed in buffer i & 1
//   mma_done[2] tcgen05.commit -> epilogue (accumulator i & 1 complete) and -> producer (the column
//               buffer i & 1 may be refilled)
//   acc_free[2] epilogue (one arrive per warp) -> producer: accumulator i & 1 has been read
// Strips are separated by CTA-wide barriers (the row block changes); inside a strip nothing but the
// four epilogue warps' own named barrier (the tile maximum) synchronises more than it must.
constexpr uint32_t DM_TMEM_COLS = 256;
constexpr int DM_EPI_THREADS = 128;
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__global__ void __launch_bounds__(DM_THREADS, 1) k_diameter_filter(const __grid_constant__ DiamArgs a) {
    extern __shared__ __align__(1024) unsigned char sm[];
    const uint32_t opb = (uint32_t)a.ksteps * 4096u;
    unsigned char *sA = sm;                       // hi, lo
    unsigned char *sB = sm + 2 * opb;             // two buffers of (hi, lo)
    float *nA = reinterpret_cast<float *>(sm + 6 * (size_t)opb);
    float *nB = nA + TILE;                        // [2][128]
    uint64_t *bars = reinterpret_cast<uint64_t *>(nB + 2 * TILE);  // a_full, b_full[2], mma_done[2], acc_free[2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 7);
    int *s_strip = reinterpret_cast<int *>(tmem_slot + 1);         // bi, bj0, count, first tile
    float *s_red = reinterpret_cast<float *>(s_strip + 4);         // [4]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint64_t *a_full = bars, *b_full = bars + 1, *mma_done = bars + 3, *acc_free = bars + 5;
    const bool producer = tid == DM_EPI_THREADS;  // lane 0 of warp 4
    if (tid == 0) {
        for (int i = 0; i < 5; ++i) mbar_init(bars + i, 1);
        for (int i = 5; i < 7; ++i) mbar_init(bars + i, 4);  // one arrive per epilogue warp
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(tmem_slot)),
                     "r"(DM_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    // instruction descriptor: D = F32 (1<<4), A = B = BF16 (1<<7, 1<<10), K-major both, N = 128
    // (N>>3 at [17,23)), M = 128 (M>>4 at [24,29))
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t pol = l2_policy(true);  // blocks are re-read by many CTAs: keep them in L2
    // phase bits, one per buffer.  Producer: ph_a, ph_b (b_full), ph_m (mma_done), ph_f (acc_free);
    // epilogue threads: ph_m only.
    uint32_t ph_a = 0, ph_b = 0, ph_m = 0, ph_f = 0;
    float lmax = 0.0f;
    const size_t blk_bytes = 2 * (size_t)opb;
    while (true) {
        if (producer) {
            const unsigned s = atomicAdd(a.strip_counter, 1u);
            int4 st = make_int4(-1, 0, 0, 0);
            if (s < (unsigned)a.n_strips) st = a.strips[s];
            s_strip[0] = st.x; s_strip[1] = st.y; s_strip[2] = st.z; s_strip[3] = st.w;
        }
        __syncthreads();
        const int bi = s_strip[0], bj0 = s_strip[1], cnt = s_strip[2], t0 = s_strip[3];
        if (bi < 0) break;
        if (producer) {
            auto request_b = [&](int i) {  // column block of tile i into buffer i & 1
                const int buf = i & 1;
                mbar_expect_tx(b_full + buf, 2 * opb);
                bulk_g2s(sB + (size_t)buf * blk_bytes, a.packed + (size_t)(bj0 + i) * blk_bytes, 2 * opb,
                         b_full + buf, pol);
            };
            // the previous strip is fully drained (barrier at its end): every buffer is free
            mbar_expect_tx(a_full, 2 * opb);
            bulk_g2s(sA, a.packed + (size_t)bi * blk_bytes, 2 * opb, a_full, pol);
            request_b(0);
            if (cnt > 1) request_b(1);
            mbar_wait(a_full, ph_a);
            ph_a ^= 1u;
            const uint32_t a_hi = smem_u32(sA), a_lo = a_hi + opb;
            for (int i = 0; i < cnt; ++i) {
                const int buf = i & 1;
                if (i >= 2) {  // accumulator i & 1 was last used by tile i-2: wait until it has been read
                    mbar_wait(acc_free + buf, (ph_f >> buf) & 1u);
                    ph_f ^= 1u << buf;
                }
                mbar_wait(b_full + buf, (ph_b >> buf) & 1u);
                ph_b ^= 1u << buf;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t b_hi = smem_u32(sB + (size_t)buf * blk_bytes), b_lo = b_hi + opb;
                const uint32_t acc_addr = tmem + (uint32_t)buf * 128u;
                uint32_t acc = 0;
                // Gram tile = hi.hi + hi.lo + lo.hi ; K step s covers k-chunks 2s, 2s+1
                for (int term = 0; term < 3; ++term) {
                    const uint32_t pa = term == 2 ? a_lo : a_hi, pb = term == 1 ? b_lo : b_hi;
                    for (int s = 0; s < a.ksteps; ++s) {
                        umma_bf16(acc_addr, umma_desc(pa + (uint32_t)s * 4096u, 2048u, 128u),
                                  umma_desc(pb + (uint32_t)s * 4096u, 2048u, 128u), idesc, acc);
                        acc = 1;
                    }
                }
                umma_commit(mma_done + buf);
                // tile i-1's MMAs (the last readers of the other column buffer) complete while tile i's
                // run: refill that buffer with the block of tile i+1
                if (i >= 1) {
                    mbar_wait(mma_done + (buf ^ 1), (ph_m >> (buf ^ 1)) & 1u);
                    ph_m ^= 1u << (buf ^ 1);
                    if (i + 1 < cnt) request_b(i + 1);
                }
            }
            // drain: the last tile's completion (keeps the producer's mma_done phases in step) and the
            // acc_free arrivals of the last two tiles
            {
                const int buf = (cnt - 1) & 1;
                mbar_wait(mma_done + buf, (ph_m >> buf) & 1u);
                ph_m ^= 1u << buf;
            }
            for (int i = cnt < 2 ? 0 : cnt - 2; i < cnt; ++i) {
                const int buf = i & 1;
                mbar_wait(acc_free + buf, (ph_f >> buf) & 1u);
                ph_f ^= 1u << buf;
            }
        } else if (tid < DM_EPI_THREADS) {
            nA[tid] = a.norms[(size_t)bi * TILE + tid];
            nB[tid] = a.norms[(size_t)bj0 * TILE + tid];
            epi_bar();
            for (int i = 0; i < cnt; ++i) {
                const int buf = i & 1;
                if (i + 1 < cnt) nB[(buf ^ 1) * TILE + tid] = a.norms[(size_t)(bj0 + i + 1) * TILE + tid];
                mbar_wait(mma_done + buf, (ph_m >> buf) & 1u);
                ph_m ^= 1u << buf;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                // thread tid owns accumulator row tid (TMEM lane), 128 FP32 columns
                const float *nb_ = nB + buf * TILE;
                float best = -3.0e38f;
#pragma unroll
                for (int c0 = 0; c0 < TILE; c0 += 32) {
                    float v[32];
                    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * 128 + c0), v);
#pragma unroll
                    for (int j = 0; j < 32; ++j) best = fmaxf(best, fmaf(-2.0f, v[j], nb_[c0 + j]));
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_free + buf);  // this warp's 32 rows have been read
                best += nA[tid];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) best = fmaxf(best, __shfl_xor_sync(FULL, best, off));
                if (lane == 0) s_red[warp] = best;
                epi_bar();  // s_red complete; also orders the nB write above before the next tile's reads
                if (tid == 0) {
                    const float m = fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3]));
                    a.tile_max[(size_t)t0 + i] = m;
                    lmax = fmaxf(lmax, m);
                }
                epi_bar();  // s_red may be overwritten
            }
        }
        __syncthreads();  // strip drained: operand buffers, accumulators and s_strip are free
    }
    if (tid == 0) atomicMax(a.gmax_bits, __float_as_uint(fmaxf(lmax, 0.0f)));
    __syncthreads();
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(DM_TMEM_COLS)
                     : "memory");
    }
}

// tiles whose approximate maximum could hide the true maximum
__global__ void k_select_candidates(const float *__restrict__ tile_max, long long n_tiles,
                                    const unsigned *__restrict__ gmax_bits,
                                    const unsigned *__restrict__ rmax_bits, float rel, int *list,
                                    int cap, int *count) {
    const float gmax = __uint_as_float(*gmax_bits);
    const float thr = gmax - 2.0f * rel * __uint_as_float(*rmax_bits);
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_tiles;
         t += (long long)gridDim.x * blockDim.x) {
        if (tile_max[t] >= thr) {
            const int slot = atomicAdd(count, 1);
            if (slot < cap) list[slot] = (int)t;
        }
    }
}

size_t diameter_smem_bytes(int ksteps) { return 6 * (size_t)ksteps * 4096 + 3 * TILE * 4 + 128; }  // 7 barriers + 9 words

void launch_pack_bf16(const double *emb, const double *mean, int dp, int n, int d_true, int nb,
                      unsigned char *packed, float *norms, unsigned *rmax_bits, cudaStream_t st) {
    const long long threads = (long long)nb * TILE * (dp / 8);
    k_pack_bf16<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(
        emb, mean, dp, n, d_true, (long long)nb * TILE, packed, norms, rmax_bits);
}

cudaError_t launch_diameter_filter(const DiamArgs &a, int grid, cudaStream_t st) {
    const size_t smem = diameter_smem_bytes(a.ksteps);
    cudaError_t e = cudaFuncSetAttribute(k_diameter_filter,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_diameter_filter<<<grid, DM_THREADS, smem, st>>>(a);
    return cudaGetLastError();
}

void launch_select_candidates(const float *tile_max, long long n_tiles, const unsigned *gmax_bits,
                              const unsigned *rmax_bits, float rel, int *list, int cap, int *count,
                              cudaStream_t st) {
    k_select_candidates<<<1184, 256, 0, st>>>(tile_max, n_tiles, gmax_bits, rmax_bits, rel, list, cap,
                                               count);
}

}  // namespace cge
