// cge_ring.cuh -- the fixed point of one alpha with the matrix streamed through a shared-memory
// ring by the TMA engine (cp.async.bulk + mbarrier, SASS: UBLKCP / SYNCS).
//
// Why: with plain coalesced loads (k_fixed_point) every warp drains its loads at each tile
// boundary (reductions + barrier) and the SM falls below the ~42 KB in flight that 6.5 TB/s over
// 148 SMs needs.  Here one elected producer thread per CTA keeps NST x 32 KB of bulk copies in
// flight independently of what the eight consumer warps are doing, so HBM stays busy across tile
// boundaries.  One CTA per SM (cooperative launch), 8 consumer warps + 1 producer warp.
//
// Stage = 32 rows x 128 columns of a tile (32 KB, contiguous in HBM).  Consumer warp w owns rows
// 4w..4w+3 of every stage, i.e. tile rows q*32 + 4w + r -- again 16 rows per warp and tile, lane l
// owns columns 2l, 2l+1, 64+2l, 65+2l, so the reductions are those of tile_pass_u/_d.
#pragma once
#include "cge_kernels.cuh"

namespace cge {

constexpr int RING_STAGE_ROWS = 32;
constexpr int RING_STAGE_ELEMS = RING_STAGE_ROWS * TILE;
constexpr int RING_STAGE_BYTES = RING_STAGE_ELEMS * 8;
constexpr int RING_QUARTERS = TILE / RING_STAGE_ROWS;
constexpr int RING_THREADS = NTHREADS + 32;

template <bool DIRECTED>
__host__ __device__ constexpr int ring_stages() {
    return DIRECTED ? 5 : 6;
}
template <bool DIRECTED>
__host__ __device__ constexpr size_t ring_smem_bytes() {
    return (size_t)ring_stages<DIRECTED>() * RING_STAGE_BYTES            // ring
           + (size_t)2 * (DIRECTED ? 2 : 1) * NWARPS * TILE * 8          // s_col, double buffered
           + (size_t)(DIRECTED ? 2 : 1) * NWARPS * 32 * 8                // s_red
           + 16 * 8 + 64;                                                // barriers, flags
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug must end in a trapped kernel, never in a hung GPU
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity))
        if (++spins > (1u << 26)) __trap();
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes,
                                         uint64_t *bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// grid-wide barrier of the consumer threads (the producer warp never takes part); the launch is
// cooperative, so all CTAs are resident.  counter is zero at kernel start and only grows.
__device__ __forceinline__ void grid_barrier_consumers(unsigned *counter, unsigned nblocks,
                                                       unsigned &epoch) {
    consumer_sync();
    if (threadIdx.x == 0) {
        __threadfence();
        ++epoch;
        atomicAdd(counter, 1u);
        const unsigned target = epoch * nblocks;
        unsigned spins = 0;
        while (ld_acquire_u32(counter) < target)
            if (++spins > (1u << 28)) __trap();
        __threadfence();
    }
    consumer_sync();
}

struct RingState {
    int stage;
    uint32_t phase;
};
template <int NST>
__device__ __forceinline__ void ring_advance(RingState &r) {
    if (++r.stage == NST) {
        r.stage = 0;
        r.phase ^= 1u;
    }
}

// tile row handled by accumulator idx (0..15) of consumer warp w
__device__ __forceinline__ int ring_row(int idx, int w) { return (idx >> 2) * 32 + 4 * w + (idx & 3); }

template <int M, bool DIRECTED>
__device__ __forceinline__ void ring_tile(int bi, int bj, const SweepArgs &a, double *s_col,
                                          const double *ring, uint64_t *full, uint64_t *empty,
                                          RingState &rs) {
    constexpr int NST = ring_stages<DIRECTED>();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const size_t cb = (size_t)bj * TILE, rb = (size_t)bi * TILE;
    // column factors: undirected T_c (ta); directed Tout_c for Sin rows (tb), Tin_c for Sout rows (ta)
    const double2 ta01 = __ldcg(reinterpret_cast<const double2 *>(a.Ta + cb) + lane);
    const double2 ta23 = __ldcg(reinterpret_cast<const double2 *>(a.Ta + cb + 64) + lane);
    double2 tb01 = make_double2(0.0, 0.0), tb23 = tb01;
    if (DIRECTED) {
        tb01 = __ldcg(reinterpret_cast<const double2 *>(a.Tb + cb) + lane);
        tb23 = __ldcg(reinterpret_cast<const double2 *>(a.Tb + cb + 64) + lane);
    }
    const bool ld = lane < 16;
    const double trow_a = ld ? __ldcg(a.Ta + rb + ring_row(lane, w)) : 0.0;
    const double trow_b = (DIRECTED && ld) ? __ldcg(a.Tb + rb + ring_row(lane, w)) : 0.0;
    double ra[16], rb_[DIRECTED ? 16 : 1];
    double ca0 = 0.0, ca1 = 0.0, ca2 = 0.0, ca3 = 0.0;  // undirected / directed-Sin column sums
    double cb0 = 0.0, cb1 = 0.0, cb2 = 0.0, cb3 = 0.0;  // directed-Sout column sums
#pragma unroll
    for (int qtr = 0; qtr < RING_QUARTERS; ++qtr) {
        mbar_wait(full + rs.stage, rs.phase);
        const double *sp = ring + (size_t)rs.stage * RING_STAGE_ELEMS + (size_t)(4 * w) * TILE;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int idx = qtr * 4 + r;
            const double2 v01 = *reinterpret_cast<const double2 *>(sp + r * TILE + 2 * lane);
            const double2 v23 = *reinterpret_cast<const double2 *>(sp + r * TILE + 64 + 2 * lane);
            const double g0 = powm_any<M>(v01.x, a.m), g1 = powm_any<M>(v01.y, a.m);
            const double g2 = powm_any<M>(v23.x, a.m), g3 = powm_any<M>(v23.y, a.m);
            if (!DIRECTED) {
                const double ti = __shfl_sync(FULL, trow_a, idx);
                ra[idx] = fma(g3, ta23.y, fma(g2, ta23.x, fma(g1, ta01.y, g0 * ta01.x)));
                ca0 = fma(ti, g0, ca0);
                ca1 = fma(ti, g1, ca1);
                ca2 = fma(ti, g2, ca2);
                ca3 = fma(ti, g3, ca3);
            } else {
                const double t_in = __shfl_sync(FULL, trow_a, idx);
                const double t_out = __shfl_sync(FULL, trow_b, idx);
                ra[idx] = fma(g3, tb23.y, fma(g2, tb23.x, fma(g1, tb01.y, g0 * tb01.x)));   // Sin row
                rb_[DIRECTED ? idx : 0] =
                    fma(g3, ta23.y, fma(g2, ta23.x, fma(g1, ta01.y, g0 * ta01.x)));           // Sout row
                ca0 = fma(t_out, g0, ca0);
                ca1 = fma(t_out, g1, ca1);
                ca2 = fma(t_out, g2, ca2);
                ca3 = fma(t_out, g3, ca3);
                cb0 = fma(t_in, g0, cb0);
                cb1 = fma(t_in, g1, cb1);
                cb2 = fma(t_in, g2, cb2);
                cb3 = fma(t_in, g3, cb3);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + rs.stage);  // this warp is done with the stage
        ring_advance<NST>(rs);
    }
    warp_treduce<16>(ra, lane);
    const size_t orow = (size_t)bj * a.np + rb + ring_row(treduce_index<16>(lane), w);
    if ((lane & 1) == 0) a.partA[orow] = ra[0];
    if constexpr (DIRECTED) {
        warp_treduce<16>(rb_, lane);
        if ((lane & 1) == 0) a.partB[orow] = rb_[0];
    }
    const bool offdiag = bi != bj;
    if (offdiag) {
        double2 *sa = reinterpret_cast<double2 *>(s_col + w * TILE);
        sa[lane] = make_double2(ca0, ca1);
        sa[32 + lane] = make_double2(ca2, ca3);
        if (DIRECTED) {
            double2 *sb = reinterpret_cast<double2 *>(s_col + NWARPS * TILE + w * TILE);
            sb[lane] = make_double2(cb0, cb1);
            sb[32 + lane] = make_double2(cb2, cb3);
        }
    }
    consumer_sync();
    if (offdiag && (DIRECTED || threadIdx.x < TILE)) {
        const int c = threadIdx.x & (TILE - 1), which = threadIdx.x >> 7;
        const double *src = s_col + which * NWARPS * TILE;
        double s = 0.0;
#pragma unroll
        for (int w2 = 0; w2 < NWARPS; ++w2) s += src[w2 * TILE + c];
        (which ? a.partB : a.partA)[(size_t)bi * a.np + cb + c] = s;
    }
}

// All passes of one alpha (divergence.jl:150-168 / 434-467); same arithmetic and the same
// reduction orders as k_fixed_point, only the data path differs.
template <int M, bool DIRECTED>
__global__ void __launch_bounds__(RING_THREADS, 1)
k_fixed_point_ring(const __grid_constant__ SweepArgs a) {
    constexpr int NST = ring_stages<DIRECTED>();
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    double *ring = reinterpret_cast<double *>(smem_raw);
    double *s_col = ring + (size_t)NST * RING_STAGE_ELEMS;                  // [2][(1|2)*8*128]
    double *s_red = s_col + 2 * (DIRECTED ? 2 : 1) * NWARPS * TILE;        // [(1|2)*8*32]
    uint64_t *full = reinterpret_cast<uint64_t *>(s_red + (DIRECTED ? 2 : 1) * NWARPS * 32);
    uint64_t *empty = full + 8;
    volatile int *s_go = reinterpret_cast<volatile int *>(empty + 8);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(full + s, 1);        // the producer's arrive.expect_tx
            mbar_init(empty + s, NWARPS);  // one arrive per consumer warp
        }
        *s_go = 1;  // passes authorised so far
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (w == NWARPS) {
        // ===== producer: one thread streams this CTA's tiles, pass after pass =====
        if (lane == 0) {
            RingState rs{0, 0};
            int pass = 0;
            while (true) {
                for (long long t = a.tile_begin + blockIdx.x; t < a.tile_end; t += gridDim.x) {
                    const uint64_t pol = l2_policy(t - a.tile_begin < a.resident_tiles);
                    const double *src = a.q + (size_t)(t - a.tile_begin) * TILE_ELEMS;
#pragma unroll 1
                    for (int qtr = 0; qtr < RING_QUARTERS; ++qtr) {
                        mbar_wait(empty + rs.stage, rs.phase ^ 1u);
                        mbar_expect_tx(full + rs.stage, RING_STAGE_BYTES);
                        bulk_g2s(ring + (size_t)rs.stage * RING_STAGE_ELEMS,
                                 src + (size_t)qtr * RING_STAGE_ELEMS, RING_STAGE_BYTES,
                                 full + rs.stage, pol);
                        ring_advance<NST>(rs);
                    }
                }
                ++pass;
                int go;
                unsigned spins = 0;
                while ((go = *s_go) == pass)  // wait for the consumers' verdict on this pass
                    if (++spins > (1u << 28)) __trap();
                if (go < 0) break;
            }
        }
        return;
    }

    // ===== consumers =====
    RingState rs{0, 0};
    const int ngroups = (a.n + 31) / 32;
    unsigned *gbar = reinterpret_cast<unsigned *>(a.slots + 4);
    unsigned epoch = 0;
    double diff = 1.0, eps = a.eps0;
    int it = 0, tile_it = 0;
    while (true) {
        for (long long t = a.tile_begin + blockIdx.x; t < a.tile_end; t += gridDim.x, ++tile_it) {
            const int2 ij = a.tile_ij[t];
            ring_tile<M, DIRECTED>(ij.x, ij.y, a,
                                   s_col + (tile_it & 1) * (DIRECTED ? 2 : 1) * NWARPS * TILE, ring,
                                   full, empty, rs);
        }
        grid_barrier_consumers(gbar, gridDim.x, epoch);
        double e = 0.0;
        for (int g = blockIdx.x; g < ngroups; g += gridDim.x) {
            const int v = g * 32 + lane;
            double pa = 0.0, pb = 0.0;
            if (v < a.n) {
                int b_lo, b_hi;
                part_range(a, v, b_lo, b_hi);
                for (int b = b_lo + w; b < b_hi; b += NWARPS) {
                    pa += __ldcg(a.partA + (size_t)b * a.np + v);
                    if (DIRECTED) pb += __ldcg(a.partB + (size_t)b * a.np + v);
                }
            }
            s_red[w * 32 + lane] = pa;
            if (DIRECTED) s_red[NWARPS * 32 + w * 32 + lane] = pb;
            consumer_sync();
            if (w == 0 && v < a.n) {
                double sa = 0.0, sb = 0.0;
#pragma unroll
                for (int w2 = 0; w2 < NWARPS; ++w2) {
                    sa += s_red[w2 * 32 + lane];
                    if (DIRECTED) sb += s_red[NWARPS * 32 + w2 * 32 + lane];
                }
                if (!DIRECTED) {
                    const double t = __ldcg(a.Ta + v), wv = a.w_a[v];
                    const double s = t * sa;
                    a.Tw_a[v] = t + eps * t * (wv / s - 1.0);
                    a.S_a[v] = s;
                    e = fmax(e, fabs(wv - s));
                } else {
                    const double ti = __ldcg(a.Ta + v), to = __ldcg(a.Tb + v);
                    const double gd = powm_any<M>(a.qdiag[v], a.m);
                    const double sin = ti * (sa + to * gd), sout = to * (sb + ti * gd);
                    a.S_a[v] = sin;
                    a.S_b[v] = sout;
                    const double di = a.w_a[v], dout = a.w_b[v];
                    if (di > 0.0) {
                        a.Tw_a[v] = ti + eps * ti * (di / sin - 1.0);
                        e = fmax(e, fabs(di - sin));
                    }
                    if (dout > 0.0) {
                        a.Tw_b[v] = to + eps * to * (dout / sout - 1.0);
                        e = fmax(e, fabs(dout - sout));
                    }
                }
            }
            consumer_sync();
        }
        if (w == 0) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) e = fmax(e, __shfl_xor_sync(FULL, e, off));
            if (lane == 0)
                atomicMax(a.slots + it % 3, (unsigned long long)__double_as_longlong(e));
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) a.slots[(it + 1) % 3] = 0ull;
        grid_barrier_consumers(gbar, gridDim.x, epoch);
        const double f = __longlong_as_double((long long)__ldcg(a.slots + it % 3));
        if (DIRECTED && f > diff) eps *= 0.99;
        diff = f;
        ++it;
        const bool more = diff > a.delta && it < a.max_iter;
        if (threadIdx.x == 0) *s_go = more ? it + 1 : -1;  // verdict for the producer
        if (!more) break;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *a.out_iters = it;
        *a.out_diff = diff;
    }
}

}  // namespace cge
