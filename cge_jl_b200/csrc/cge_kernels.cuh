// cge_kernels.cuh -- device code of the stored-regime pair-matrix sweeps (sm_100a).
//
// Data layout (DESIGN.md "Layout in HBM"): vertices are sorted by community and padded to
// np = nb*128; the symmetric matrix q_ij = (1 - D_ij)^(1/4) is stored as the upper-triangular
// sequence of 128x128 FP64 tiles (bi <= bj, row-major inside a tile, diagonal tiles stored as
// full symmetric squares, pad entries 0).  Because alpha = m/4, the geometric kernel of
// divergence.jl:142-148 is (1-D)^alpha = q^m, evaluated on the fly with <= 8 multiplies, so one
// pass over the matrix moves 8 bytes per unordered pair.
//
// One pass (divergence.jl:152-159) = every tile (bi,bj) produces
//     row partials  sum_c T_c q_rc^m  -> part[bj][bi*128 + r]
//     col partials  sum_r T_r q_rc^m  -> part[bi][bj*128 + c]      (bi != bj)
// Each slot part[b][v] is written by exactly one tile, so the reduction over b done by the
// finalize kernel has a fixed order (bit-reproducible, no atomics).
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

namespace cge {

constexpr int TILE = 128;
constexpr int TILE_ELEMS = TILE * TILE;
constexpr int NTHREADS = 256;
constexpr int NWARPS = NTHREADS / 32;
constexpr int ROWS_PER_WARP = TILE / NWARPS;  // 16
constexpr unsigned FULL = 0xffffffffu;

struct SweepArgs {
    const double *q;       // this rank's tiles: global tile t at q + (t - tile_begin) * TILE_ELEMS
    const int2 *tile_ij;   // [n_tiles] (bi, bj) of every global tile
    long long tile_begin, tile_end;
    int row_begin, row_end;  // tile rows bi touched by [tile_begin, tile_end)
    int nb, np, n, k;
    const double *Ta;      // undirected: T        directed: Tin
    const double *Tb;      //                      directed: Tout
    double *partA;         // [nb][np] undirected: S partials   directed: Sin partials
    double *partB;         //                                   directed: Sout partials
    const int *comm;       // [np] community per (sorted) vertex, -1 on pads
    double *B;             // [k][k] expected community mass (divergence.jl:228-234 / 532-538)
    // persistent fixed-point kernel only
    double *Tw_a, *Tw_b;   // the same T arrays, writable
    const double *w_a;     // target degrees: undirected vweights, directed degree_in
    const double *w_b;     //                                     directed degree_out
    const double *qdiag;   // q_ii (directed: the diagonal term is counted twice)
    double *S_a, *S_b;     // last S (Sin / Sout), for probes
    unsigned long long *slots;  // [3] residual slots, zero on entry
    double eps0, delta;
    int max_iter;
    int *out_iters;        // passes executed
    double *out_diff;      // last residual
    long long resident_tiles;  // local tiles [0, resident_tiles) are kept in L2 (evict_last)
    int skip_first_tiles = 0;  // the partial slots of pass 1 were already filled by k_bfp (fused B + pass)
    // recompute regime only
    const double *emb;     // [np][dp] sorted, zero-padded embedding
    const double *emb_c = nullptr;  // row-norm / dot form only: the same, centred at the mean
    const double *nrm = nullptr;    //   and its squared row norms [np]
    const double *diag;    // [np] D_ii (divergence.jl:85-86)
    const unsigned long long *lohi;  // bit patterns of lo, hi (divergence.jl:92)
    int dp, m;             // padded dimension; exponent m = 4*alpha
    // multi-GPU persistent driver: per-pass exchange of the raw degree sums over NVLink peer memory
    int rank, n_ranks;     // n_ranks == 1: no exchange
    unsigned pass_base;    // passes completed before this launch (flags only grow)
    long long xcap;        // vertex capacity of one exchange slot
    uint4 *xbuf_peer[8];   // rank r's exchange buffer [2 parities][n_ranks writers][2][xcap] of 16-byte
                           // records {lo32, pass_no, hi32, pass_no}, peer-mapped (CUDA IPC)
    unsigned long long *phase_ns;  // optional [8]: time of block 0 per phase of a pass (CGE_B200_PHASES=1)
};

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// phase clock of block 0 / thread 0: adds the time since the previous mark to phase_ns[slot]
struct PhaseClock {
    unsigned long long *acc, last;
    __device__ __forceinline__ PhaseClock(unsigned long long *p)
        : acc((blockIdx.x == 0 && threadIdx.x == 0) ? p : nullptr), last(0) {
        if (acc) last = global_ns();
    }
    __device__ __forceinline__ void mark(int slot) {
        if (acc) {
            const unsigned long long now = global_ns();
            acc[slot] += now - last;
            last = now;
        }
    }
};

// Exchange records carry their own arrival flag (the protocol NCCL calls LL): a double travels
// as two 8-byte stores {lo32, pass_no} and {hi32, pass_no}; an 8-byte store is atomic, so a
// reader that sees pass_no in both halves has the value -- no fence, no separate flag, no
// barrier between the producer's stores and the consumer's loads.
__device__ __forceinline__ void xchg_store(uint4 *rec, double v, unsigned pass_no) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(rec), "r"((unsigned)b),
                 "r"(pass_no)
                 : "memory");
    asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(
                     reinterpret_cast<char *>(rec) + 8),
                 "r"((unsigned)(b >> 32)), "r"(pass_no)
                 : "memory");
}
__device__ __forceinline__ double xchg_load(const uint4 *rec, unsigned pass_no) {
    unsigned lo, f0, hi, f1, spins = 0;
    do {
        asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(lo), "=r"(f0), "=r"(hi), "=r"(f1)
                     : "l"(rec)
                     : "memory");
        if (++spins > (1u << 26)) __trap();  // a lost peer must not hang the GPU
    } while (f0 != pass_no || f1 != pass_no);
    return __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
}

// Partial slots part[b][v] that this rank's tiles can have written for vertex v (block bv):
// row sums of tiles (bv, b >= bv) when bv is one of the rank's tile rows, column sums of tiles
// (b, bv) for the rank's tile rows b < bv.  Everything else is still zero and is not read
// (at 200k vertices on 8 GPUs this removes most of the 2.5 GB per pass the reduction would read).
__device__ __forceinline__ void part_range(const SweepArgs &a, int v, int &lo, int &hi) {
    const int bv = v / TILE;
    lo = a.row_begin;
    hi = (bv >= a.row_begin && bv <= a.row_end) ? a.nb : (bv > a.row_end ? a.row_end + 1 : a.row_begin);
}

// q^M with a fixed multiplication chain (binary powering), M = 4*alpha in 1..40
template <int M>
__device__ __forceinline__ double powm(double q) {
    if constexpr (M == 1) {
        return q;
    } else if constexpr (M % 2 == 0) {
        const double h = powm<M / 2>(q);
        return h * h;
    } else {
        return powm<M - 1>(q) * q;
    }
}

__device__ __forceinline__ double powm_rt(double q, int m) {
    double r = 1.0, b = q;
    while (m) {
        if (m & 1) r *= b;
        b *= b;
        m >>= 1;
    }
    return r;
}
// M > 0: compile-time exponent (the large-problem kernels, one instantiation per alpha);
// M == 0: exponent read from the arguments -- ONE kernel for the whole alpha grid.  Small problems
// (landmark mode: a few hundred to a few thousand vertices) are latency bound, and for them the
// ~10 ms CUDA needs to load each of 80 kernel instantiations on first use would dwarf the run.
template <int M>
__device__ __forceinline__ double powm_any(double q, int m_rt) {
    if constexpr (M == 0) {
        return powm_rt(q, m_rt);
    } else {
        return powm<M>(q);
    }
}

// 16-byte load of matrix data: read once per pass, kept out of L1.  The L2 policy decides what
// survives from one pass to the next: the first `resident_tiles` tiles are loaded evict_last (they
// stay in the 126 MB L2 across passes), the rest evict_first (they stream through without
// displacing the resident set) -- a cyclic sweep under plain LRU would hit nothing.
__device__ __forceinline__ uint64_t l2_policy(bool keep) {
    uint64_t pl, pf;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pl));
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pf));
    return keep ? pl : pf;
}
// The L2-resident set is the first resident_tiles tiles of the sequence (every CTA starts a pass
// on L2 hits).  Measured alternative (r01): a staggered set (CTA c keeps its k-th tile when
// k % P == c % P) is slower for every P tried (P = 4..8: 75..69 us/pass against 67).
__device__ __forceinline__ uint64_t l2_policy_for(long long local_tile, long long resident_tiles) {
    if (resident_tiles < 0) {
        uint64_t pn;
        asm("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pn));
        return pn;
    }
    return l2_policy(local_tile < resident_tiles);
}
__device__ __forceinline__ double2 ld_stream(const double2 *p, uint64_t pol) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;"
                 : "=d"(v.x), "=d"(v.y)
                 : "l"(p), "l"(pol));
    return v;
}

// Reduce NV per-lane values across the 32 lanes of a warp with NV-1 + (5 - log2 NV) shuffle
// steps instead of 5*NV.  On return v[0] holds the warp total of value index
// treduce_index<NV>(lane); the summation tree is fixed.
template <int NV0, int NV, int OFF>
struct TReduce {
    static __device__ __forceinline__ void run(double (&v)[NV0], int lane) {
        if constexpr (OFF >= 1) {
            if constexpr (NV > 1) {
                constexpr int H = NV / 2;
                const bool up = (lane & OFF) != 0;
#pragma unroll
                for (int i = 0; i < H; ++i) {
                    const double send = up ? v[i] : v[i + H];
                    const double keep = up ? v[i + H] : v[i];
                    v[i] = keep + __shfl_xor_sync(FULL, send, OFF);
                }
                TReduce<NV0, H, OFF / 2>::run(v, lane);
            } else {
                v[0] += __shfl_xor_sync(FULL, v[0], OFF);
                TReduce<NV0, 1, OFF / 2>::run(v, lane);
            }
        }
    }
};
template <int NV>
__device__ __forceinline__ void warp_treduce(double (&v)[NV], int lane) {
    TReduce<NV, NV, 16>::run(v, lane);
}
template <int NV>
__device__ __forceinline__ int treduce_index(int lane) {
    // NV = 16 -> lane >> 1, NV = 8 -> lane >> 2
    return NV == 16 ? (lane >> 1) : NV == 8 ? (lane >> 2) : NV == 4 ? (lane >> 3) : (lane >> 4);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
    return v;
}

// ---------------------------------------------------------------------------------------------
// fixed-point pass, undirected (divergence.jl:152-159)
// warp w owns rows 16w..16w+15 of the tile, lane l owns columns 2l, 2l+1, 64+2l, 65+2l
// ---------------------------------------------------------------------------------------------
#ifndef CGE_ROWRED4
#define CGE_ROWRED4 1
#endif
// CGE_ROWRED4: the 16 row sums of a warp are reduced across the lanes four at a time (warp_treduce<4> per
// batch of four rows) instead of all at once (warp_treduce<16>).  Every row's sum runs through the same
// tree either way -- partners lane ^ 16, ^ 8, ^ 4, ^ 2, ^ 1 in that order, and a + b == b + a -- so the
// partial slots are bit-identical; what changes is that only 4 row sums are live at a time, which leaves
// the compiler room to keep more of a tile's 32 loads in flight under the 128-register cap.  Measured
// (r02, same box, config 4): 27.33 -> 26.79 ms per pass (0.894 -> 0.912 of the HBM peak); 10k example
// 69.7 -> 68.1 us.
template <int M>
__device__ __forceinline__ void tile_pass_u(const double *__restrict__ qt, int bi, int bj,
                                            const SweepArgs &a, double *s_col, uint64_t pol) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int row0 = w * ROWS_PER_WARP;
    const double *Tc = a.Ta + (size_t)bj * TILE;
    const double2 tc01 = __ldcg(reinterpret_cast<const double2 *>(Tc) + lane);
    const double2 tc23 = __ldcg(reinterpret_cast<const double2 *>(Tc + 64) + lane);
    const double trow =
        lane < ROWS_PER_WARP ? __ldcg(a.Ta + (size_t)bi * TILE + row0 + lane) : 0.0;
    const double2 *base = reinterpret_cast<const double2 *>(qt + (size_t)row0 * TILE);
    double c0 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0;
#if CGE_ROWRED4
#pragma unroll
    for (int batch = 0; batch < ROWS_PER_WARP / 4; ++batch) {
        double r4[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int rr = batch * 4 + r;
            const double2 v01 = ld_stream(base + rr * (TILE / 2) + lane, pol);
            const double2 v23 = ld_stream(base + rr * (TILE / 2) + 32 + lane, pol);
            const double ti = __shfl_sync(FULL, trow, rr);
            const double g0 = powm_any<M>(v01.x, a.m), g1 = powm_any<M>(v01.y, a.m);
            const double g2 = powm_any<M>(v23.x, a.m), g3 = powm_any<M>(v23.y, a.m);
            r4[r] = fma(g3, tc23.y, fma(g2, tc23.x, fma(g1, tc01.y, g0 * tc01.x)));
            c0 = fma(ti, g0, c0);
            c1 = fma(ti, g1, c1);
            c2 = fma(ti, g2, c2);
            c3 = fma(ti, g3, c3);
        }
        warp_treduce<4>(r4, lane);
        if ((lane & 7) == 0)
            a.partA[(size_t)bj * a.np + (size_t)bi * TILE + row0 + batch * 4 + treduce_index<4>(lane)] = r4[0];
    }
#else
    double racc[ROWS_PER_WARP];
#pragma unroll
    for (int rr = 0; rr < ROWS_PER_WARP; ++rr) {
        const double2 v01 = ld_stream(base + rr * (TILE / 2) + lane, pol);
        const double2 v23 = ld_stream(base + rr * (TILE / 2) + 32 + lane, pol);
        const double ti = __shfl_sync(FULL, trow, rr);
        const double g0 = powm_any<M>(v01.x, a.m), g1 = powm_any<M>(v01.y, a.m);
        const double g2 = powm_any<M>(v23.x, a.m), g3 = powm_any<M>(v23.y, a.m);
        racc[rr] = fma(g3, tc23.y, fma(g2, tc23.x, fma(g1, tc01.y, g0 * tc01.x)));
        c0 = fma(ti, g0, c0);
        c1 = fma(ti, g1, c1);
        c2 = fma(ti, g2, c2);
        c3 = fma(ti, g3, c3);
    }
    warp_treduce<ROWS_PER_WARP>(racc, lane);
    if ((lane & 1) == 0)
        a.partA[(size_t)bj * a.np + (size_t)bi * TILE + row0 + treduce_index<16>(lane)] = racc[0];
#endif
    const bool offdiag = bi != bj;
    if (offdiag) {
        double2 *sc = reinterpret_cast<double2 *>(s_col + w * TILE);
        sc[lane] = make_double2(c0, c1);
        sc[32 + lane] = make_double2(c2, c3);
    }
    __syncthreads();
    if (offdiag && threadIdx.x < TILE) {
        double s = 0.0;
#pragma unroll
        for (int w2 = 0; w2 < NWARPS; ++w2) s += s_col[w2 * TILE + threadIdx.x];
        a.partA[(size_t)bi * a.np + (size_t)bj * TILE + threadIdx.x] = s;
    }
}

// ---------------------------------------------------------------------------------------------
// fixed-point pass, directed (divergence.jl:437-449)
//   Sin_i  += Tin_i * Tout_j * g     Sin_j  += Tin_j * Tout_i * g
//   Sout_i += Tin_j * Tout_i * g     Sout_j += Tin_i * Tout_j * g
// partA collects the Sin sums without their own Tin factor, partB the Sout sums without Tout.
// The second copy of the diagonal term (i == j is added twice at :444-447) is added by the
// finalize kernel.  Four batches of 4 rows at two CTAs per SM; a single 16-row loop with 32 row
// accumulators needs 255 registers (one CTA per SM) and measured slower (config 3: 2.32 ms per
// pass against 2.00).
// ---------------------------------------------------------------------------------------------
template <int M>
__device__ __forceinline__ void tile_pass_d(const double *__restrict__ qt, int bi, int bj,
                                            const SweepArgs &a, double *s_col, uint64_t pol) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int row0 = w * ROWS_PER_WARP;
    const double *Tic = a.Ta + (size_t)bj * TILE, *Toc = a.Tb + (size_t)bj * TILE;
    const double2 ti01 = __ldcg(reinterpret_cast<const double2 *>(Tic) + lane);
    const double2 ti23 = __ldcg(reinterpret_cast<const double2 *>(Tic + 64) + lane);
    const double2 to01 = __ldcg(reinterpret_cast<const double2 *>(Toc) + lane);
    const double2 to23 = __ldcg(reinterpret_cast<const double2 *>(Toc + 64) + lane);
    const bool ld = lane < ROWS_PER_WARP;
    const double trow_in = ld ? __ldcg(a.Ta + (size_t)bi * TILE + row0 + lane) : 0.0;
    const double trow_out = ld ? __ldcg(a.Tb + (size_t)bi * TILE + row0 + lane) : 0.0;
    const double2 *base = reinterpret_cast<const double2 *>(qt + (size_t)row0 * TILE);
    double ci0 = 0.0, ci1 = 0.0, ci2 = 0.0, ci3 = 0.0;  // Sin column sums  (Tout_r * g)
    double co0 = 0.0, co1 = 0.0, co2 = 0.0, co3 = 0.0;  // Sout column sums (Tin_r * g)
    // four batches of four rows, the loads of the next batch issued before the current one is
    // consumed (register double buffering, as in tile_bpass).  Leaving the loads to the compiler inside
    // the row loop, which is what tile_pass_u does, measured the same (r02, config 3: 1.885 -> 1.875 ms).
    constexpr int BR = 4, NBAT = ROWS_PER_WARP / BR;
    double2 v01[2][BR], v23[2][BR];
#pragma unroll
    for (int r4 = 0; r4 < BR; ++r4) {
        v01[0][r4] = ld_stream(base + r4 * (TILE / 2) + lane, pol);
        v23[0][r4] = ld_stream(base + r4 * (TILE / 2) + 32 + lane, pol);
    }
#pragma unroll
    for (int batch = 0; batch < NBAT; ++batch) {
        const int cb = batch & 1, nx = cb ^ 1;
        if (batch + 1 < NBAT) {
#pragma unroll
            for (int r4 = 0; r4 < BR; ++r4) {
                v01[nx][r4] = ld_stream(base + ((batch + 1) * BR + r4) * (TILE / 2) + lane, pol);
                v23[nx][r4] = ld_stream(base + ((batch + 1) * BR + r4) * (TILE / 2) + 32 + lane, pol);
            }
        }
        double rin[BR], rout[BR];
#pragma unroll
        for (int r4 = 0; r4 < BR; ++r4) {
            const int rr = batch * BR + r4;
            const double t_in = __shfl_sync(FULL, trow_in, rr);
            const double t_out = __shfl_sync(FULL, trow_out, rr);
            const double2 q01 = v01[cb][r4], q23 = v23[cb][r4];
            const double g0 = powm_any<M>(q01.x, a.m), g1 = powm_any<M>(q01.y, a.m);
            const double g2 = powm_any<M>(q23.x, a.m), g3 = powm_any<M>(q23.y, a.m);
            rin[r4] = fma(g3, to23.y, fma(g2, to23.x, fma(g1, to01.y, g0 * to01.x)));
            rout[r4] = fma(g3, ti23.y, fma(g2, ti23.x, fma(g1, ti01.y, g0 * ti01.x)));
            ci0 = fma(t_out, g0, ci0);
            ci1 = fma(t_out, g1, ci1);
            ci2 = fma(t_out, g2, ci2);
            ci3 = fma(t_out, g3, ci3);
            co0 = fma(t_in, g0, co0);
            co1 = fma(t_in, g1, co1);
            co2 = fma(t_in, g2, co2);
            co3 = fma(t_in, g3, co3);
        }
        warp_treduce<BR>(rin, lane);
        warp_treduce<BR>(rout, lane);
        if ((lane & 7) == 0) {
            const size_t o = (size_t)bj * a.np + (size_t)bi * TILE + row0 + batch * BR +
                             treduce_index<BR>(lane);
            a.partA[o] = rin[0];
            a.partB[o] = rout[0];
        }
    }
    const bool offdiag = bi != bj;
    if (offdiag) {
        double2 *sa = reinterpret_cast<double2 *>(s_col + w * TILE);
        double2 *sb = reinterpret_cast<double2 *>(s_col + NWARPS * TILE + w * TILE);
        sa[lane] = make_double2(ci0, ci1);
        sa[32 + lane] = make_double2(ci2, ci3);
        sb[lane] = make_double2(co0, co1);
        sb[32 + lane] = make_double2(co2, co3);
    }
    __syncthreads();
    if (offdiag) {
        const int c = threadIdx.x & (TILE - 1);
        const double *src = s_col + (threadIdx.x >> 7) * NWARPS * TILE;
        double s = 0.0;
#pragma unroll
        for (int w2 = 0; w2 < NWARPS; ++w2) s += src[w2 * TILE + c];
        double *dst = (threadIdx.x >> 7) ? a.partB : a.partA;
        dst[(size_t)bi * a.np + (size_t)bj * TILE + c] = s;
    }
}

// ---------------------------------------------------------------------------------------------
// expected community mass B (divergence.jl:228-234 undirected, 532-538 directed).
// Vertices are sorted by community, so along the 16 rows of a warp the row community changes
// rarely: column accumulators are flushed (warp-reduced per column community, one FP64 atomic
// per bin) only when it does.  B does not feed back into the fixed point, so the atomics'
// summation order only perturbs the score at the 1e-16 level.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void flush_bins(double v, int bin_col, long long bin_base,
                                           long long col_stride, double *B, int lane) {
    // v: this lane's contribution for community bin_col (or -1 on pads).  Columns are sorted by
    // community, so a warp sees one or two communities: one masked warp reduction and one FP64
    // atomic per community present; beyond four communities the rest goes lane by lane.
    unsigned rem = __ballot_sync(FULL, bin_col >= 0);
    int round = 0;
    while (rem) {  // warp-uniform
        if (round++ == 4) {
            if (((rem >> lane) & 1u) && v != 0.0)
                atomicAdd(B + bin_base + (long long)bin_col * col_stride, v);
            break;
        }
        const int leader = __ffs(rem) - 1;
        const int key = __shfl_sync(FULL, bin_col, leader);
        const bool in = bin_col == key;
        const double s = warp_sum(in ? v : 0.0);
        if (lane == leader) atomicAdd(B + bin_base + (long long)key * col_stride, s);
        rem &= ~__ballot_sync(FULL, in);
    }
}

template <int M, bool DIRECTED>
__device__ __forceinline__ void tile_bpass(const double *__restrict__ qt, int bi, int bj,
                                           const SweepArgs &a, uint64_t pol) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int row0 = w * ROWS_PER_WARP;
    const int gc0 = bj * TILE + 2 * lane, gc2 = bj * TILE + 64 + 2 * lane;
    const int2 cc01 = __ldcg(reinterpret_cast<const int2 *>(a.comm + gc0));
    const int2 cc23 = __ldcg(reinterpret_cast<const int2 *>(a.comm + gc2));
    const int cc[4] = {cc01.x, cc01.y, cc23.x, cc23.y};
    const int gc[4] = {gc0, gc0 + 1, gc2, gc2 + 1};
    // column factors: undirected T_c; directed Tin_c (for B[cr][cc]) and Tout_c (for B[cc][cr])
    const double2 ta01 = __ldcg(reinterpret_cast<const double2 *>(a.Ta + gc0));
    const double2 ta23 = __ldcg(reinterpret_cast<const double2 *>(a.Ta + gc2));
    const double tca[4] = {ta01.x, ta01.y, ta23.x, ta23.y};
    double tcb[4] = {0.0, 0.0, 0.0, 0.0};
    if (DIRECTED) {
        const double2 tb01 = __ldcg(reinterpret_cast<const double2 *>(a.Tb + gc0));
        const double2 tb23 = __ldcg(reinterpret_cast<const double2 *>(a.Tb + gc2));
        tcb[0] = tb01.x; tcb[1] = tb01.y; tcb[2] = tb23.x; tcb[3] = tb23.y;
    }
    const bool ld = lane < ROWS_PER_WARP;
    const int grow = bi * TILE + row0 + lane;
    const int crow = ld ? __ldcg(a.comm + grow) : -1;
    // row factors: undirected T_r; directed Tout_r (rowA) and Tin_r (rowB)
    const double trow_a = ld ? __ldcg((DIRECTED ? a.Tb : a.Ta) + grow) : 0.0;
    const double trow_b = (DIRECTED && ld) ? __ldcg(a.Ta + grow) : 0.0;
    const bool diag = bi == bj;
    const double2 *base = reinterpret_cast<const double2 *>(qt + (size_t)row0 * TILE);
    double accA[4] = {0.0, 0.0, 0.0, 0.0}, accB[4] = {0.0, 0.0, 0.0, 0.0};
    int cur = __shfl_sync(FULL, crow, 0);

    auto flush = [&](int cr) {
        if (cr >= 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                // B[cr][cc] += (sum_r rowA_r g) * colA_c
                flush_bins(accA[k] * tca[k], cc[k], (long long)cr * a.k, 1, a.B, lane);
                if (DIRECTED && !diag)  // B[cc][cr] += (sum_r Tin_r g) * Tout_c
                    flush_bins(accB[k] * tcb[k], cc[k], (long long)cr, a.k, a.B, lane);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) accA[k] = accB[k] = 0.0;
    };

    // rows are taken in batches of four with the loads of the next batch issued before the
    // current one is consumed (register double buffering); a batch whose rows all share the
    // current row community (the common case after the community sort) runs without any per-row
    // check
    constexpr int BR = 4, NB = ROWS_PER_WARP / BR;
    double2 v01[2][BR], v23[2][BR];
#pragma unroll
    for (int r8 = 0; r8 < BR; ++r8) {
        v01[0][r8] = ld_stream(base + r8 * (TILE / 2) + lane, pol);
        v23[0][r8] = ld_stream(base + r8 * (TILE / 2) + 32 + lane, pol);
    }
#pragma unroll
    for (int batch = 0; batch < NB; ++batch) {
        const int cb = batch & 1, nb_ = cb ^ 1;
        if (batch + 1 < NB) {
#pragma unroll
            for (int r8 = 0; r8 < BR; ++r8) {
                v01[nb_][r8] = ld_stream(base + ((batch + 1) * BR + r8) * (TILE / 2) + lane, pol);
                v23[nb_][r8] =
                    ld_stream(base + ((batch + 1) * BR + r8) * (TILE / 2) + 32 + lane, pol);
            }
        }
        const unsigned bm = ((1u << BR) - 1u) << (batch * BR);
        const bool uniform = (__ballot_sync(FULL, crow != cur) & bm) == 0u;
#pragma unroll
        for (int r8 = 0; r8 < BR; ++r8) {
            const int rr = batch * BR + r8;
            if (!uniform) {
                const int cr = __shfl_sync(FULL, crow, rr);
                if (cr != cur) {  // warp-uniform
                    flush(cur);
                    cur = cr;
                }
            }
            double g[4] = {powm_any<M>(v01[cb][r8].x, a.m), powm_any<M>(v01[cb][r8].y, a.m),
                           powm_any<M>(v23[cb][r8].x, a.m), powm_any<M>(v23[cb][r8].y, a.m)};
            if (!DIRECTED && diag) {  // unordered pairs once: keep col >= row (divergence.jl:229-230)
                const int gr = bi * TILE + row0 + rr;
#pragma unroll
                for (int k = 0; k < 4; ++k) g[k] = gc[k] >= gr ? g[k] : 0.0;
            }
            const double ra = __shfl_sync(FULL, trow_a, rr);
#pragma unroll
            for (int k = 0; k < 4; ++k) accA[k] = fma(ra, g[k], accA[k]);
            if (DIRECTED) {
                const double rb = __shfl_sync(FULL, trow_b, rr);
#pragma unroll
                for (int k = 0; k < 4; ++k) accB[k] = fma(rb, g[k], accB[k]);
            }
        }
    }
    flush(cur);
}

// ---------------------------------------------------------------------------------------------
// Fused pass (undirected): B sweep of alpha = (M-1)/4 + FIRST fixed-point pass of alpha = M/4.
// T is warm-started (divergence.jl:33, 139-168: the T an alpha starts from is the T the previous
// alpha's P and B were built from), so the two sweeps read the same matrix with the same T and
// differ only in the exponent: q^(M-1) feeds the community bins (divergence.jl:228-234), q^M the
// degree sums (:152-159).  One read of the matrix instead of two.  The degree-sum arithmetic is
// tile_pass_u's operation for operation (same FMA chains, same reduction trees: bit-identical
// partial slots), the bin arithmetic is tile_bpass<M-1>'s.
// ---------------------------------------------------------------------------------------------
template <int M>
__device__ __forceinline__ void tile_pass_ub(const double *__restrict__ qt, int bi, int bj,
                                             const SweepArgs &a, double *s_col, uint64_t pol) {
    static_assert(M >= 2, "the fused pass needs a previous alpha");
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int row0 = w * ROWS_PER_WARP;
    const int gc0 = bj * TILE + 2 * lane, gc2 = bj * TILE + 64 + 2 * lane;
    const double2 tc01 = __ldcg(reinterpret_cast<const double2 *>(a.Ta + gc0));
    const double2 tc23 = __ldcg(reinterpret_cast<const double2 *>(a.Ta + gc2));
    const int2 cc01 = __ldcg(reinterpret_cast<const int2 *>(a.comm + gc0));
    const int2 cc23 = __ldcg(reinterpret_cast<const int2 *>(a.comm + gc2));
    const bool ld = lane < ROWS_PER_WARP;
    const int grow = bi * TILE + row0 + lane;
    const double trow = ld ? __ldcg(a.Ta + grow) : 0.0;
    const int crow = ld ? __ldcg(a.comm + grow) : -1;
    const bool diag = bi == bj;
    const double2 *base = reinterpret_cast<const double2 *>(qt + (size_t)row0 * TILE);
    double c0 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0;
    double b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;  // bins: sum_r T_r q^(M-1) per column
    int cur = __shfl_sync(FULL, crow, 0);

    auto flush = [&](int cr) {
        if (cr >= 0) {
            const long long bb = (long long)cr * a.k;
            flush_bins(b0 * tc01.x, cc01.x, bb, 1, a.B, lane);
            flush_bins(b1 * tc01.y, cc01.y, bb, 1, a.B, lane);
            flush_bins(b2 * tc23.x, cc23.x, bb, 1, a.B, lane);
            flush_bins(b3 * tc23.y, cc23.y, bb, 1, a.B, lane);
        }
        b0 = b1 = b2 = b3 = 0.0;
    };
    // The 16 rows of a warp almost always share one community (vertices are sorted by community), and then
    // the row loop has no branch in it -- the loads of the whole tile can be scheduled ahead, as in
    // tile_pass_u; a warp whose rows straddle a community boundary takes the same loop with the per-row
    // check.  Row sums are reduced four rows at a time (see CGE_ROWRED4: bit-identical to tile_pass_u).
    auto rows = [&](auto check_tag) {
        constexpr bool CHECK = decltype(check_tag)::value;
#pragma unroll
        for (int batch = 0; batch < ROWS_PER_WARP / 4; ++batch) {
            double r4[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int rr = batch * 4 + r;
                if constexpr (CHECK) {
                    const int cr = __shfl_sync(FULL, crow, rr);
                    if (cr != cur) {  // warp-uniform
                        flush(cur);
                        cur = cr;
                    }
                }
                const double2 v01 = ld_stream(base + rr * (TILE / 2) + lane, pol);
                const double2 v23 = ld_stream(base + rr * (TILE / 2) + 32 + lane, pol);
                const double q0 = v01.x, q1 = v01.y, q2 = v23.x, q3 = v23.y;
                const double ti = __shfl_sync(FULL, trow, rr);
                // fixed-point pass, exponent M (as tile_pass_u)
                const double g0 = powm<M>(q0), g1 = powm<M>(q1), g2 = powm<M>(q2), g3 = powm<M>(q3);
                r4[r] = fma(g3, tc23.y, fma(g2, tc23.x, fma(g1, tc01.y, g0 * tc01.x)));
                c0 = fma(ti, g0, c0);
                c1 = fma(ti, g1, c1);
                c2 = fma(ti, g2, c2);
                c3 = fma(ti, g3, c3);
                // community bins, exponent M-1 (as tile_bpass<M-1, false>)
                double h0 = powm<M - 1>(q0), h1 = powm<M - 1>(q1), h2 = powm<M - 1>(q2), h3 = powm<M - 1>(q3);
                if (diag) {  // unordered pairs once: keep col >= row (divergence.jl:229-230)
                    const int gr = bi * TILE + row0 + rr;
                    h0 = gc0 >= gr ? h0 : 0.0;
                    h1 = gc0 + 1 >= gr ? h1 : 0.0;
                    h2 = gc2 >= gr ? h2 : 0.0;
                    h3 = gc2 + 1 >= gr ? h3 : 0.0;
                }
                b0 = fma(ti, h0, b0);
                b1 = fma(ti, h1, b1);
                b2 = fma(ti, h2, b2);
                b3 = fma(ti, h3, b3);
            }
            warp_treduce<4>(r4, lane);
            if ((lane & 7) == 0)
                a.partA[(size_t)bj * a.np + (size_t)bi * TILE + row0 + batch * 4 + treduce_index<4>(lane)] = r4[0];
        }
    };
    if ((__ballot_sync(FULL, crow != cur) & 0xffffu) == 0u)  // lanes 0..15 hold the rows' communities
        rows(std::false_type{});
    else
        rows(std::true_type{});
    flush(cur);
    if (!diag) {
        double2 *sc = reinterpret_cast<double2 *>(s_col + w * TILE);
        sc[lane] = make_double2(c0, c1);
        sc[32 + lane] = make_double2(c2, c3);
    }
    __syncthreads();
    if (!diag && threadIdx.x < TILE) {
        double s = 0.0;
#pragma unroll
        for (int w2 = 0; w2 < NWARPS; ++w2) s += s_col[w2 * TILE + threadIdx.x];
        a.partA[(size_t)bi * a.np + (size_t)bj * TILE + threadIdx.x] = s;
    }
}

// ---------------------------------------------------------------------------------------------
// kernels: grid-stride over this rank's tiles
// ---------------------------------------------------------------------------------------------
template <int M, bool DIRECTED>
__global__ void __launch_bounds__(NTHREADS, 2) k_sweep(const __grid_constant__ SweepArgs a) {
    __shared__ __align__(16) double s_col[2 * 2 * NWARPS * TILE];
    int it = 0;
    for (long long t = a.tile_begin + blockIdx.x; t < a.tile_end; t += gridDim.x, ++it) {
        const int2 ij = a.tile_ij[t];
        const double *qt = a.q + (size_t)(t - a.tile_begin) * TILE_ELEMS;
        const uint64_t pol = l2_policy_for(t - a.tile_begin, a.resident_tiles);
        double *sc = s_col + (it & 1) * 2 * NWARPS * TILE;
        if constexpr (DIRECTED)
            tile_pass_d<M>(qt, ij.x, ij.y, a, sc, pol);
        else
            tile_pass_u<M>(qt, ij.x, ij.y, a, sc, pol);
    }
}

template <int M, bool DIRECTED>
__global__ void __launch_bounds__(NTHREADS, 2) k_bsweep(const __grid_constant__ SweepArgs a) {
    for (long long t = a.tile_begin + blockIdx.x; t < a.tile_end; t += gridDim.x) {
        const int2 ij = a.tile_ij[t];
        tile_bpass<M, DIRECTED>(a.q + (size_t)(t - a.tile_begin) * TILE_ELEMS, ij.x, ij.y, a,
                                l2_policy_for(t - a.tile_begin, a.resident_tiles));
    }
}

// B sweep of alpha (M-1)/4 fused with the first fixed-point pass of alpha M/4 (tile_pass_ub): fills
// B and the partial slots; k_fixed_point<M> launched behind it with skip_first_tiles = 1 starts at
// the reduction of these slots.
template <int M>
__global__ void __launch_bounds__(NTHREADS, 2) k_bfp(const __grid_constant__ SweepArgs a) {
    __shared__ __align__(16) double s_col[2 * NWARPS * TILE];
    int it = 0;
    for (long long t = a.tile_begin + blockIdx.x; t < a.tile_end; t += gridDim.x, ++it) {
        const int2 ij = a.tile_ij[t];
        tile_pass_ub<M>(a.q + (size_t)(t - a.tile_begin) * TILE_ELEMS, ij.x, ij.y, a,
                        s_col + (it & 1) * NWARPS * TILE,
                        l2_policy_for(t - a.tile_begin, a.resident_tiles));
    }
}
template <int M>
inline void launch_bfp(int grid, cudaStream_t stream, const SweepArgs &a) {
    if constexpr (M >= 2) k_bfp<M><<<grid, NTHREADS, 0, stream>>>(a);
}

// divergence.jl:160-165 (undirected) / 451-461 with the doubled diagonal of :442-447 (directed)
// for vertex v given the summed raw degree sums; returns the vertex's residual.
template <int M, bool DIRECTED>
__device__ __forceinline__ double fp_update(const SweepArgs &a, int v, double sa, double sb,
                                            double eps) {
    double e = 0.0;
    if (!DIRECTED) {
        const double t = __ldcg(a.Ta + v), wv = a.w_a[v];
        const double s = t * sa;
        a.Tw_a[v] = t + eps * t * (wv / s - 1.0);
        a.S_a[v] = s;
        e = fabs(wv - s);
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
    return e;
}

// ---------------------------------------------------------------------------------------------
// Persistent fixed point of one alpha (divergence.jl:150-168 / 434-467) as ONE cooperative
// launch: every pass is [tiles] -> grid.sync -> [reduce partials, update T, residual] ->
// grid.sync, so the host is out of the loop (no launch or D2H per pass).  All control flow is
// computed identically by every thread from the residual slot.
// ---------------------------------------------------------------------------------------------
template <int M, bool DIRECTED>
__global__ void __launch_bounds__(NTHREADS, 2) k_fixed_point(const __grid_constant__ SweepArgs a) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ __align__(16) double s_col[2 * 2 * NWARPS * TILE];
    __shared__ double s_red[(DIRECTED ? 2 : 1) * NWARPS * 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int ngroups = (a.n + 31) / 32;
    double diff = 1.0, eps = a.eps0;
    int it = 0, tile_it = 0;
    PhaseClock clk(a.phase_ns);
    while (diff > a.delta && it < a.max_iter) {
        // tiles are dealt round-robin (t = blockIdx.x + k*gridDim.x): at any moment the grid reads
        // one contiguous ~38 MB window of the matrix.  Measured alternatives (r01, 10k example):
        // per-CTA contiguous ranges 86 us/pass, one barrier per two tiles 75, claiming tiles from
        // a global counter 70, this 66-67.
        // (pass 1 of an alpha whose partial slots k_bfp has just filled has no tile phase)
        if (!(it == 0 && a.skip_first_tiles)) {
            for (long long t = a.tile_begin + blockIdx.x; t < a.tile_end; t += gridDim.x, ++tile_it) {
                const int2 ij = a.tile_ij[t];
                const double *qt = a.q + (size_t)(t - a.tile_begin) * TILE_ELEMS;
                const uint64_t pol = l2_policy_for(t - a.tile_begin, a.resident_tiles);
                double *sc = s_col + (tile_it & 1) * 2 * NWARPS * TILE;
                if constexpr (DIRECTED)
                    tile_pass_d<M>(qt, ij.x, ij.y, a, sc, pol);
                else
                    tile_pass_u<M>(qt, ij.x, ij.y, a, sc, pol);
            }
        }
        clk.mark(0);  // tiles of block 0
        grid.sync();
        clk.mark(1);  // waiting for the slowest CTA + barrier
        // 32 vertices per CTA step: warp w sums the partial slots b = w, w+8, ..., warp 0 adds
        // the eight sub-sums in fixed order.  Single GPU: warp 0 applies divergence.jl:160-165 /
        // 451-461 at once.  Multi GPU: the sums of this rank's tiles go to every rank's exchange
        // buffer over NVLink as self-flagged records (xchg_store), and every rank adds the n_ranks
        // contributions in rank order as they arrive -- identical T on all ranks, no host round
        // trip, no separate collective, no extra barrier.
        double e = 0.0;
        const bool multi = a.n_ranks > 1;
        const unsigned pass_no = a.pass_base + (unsigned)it + 1u;
        const size_t xpar = (size_t)(pass_no & 1u) * a.n_ranks;
        // two 32-vertex groups per CTA step, four warps each (sub-warp q sums slots b = q, q+4, ..):
        // 592 group slots per round cover the 313 groups of the 10k example in one round
        constexpr int RW = NWARPS / 2;  // warps per group
        const int half = w / RW, q4 = w % RW;
        for (int g0 = blockIdx.x * 2; g0 < ngroups; g0 += gridDim.x * 2) {
            const int g = g0 + half;
            const int v = g * 32 + lane;
            const bool live = g < ngroups && v < a.n;
            double pa = 0.0, pb = 0.0;
            if (live) {
                int b_lo, b_hi;
                part_range(a, v, b_lo, b_hi);
                for (int b = b_lo + q4; b < b_hi; b += RW) {
                    pa += __ldcg(a.partA + (size_t)b * a.np + v);
                    if (DIRECTED) pb += __ldcg(a.partB + (size_t)b * a.np + v);
                }
            }
            s_red[w * 32 + lane] = pa;
            if (DIRECTED) s_red[NWARPS * 32 + w * 32 + lane] = pb;
            __syncthreads();
            if (q4 == 0 && live) {
                double sa = 0.0, sb = 0.0;
#pragma unroll
                for (int w2 = 0; w2 < RW; ++w2) {
                    sa += s_red[(half * RW + w2) * 32 + lane];
                    if (DIRECTED) sb += s_red[NWARPS * 32 + (half * RW + w2) * 32 + lane];
                }
                if (multi) {
                    const size_t o = ((xpar + a.rank) * 2) * (size_t)a.xcap + v;
                    for (int r = 0; r < a.n_ranks; ++r) {
                        xchg_store(a.xbuf_peer[r] + o, sa, pass_no);
                        if (DIRECTED) xchg_store(a.xbuf_peer[r] + o + a.xcap, sb, pass_no);
                    }
                } else {
                    e = fmax(e, fp_update<M, DIRECTED>(a, v, sa, sb, eps));
                }
            }
            __syncthreads();
        }
        clk.mark(2);  // partial-slot reduction (+ update, or peer stores)
        if (multi) {
            // every rank adds the n_ranks contributions in rank order as they arrive
            const uint4 *mine = a.xbuf_peer[a.rank];
            for (int g = blockIdx.x * NWARPS + w; g < ngroups; g += gridDim.x * NWARPS) {
                const int v = g * 32 + lane;
                if (v < a.n) {
                    double sa = 0.0, sb = 0.0;
                    for (int r = 0; r < a.n_ranks; ++r) {
                        const size_t o = ((xpar + r) * 2) * (size_t)a.xcap + v;
                        sa += xchg_load(mine + o, pass_no);
                        if (DIRECTED) sb += xchg_load(mine + o + a.xcap, pass_no);
                    }
                    e = fmax(e, fp_update<M, DIRECTED>(a, v, sa, sb, eps));
                }
            }
            clk.mark(5);  // wait for the peers' sums + update
        }
        if (w % (NWARPS / 2) == 0 || multi) {  // max is order independent: one atomic per contributing warp
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) e = fmax(e, __shfl_xor_sync(FULL, e, off));
            if (lane == 0 && e > 0.0)
                atomicMax(a.slots + it % 3, (unsigned long long)__double_as_longlong(e));
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) a.slots[(it + 1) % 3] = 0ull;
        grid.sync();
        clk.mark(6);  // residual barrier
        const double f = __longlong_as_double((long long)__ldcg(a.slots + it % 3));
        if (DIRECTED && f > diff) eps *= 0.99;  // divergence.jl:462-464
        diff = f;
        ++it;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *a.out_iters = it;
        *a.out_diff = diff;
    }
}

// host-side dispatch over the compile-time exponent; defined in cge_inst_*.cu
// kind: 0 = sweep undirected, 1 = sweep directed, 2 = B undirected, 3 = B directed,
//       4 = B of exponent m-1 fused with the first pass of exponent m (undirected, m >= 2)
// fp_kernel(m, directed) returns the cooperative fixed-point kernel for cudaLaunchCooperativeKernel
void launch_tiles(int m, int kind, int grid, cudaStream_t stream, const SweepArgs &a);
const void *fp_kernel(int m, int directed);
const void *fp_kernel_part0(int m, int directed);
const void *fp_kernel_part1(int m, int directed);
const void *fp_kernel_part2(int m, int directed);
const void *fp_kernel_part3(int m, int directed);
const void *fp_kernel_part4(int m, int directed);
const void *fp_kernel_part5(int m, int directed);
const void *fp_kernel_part6(int m, int directed);
const void *fp_kernel_part7(int m, int directed);
// the same fixed point with the matrix streamed by cp.async.bulk through a shared-memory ring
const void *fp_ring_kernel(int m, int directed);
const void *fp_ring_kernel_part0(int m, int directed);
const void *fp_ring_kernel_part1(int m, int directed);
const void *fp_ring_kernel_part2(int m, int directed);
const void *fp_ring_kernel_part3(int m, int directed);
const void *fp_ring_kernel_part4(int m, int directed);
const void *fp_ring_kernel_part5(int m, int directed);
const void *fp_ring_kernel_part6(int m, int directed);
const void *fp_ring_kernel_part7(int m, int directed);
size_t fp_ring_smem_bytes(int directed);
// stored regime with a run-time exponent (small problems): kind as in launch_tiles
void launch_tiles_rt(int kind, int grid, cudaStream_t stream, const SweepArgs &a);
const void *fp_kernel_rt(int directed);
// tensor-core diameter filter (cge_diameter.cu)
struct DiamArgs {
    const unsigned char *packed;  // per 128-row block: hi part then lo part, each ksteps*4096 B
    const float *norms;           // [nb*128] squared norms, a large negative value on pads
    int nb, ksteps;               // blocks; dp/16
    const int4 *strips;           // (bi, bj0, count, first tile index): runs of tiles in one tile row
    int n_strips;
    unsigned *strip_counter;
    float *tile_max;              // [n_tiles] largest approximate d^2 of each tile
    unsigned *gmax_bits;          // bit pattern of the largest approximate d^2 (non-negative float)
};
size_t diameter_smem_bytes(int ksteps);
void launch_pack_bf16(const double *emb, const double *mean, int dp, int n, int d_true, int nb,
                      unsigned char *packed, float *norms, unsigned *rmax_bits, cudaStream_t st);
cudaError_t launch_diameter_filter(const DiamArgs &a, int grid, cudaStream_t st);
void launch_select_candidates(const float *tile_max, long long n_tiles, const unsigned *gmax_bits,
                              const unsigned *rmax_bits, float rel, int *list, int cap, int *count,
                              cudaStream_t st);
// recompute regime: see cge_rc.cuh
void launch_selftest_math(long long n, unsigned long long seed, unsigned long long *out, int grid,
                          cudaStream_t st);
double measure_fp64_peak_tflops(int sm_count, cudaStream_t st);
int measure_fp64_pipes(int sm_count, cudaStream_t st, double *out);  // cge_microbench.cu
// device sampler of non-edges (cge_sampler.cu)
void launch_edge_set_insert(const long long *src, const long long *dst, long long m, long long n,
                            int index_base, int directed, unsigned long long *table,
                            unsigned long long mask, unsigned long long *counts, int grid,
                            cudaStream_t st);
void launch_sample_non_edges(long long n, int directed, int index_base,
                             const unsigned long long *table, unsigned long long mask,
                             unsigned long long seed, long long total, long long *out_i,
                             long long *out_j, unsigned long long *counts, int grid,
                             cudaStream_t st);
int fp_ring_threads();
void launch_tiles_part0(int m, int kind, int grid, cudaStream_t stream, const SweepArgs &a);
void launch_tiles_part1(int m, int kind, int grid, cudaStream_t stream, const SweepArgs &a);
void launch_tiles_part2(int m, int kind, int grid, cudaStream_t stream, const SweepArgs &a);
void launch_tiles_part3(int m, int kind, int grid, cudaStream_t stream, const SweepArgs &a);
void launch_tiles_part4(int m, int kind, int grid, cudaStream_t stream, const SweepArgs &a);
void launch_tiles_part5(int m, int kind, int grid, cudaStream_t stream, const SweepArgs &a);
void launch_tiles_part6(int m, int kind, int grid, cudaStream_t stream, const SweepArgs &a);
void launch_tiles_part7(int m, int kind, int grid, cudaStream_t stream, const SweepArgs &a);

}  // namespace cge
