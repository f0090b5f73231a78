// cge_rc.cuh -- interface of the recompute regime (cge_recompute.cu): no stored pair matrix, every
// pass re-derives q_ij from the embedding rows in FP64.  See the header of cge_recompute.cu.
#pragma once
#include "cge_kernels.cuh"

namespace cge {

constexpr int RC_DK = 16;                    // embedding dimensions per staged chunk
constexpr int RC_LD = TILE + 4;              // doubles per kk row of a chunk: 128 rows + 4 pad (bank-conflict-free
                                             // fragment loads: 132 * 8 B shifts consecutive kk by 8 banks)
constexpr int RC_CHUNK = RC_DK * RC_LD;      // doubles of one operand chunk: [kk][132], 16.5 KB
constexpr int RC_MAX_SB = 8;                 // largest super-block (tile rows / columns per super-tile)

// Work of the recompute regime is dealt in SUPER-TILES: sb x sb tiles (sb = 1, 2, 4 or 8; only the
// tiles bi <= bj of a diagonal super-tile).  A CTA runs the tiles of a super-tile one after the
// other and keeps their row / column sums in shared-memory accumulators, so the partial-sum slots in
// HBM are part[nsb][np] (one per super-block and vertex) instead of the stored regime's part[nb][np]:
// 7.8 GB instead of 62 GB at 10^6 vertices.  Every slot still has exactly one writer and every sum a
// fixed order: bit-reproducible, no atomics in the fixed point.
struct RcArgs : SweepArgs {
    const int2 *st_ij;            // [n_st] (I, J), I <= J, row-major upper triangle of super-blocks
    long long st_begin, st_end;   // this rank's share of the super-tile sequence
    int sb, nsb;                  // tiles per super-block side; super-blocks
    int srow_begin, srow_end;     // super-rows I touched by [st_begin, st_end) (end < begin: none)
    int nchunk;                   // dp / RC_DK
    const double *opT;            // operand image of the Gram / difference loop: per 128-row block and
                                  // 16-dimension chunk one contiguous [kk][132] block of 16.5 KB (what one
                                  // cp.async.bulk brings into shared memory); the embedding centred at its
                                  // mean for the row-norm / dot form, the raw embedding otherwise
    // "store what fits": the super-tiles [st_begin, st_store_end) of this rank keep their q tiles in
    // HBM (row-major 128 x 128, in the order the CTA walks them) and are read instead of recomputed
    double *qst;                  // tile t of super-tile st at qst + (st_pre[st] - st_pre[st_begin] + t) * TILE_ELEMS
    const long long *st_pre;      // [n_st + 1] tiles before super-tile st in the global sequence
    long long st_store_end;       // <= st_begin: nothing stored
};

size_t rc_smem_bytes();
// kind: 0 = fixed-point pass undirected, 1 = directed, 2 = B undirected, 3 = B directed
void launch_tiles_rc(int kind, int grid, cudaStream_t stream, const RcArgs &a, bool dot);
const void *fp_kernel_rc(int directed, int dot);
void launch_extrema_rc(int grid, cudaStream_t stream, const RcArgs &a, unsigned long long *lohi, bool dot);
// fills the q tiles of the super-tiles [a.st_begin, a.st_end) (called with st_end = the stored prefix)
void launch_store_rc(int grid, cudaStream_t stream, const RcArgs &a, bool dot);
void launch_reduce_part_rc(const RcArgs &a, const double *part, double *sraw, cudaStream_t stream);
void launch_rc_pack(const double *emb, const double *mean, int n, int np, int dp, double *opT,
                    double *nrm, cudaStream_t stream);
// q of sampled pairs in the regime's arithmetic; nrm == nullptr selects the difference form
void launch_sample_q_dot(const double *opT, int nchunk, const double *nrm, const double *emb, int dp,
                         const int *ia, const int *ib, const double *diag,
                         const unsigned long long *lohi, long long count, double *out,
                         cudaStream_t stream);

}  // namespace cge
