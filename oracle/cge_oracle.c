/*
 * cge_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A scalar, single-threaded CPU restatement of CGE.jl's scoring hot path, written to follow
 * the reference loop by loop (same packed n(n+1)/2 arrays, same iteration order, same
 * comparisons), so that the CUDA path in cge_jl_b200/csrc can be checked against it.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; the product path never does.
 *
 * Reference (all paths under /root/reference):
 *   src/divergence.jl:27-257   wGCL            -> cge_oracle_wgcl
 *   src/divergence.jl:282-561  wGCL_directed   -> cge_oracle_wgcl_directed
 *   src/auxilary.jl:14-20      dist            -> o_dist
 *   src/auxilary.jl:34-52      JS              -> cge_oracle_js
 *   src/auxilary.jl:57-59      idx             -> cge_oracle_idx
 *
 * Pinning: Julia is not installed in the build image, so the reference itself cannot run
 * here.  The oracle is pinned by the only numerical known-answer the reference publishes,
 * README.md:99 (10k example, -l 200 --seed 42): elements 1-2 (best alpha, global score) are
 * RNG-free and are reproduced by tests/test_oracle_golden.py.  Elements 5-7 (local score)
 * depend on Julia's RNG stream and Set iteration order (divergence.jl:137,184-210) and are
 * PARITY UNPINNED against the reference; the oracle takes the sample index arrays as inputs
 * so that the CUDA path and the oracle always see identical sample sets.
 *
 * Deviations from the Julia text, all outside the arithmetic:
 *   - the sampled edge / non-edge index arrays are inputs (the Julia code draws them with
 *     StatsBase.sample at divergence.jl:185,194,203,210,485,495,505,510,513);
 *   - embed is row-major here (Julia's Matrix is column-major); indices are 1-based at the
 *     interface exactly like the Julia arrays;
 *   - `max_alphas` (<= 40) lets bench.py time a bounded prefix of the alpha grid.
 *   - Julia's sum() is pairwise/SIMD; here sums are sequential (differences ~1e-16 relative).
 */
#define _POSIX_C_SOURCE 200809L /* clock_gettime under -std=c11 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <time.h>
#include <string.h>

#define N_ALPHA 40

typedef struct {
    int32_t n_alpha_run;        /* number of alpha values for which the fixed point ran */
    int32_t iters[N_ALPHA];     /* fixed-point passes per alpha */
    double div[N_ALPHA];        /* global score per alpha (NaN when skipped) */
    double auc[N_ALPHA];        /* local score per alpha (NaN when skipped) */
    double lo, hi;              /* extrema of the raw distance vector (divergence.jl:92) */
    double hi_full;             /* landmark mode: max distance of the full graph (:113) */
    double final_diff;          /* last max|w-S| */
    /* wGCL only: seconds spent in the O(n^2) phases -- D build (:79-93), GD = (1-D)^alpha (:142-148),
     * fixed-point passes (:150-168), P (:170-176), B (:228-234) -- so that bench.py can extrapolate a
     * bounded sample to a configuration whose arrays fit no host, phase by phase */
    double t_phase[5];
} cge_oracle_trace;

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* auxilary.jl:57-59 (1-based, i <= j) */
int64_t cge_oracle_idx(int64_t n, int64_t i, int64_t j) {
    return n * (i - 1) - (i - 1) * (i - 2) / 2 + j - i + 1;
}

/* auxilary.jl:14-20 (1-based rows of a row-major n x d matrix) */
static double o_dist(int64_t v1, int64_t v2, const double *embed, int64_t d) {
    if (v1 == v2) return 0.0;
    const double *a = embed + (v1 - 1) * d, *b = embed + (v2 - 1) * d;
    double s = 0.0;
    for (int64_t c = 0; c < d; ++c) {
        double t = a[c] - b[c];
        s += t * t;
    }
    return sqrt(s);
}
double cge_oracle_dist(int64_t v1, int64_t v2, const double *embed, int64_t d) {
    return o_dist(v1, v2, embed, d);
}

/* auxilary.jl:34-52.  vI == NULL means "no mask" (the Bool[] call at divergence.jl:236). */
double cge_oracle_js(const double *vC, const double *vB, const uint8_t *vI, int internal,
                     int64_t len) {
    double sp1 = 0.0, sp2 = 0.0;
    int64_t cnt = 0;
    for (int64_t i = 0; i < len; ++i) {
        if (vI && ((vI[i] != 0) != (internal != 0))) continue;
        sp1 += vC[i];
        sp2 += vB[i];
        ++cnt;
    }
    sp1 += (double)cnt;
    sp2 += (double)cnt;
    double f = 0.0;
    for (int64_t i = 0; i < len; ++i) {
        if (vI && ((vI[i] != 0) != (internal != 0))) continue;
        double p = (vC[i] + 1.0) / sp1;
        double q = (vB[i] + 1.0) / sp2;
        double m = (p + q) / 2.0;
        f += p * log(p / m) + q * log(q / m);
    }
    return f / 2.0;
}

static int64_t max_i64(const int64_t *a, int64_t n) {
    int64_t m = a[0];
    for (int64_t i = 1; i < n; ++i)
        if (a[i] > m) m = a[i];
    return m;
}

/* packed distance vector + min-max normalisation, divergence.jl:79-93 / 359-375 */
static double *build_D(int64_t n, const double *embed, int64_t d, const double *distances,
                       double *lo_out, double *hi_out) {
    int64_t p_len = n * (n + 1) / 2;
    double *D = (double *)malloc(sizeof(double) * (size_t)p_len);
    if (!D) return NULL;
    for (int64_t i = 1; i <= n; ++i)
        for (int64_t j = i; j <= n; ++j) {
            int64_t l = cge_oracle_idx(n, i, j);
            D[l - 1] = (i == j) ? distances[i - 1] : o_dist(i, j, embed, d);
        }
    double lo = D[0], hi = D[0];
    for (int64_t l = 1; l < p_len; ++l) {
        if (D[l] < lo) lo = D[l];
        if (D[l] > hi) hi = D[l];
    }
    for (int64_t l = 0; l < p_len; ++l) D[l] = (D[l] - lo) / (hi - lo);
    *lo_out = lo;
    *hi_out = hi;
    return D;
}

/* full-graph distances for the landmark-mode local score, divergence.jl:104-115 / 386-397:
 * diagonal entries stay 0 and take part in extrema(). */
static double *build_full_D(int64_t n, const double *embed, int64_t d, double *hi_out) {
    int64_t p_len = n * (n + 1) / 2;
    double *D = (double *)calloc((size_t)p_len, sizeof(double));
    if (!D) return NULL;
    for (int64_t i = 1; i <= n; ++i)
        for (int64_t j = i + 1; j <= n; ++j)
            D[cge_oracle_idx(n, i, j) - 1] = o_dist(i, j, embed, d);
    double lo = D[0], hi = D[0];
    for (int64_t l = 1; l < p_len; ++l) {
        if (D[l] < lo) lo = D[l];
        if (D[l] > hi) hi = D[l];
    }
    for (int64_t l = 0; l < p_len; ++l) D[l] = (D[l] - lo) / (hi - lo);
    *hi_out = hi;
    return D;
}

static void trace_init(cge_oracle_trace *tr) {
    if (!tr) return;
    memset(tr, 0, sizeof(*tr));
    for (int a = 0; a < N_ALPHA; ++a) tr->div[a] = tr->auc[a] = NAN;
}

/*
 * wGCL, divergence.jl:27-257.  Returns 0 on success; out[7] as the Julia return vector.
 * edges (m x 2, given as two 1-based columns), comm 1-based, embed row-major n x d.
 * Landmark mode iff n_full > 0 (v_to_l non-empty at :44).  Sample arrays are K x n_sets,
 * 1-based vertex ids of the ORIGINAL graph; n_sets is 1 (seeded) or the number of alphas.
 */
int cge_oracle_wgcl(int64_t m, const int64_t *e_src, const int64_t *e_dst, const double *eweights,
                    const int64_t *comm, int64_t n_comm_rows, const double *embed, int64_t d,
                    const double *distances, int64_t n_distances, const double *vweights,
                    int64_t n_full, const double *init_vweights, const int64_t *v_to_l,
                    const double *init_embed, int split, int64_t K, int64_t n_sets,
                    const int64_t *pos_i, const int64_t *pos_j, const double *pos_w,
                    const int64_t *neg_i, const int64_t *neg_j, int max_alphas, double *out,
                    cge_oracle_trace *tr) {
    const double epsilon = 0.25, delta = 0.001, AlphaMax = 10.0, AlphaStep = 0.25; /* :34-37 */
    int alpha_div_counter = 5, alpha_auc_counter = 5;                              /* :38 */
    int skip_div = 0, skip_auc = 0;                                                /* :39 */
    trace_init(tr);

    int64_t n = max_i64(e_src, m);                                                 /* :41 */
    {
        int64_t t = max_i64(e_dst, m);
        if (t > n) n = t;
    }
    int landmarks = n_full > 0;                                                    /* :44 */
    if (n_comm_rows != n) return -2;                                               /* :50 */
    int64_t n_parts = max_i64(comm, n);                                            /* :51 */
    int64_t vect_len = n_parts * (n_parts + 1) / 2;                                /* :55 */
    double *vect_C = (double *)calloc((size_t)vect_len, sizeof(double));
    double *vect_B = (double *)calloc((size_t)vect_len, sizeof(double));
    for (int64_t i = 0; i < m; ++i) {                                              /* :59-63 */
        int64_t c1 = comm[e_src[i] - 1], c2 = comm[e_dst[i] - 1];
        int64_t j = c1 < c2 ? c1 : c2, k = c1 < c2 ? c2 : c1;
        vect_C[cge_oracle_idx(n_parts, j, k) - 1] += eweights[i];
    }
    uint8_t *vect_I = (uint8_t *)calloc((size_t)vect_len, 1);                      /* :66-71 */
    {
        int64_t j = 1;
        for (int64_t i = 1; i <= n_parts; ++i) {
            vect_I[j - 1] = 1;
            j += n_parts - i + 1;
        }
    }
    double best_div = INFINITY, best_div_ext = INFINITY, best_div_int = INFINITY;  /* :72-73 */
    double best_auc_err = INFINITY, best_auc = INFINITY;
    double best_alpha = -1.0, best_alpha_auc = -1.0;

    if (n_distances != n) return -3;                                               /* :81 */
    int64_t p_len = n * (n + 1) / 2;
    double lo, hi;
    double tph = now_s();
    double *D = build_D(n, embed, d, distances, &lo, &hi);                         /* :79-93 */
    if (!D) return -4;
    if (tr) { tr->lo = lo; tr->hi = hi; tr->t_phase[0] += now_s() - tph; }

    int64_t adj_n = landmarks ? n_full : n;                                        /* :95-102 */
    double *full_D = NULL;
    if (landmarks) {                                                               /* :104-115 */
        double hf;
        full_D = build_full_D(adj_n, init_embed, d, &hf);
        if (!full_D) return -4;
        if (tr) tr->hi_full = hf;
    }
    double *T = (double *)malloc(sizeof(double) * (size_t)n);                      /* :118 */
    for (int64_t i = 0; i < n; ++i) T[i] = 1.0;
    double *GD = (double *)malloc(sizeof(double) * (size_t)p_len);
    double *P = (double *)malloc(sizeof(double) * (size_t)p_len);
    double *S = (double *)malloc(sizeof(double) * (size_t)n);
    double *pos = (double *)malloc(sizeof(double) * (size_t)(K > 0 ? K : 1));
    double *neg = (double *)malloc(sizeof(double) * (size_t)(K > 0 ? K : 1));

    int a = 0;
    /* alpha in 0.25:0.25:10.001 -> 0.25*a, a = 1..40 (:139) */
    for (a = 1; a <= N_ALPHA && a <= max_alphas; ++a) {
        double alpha = AlphaStep * (double)a;
        (void)AlphaMax;
        tph = now_s();
        for (int64_t k = 0; k < p_len; ++k) GD[k] = pow(1.0 - D[k], alpha);        /* :142-148 */
        if (tr) tr->t_phase[1] += now_s() - tph;
        double diff = 1.0;                                                         /* :150 */
        int it = 0;
        tph = now_s();
        while (diff > delta) {                                                     /* :151-168 */
            for (int64_t i = 0; i < n; ++i) S[i] = 0.0;
            for (int64_t i = 1; i <= n; ++i) {
                int64_t base = cge_oracle_idx(n, i, i) - 1;
                for (int64_t j = i; j <= n; ++j) {
                    double tmp = T[i - 1] * T[j - 1] * GD[base + (j - i)];
                    S[i - 1] += tmp;
                    if (i != j) S[j - 1] += tmp;
                }
            }
            double f = 0.0;
            for (int64_t i = 0; i < n; ++i) {
                double move = epsilon * T[i] * (vweights[i] / S[i] - 1.0);
                T[i] += move;
                double e = fabs(vweights[i] - S[i]);
                if (e > f) f = e;
            }
            diff = f;
            ++it;
        }
        if (tr) { tr->iters[a - 1] = it; tr->n_alpha_run = a; tr->final_diff = diff; tr->t_phase[2] += now_s() - tph; }
        tph = now_s();
        for (int64_t i = 1; i <= n; ++i) {                                         /* :170-176 */
            int64_t base = cge_oracle_idx(n, i, i) - 1;
            for (int64_t j = i; j <= n; ++j)
                P[base + (j - i)] = T[i - 1] * T[j - 1] * GD[base + (j - i)];
        }
        if (tr) tr->t_phase[3] += now_s() - tph;
        if (!skip_auc && K > 0) {                                                  /* :178-224 */
            int64_t off = (n_sets > 1 ? (int64_t)(a - 1) : 0) * K;
            double sw = 0.0, swin = 0.0;
            for (int64_t s = 0; s < K; ++s) {
                int64_t pi = pos_i[off + s], pj = pos_j[off + s];
                int64_t ni = neg_i[off + s], nj = neg_j[off + s];
                if (pi > pj) { int64_t t = pi; pi = pj; pj = t; }
                if (ni > nj) { int64_t t = ni; ni = nj; nj = t; }
                if (landmarks) {                                                   /* :184-199 */
                    double ti = T[v_to_l[pi - 1] - 1] * init_vweights[pi - 1] / vweights[v_to_l[pi - 1] - 1];
                    double tj = T[v_to_l[pj - 1] - 1] * init_vweights[pj - 1] / vweights[v_to_l[pj - 1] - 1];
                    pos[s] = ti * tj * pow(1.0 - full_D[cge_oracle_idx(adj_n, pi, pj) - 1], alpha);
                    ti = T[v_to_l[ni - 1] - 1] * init_vweights[ni - 1] / vweights[v_to_l[ni - 1] - 1];
                    tj = T[v_to_l[nj - 1] - 1] * init_vweights[nj - 1] / vweights[v_to_l[nj - 1] - 1];
                    neg[s] = ti * tj * pow(1.0 - full_D[cge_oracle_idx(adj_n, ni, nj) - 1], alpha);
                } else {                                                           /* :201-210 */
                    pos[s] = P[cge_oracle_idx(n, pi, pj) - 1];
                    neg[s] = P[cge_oracle_idx(n, ni, nj) - 1];
                }
                swin += (pos[s] > neg[s] ? 1.0 : 0.0) * pos_w[off + s];            /* :213 */
                sw += pos_w[off + s];
            }
            double auc = 1.0 - swin / sw;
            if (tr) tr->auc[a - 1] = auc;
            if (auc < best_auc) {                                                  /* :215-223 */
                best_auc = auc;
                best_auc_err = 1.96 * sqrt(auc * (1.0 - auc) / (double)K);
                best_alpha_auc = alpha;
                alpha_auc_counter = 5;
            } else {
                alpha_auc_counter -= 1;
                skip_auc = alpha_auc_counter == 0;
            }
        }
        if (!skip_div) {                                                           /* :226-252 */
            tph = now_s();
            for (int64_t k = 0; k < vect_len; ++k) vect_B[k] = 0.0;
            for (int64_t i = 1; i <= n; ++i) {
                int64_t base = cge_oracle_idx(n, i, i) - 1;
                for (int64_t j = i; j <= n; ++j) {
                    int64_t c1 = comm[i - 1], c2 = comm[j - 1];
                    int64_t k = c1 < c2 ? c1 : c2, l = c1 < c2 ? c2 : c1;
                    vect_B[cge_oracle_idx(n_parts, k, l) - 1] += P[base + (j - i)];
                }
            }
            if (tr) tr->t_phase[4] += now_s() - tph;
            double f, div_int = 0.0, div_ext = 0.0;
            if (!split) {
                f = cge_oracle_js(vect_C, vect_B, NULL, 1, vect_len);
            } else {
                div_int = cge_oracle_js(vect_C, vect_B, vect_I, 1, vect_len);
                div_ext = cge_oracle_js(vect_C, vect_B, vect_I, 0, vect_len);
                f = (div_int + div_ext) / 2.0;
            }
            if (tr) tr->div[a - 1] = f;
            if (f < best_div) {
                best_div = f;
                best_alpha = alpha;
                best_div_ext = !split ? 0.0 : div_ext;
                best_div_int = !split ? 0.0 : div_int;
                alpha_div_counter = 5;
            } else {
                alpha_div_counter -= 1;
                skip_div = alpha_div_counter == 0;
            }
        }
        if (skip_div && (skip_auc || K <= 0)) break;                               /* :253 */
    }
    out[0] = best_alpha; out[1] = best_div; out[2] = best_div_ext; out[3] = best_div_int;
    out[4] = best_alpha_auc; out[5] = best_auc; out[6] = best_auc_err;             /* :256 */
    free(vect_C); free(vect_B); free(vect_I); free(D); free(full_D); free(T);
    free(GD); free(P); free(S); free(pos); free(neg);
    return 0;
}

/*
 * wGCL_directed, divergence.jl:282-561.  Returns 0 and *out_len = 7, or *out_len = 6 with
 * out = [-1,0,0,0,0,0] for the star-graph early exit (:332-334).
 * In exact mode the caller passes pos_i/pos_j from the SECOND positive draw and pos_w from the
 * first one (the overwrite at :510).
 */
int cge_oracle_wgcl_directed(int64_t m, const int64_t *e_src, const int64_t *e_dst,
                             const double *eweights, const int64_t *comm, int64_t n_comm_rows,
                             const double *embed, int64_t d, const double *distances,
                             int64_t n_distances, const double *vweights, int64_t n_full,
                             const double *init_vweights, const int64_t *v_to_l,
                             const double *init_embed, int split, int64_t K, int64_t n_sets,
                             const int64_t *pos_i, const int64_t *pos_j, const double *pos_w,
                             const int64_t *neg_i, const int64_t *neg_j, int max_alphas,
                             double *out, int *out_len, cge_oracle_trace *tr) {
    const double delta = 0.001, AlphaStep = 0.25;                                  /* :288-290 */
    int alpha_div_counter = 5, alpha_auc_counter = 5, skip_div = 0, skip_auc = 0;  /* :291-292 */
    trace_init(tr);
    *out_len = 7;
    int64_t n = max_i64(e_src, m);                                                 /* :294 */
    {
        int64_t t = max_i64(e_dst, m);
        if (t > n) n = t;
    }
    int landmarks = n_full > 0;
    if (n_comm_rows != n) return -2;                                               /* :303 */
    int64_t n_parts = max_i64(comm, n);

    double *degree_in = (double *)calloc((size_t)n, sizeof(double));               /* :308-319 */
    double *degree_out = (double *)calloc((size_t)n, sizeof(double));
    int64_t *star_check = (int64_t *)calloc((size_t)n, sizeof(int64_t));
    for (int64_t i = 0; i < m; ++i) {
        degree_out[e_src[i] - 1] += eweights[i];
        degree_in[e_dst[i] - 1] += eweights[i];
        star_check[e_src[i] - 1] += 1;
        star_check[e_dst[i] - 1] += 1;
    }
    {                                                                              /* :322-334 */
        int has_nm1 = 0, has_2nm1 = 0;
        int64_t sum = 0, cnt2 = 0;
        for (int64_t i = 0; i < n; ++i) {
            if (star_check[i] == n - 1) has_nm1 = 1;
            if (star_check[i] == 2 * (n - 1)) has_2nm1 = 1;
            if (star_check[i] == 2) ++cnt2;
            sum += star_check[i];
        }
        int is_star = 0;
        if (has_nm1 && sum == 2 * (n - 1)) is_star = 1;
        else if (has_2nm1 && cnt2 == n - 1) is_star = 1;
        if (is_star) {
            out[0] = -1.0;
            for (int i = 1; i < 6; ++i) out[i] = 0.0;
            *out_len = 6;
            free(degree_in); free(degree_out); free(star_check);
            return 0;
        }
    }
    int64_t vect_len = n_parts * n_parts;                                          /* :337-345 */
    double *vect_C = (double *)calloc((size_t)vect_len, sizeof(double));
    double *vect_B = (double *)calloc((size_t)vect_len, sizeof(double));
    for (int64_t i = 0; i < m; ++i) {
        int64_t j = comm[e_src[i] - 1], k = comm[e_dst[i] - 1];
        vect_C[(j - 1) * n_parts + k - 1] += eweights[i];
    }
    uint8_t *vect_I = (uint8_t *)calloc((size_t)vect_len, 1);                      /* :348-351 */
    for (int64_t i = 1; i <= vect_len; i += n_parts + 1) vect_I[i - 1] = 1;
    double best_div = INFINITY, best_div_ext = INFINITY, best_div_int = INFINITY;
    double best_auc_err = INFINITY, best_auc = INFINITY;
    double best_alpha = -1.0, best_alpha_auc = -1.0;

    if (n_distances != n) return -3;                                               /* :363 */
    int64_t p_len = n * (n + 1) / 2;
    double lo, hi;
    double *D = build_D(n, embed, d, distances, &lo, &hi);                         /* :359-375 */
    if (!D) return -4;
    if (tr) { tr->lo = lo; tr->hi = hi; }
    int64_t adj_n = landmarks ? n_full : n;
    double *full_D = NULL;
    if (landmarks) {                                                               /* :386-397 */
        double hf;
        full_D = build_full_D(adj_n, init_embed, d, &hf);
        if (!full_D) return -4;
        if (tr) tr->hi_full = hf;
    }
    double *Tin = (double *)malloc(sizeof(double) * (size_t)n);                    /* :399-402 */
    double *Tout = (double *)malloc(sizeof(double) * (size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        Tin[i] = degree_in[i] == 0.0 ? 0.0 : 1.0;
        Tout[i] = degree_out[i] == 0.0 ? 0.0 : 1.0;
    }
    double *GD = (double *)malloc(sizeof(double) * (size_t)p_len);
    double *P = (double *)malloc(sizeof(double) * (size_t)(n * n));
    double *Sin = (double *)malloc(sizeof(double) * (size_t)n);
    double *Sout = (double *)malloc(sizeof(double) * (size_t)n);

    for (int a = 1; a <= N_ALPHA && a <= max_alphas; ++a) {                        /* :423 */
        double alpha = AlphaStep * (double)a;
        for (int64_t k = 0; k < p_len; ++k) GD[k] = pow(1.0 - D[k], alpha);        /* :426-432 */
        double diff = 1.0, epsilon = 0.9;                                          /* :434-435 */
        int it = 0;
        while (diff > delta) {                                                     /* :436-467 */
            for (int64_t i = 0; i < n; ++i) Sin[i] = Sout[i] = 0.0;
            for (int64_t i = 1; i <= n; ++i) {
                int64_t base = cge_oracle_idx(n, i, i) - 1;
                for (int64_t j = i; j <= n; ++j) {
                    double g = GD[base + (j - i)];
                    double tmp1 = Tin[i - 1] * Tout[j - 1] * g;
                    double tmp2 = Tin[j - 1] * Tout[i - 1] * g;
                    Sin[i - 1] += tmp1;
                    Sin[j - 1] += tmp2;
                    Sout[i - 1] += tmp2;
                    Sout[j - 1] += tmp1;
                }
            }
            double f = 0.0;
            for (int64_t i = 0; i < n; ++i) {
                if (degree_in[i] > 0) {
                    Tin[i] += epsilon * Tin[i] * (degree_in[i] / Sin[i] - 1.0);
                    double e = fabs(degree_in[i] - Sin[i]);
                    if (e > f) f = e;
                }
                if (degree_out[i] > 0) {
                    Tout[i] += epsilon * Tout[i] * (degree_out[i] / Sout[i] - 1.0);
                    double e = fabs(degree_out[i] - Sout[i]);
                    if (e > f) f = e;
                }
            }
            if (f > diff) epsilon *= 0.99;                                         /* :462-464 */
            diff = f;
            ++it;
        }
        if (tr) { tr->iters[a - 1] = it; tr->n_alpha_run = a; tr->final_diff = diff; }
        for (int64_t i = 1; i <= n; ++i)                                           /* :470-476 */
            for (int64_t j = 1; j <= n; ++j) {
                int64_t lo_ = i < j ? i : j, hi_ = i < j ? j : i;
                P[n * (i - 1) + j - 1] = Tout[i - 1] * Tin[j - 1] * GD[cge_oracle_idx(n, lo_, hi_) - 1];
            }
        if (!skip_auc && K > 0) {                                                  /* :478-528 */
            int64_t off = (n_sets > 1 ? (int64_t)(a - 1) : 0) * K;
            double sw = 0.0, swin = 0.0;
            for (int64_t s = 0; s < K; ++s) {
                int64_t pi = pos_i[off + s], pj = pos_j[off + s];
                int64_t ni = neg_i[off + s], nj = neg_j[off + s];
                double pv, nv;
                if (landmarks) {                                                   /* :484-501 */
                    int64_t lo_ = pi < pj ? pi : pj, hi_ = pi < pj ? pj : pi;
                    double to = Tout[v_to_l[pi - 1] - 1] * init_vweights[pi - 1] / vweights[v_to_l[pi - 1] - 1];
                    double ti = Tin[v_to_l[pj - 1] - 1] * init_vweights[pj - 1] / vweights[v_to_l[pj - 1] - 1];
                    pv = to * ti * pow(1.0 - full_D[cge_oracle_idx(adj_n, lo_, hi_) - 1], alpha);
                    lo_ = ni < nj ? ni : nj; hi_ = ni < nj ? nj : ni;
                    to = Tout[v_to_l[ni - 1] - 1] * init_vweights[ni - 1] / vweights[v_to_l[ni - 1] - 1];
                    ti = Tin[v_to_l[nj - 1] - 1] * init_vweights[nj - 1] / vweights[v_to_l[nj - 1] - 1];
                    nv = to * ti * pow(1.0 - full_D[cge_oracle_idx(adj_n, lo_, hi_) - 1], alpha);
                } else {                                                           /* :503-513 */
                    pv = P[n * (pi - 1) + pj - 1];
                    nv = P[n * (ni - 1) + nj - 1];
                }
                swin += (pv > nv ? 1.0 : 0.0) * pos_w[off + s];                    /* :517 */
                sw += pos_w[off + s];
            }
            double auc = 1.0 - swin / sw;
            if (tr) tr->auc[a - 1] = auc;
            if (auc < best_auc) {                                                  /* :519-527 */
                best_auc = auc;
                best_auc_err = 1.96 * sqrt(auc * (1.0 - auc) / (double)K);
                best_alpha_auc = alpha;
                alpha_auc_counter = 5;
            } else {
                alpha_auc_counter -= 1;
                skip_auc = alpha_auc_counter == 0;
            }
        }
        if (!skip_div) {                                                           /* :530-556 */
            for (int64_t k = 0; k < vect_len; ++k) vect_B[k] = 0.0;
            for (int64_t i = 1; i <= n; ++i)
                for (int64_t j = 1; j <= n; ++j)
                    vect_B[(comm[i - 1] - 1) * n_parts + comm[j - 1] - 1] += P[n * (i - 1) + j - 1];
            double f, div_int = 0.0, div_ext = 0.0;
            if (!split) {
                f = cge_oracle_js(vect_C, vect_B, NULL, 1, vect_len);
            } else {
                div_int = cge_oracle_js(vect_C, vect_B, vect_I, 1, vect_len);
                div_ext = cge_oracle_js(vect_C, vect_B, vect_I, 0, vect_len);
                f = (div_int + div_ext) / 2.0;
            }
            if (tr) tr->div[a - 1] = f;
            if (f < best_div) {
                best_div = f;
                best_alpha = alpha;
                best_div_ext = !split ? 0.0 : div_ext;
                best_div_int = !split ? 0.0 : div_int;
                alpha_div_counter = 5;
            } else {
                alpha_div_counter -= 1;
                skip_div = alpha_div_counter == 0;
            }
        }
        if (skip_div && (skip_auc || K <= 0)) break;                               /* :557 */
    }
    out[0] = best_alpha; out[1] = best_div; out[2] = best_div_ext; out[3] = best_div_int;
    out[4] = best_alpha_auc; out[5] = best_auc; out[6] = best_auc_err;             /* :560 */
    free(degree_in); free(degree_out); free(star_check); free(vect_C); free(vect_B);
    free(vect_I); free(D); free(full_D); free(Tin); free(Tout); free(GD); free(P);
    free(Sin); free(Sout);
    return 0;
}
