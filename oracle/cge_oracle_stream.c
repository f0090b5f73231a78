/*
 * cge_oracle_stream.c -- TEST / MEASUREMENT INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * SURVEY.md section 7 "(ii) streaming-oracle agreement at 50k" and VERDICT r01 item 4: the exact-mode
 * algorithm of wGCL (src/divergence.jl:27-257) and wGCL_directed (:282-561) WITHOUT any O(n^2) array,
 * on all host cores, so that BASELINE configs 3 (50 000 vertices, directed, d = 64) and 4 (200 000
 * vertices, d = 128) -- where the reference's packed D, GD and P do not fit any host -- still have a
 * CPU answer to hold the GPU to.  Only tests/ and tests/golden/make_size_goldens.py load it.
 *
 * What is kept from the reference (line by line in cge_oracle.c, which this file is tested against):
 *   - dist (auxilary.jl:14-20): sum over the dimensions in ascending order of (a - b)^2, multiply and
 *     add rounded separately, then sqrt -- bit-identical to cge_oracle.c for every pair;
 *   - D = (D - lo)/(hi - lo) with lo = 0 (the zero diagonal of exact mode, divergence.jl:85-93);
 *   - the Jacobi pass, T update, residual, epsilon decay, patience counters, sampled-pair local
 *     score and JS exactly as cge_oracle.c (divergence.jl:139-254 / 423-558).
 * What differs, all documented deviations of the order of an ulp:
 *   - (1 - D)^alpha is evaluated as q^m with q = sqrt(sqrt(1 - D)), m = 4 alpha and binary powering
 *     (the reference calls pow once per alpha and pair and stores GD; a streaming pass would call it
 *     once per PASS and pair).  |q^m - pow| <= ~m ulp; tests/test_oracle_stream.py holds the result
 *     to cge_oracle.c: identical pass counts and best alphas, scores within 1e-11;
 *   - the degree sums are accumulated per thread over blocks of 64 x 64 pairs and added in thread
 *     order (deterministic for a given thread count, ~1e-13 relative from the sequential order);
 *   - P is not materialised (B and the sampled pairs use T_i T_j GD_ij directly).
 * When 8 n(n+1)/2 bytes fit in `mem_budget` the per-pair q is kept after the first sweep ("cached":
 * same values, same order of additions, only faster); otherwise every sweep recomputes it.
 */
#define _POSIX_C_SOURCE 200809L
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#define N_ALPHA 40
#define BS 64 /* block side */

typedef struct {
    int32_t n_alpha_run;
    int32_t iters[N_ALPHA];
    double div[N_ALPHA];
    double auc[N_ALPHA];
    double lo, hi;
    int32_t threads, cached;
    double seconds;
} cge_oracle_stream_trace;

double cge_oracle_js(const double *vC, const double *vB, const uint8_t *vI, int internal,
                     int64_t len); /* cge_oracle.c (auxilary.jl:34-52) */

static inline double powm(double q, int m) {
    double r = 1.0, b = q;
    while (m) {
        if (m & 1) r *= b;
        b *= b;
        m >>= 1;
    }
    return r;
}

typedef struct {
    int nt, directed;
    pthread_barrier_t bar;
    int64_t n, d, nb, K, n_sets, k;
    const double *embed; /* n x d row-major */
    double *xt;          /* per block of 64 rows: [k][64] transposed copy (vector loads over j) */
    double *qcache;      /* NULL, or per block pair (bi <= bj) a 64 x 64 tile of q */
    int64_t *tile_off;   /* [nb] offset (in tiles) of block row bi in the upper-triangular sequence */
    const int64_t *comm;
    double hi;
    int m, phase, have_cache; /* phase: 0 extrema, 1 pass, 2 B */
    double *Ta, *Tb;          /* T (or Tin), Tout */
    double *Sp;               /* [nt][2][n] private degree sums */
    double *Bp;               /* [nt][k*k] private community mass */
    double *hip;              /* [nt] private maxima */
    int quit;
} sctx;

typedef struct {
    sctx *c;
    int t;
} sarg;

/* squared distances of the 64 rows of block bi against the 64 columns of block bj:
 * acc[jj] += (x_i[k] - x_j[k])^2, k ascending, mul and add rounded separately (o_dist).
 * The (a - b)^2 sums stay in registers: 2 rows x 32 columns per sweep over the dimensions (the
 * compiler turns the fixed-trip inner loops into vector code; every pair still adds its dimensions
 * in ascending order, one rounding per multiply and per add). */
#define JB 32
__attribute__((target_clones("avx512f", "avx2", "default"))) static void
dist_block(const sctx *c, int64_t bi, int64_t bj, double *d2 /* [64][64] */) {
    const int64_t n = c->n, d = c->d;
    const double *xt = c->xt + (size_t)bj * (size_t)d * BS;
    for (int ii = 0; ii < BS; ii += 2) {
        const int64_t i0 = bi * BS + ii, i1 = i0 + 1;
        if (i0 >= n) {
            for (int jj = 0; jj < 2 * BS; ++jj) d2[ii * BS + jj] = 0.0;
            continue;
        }
        const double *x0 = c->embed + (size_t)i0 * (size_t)d;
        const double *x1 = c->embed + (size_t)(i1 < n ? i1 : i0) * (size_t)d;
        for (int jb = 0; jb < BS; jb += JB) {
            double a0[JB], a1[JB];
            for (int u = 0; u < JB; ++u) a0[u] = a1[u] = 0.0;
            for (int64_t k = 0; k < d; ++k) {
                const double p0 = x0[k], p1 = x1[k];
                const double *col = xt + (size_t)k * BS + jb;
                for (int u = 0; u < JB; ++u) {
                    const double t0 = p0 - col[u], t1 = p1 - col[u];
                    a0[u] += t0 * t0;
                    a1[u] += t1 * t1;
                }
            }
            for (int u = 0; u < JB; ++u) {
                d2[ii * BS + jb + u] = a0[u];
                d2[(ii + 1) * BS + jb + u] = i1 < n ? a1[u] : 0.0;
            }
        }
    }
}

/* q = ((1 - D/hi))^(1/4) of a block pair; pairs outside i < j (and pads) are left unspecified
 * (sqrt and divide are correctly rounded in scalar and vector form alike) */
__attribute__((target_clones("avx512f", "avx2", "default"))) static void q_block(const sctx *c, int64_t bi, int64_t bj, double *buf) {
    const double hi = c->hi;
    dist_block(c, bi, bj, buf);
    for (int e = 0; e < BS * BS; ++e) {
        const double D = sqrt(buf[e]);
        const double Dn = (D - 0.0) / (hi - 0.0); /* divergence.jl:93 with lo = 0 */
        buf[e] = sqrt(sqrt(1.0 - Dn));
    }
}

static void *worker(void *argp) {
    sctx *c = ((sarg *)argp)->c;
    const int t = ((sarg *)argp)->t, nt = c->nt;
    const int64_t n = c->n, nb = c->nb, k = c->k;
    double *buf = (double *)aligned_alloc(64, sizeof(double) * BS * BS);
    for (;;) {
        pthread_barrier_wait(&c->bar); /* work published */
        if (c->quit) break;
        const int phase = c->phase, m = c->m;
        double *sa = c->Sp + (size_t)t * 2 * (size_t)n, *sb = sa + n;
        double *bp = c->Bp + (size_t)t * (size_t)(k * k);
        if (phase == 1) memset(sa, 0, sizeof(double) * 2 * (size_t)n);
        if (phase == 2) memset(bp, 0, sizeof(double) * (size_t)(k * k));
        double hi = 0.0;
        /* block pairs dealt cyclically by their index in the upper-triangular sequence */
        for (int64_t bi = 0; bi < nb; ++bi)
            for (int64_t bj = bi; bj < nb; ++bj) {
                const int64_t tile = c->tile_off[bi] + (bj - bi);
                if (tile % nt != t) continue;
                double *q = buf;
                if (phase == 0) {
                    dist_block(c, bi, bj, buf);
                } else if (c->qcache) {
                    q = c->qcache + (size_t)tile * BS * BS;
                    if (!c->have_cache) q_block(c, bi, bj, q);
                } else {
                    q_block(c, bi, bj, buf);
                }
                for (int ii = 0; ii < BS; ++ii) {
                    const int64_t i = bi * BS + ii;
                    if (i >= n) break;
                    const double *row = q + ii * BS;
                    for (int jj = (bi == bj ? ii + 1 : 0); jj < BS; ++jj) {
                        const int64_t j = bj * BS + jj;
                        if (j >= n) break;
                        if (phase == 0) {
                            const double v = sqrt(row[jj]);
                            if (v > hi) hi = v;
                        } else if (phase == 1) {
                            const double g = powm(row[jj], m);
                            if (!c->directed) {
                                const double tmp = c->Ta[i] * c->Ta[j] * g; /* :155 */
                                sa[i] += tmp;
                                sa[j] += tmp;
                            } else {
                                const double tmp1 = c->Ta[i] * c->Tb[j] * g; /* :442-447 */
                                const double tmp2 = c->Ta[j] * c->Tb[i] * g;
                                sa[i] += tmp1;
                                sa[j] += tmp2;
                                sb[i] += tmp2;
                                sb[j] += tmp1;
                            }
                        } else {
                            const double g = powm(row[jj], m);
                            const int64_t ci = c->comm[i] - 1, cj = c->comm[j] - 1;
                            if (!c->directed) { /* :228-234: bin (min c, max c) */
                                const int64_t a = ci < cj ? ci : cj, b = ci < cj ? cj : ci;
                                bp[a * k + b] += c->Ta[i] * c->Ta[j] * g;
                            } else { /* :532-538: P[i][j] = Tout_i Tin_j GD, all ordered pairs */
                                bp[ci * k + cj] += c->Tb[i] * c->Ta[j] * g;
                                bp[cj * k + ci] += c->Tb[j] * c->Ta[i] * g;
                            }
                        }
                    }
                }
            }
        c->hip[t] = hi;
        pthread_barrier_wait(&c->bar); /* work done */
    }
    free(buf);
    return NULL;
}

static void run_phase(sctx *c, int phase, int m) {
    c->phase = phase;
    c->m = m;
    pthread_barrier_wait(&c->bar);
    pthread_barrier_wait(&c->bar);
    if (phase != 0 && c->qcache) c->have_cache = 1;
}

static double pair_q(const sctx *c, int64_t i, int64_t j) { /* 0-based, i != j */
    const double *a = c->embed + (size_t)i * c->d, *b = c->embed + (size_t)j * c->d;
    double s = 0.0;
    for (int64_t k = 0; k < c->d; ++k) {
        const double t = a[k] - b[k];
        s += t * t;
    }
    const double Dn = (sqrt(s) - 0.0) / (c->hi - 0.0);
    return sqrt(sqrt(1.0 - Dn));
}

/*
 * Exact mode, no --split-global.  Arrays as in cge_oracle.c (1-based ids, row-major embed);
 * vweights is used by the undirected model only.  n_threads = 0: all cores.  Returns 0, or -1 on a
 * bad argument / allocation failure, or 1 for the directed star-graph exit (not scored here).
 */
int cge_oracle_stream(int directed, int64_t m_edges, const int64_t *e_src, const int64_t *e_dst,
                      const double *eweights, int64_t n, const int64_t *comm, const double *embed,
                      int64_t d, const double *vweights, int64_t K, int64_t n_sets,
                      const int64_t *pos_i, const int64_t *pos_j, const double *pos_w,
                      const int64_t *neg_i, const int64_t *neg_j, int max_alphas, int n_threads,
                      int64_t mem_budget, double *out, cge_oracle_stream_trace *tr) {
    struct timespec ts0, ts1;
    clock_gettime(CLOCK_MONOTONIC, &ts0);
    const double delta = 0.001;
    if (n < 2 || d < 1 || max_alphas < 1) return -1;
    int nt = n_threads > 0 ? n_threads : (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (nt < 1) nt = 1;
    int64_t k = 0;
    for (int64_t i = 0; i < n; ++i)
        if (comm[i] > k) k = comm[i];
    sctx c;
    memset(&c, 0, sizeof(c));
    c.nt = nt;
    c.directed = directed;
    c.n = n;
    c.d = d;
    c.nb = (n + BS - 1) / BS;
    c.K = K;
    c.n_sets = n_sets;
    c.k = k;
    c.embed = embed;
    c.comm = comm;
    const int64_t n_tiles = c.nb * (c.nb + 1) / 2;
    c.xt = (double *)calloc((size_t)c.nb * (size_t)d * BS, sizeof(double));
    c.tile_off = (int64_t *)malloc(sizeof(int64_t) * (size_t)c.nb);
    c.Ta = (double *)malloc(sizeof(double) * (size_t)n);
    c.Tb = (double *)malloc(sizeof(double) * (size_t)n);
    c.Sp = (double *)malloc(sizeof(double) * (size_t)nt * 2 * (size_t)n);
    c.Bp = (double *)malloc(sizeof(double) * (size_t)nt * (size_t)(k * k));
    c.hip = (double *)malloc(sizeof(double) * (size_t)nt);
    double *Sa = (double *)malloc(sizeof(double) * (size_t)n), *Sb = (double *)malloc(sizeof(double) * (size_t)n);
    double *wa = (double *)calloc((size_t)n, sizeof(double)), *wb = (double *)calloc((size_t)n, sizeof(double));
    double *vC = (double *)calloc((size_t)(k * k), sizeof(double)), *vB = (double *)calloc((size_t)(k * k), sizeof(double));
    double *binC = (double *)malloc(sizeof(double) * (size_t)(k * k)), *binB = (double *)malloc(sizeof(double) * (size_t)(k * k));
    if (!c.xt || !c.tile_off || !c.Ta || !c.Tb || !c.Sp || !c.Bp || !c.hip || !Sa || !Sb || !wa || !wb || !vC || !vB || !binC || !binB)
        return -1;
    if ((int64_t)sizeof(double) * BS * BS * n_tiles <= mem_budget)
        c.qcache = (double *)malloc(sizeof(double) * BS * BS * (size_t)n_tiles);
    for (int64_t bi = 0, off = 0; bi < c.nb; ++bi) {
        c.tile_off[bi] = off;
        off += c.nb - bi;
    }
    for (int64_t i = 0; i < n; ++i)
        for (int64_t kk = 0; kk < d; ++kk)
            c.xt[((size_t)(i / BS) * (size_t)d + (size_t)kk) * BS + (size_t)(i % BS)] = embed[(size_t)i * d + kk];
    /* degrees / targets, C (divergence.jl:55-63 / 308-319, 337-345) */
    int64_t *star = (int64_t *)calloc((size_t)n, sizeof(int64_t));
    for (int64_t e = 0; e < m_edges; ++e) {
        const int64_t u = e_src[e] - 1, v = e_dst[e] - 1, cu = comm[u] - 1, cv = comm[v] - 1;
        if (directed) {
            wb[u] += eweights[e]; /* degree_out */
            wa[v] += eweights[e]; /* degree_in */
            star[u] += 1;
            star[v] += 1;
            vC[cu * k + cv] += eweights[e];
        } else {
            vC[(cu < cv ? cu : cv) * k + (cu < cv ? cv : cu)] += eweights[e];
        }
    }
    if (directed) { /* :322-334 */
        int has_nm1 = 0, has_2nm1 = 0;
        int64_t sum = 0, cnt2 = 0;
        for (int64_t i = 0; i < n; ++i) {
            has_nm1 |= star[i] == n - 1;
            has_2nm1 |= star[i] == 2 * (n - 1);
            cnt2 += star[i] == 2;
            sum += star[i];
        }
        if ((has_nm1 && sum == 2 * (n - 1)) || (has_2nm1 && cnt2 == n - 1)) return 1;
    } else {
        for (int64_t i = 0; i < n; ++i) wa[i] = vweights[i];
    }
    free(star);
    for (int64_t i = 0; i < n; ++i) { /* :118 / :399-402 */
        c.Ta[i] = directed ? (wa[i] == 0.0 ? 0.0 : 1.0) : 1.0;
        c.Tb[i] = directed ? (wb[i] == 0.0 ? 0.0 : 1.0) : 0.0;
    }
    /* bins that enter JS, in the reference's order: packed upper triangle / full k x k */
    int64_t n_bins = 0;
    int64_t *bins = (int64_t *)malloc(sizeof(int64_t) * (size_t)(k * k));
    for (int64_t a = 0; a < k; ++a)
        for (int64_t b = directed ? 0 : a; b < k; ++b) bins[n_bins++] = a * k + b;
    for (int64_t b = 0; b < n_bins; ++b) binC[b] = vC[bins[b]];

    pthread_barrier_init(&c.bar, NULL, (unsigned)nt + 1);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nt);
    sarg *args = (sarg *)malloc(sizeof(sarg) * (size_t)nt);
    for (int t = 0; t < nt; ++t) {
        args[t].c = &c;
        args[t].t = t;
        pthread_create(&th[t], NULL, worker, &args[t]);
    }
    /* extrema (divergence.jl:92): lo = 0 (diagonal), hi = largest distance */
    run_phase(&c, 0, 0);
    c.hi = 0.0;
    for (int t = 0; t < nt; ++t)
        if (c.hip[t] > c.hi) c.hi = c.hip[t];
    if (tr) {
        memset(tr, 0, sizeof(*tr));
        for (int a = 0; a < N_ALPHA; ++a) tr->div[a] = tr->auc[a] = NAN;
        tr->lo = 0.0;
        tr->hi = c.hi;
        tr->threads = nt;
        tr->cached = c.qcache != NULL;
    }
    /* q of the sampled pairs, once */
    const int64_t SK = K > 0 ? K * (n_sets > 1 ? n_sets : 1) : 0;
    double *pq = (double *)malloc(sizeof(double) * (size_t)(SK + 1)), *nq = (double *)malloc(sizeof(double) * (size_t)(SK + 1));
    for (int64_t s = 0; s < SK; ++s) {
        pq[s] = pos_i[s] == pos_j[s] ? 1.0 : pair_q(&c, pos_i[s] - 1, pos_j[s] - 1);
        nq[s] = neg_i[s] == neg_j[s] ? 1.0 : pair_q(&c, neg_i[s] - 1, neg_j[s] - 1);
    }
    int alpha_div_counter = 5, alpha_auc_counter = 5, skip_div = 0, skip_auc = K <= 0;
    double best_div = INFINITY, best_auc = INFINITY, best_auc_err = INFINITY, best_alpha = -1.0,
           best_alpha_auc = -1.0;
    for (int a = 1; a <= N_ALPHA && a <= max_alphas; ++a) {
        const double alpha = 0.25 * (double)a;
        double diff = 1.0, epsilon = directed ? 0.9 : 0.25;
        int it = 0;
        while (diff > delta) { /* :150-168 / :436-467 */
            run_phase(&c, 1, a);
            for (int64_t i = 0; i < n; ++i) {
                double s1 = 0.0, s2 = 0.0;
                for (int t = 0; t < nt; ++t) {
                    s1 += c.Sp[(size_t)t * 2 * (size_t)n + (size_t)i];
                    s2 += c.Sp[(size_t)t * 2 * (size_t)n + (size_t)n + (size_t)i];
                }
                Sa[i] = s1;
                Sb[i] = s2;
            }
            double f = 0.0;
            if (!directed) {
                for (int64_t i = 0; i < n; ++i) { /* diagonal: T_i T_i GD_ii = T_i^2 (GD_ii = 1), :153-159 */
                    Sa[i] += c.Ta[i] * c.Ta[i] * 1.0;
                    const double e = fabs(wa[i] - Sa[i]);
                    if (e > f) f = e;
                }
                for (int64_t i = 0; i < n; ++i) c.Ta[i] += epsilon * c.Ta[i] * (wa[i] / Sa[i] - 1.0);
            } else {
                for (int64_t i = 0; i < n; ++i) { /* i == j adds tmp1 = tmp2 to both sums twice, :442-447 */
                    const double tmp = c.Ta[i] * c.Tb[i] * 1.0;
                    Sa[i] += tmp;
                    Sa[i] += tmp;
                    Sb[i] += tmp;
                    Sb[i] += tmp;
                }
                for (int64_t i = 0; i < n; ++i) {
                    if (wa[i] > 0) {
                        c.Ta[i] += epsilon * c.Ta[i] * (wa[i] / Sa[i] - 1.0);
                        const double e = fabs(wa[i] - Sa[i]);
                        if (e > f) f = e;
                    }
                    if (wb[i] > 0) {
                        c.Tb[i] += epsilon * c.Tb[i] * (wb[i] / Sb[i] - 1.0);
                        const double e = fabs(wb[i] - Sb[i]);
                        if (e > f) f = e;
                    }
                }
                if (f > diff) epsilon *= 0.99;
            }
            diff = f;
            ++it;
        }
        if (tr) {
            tr->iters[a - 1] = it;
            tr->n_alpha_run = a;
        }
        if (!skip_auc) { /* :178-224 / :478-528 */
            const int64_t off = (n_sets > 1 ? (int64_t)(a - 1) : 0) * K;
            double sw = 0.0, swin = 0.0;
            for (int64_t s = 0; s < K; ++s) {
                const int64_t pi = pos_i[off + s] - 1, pj = pos_j[off + s] - 1;
                const int64_t ni = neg_i[off + s] - 1, nj = neg_j[off + s] - 1;
                double pv, nv;
                if (!directed) {
                    pv = c.Ta[pi] * c.Ta[pj] * powm(pq[off + s], a);
                    nv = c.Ta[ni] * c.Ta[nj] * powm(nq[off + s], a);
                } else { /* P[i][j] = Tout_i Tin_j GD */
                    pv = c.Tb[pi] * c.Ta[pj] * powm(pq[off + s], a);
                    nv = c.Tb[ni] * c.Ta[nj] * powm(nq[off + s], a);
                }
                swin += (pv > nv ? 1.0 : 0.0) * pos_w[off + s];
                sw += pos_w[off + s];
            }
            const double auc = 1.0 - swin / sw;
            if (tr) tr->auc[a - 1] = auc;
            if (auc < best_auc) {
                best_auc = auc;
                best_auc_err = 1.96 * sqrt(auc * (1.0 - auc) / (double)K);
                best_alpha_auc = alpha;
                alpha_auc_counter = 5;
            } else {
                alpha_auc_counter -= 1;
                skip_auc = alpha_auc_counter == 0;
            }
        }
        if (!skip_div) { /* :226-252 / :530-556 */
            run_phase(&c, 2, a);
            for (int64_t b = 0; b < k * k; ++b) {
                double s = 0.0;
                for (int t = 0; t < nt; ++t) s += c.Bp[(size_t)t * (size_t)(k * k) + (size_t)b];
                vB[b] = s;
            }
            for (int64_t i = 0; i < n; ++i) { /* the diagonal pairs (i, i), once (:229-233 / :533-537) */
                const int64_t ci = comm[i] - 1;
                vB[ci * k + ci] += directed ? c.Tb[i] * c.Ta[i] * 1.0 : c.Ta[i] * c.Ta[i] * 1.0;
            }
            for (int64_t b = 0; b < n_bins; ++b) binB[b] = vB[bins[b]];
            const double f = cge_oracle_js(binC, binB, NULL, 1, n_bins);
            if (tr) tr->div[a - 1] = f;
            if (f < best_div) {
                best_div = f;
                best_alpha = alpha;
                alpha_div_counter = 5;
            } else {
                alpha_div_counter -= 1;
                skip_div = alpha_div_counter == 0;
            }
        }
        if (skip_div && skip_auc) break;
    }
    c.quit = 1;
    pthread_barrier_wait(&c.bar);
    for (int t = 0; t < nt; ++t) pthread_join(th[t], NULL);
    pthread_barrier_destroy(&c.bar);
    out[0] = best_alpha; out[1] = best_div; out[2] = 0.0; out[3] = 0.0;
    out[4] = best_alpha_auc; out[5] = best_auc; out[6] = best_auc_err;
    clock_gettime(CLOCK_MONOTONIC, &ts1);
    if (tr) tr->seconds = (double)(ts1.tv_sec - ts0.tv_sec) + 1e-9 * (double)(ts1.tv_nsec - ts0.tv_nsec);
    free(c.xt); free(c.tile_off); free(c.Ta); free(c.Tb); free(c.Sp); free(c.Bp); free(c.hip);
    free(c.qcache); free(Sa); free(Sb); free(wa); free(wb); free(vC); free(vB); free(binC); free(binB);
    free(bins); free(pq); free(nq); free(th); free(args);
    return 0;
}
