/*
 * cge_oracle_mt.c -- TEST / MEASUREMENT INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * SURVEY.md section 8(d), "CPU baseline beside it", item (ii): the reference's exact-mode
 * undirected algorithm (src/divergence.jl:27-257, restated loop by loop in cge_oracle.c) with its
 * O(n^2) loops spread over all host cores (pthreads: the image's gcc wrapper lacks the OpenMP spec
 * file).  The reference itself is single-threaded (no @threads / @spawn / Distributed in src/), so cge_oracle.c stays the faithful baseline; this
 * file answers "what would a parallel CPU implementation of the same algorithm do on this box?"
 * and is reported next to it by bench.py (cpu_baseline.parallel_port).  Only bench.py and
 * tests/test_oracle_mt.py load it.
 *
 * Same arithmetic per pair as cge_oracle.c; differences, all in the order of additions:
 *   - rows are dealt to threads cyclically; every thread accumulates the degree sums S (and the
 *     community matrix B) privately and the partial vectors are added in thread order, so the
 *     result is deterministic for a given thread count and differs from the sequential order by
 *     ~1e-13 relative (tests/test_oracle_mt.py: pass counts identical, scores within 1e-11);
 *   - P (divergence.jl:170-176) is not materialised: B and the sampled pairs use T_i*T_j*GD_ij
 *     directly -- one O(n^2) sweep per alpha less than the reference.
 * Exact mode only (no landmarks, no --split-global): the shapes bench.py measures.
 */
#define _POSIX_C_SOURCE 200809L /* pthread_barrier_t under -std=c11 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define N_ALPHA 40

typedef struct {
    int32_t n_alpha_run;
    int32_t iters[N_ALPHA];
    double div[N_ALPHA];
    double auc[N_ALPHA];
    double lo, hi;
    int32_t threads;
} cge_oracle_mt_trace;

double cge_oracle_js(const double *vC, const double *vB, const uint8_t *vI, int internal,
                     int64_t len); /* cge_oracle.c (auxilary.jl:34-52) */

static inline int64_t row_base(int64_t n, int64_t i) { /* 0-based offset of (i, i), i 1-based */
    return n * (i - 1) - (i - 1) * (i - 2) / 2;
}

static inline int64_t tri(int64_t k, int64_t a, int64_t b) { /* idx(k, min, max) - 1 */
    int64_t i = a < b ? a : b, j = a < b ? b : a;
    return k * (i - 1) - (i - 1) * (i - 2) / 2 + j - i;
}

/* Numerics experiment for the recompute regime's planned row-norm / dot form:
 * 0 = difference form (the reference, default); 1 = centred embedding, d^2 = n_i + n_j - 2 x_i.x_j
 * with FMA accumulation, pairs under cancellation (d^2 < 2^-13 (n_i + n_j)) redone in the
 * difference form, extrema taken from the same arithmetic.  tests/test_oracle_mt.py compares. */
static int g_dist_form = 0;
void cge_oracle_mt_set_dist_form(int form) { g_dist_form = form; }

int cge_oracle_mt_threads(void) {
    long c = sysconf(_SC_NPROCESSORS_ONLN);
    return c > 0 ? (int)c : 1;
}

/* everything the threads share; thread 0 runs the serial steps between barriers */
typedef struct {
    int nt;
    pthread_barrier_t bar;
    int64_t m, n, d, K, n_sets, n_parts, vect_len, p_len;
    const int64_t *e_src, *e_dst, *comm, *pos_i, *pos_j, *neg_i, *neg_j;
    const double *eweights, *embed, *vweights, *pos_w;
    int max_alphas;
    double *vect_C, *vect_B, *D, *GD, *T, *S, *Sp, *Bp, *red;
    double *cen, *nrm; /* dist form 1: centred embedding and squared row norms */
    double hi, alpha, diff;
    int stop, do_div;
    double *out;
    cge_oracle_mt_trace *tr;
} mt_ctx;

typedef struct {
    mt_ctx *c;
    int t;
} mt_arg;

static void *mt_worker(void *argp) {
    mt_ctx *c = ((mt_arg *)argp)->c;
    const int t = ((mt_arg *)argp)->t, nt = c->nt;
    const int64_t n = c->n, d = c->d, p_len = c->p_len, vect_len = c->vect_len;
    const double epsilon = 0.25, delta = 0.001, AlphaStep = 0.25;
    const int64_t l0 = p_len * t / nt, l1 = p_len * (t + 1) / nt;     /* contiguous share of the packed array */
    const int64_t v0 = n * t / nt, v1 = n * (t + 1) / nt;             /* contiguous share of the vertices */
    double *s = c->Sp + (size_t)t * (size_t)n, *b = c->Bp + (size_t)t * (size_t)vect_len;

    /* D build + extrema + normalisation, :79-93 (diagonal 0 in exact mode); rows dealt cyclically */
    double hi = 0.0;
    for (int64_t i = 1 + t; i <= n; i += nt) {
        const int64_t base = row_base(n, i);
        c->D[base] = 0.0;
        for (int64_t j = i + 1; j <= n; ++j) {
            double acc = 0.0;
            int diff_form = c->cen == NULL;
            if (!diff_form) {
                double g = 0.0;
                for (int64_t k = 0; k < d; ++k) g = fma(c->cen[(i - 1) * d + k], c->cen[(j - 1) * d + k], g);
                const double nn = c->nrm[i - 1] + c->nrm[j - 1];
                acc = fma(-2.0, g, nn);
                if (acc < nn * 0x1.0p-13) diff_form = 1; /* cancellation: <= 40 bits left */
            }
            if (diff_form) {
                acc = 0.0;
                for (int64_t k = 0; k < d; ++k) {
                    const double df = c->embed[(i - 1) * d + k] - c->embed[(j - 1) * d + k];
                    acc += df * df;
                }
            }
            const double v = sqrt(acc);
            c->D[base + (j - i)] = v;
            if (v > hi) hi = v;
        }
    }
    c->red[t] = hi;
    pthread_barrier_wait(&c->bar);
    if (t == 0) {
        double h = 0.0;
        for (int k = 0; k < nt; ++k)
            if (c->red[k] > h) h = c->red[k];
        c->hi = h;
        if (c->tr) {
            c->tr->lo = 0.0;
            c->tr->hi = h;
        }
        for (int64_t i = 0; i < n; ++i) c->T[i] = 1.0; /* :118, warm-started across alpha */
    }
    pthread_barrier_wait(&c->bar);
    for (int64_t l = l0; l < l1; ++l) c->D[l] = (c->D[l] - 0.0) / (c->hi - 0.0);

    int alpha_div_counter = 5, alpha_auc_counter = 5, skip_div = 0, skip_auc = c->K <= 0; /* thread 0 only */
    double best_div = INFINITY, best_auc = INFINITY, best_auc_err = INFINITY;
    double best_alpha = -1.0, best_alpha_auc = -1.0;
    for (int a = 1; a <= N_ALPHA && a <= c->max_alphas; ++a) {
        const double alpha = AlphaStep * (double)a;
        pthread_barrier_wait(&c->bar); /* D normalised / previous alpha done with GD */
        for (int64_t l = l0; l < l1; ++l) c->GD[l] = pow(1.0 - c->D[l], alpha); /* :142-148 */
        int it = 0;
        while (1) { /* :151-168 */
            pthread_barrier_wait(&c->bar); /* GD and T ready */
            memset(s, 0, sizeof(double) * (size_t)n);
            for (int64_t i = 1 + t; i <= n; i += nt) {
                const double *g = c->GD + row_base(n, i);
                const double ti = c->T[i - 1];
                double si = ti * ti * g[0];
                for (int64_t j = i + 1; j <= n; ++j) {
                    const double tmp = ti * c->T[j - 1] * g[j - i];
                    si += tmp;
                    s[j - 1] += tmp;
                }
                s[i - 1] += si;
            }
            pthread_barrier_wait(&c->bar); /* all partial S complete, nobody reads T any more */
            double f = 0.0;
            for (int64_t i = v0; i < v1; ++i) {
                double acc = 0.0;
                for (int k = 0; k < nt; ++k) acc += c->Sp[(size_t)k * (size_t)n + (size_t)i];
                c->S[i] = acc;
                const double e = fabs(c->vweights[i] - acc);
                if (e > f) f = e;
                c->T[i] += epsilon * c->T[i] * (c->vweights[i] / acc - 1.0);
            }
            c->red[t] = f;
            pthread_barrier_wait(&c->bar);
            double diff = 0.0; /* every thread derives the same residual */
            for (int k = 0; k < nt; ++k)
                if (c->red[k] > diff) diff = c->red[k];
            ++it;
            if (!(diff > delta)) break;
        }
        if (t == 0) {
            if (c->tr) {
                c->tr->iters[a - 1] = it;
                c->tr->n_alpha_run = a;
            }
            if (!skip_auc) { /* :178-224, exact mode: P at the sampled pairs */
                const int64_t K = c->K, off = (c->n_sets > 1 ? (int64_t)(a - 1) : 0) * K;
                double sw = 0.0, swin = 0.0;
                for (int64_t q = 0; q < K; ++q) {
                    int64_t pi = c->pos_i[off + q], pj = c->pos_j[off + q];
                    int64_t ni = c->neg_i[off + q], nj = c->neg_j[off + q];
                    if (pi > pj) { int64_t x = pi; pi = pj; pj = x; }
                    if (ni > nj) { int64_t x = ni; ni = nj; nj = x; }
                    const double pp = c->T[pi - 1] * c->T[pj - 1] * c->GD[row_base(n, pi) + (pj - pi)];
                    const double nn = c->T[ni - 1] * c->T[nj - 1] * c->GD[row_base(n, ni) + (nj - ni)];
                    swin += (pp > nn ? 1.0 : 0.0) * c->pos_w[off + q];
                    sw += c->pos_w[off + q];
                }
                const double auc = 1.0 - swin / sw;
                if (c->tr) c->tr->auc[a - 1] = auc;
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
            c->do_div = !skip_div;
        }
        pthread_barrier_wait(&c->bar);
        if (c->do_div) { /* :226-252 */
            memset(b, 0, sizeof(double) * (size_t)vect_len);
            for (int64_t i = 1 + t; i <= n; i += nt) {
                const double *g = c->GD + row_base(n, i);
                const double ti = c->T[i - 1];
                const int64_t ci = c->comm[i - 1];
                for (int64_t j = i; j <= n; ++j)
                    b[tri(c->n_parts, ci, c->comm[j - 1])] += ti * c->T[j - 1] * g[j - i];
            }
            pthread_barrier_wait(&c->bar);
            if (t == 0) {
                for (int64_t k = 0; k < vect_len; ++k) {
                    double acc = 0.0;
                    for (int q = 0; q < nt; ++q) acc += c->Bp[(size_t)q * (size_t)vect_len + (size_t)k];
                    c->vect_B[k] = acc;
                }
                const double f = cge_oracle_js(c->vect_C, c->vect_B, NULL, 1, vect_len);
                if (c->tr) c->tr->div[a - 1] = f;
                if (f < best_div) {
                    best_div = f;
                    best_alpha = alpha;
                    alpha_div_counter = 5;
                } else {
                    alpha_div_counter -= 1;
                    skip_div = alpha_div_counter == 0;
                }
            }
        }
        if (t == 0) c->stop = skip_div && skip_auc; /* :253 */
        pthread_barrier_wait(&c->bar);
        if (c->stop) break;
    }
    if (t == 0) {
        double *out = c->out;
        out[0] = best_alpha; out[1] = best_div; out[2] = 0.0; out[3] = 0.0;
        out[4] = best_alpha_auc; out[5] = best_auc; out[6] = best_auc_err;
    }
    return NULL;
}

int cge_oracle_wgcl_mt(int64_t m, const int64_t *e_src, const int64_t *e_dst,
                       const double *eweights, const int64_t *comm, int64_t n, const double *embed,
                       int64_t d, const double *vweights, int64_t K, int64_t n_sets,
                       const int64_t *pos_i, const int64_t *pos_j, const double *pos_w,
                       const int64_t *neg_i, const int64_t *neg_j, int max_alphas, int n_threads,
                       double *out, cge_oracle_mt_trace *tr) {
    mt_ctx c;
    memset(&c, 0, sizeof(c));
    c.nt = n_threads > 0 ? n_threads : cge_oracle_mt_threads();
    if (c.nt > 256) c.nt = 256;
    if (tr) {
        memset(tr, 0, sizeof(*tr));
        for (int a = 0; a < N_ALPHA; ++a) tr->div[a] = tr->auc[a] = NAN;
        tr->threads = c.nt;
    }
    c.m = m; c.n = n; c.d = d; c.K = K; c.n_sets = n_sets;
    c.e_src = e_src; c.e_dst = e_dst; c.comm = comm; c.eweights = eweights; c.embed = embed;
    c.vweights = vweights; c.pos_i = pos_i; c.pos_j = pos_j; c.pos_w = pos_w; c.neg_i = neg_i;
    c.neg_j = neg_j; c.max_alphas = max_alphas; c.out = out; c.tr = tr;
    for (int64_t i = 0; i < n; ++i)
        if (comm[i] > c.n_parts) c.n_parts = comm[i];
    c.vect_len = c.n_parts * (c.n_parts + 1) / 2;
    c.p_len = n * (n + 1) / 2;
    c.vect_C = (double *)calloc((size_t)c.vect_len, sizeof(double));
    c.vect_B = (double *)calloc((size_t)c.vect_len, sizeof(double));
    c.D = (double *)malloc(sizeof(double) * (size_t)c.p_len);
    c.GD = (double *)malloc(sizeof(double) * (size_t)c.p_len);
    c.T = (double *)malloc(sizeof(double) * (size_t)n);
    c.S = (double *)malloc(sizeof(double) * (size_t)n);
    c.Sp = (double *)malloc(sizeof(double) * (size_t)n * (size_t)c.nt);
    c.Bp = (double *)malloc(sizeof(double) * (size_t)c.vect_len * (size_t)c.nt);
    c.red = (double *)calloc((size_t)c.nt, sizeof(double));
    if (g_dist_form == 1) {
        c.cen = (double *)malloc(sizeof(double) * (size_t)n * (size_t)d);
        c.nrm = (double *)calloc((size_t)n, sizeof(double));
        if (c.cen && c.nrm) {
            for (int64_t k = 0; k < d; ++k) {
                double mean = 0.0;
                for (int64_t i = 0; i < n; ++i) mean += embed[i * d + k];
                mean /= (double)n;
                for (int64_t i = 0; i < n; ++i) c.cen[i * d + k] = embed[i * d + k] - mean;
            }
            for (int64_t i = 0; i < n; ++i)
                for (int64_t k = 0; k < d; ++k) c.nrm[i] = fma(c.cen[i * d + k], c.cen[i * d + k], c.nrm[i]);
        }
    }
    int rc = 0;
    if (g_dist_form == 1 && (!c.cen || !c.nrm)) rc = -4;
    if (!c.vect_C || !c.vect_B || !c.D || !c.GD || !c.T || !c.S || !c.Sp || !c.Bp || !c.red) rc = -4;
    if (!rc) {
        for (int64_t i = 0; i < m; ++i) /* :59-63 */
            c.vect_C[tri(c.n_parts, comm[e_src[i] - 1], comm[e_dst[i] - 1])] += eweights[i];
        pthread_barrier_init(&c.bar, NULL, (unsigned)c.nt);
        pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)c.nt);
        mt_arg *args = (mt_arg *)malloc(sizeof(mt_arg) * (size_t)c.nt);
        int started = 0;
        for (int t = 1; t < c.nt; ++t) {
            args[t].c = &c;
            args[t].t = t;
            if (pthread_create(&th[t], NULL, mt_worker, &args[t]) != 0) break;
            ++started;
        }
        if (started == c.nt - 1) {
            args[0].c = &c;
            args[0].t = 0;
            mt_worker(&args[0]);
            for (int t = 1; t < c.nt; ++t) pthread_join(th[t], NULL);
        } else {
            rc = -5; /* could not start the threads; the started ones wait at the first barrier */
            for (int t = 1; t <= started; ++t) pthread_cancel(th[t]);
            for (int t = 1; t <= started; ++t) pthread_join(th[t], NULL);
        }
        pthread_barrier_destroy(&c.bar);
        free(th);
        free(args);
    }
    free(c.vect_C); free(c.vect_B); free(c.D); free(c.GD); free(c.T); free(c.S); free(c.Sp);
    free(c.Bp); free(c.red); free(c.cen); free(c.nrm);
    return rc;
}
