/*
 * comet_oracle.c -- CPU restatement of wizenheimer/comet's vector distance + ANN search path.
 * TEST INFRASTRUCTURE ONLY -- see comet_oracle.h for the contract, the arithmetic model and
 * how this oracle is pinned.  Build: oracle/Makefile (gcc -O2 -ffp-contract=off).
 *
 * Layout differences from the reference that do NOT change results: rows live in one
 * contiguous n x dim array instead of one heap slice per VectorNode (node.go:30-33), and the
 * roaring bitmaps (deleted set, document filter, HNSW visited set) are sorted arrays / byte
 * maps -- they are used purely as sets by the reference.
 */
#define _GNU_SOURCE
#include "comet_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------ */
/* small utilities                                                                      */
/* ------------------------------------------------------------------------------------ */

static int g_fma = 0;
static int g_threads = 1;
void co_set_fma(int on) { g_fma = on ? 1 : 0; }
int co_get_fma(void) { return g_fma; }
void co_set_threads(int n) { g_threads = n < 1 ? 1 : n; }

static void *xmalloc(size_t n) {
    void *p = malloc(n ? n : 1);
    if (!p) abort();
    return p;
}
static void *xrealloc(void *p, size_t n) {
    p = realloc(p, n ? n : 1);
    if (!p) abort();
    return p;
}

/* sorted-array set of uint32 (stands in for roaring.Bitmap used as a set) */
typedef struct {
    uint32_t *v;
    long n, cap;
} idset;

static int cmp_u32(const void *a, const void *b) {
    uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}
static int idset_contains(const idset *s, uint32_t id) {
    long lo = 0, hi = s->n;
    while (lo < hi) {
        long mid = (lo + hi) >> 1;
        if (s->v[mid] < id) lo = mid + 1; else hi = mid;
    }
    return lo < s->n && s->v[lo] == id;
}
static void idset_add(idset *s, uint32_t id) {
    if (idset_contains(s, id)) return;
    if (s->n == s->cap) {
        s->cap = s->cap ? s->cap * 2 : 16;
        s->v = xrealloc(s->v, (size_t)s->cap * sizeof(uint32_t));
    }
    long i = s->n++;
    while (i > 0 && s->v[i - 1] > id) { s->v[i] = s->v[i - 1]; i--; }
    s->v[i] = id;
}
static void idset_clear(idset *s) { s->n = 0; }
static void idset_free(idset *s) { free(s->v); s->v = NULL; s->n = s->cap = 0; }
/* document_filter.go:27-41: nil filter (== everything eligible) when the list is empty */
static void idset_from(idset *s, const uint32_t *ids, long n) {
    s->v = NULL; s->n = s->cap = 0;
    if (n <= 0) return;
    s->v = xmalloc((size_t)n * sizeof(uint32_t));
    memcpy(s->v, ids, (size_t)n * sizeof(uint32_t));
    qsort(s->v, (size_t)n, sizeof(uint32_t), cmp_u32);
    long w = 0;
    for (long i = 0; i < n; i++) if (w == 0 || s->v[w - 1] != s->v[i]) s->v[w++] = s->v[i];
    s->n = s->cap = w;
}
/* document_filter.go:58-67 ShouldSkip: nil filter never skips */
static int filter_should_skip(const idset *f, long nfilter, uint32_t id) {
    if (nfilter <= 0) return 0;
    return !idset_contains(f, id);
}

/* (score, scan order) pairs and their stable ordering == reference sort.Slice up to ties */
typedef struct {
    float score;
    long pos;      /* scan / collection order */
    uint32_t id;
} scored;

static int cmp_scored(const void *a, const void *b) {
    const scored *x = a, *y = b;
    if (x->score < y->score) return -1;
    if (y->score < x->score) return 1;
    return x->pos < y->pos ? -1 : (x->pos > y->pos ? 1 : 0);
}

/* ------------------------------------------------------------------------------------ */
/* distance.go                                                                          */
/* ------------------------------------------------------------------------------------ */

/* distance.go:158-165 (l2Squared.Calculate) and the loop of :114-121 (euclidean.Calculate) */
static float l2sq_seq(const float *a, const float *b, int d) {
    float sum = 0.0f;
    if (g_fma) {
        for (int i = 0; i < d; i++) { float diff = a[i] - b[i]; sum = fmaf(diff, diff, sum); }
    } else {
        for (int i = 0; i < d; i++) { float diff = a[i] - b[i]; sum += diff * diff; }
    }
    return sum;
}
/* distance.go:201-216 (cosine.Calculate): assumes both operands pre-normalised */
static float cosine_seq(const float *a, const float *b, int d) {
    float dot = 0.0f;
    if (g_fma) {
        for (int i = 0; i < d; i++) dot = fmaf(a[i], b[i], dot);
    } else {
        for (int i = 0; i < d; i++) dot += a[i] * b[i];
    }
    if (dot > 1.0f) dot = 1.0f; else if (dot < -1.0f) dot = -1.0f;
    return 1.0f - dot;
}
/* float32(math.Sqrt(float64(x))): double rounding through binary64 is innocuous for sqrt,
 * so this equals the correctly rounded binary32 sqrt (what __fsqrt_rn computes on device). */
static float sqrt_go(float x) { return (float)sqrt((double)x); }

float co_distance(int metric, const float *a, const float *b, int d) {
    switch (metric) {
    case CO_L2:     return sqrt_go(l2sq_seq(a, b, d));   /* distance.go:114-121 */
    case CO_L2SQ:   return l2sq_seq(a, b, d);            /* distance.go:158-165 */
    case CO_COSINE: return cosine_seq(a, b, d);          /* distance.go:201-216 */
    }
    return NAN;
}

/* distance.go:244-264 / 269-290: sum of squares sequential, norm = float32(sqrt(float64)),
 * scale = 1.0/norm in float32, x*scale.  (Norm(), distance.go:303-311, is the same loop.) */
float co_norm(const float *v, int d) {
    float sum = 0.0f;
    if (g_fma) { for (int i = 0; i < d; i++) sum = fmaf(v[i], v[i], sum); }
    else       { for (int i = 0; i < d; i++) sum += v[i] * v[i]; }
    return sqrt_go(sum);
}
int co_normalize(const float *in, float *out, int d) {
    float norm = co_norm(in, d);
    if (norm == 0.0f) return CO_ERR_ZERO_VECTOR;
    float scale = 1.0f / norm;
    for (int i = 0; i < d; i++) out[i] = in[i] * scale;
    return CO_OK;
}
/* Distance.Preprocess / PreprocessInPlace (distance.go:50-81): no-op except for cosine */
int co_preprocess(int metric, const float *in, float *out, int d) {
    if (metric == CO_COSINE) return co_normalize(in, out, d);
    if (in != out) memcpy(out, in, (size_t)d * sizeof(float));
    return CO_OK;
}

/* ------------------------------------------------------------------------------------ */
/* limiter.go / aggregation.go                                                          */
/* ------------------------------------------------------------------------------------ */

/* limiter.go:12-17 */
long co_sanitize_k(long k, long max_results) {
    if (k <= 0 || k > max_results) return max_results;
    return k;
}

/* aggregation.go:107-141 (sum), :161-199 (max), :217-255 (mean).  The reference groups through
 * a Go map and sorts unstably; the oracle keeps first-appearance order for equal scores. */
long co_aggregate(int kind, const uint32_t *ids, const float *scores, long n,
                  uint32_t *out_ids, float *out_scores) {
    if (n <= 0) return 0;
    /* group in first-appearance order; scores of a group are combined in arrival order */
    uint32_t *gid = xmalloc((size_t)n * sizeof(uint32_t));
    float *gacc = xmalloc((size_t)n * sizeof(float));
    long *gcnt = xmalloc((size_t)n * sizeof(long));
    long ng = 0;
    /* simple open-addressing map id -> group */
    long cap = 16; while (cap < 2 * n) cap <<= 1;
    long *slot = xmalloc((size_t)cap * sizeof(long));
    for (long i = 0; i < cap; i++) slot[i] = -1;
    for (long i = 0; i < n; i++) {
        uint32_t id = ids[i];
        long h = (long)((id * 2654435761u) & (uint32_t)(cap - 1));
        while (slot[h] >= 0 && gid[slot[h]] != id) h = (h + 1) & (cap - 1);
        if (slot[h] < 0) {
            slot[h] = ng; gid[ng] = id; gcnt[ng] = 1;
            /* sum: float32(0) + s ; max: scores[0] */
            gacc[ng] = (kind == CO_AGG_MAX) ? scores[i] : 0.0f + scores[i];
            ng++;
        } else {
            long g = slot[h];
            if (kind == CO_AGG_MAX) { if (scores[i] > gacc[g]) gacc[g] = scores[i]; }
            else gacc[g] += scores[i];
            gcnt[g]++;
        }
    }
    scored *r = xmalloc((size_t)ng * sizeof(scored));
    for (long g = 0; g < ng; g++) {
        float s = gacc[g];
        if (kind == CO_AGG_MEAN) s = s / (float)gcnt[g];
        r[g].score = s; r[g].pos = g; r[g].id = gid[g];
    }
    qsort(r, (size_t)ng, sizeof(scored), cmp_scored);
    for (long g = 0; g < ng; g++) { out_ids[g] = r[g].id; out_scores[g] = r[g].score; }
    free(r); free(slot); free(gcnt); free(gacc); free(gid);
    return ng;
}

/* limiter.go:84-118 Autocut (AutocutResults, :52-69, is a no-op for cutoff == -1 or n == 0) */
long co_autocut(const float *y, long n, int cutoff) {
    if (cutoff == -1 || n == 0) return n;
    if (n <= 1) return n;
    float *diff = xmalloc((size_t)n * sizeof(float));
    float step = 1.0f / ((float)n - 1.0f);
    for (long i = 0; i < n; i++) {
        float x = 0.0f + (float)i * step;
        float yn = (y[i] - y[0]) / (y[n - 1] - y[0]);
        diff[i] = yn - x;
    }
    int extrema = 0;
    long cut = n;
    for (long i = 1; i < n; i++) {
        if (i == n - 1) {
            /* reference reads diff[i-2] after the first test passes; with n == 2 that index is
             * -1 and Go would panic -- the oracle treats it as "no extremum". */
            if (diff[i] > diff[i - 1] && i - 2 >= 0 && diff[i] > diff[i - 2]) {
                if (++extrema >= cutoff) { cut = i; break; }
            }
        } else {
            if (diff[i] > diff[i - 1] && diff[i] > diff[i + 1]) {
                if (++extrema >= cutoff) { cut = i; break; }
            }
        }
    }
    free(diff);
    return cut;
}

/* ------------------------------------------------------------------------------------ */
/* clustering.go                                                                        */
/* ------------------------------------------------------------------------------------ */

/* clustering.go:256-272 FindNearestCentroidIndex: strict '<', first minimum wins */
int co_nearest_centroid(const float *v, const float *centroids, int k, int d, int metric) {
    float min_dist = INFINITY;
    int min_idx = 0;
    for (int i = 0; i < k; i++) {
        float dist = co_distance(metric, v, centroids + (size_t)i * d, d);
        if (dist < min_dist) { min_dist = dist; min_idx = i; }
    }
    return min_idx;
}

/* clustering.go:119-253 kmeansInternal.  Deterministic: centroid c starts at vector c*(n/k);
 * assignment with strict '<'; update = sequential float32 sums in vector order / float32(count);
 * an empty cluster keeps its centroid; stop when no assignment changed or after max_iter. */
int co_kmeans(const float *vectors, long n, int d, long ld, int k, int metric, int max_iter,
              float *centroids, int *assign_out) {
    if (n <= 0 || k <= 0) return 0;
    if (k > n) k = (int)n;
    if (max_iter <= 0) max_iter = 20;            /* clustering.go:13-14 DefaultMaxIter */
    long step = n / k;
    if (step == 0) step = 1;
    for (int c = 0; c < k; c++) {
        long vi = (long)c * step;
        if (vi >= n) vi = n - 1;
        memcpy(centroids + (size_t)c * d, vectors + (size_t)vi * ld, (size_t)d * sizeof(float));
    }
    int *assign = assign_out ? assign_out : xmalloc((size_t)n * sizeof(int));
    for (long i = 0; i < n; i++) assign[i] = -1;  /* UnassignedCluster */
    float *sums = xmalloc((size_t)k * d * sizeof(float));
    long *sizes = xmalloc((size_t)k * sizeof(long));
    for (int it = 0; it < max_iter; it++) {
        int changed = 0;
        for (long i = 0; i < n; i++) {
            int c = co_nearest_centroid(vectors + (size_t)i * ld, centroids, k, d, metric);
            if (assign[i] != c) { changed = 1; assign[i] = c; }
        }
        if (!changed) break;
        memset(sums, 0, (size_t)k * d * sizeof(float));
        memset(sizes, 0, (size_t)k * sizeof(long));
        for (long i = 0; i < n; i++) {
            int c = assign[i];
            if (c < 0) continue;
            float *s = sums + (size_t)c * d;
            const float *v = vectors + (size_t)i * ld;
            for (int j = 0; j < d; j++) s[j] += v[j];
            sizes[c]++;
        }
        for (int c = 0; c < k; c++) {
            if (sizes[c] > 0) {
                float cnt = (float)sizes[c];
                for (int j = 0; j < d; j++)
                    centroids[(size_t)c * d + j] = sums[(size_t)c * d + j] / cnt;
            }
        }
    }
    free(sizes); free(sums);
    if (!assign_out) free(assign);
    return k;
}

/* ------------------------------------------------------------------------------------ */
/* flat_index.go / flat_index_search.go                                                 */
/* ------------------------------------------------------------------------------------ */

struct co_flat {
    int dim, metric;
    long n, cap;
    float *rows;       /* n x dim, preprocessed (flat_index.go:169-189) */
    uint32_t *ids;
    idset deleted;     /* flat_index.go:87 deletedNodes */
};

co_flat *co_flat_new(int dim, int metric) {
    if (dim <= 0 || metric < 0 || metric > 2) return NULL;   /* flat_index.go:118-139 */
    co_flat *f = calloc(1, sizeof(*f));
    f->dim = dim; f->metric = metric;
    return f;
}
void co_flat_free(co_flat *f) {
    if (!f) return;
    free(f->rows); free(f->ids); idset_free(&f->deleted); free(f);
}
static void flat_reserve(co_flat *f, long need) {
    if (need <= f->cap) return;
    long cap = f->cap ? f->cap : 64;
    while (cap < need) cap *= 2;
    f->rows = xrealloc(f->rows, (size_t)cap * f->dim * sizeof(float));
    f->ids = xrealloc(f->ids, (size_t)cap * sizeof(uint32_t));
    f->cap = cap;
}
/* flat_index.go:169-189 Add: PreprocessInPlace on the caller's slice, then append */
int co_flat_add(co_flat *f, uint32_t id, float *vec) {
    int rc = co_preprocess(f->metric, vec, vec, f->dim);
    if (rc) return rc;
    flat_reserve(f, f->n + 1);
    memcpy(f->rows + (size_t)f->n * f->dim, vec, (size_t)f->dim * sizeof(float));
    f->ids[f->n++] = id;
    return CO_OK;
}
int co_flat_add_batch(co_flat *f, const uint32_t *ids, float *rows, long n) {
    flat_reserve(f, f->n + n);
    for (long i = 0; i < n; i++) {
        int rc = co_flat_add(f, ids[i], rows + (size_t)i * f->dim);
        if (rc) return rc;
    }
    return CO_OK;
}
/* flat_index.go:219-250 Remove: error if absent or already deleted */
int co_flat_remove(co_flat *f, uint32_t id) {
    int exists = 0;
    for (long i = 0; i < f->n; i++) if (f->ids[i] == id) { exists = 1; break; }
    if (!exists) return CO_ERR_NOT_FOUND;
    if (idset_contains(&f->deleted, id)) return CO_ERR_NOT_FOUND;
    idset_add(&f->deleted, id);
    return CO_OK;
}
/* flat_index.go:266-299 Flush */
int co_flat_flush(co_flat *f) {
    if (f->deleted.n == 0) return CO_OK;
    long w = 0;
    for (long i = 0; i < f->n; i++) {
        if (idset_contains(&f->deleted, f->ids[i])) continue;
        if (w != i) {
            memmove(f->rows + (size_t)w * f->dim, f->rows + (size_t)i * f->dim,
                    (size_t)f->dim * sizeof(float));
            f->ids[w] = f->ids[i];
        }
        w++;
    }
    f->n = w;
    idset_clear(&f->deleted);
    return CO_OK;
}
long co_flat_size(const co_flat *f) { return f->n; }
const float *co_flat_rows(const co_flat *f) { return f->rows; }
const uint32_t *co_flat_ids(const co_flat *f) { return f->ids; }

/* flat_index_search.go:221-294 searchSingleQuery */
long co_flat_search(const co_flat *f, const float *query, long k_req, float threshold,
                    const uint32_t *filter_ids, long nfilter,
                    uint32_t *out_ids, float *out_scores, long *out_pos) {
    int d = f->dim;
    long k = co_sanitize_k(k_req, f->n);                       /* :231 */
    float *q = xmalloc((size_t)d * sizeof(float));
    int rc = co_preprocess(f->metric, query, q, d);            /* :236 */
    if (rc) { free(q); return rc; }
    idset filt; idset_from(&filt, filter_ids, nfilter);        /* :242 */
    scored *res = xmalloc((size_t)(f->n ? f->n : 1) * sizeof(scored));
    long nres = 0;
    for (long i = 0; i < f->n; i++) {                          /* :254-274 */
        uint32_t id = f->ids[i];
        if (f->deleted.n && idset_contains(&f->deleted, id)) continue;
        if (filter_should_skip(&filt, nfilter, id)) continue;
        float dist = co_distance(f->metric, q, f->rows + (size_t)i * d, d);
        if (threshold > 0 && dist > threshold) continue;
        res[nres].score = dist; res[nres].pos = i; res[nres].id = id; nres++;
    }
    qsort(res, (size_t)nres, sizeof(scored), cmp_scored);      /* :277-279 */
    k = co_sanitize_k(k, nres);                                /* :282 */
    for (long i = 0; i < k; i++) {
        out_ids[i] = res[i].id; out_scores[i] = res[i].score;
        if (out_pos) out_pos[i] = res[i].pos;
    }
    free(res); idset_free(&filt); free(q);
    return k;
}

typedef struct {
    const co_flat *f;
    const float *queries;
    long nq, k, kcap;
    float threshold;
    uint32_t *out_ids; float *out_scores; long *counts;
    long next; pthread_mutex_t mu; int err;
} flat_batch_job;

static void *flat_batch_worker(void *arg) {
    flat_batch_job *j = arg;
    long kbuf = co_sanitize_k(j->k, j->f->n);
    uint32_t *ids = xmalloc((size_t)(kbuf ? kbuf : 1) * sizeof(uint32_t));
    float *sc = xmalloc((size_t)(kbuf ? kbuf : 1) * sizeof(float));
    for (;;) {
        pthread_mutex_lock(&j->mu);
        long qi = j->next++;
        pthread_mutex_unlock(&j->mu);
        if (qi >= j->nq) break;
        long c = co_flat_search(j->f, j->queries + (size_t)qi * j->f->dim, j->k, j->threshold,
                                NULL, 0, ids, sc, NULL);
        if (c < 0) { j->err = (int)c; j->counts[qi] = 0; continue; }
        long w = c < j->kcap ? c : j->kcap;
        memcpy(j->out_ids + (size_t)qi * j->kcap, ids, (size_t)w * sizeof(uint32_t));
        memcpy(j->out_scores + (size_t)qi * j->kcap, sc, (size_t)w * sizeof(float));
        j->counts[qi] = w;
    }
    free(sc); free(ids);
    return NULL;
}
/* nq independent Execute()s -- what a reference user gets from one goroutine per query under
 * the RLock (flat_index_search.go:222); see SURVEY.md F6. */
int co_flat_search_batch(const co_flat *f, const float *queries, long nq, long k, float threshold,
                         long kcap, uint32_t *out_ids, float *out_scores, long *counts) {
    flat_batch_job j = { f, queries, nq, k, kcap, threshold, out_ids, out_scores, counts, 0,
                         PTHREAD_MUTEX_INITIALIZER, 0 };
    int nt = g_threads; if (nt > nq) nt = (int)nq; if (nt < 1) nt = 1;
    pthread_t *th = xmalloc((size_t)nt * sizeof(pthread_t));
    for (int t = 1; t < nt; t++) pthread_create(&th[t], NULL, flat_batch_worker, &j);
    flat_batch_worker(&j);
    for (int t = 1; t < nt; t++) pthread_join(th[t], NULL);
    free(th);
    return j.err;
}

/* ------------------------------------------------------------------------------------ */
/* ivf_index.go / ivf_index_search.go                                                   */
/* ------------------------------------------------------------------------------------ */

typedef struct {
    long n, cap;
    uint32_t *ids;
    float *rows;     /* IVF */
    uint8_t *codes;  /* IVFPQ */
} inv_list;

struct co_ivf {
    int dim, nlist, metric, trained;
    float *centroids;  /* nlist x dim */
    inv_list *lists;
    idset deleted;
};

co_ivf *co_ivf_new(int dim, int nlist, int metric) {
    if (dim <= 0 || nlist <= 0 || metric < 0 || metric > 2) return NULL;  /* ivf_index.go:147-175 */
    co_ivf *x = calloc(1, sizeof(*x));
    x->dim = dim; x->nlist = nlist; x->metric = metric;
    x->lists = calloc((size_t)nlist, sizeof(inv_list));
    return x;
}
void co_ivf_free(co_ivf *x) {
    if (!x) return;
    for (int l = 0; l < x->nlist; l++) { free(x->lists[l].ids); free(x->lists[l].rows); }
    free(x->lists); free(x->centroids); idset_free(&x->deleted); free(x);
}
/* ivf_index.go:206-236 Train: KMeans(raw, nlist, idx.distance, 20) on the raw vectors */
int co_ivf_train(co_ivf *x, const float *rows, long n) {
    if (n < x->nlist) return CO_ERR_TOO_FEW;
    free(x->centroids);
    x->centroids = xmalloc((size_t)x->nlist * x->dim * sizeof(float));
    int k = co_kmeans(rows, n, x->dim, x->dim, x->nlist, x->metric, 20, x->centroids, NULL);
    if (k != x->nlist) return CO_ERR_ARG;
    x->trained = 1;
    return CO_OK;
}
int co_ivf_set_centroids(co_ivf *x, const float *c) {
    free(x->centroids);
    x->centroids = xmalloc((size_t)x->nlist * x->dim * sizeof(float));
    memcpy(x->centroids, c, (size_t)x->nlist * x->dim * sizeof(float));
    x->trained = 1;
    return CO_OK;
}
static void list_push(inv_list *L, uint32_t id, const float *row, int dim, const uint8_t *code, int M) {
    if (L->n == L->cap) {
        L->cap = L->cap ? L->cap * 2 : 8;
        L->ids = xrealloc(L->ids, (size_t)L->cap * sizeof(uint32_t));
        if (row) L->rows = xrealloc(L->rows, (size_t)L->cap * dim * sizeof(float));
        if (code) L->codes = xrealloc(L->codes, (size_t)L->cap * M);
    }
    L->ids[L->n] = id;
    if (row) memcpy(L->rows + (size_t)L->n * dim, row, (size_t)dim * sizeof(float));
    if (code) memcpy(L->codes + (size_t)L->n * M, code, (size_t)M);
    L->n++;
}
/* ivf_index.go:251-280 Add */
int co_ivf_add(co_ivf *x, uint32_t id, float *vec) {
    if (!x->trained) return CO_ERR_NOT_TRAINED;
    int rc = co_preprocess(x->metric, vec, vec, x->dim);
    if (rc) return rc;
    int l = co_nearest_centroid(vec, x->centroids, x->nlist, x->dim, x->metric);
    list_push(&x->lists[l], id, vec, x->dim, NULL, 0);
    return CO_OK;
}
int co_ivf_add_batch(co_ivf *x, const uint32_t *ids, float *rows, long n) {
    for (long i = 0; i < n; i++) {
        int rc = co_ivf_add(x, ids[i], rows + (size_t)i * x->dim);
        if (rc) return rc;
    }
    return CO_OK;
}
static int lists_contain(const inv_list *lists, int nlist, uint32_t id) {
    for (int l = 0; l < nlist; l++)
        for (long i = 0; i < lists[l].n; i++) if (lists[l].ids[i] == id) return 1;
    return 0;
}
/* ivf_index.go:307-345 Remove */
int co_ivf_remove(co_ivf *x, uint32_t id) {
    if (!lists_contain(x->lists, x->nlist, id)) return CO_ERR_NOT_FOUND;
    if (idset_contains(&x->deleted, id)) return CO_ERR_NOT_FOUND;
    idset_add(&x->deleted, id);
    return CO_OK;
}
/* ivf_index.go:362-400 Flush */
int co_ivf_flush(co_ivf *x) {
    if (x->deleted.n == 0) return CO_OK;
    for (int l = 0; l < x->nlist; l++) {
        inv_list *L = &x->lists[l];
        long w = 0;
        for (long i = 0; i < L->n; i++) {
            if (idset_contains(&x->deleted, L->ids[i])) continue;
            if (w != i) {
                L->ids[w] = L->ids[i];
                memmove(L->rows + (size_t)w * x->dim, L->rows + (size_t)i * x->dim,
                        (size_t)x->dim * sizeof(float));
            }
            w++;
        }
        L->n = w;
    }
    idset_clear(&x->deleted);
    return CO_OK;
}
/* ivf_index.go:406-413 default nprobes = int(sqrt(nlist)) */
int co_ivf_default_nprobes(const co_ivf *x) { return (int)sqrt((double)x->nlist); }
const float *co_ivf_centroids(const co_ivf *x) { return x->centroids; }
long co_ivf_list_len(const co_ivf *x, int l) { return x->lists[l].n; }
void co_ivf_list_get(const co_ivf *x, int l, uint32_t *ids, float *rows) {
    const inv_list *L = &x->lists[l];
    if (ids) memcpy(ids, L->ids, (size_t)L->n * sizeof(uint32_t));
    if (rows) memcpy(rows, L->rows, (size_t)L->n * x->dim * sizeof(float));
}

/* coarse quantiser step shared by IVF and IVFPQ: ivf_index_search.go:252-261,
 * ivfpq_index_search.go:263-272.  Full sort of nlist (stable here: ties by list index). */
static scored *coarse_sort(const float *q, const float *centroids, int nlist, int d, int metric) {
    scored *cd = xmalloc((size_t)nlist * sizeof(scored));
    for (int i = 0; i < nlist; i++) {
        cd[i].score = co_distance(metric, q, centroids + (size_t)i * d, d);
        cd[i].pos = i; cd[i].id = (uint32_t)i;
    }
    qsort(cd, (size_t)nlist, sizeof(scored), cmp_scored);
    return cd;
}

/* ivf_index_search.go:217-322 searchSingleQuery */
long co_ivf_search(const co_ivf *x, const float *query, long k_req, int nprobes, float threshold,
                   const uint32_t *filter_ids, long nfilter,
                   uint32_t *out_ids, float *out_scores) {
    if (!x->trained) return CO_ERR_NOT_TRAINED;                   /* :222-224 */
    int d = x->dim;
    if (nprobes <= 0 || nprobes > x->nlist) nprobes = x->nlist;   /* :233-236 */
    float *q = xmalloc((size_t)d * sizeof(float));
    int rc = co_preprocess(x->metric, query, q, d);
    if (rc) { free(q); return rc; }
    scored *cd = coarse_sort(q, x->centroids, x->nlist, d, x->metric);
    idset filt; idset_from(&filt, filter_ids, nfilter);
    long total = 0;
    for (int i = 0; i < nprobes; i++) total += x->lists[cd[i].id].n;
    scored *cand = xmalloc((size_t)(total ? total : 1) * sizeof(scored));
    long nc = 0, pos = 0;
    for (int i = 0; i < nprobes; i++) {                           /* :277-308 */
        const inv_list *L = &x->lists[cd[i].id];
        for (long j = 0; j < L->n; j++, pos++) {
            uint32_t id = L->ids[j];
            if (x->deleted.n && idset_contains(&x->deleted, id)) continue;
            if (filter_should_skip(&filt, nfilter, id)) continue;
            float dist = co_distance(x->metric, q, L->rows + (size_t)j * d, d);
            if (threshold > 0 && dist > threshold) continue;
            cand[nc].score = dist; cand[nc].pos = pos; cand[nc].id = id; nc++;
        }
    }
    qsort(cand, (size_t)nc, sizeof(scored), cmp_scored);          /* :313-315 */
    long k = co_sanitize_k(k_req, nc);                            /* :318 */
    for (long i = 0; i < k; i++) { out_ids[i] = cand[i].id; out_scores[i] = cand[i].score; }
    free(cand); idset_free(&filt); free(cd); free(q);
    return k;
}

/* ------------------------------------------------------------------------------------ */
/* pq_index.go / pq_index_search.go                                                     */
/* ------------------------------------------------------------------------------------ */

struct co_pq {
    int dim, metric, M, nbits, Ksub, dsub, trained;
    float *codebooks;   /* M x Ksub x dsub (pq_index.go:98-104: codebooks[m][k*dsub + j]) */
    long n, cap;
    uint8_t *codes;     /* n x M */
    uint32_t *ids;
    idset deleted;
};

/* pq_index.go:439-473 encode / ivfpq_index.go:467-500 encodeResidual: per sub-space argmin of
 * sequential L2^2 over Ksub centroids, strict '<', stored as uint8(minIdx) (truncating!) */
static void pq_encode(const float *codebooks, int M, int Ksub, int dsub, const float *v, uint8_t *code) {
    for (int m = 0; m < M; m++) {
        const float *sub = v + (size_t)m * dsub;
        float min_dist = INFINITY;
        int min_idx = 0;
        for (int ks = 0; ks < Ksub; ks++) {
            const float *c = codebooks + ((size_t)m * Ksub + ks) * dsub;
            float dist = l2sq_seq(sub, c, dsub);
            if (dist < min_dist) { min_dist = dist; min_idx = ks; }
        }
        code[m] = (uint8_t)min_idx;
    }
}
/* per-subspace codebook training: pq_index.go:220-243 / ivfpq_index.go:235-255.
 * KMeansSubspace(sub, Ksub, 20) == kmeansInternal with L2Squared (clustering.go:112-115) */
static int pq_train_codebooks(const float *rows, long n, int dim, int M, int Ksub, int dsub, float *codebooks) {
    for (int m = 0; m < M; m++) {
        int k = co_kmeans(rows + (size_t)m * dsub, n, dsub, dim, Ksub, CO_L2SQ, 20,
                          codebooks + (size_t)m * Ksub * dsub, NULL);
        if (k != Ksub) return CO_ERR_ARG;
    }
    return CO_OK;
}
/* K7: pq_index_search.go:243-264 / ivfpq_index_search.go:350-375 */
static void pq_build_lut(const float *codebooks, int M, int Ksub, int dsub, const float *q, float *lut) {
    for (int m = 0; m < M; m++)
        for (int ks = 0; ks < Ksub; ks++)
            lut[(size_t)m * Ksub + ks] =
                l2sq_seq(q + (size_t)m * dsub, codebooks + ((size_t)m * Ksub + ks) * dsub, dsub);
}
/* K8: pq_index_search.go:290-295 / ivfpq_index_search.go:384-390 */
static float pq_adc(const float *lut, int M, int Ksub, const uint8_t *code) {
    float dist = 0.0f;
    for (int m = 0; m < M; m++) dist += lut[(size_t)m * Ksub + code[m]];
    return sqrt_go(dist);
}

co_pq *co_pq_new(int dim, int metric, int M, int nbits) {
    /* pq_index.go:135-176 */
    if (dim <= 0 || M <= 0 || dim % M != 0 || nbits <= 0 || nbits > 16) return NULL;
    if (metric < 0 || metric > 2) return NULL;
    co_pq *p = calloc(1, sizeof(*p));
    p->dim = dim; p->metric = metric; p->M = M; p->nbits = nbits;
    p->Ksub = 1 << nbits; p->dsub = dim / M;
    return p;
}
void co_pq_free(co_pq *p) {
    if (!p) return;
    free(p->codebooks); free(p->codes); free(p->ids); idset_free(&p->deleted); free(p);
}
/* pq_index.go:193-247 Train: raw (NOT preprocessed) vectors */
int co_pq_train(co_pq *p, const float *rows, long n) {
    if (n < p->Ksub) return CO_ERR_TOO_FEW;
    free(p->codebooks);
    p->codebooks = xmalloc((size_t)p->M * p->Ksub * p->dsub * sizeof(float));
    int rc = pq_train_codebooks(rows, n, p->dim, p->M, p->Ksub, p->dsub, p->codebooks);
    if (rc) return rc;
    p->trained = 1;
    return CO_OK;
}
int co_pq_set_codebooks(co_pq *p, const float *cb) {
    size_t sz = (size_t)p->M * p->Ksub * p->dsub * sizeof(float);
    free(p->codebooks);
    p->codebooks = xmalloc(sz);
    memcpy(p->codebooks, cb, sz);
    p->trained = 1;
    return CO_OK;
}
/* pq_index.go:262-293 Add */
int co_pq_add(co_pq *p, uint32_t id, float *vec) {
    if (!p->trained) return CO_ERR_NOT_TRAINED;
    int rc = co_preprocess(p->metric, vec, vec, p->dim);
    if (rc) return rc;
    if (p->n == p->cap) {
        p->cap = p->cap ? p->cap * 2 : 64;
        p->codes = xrealloc(p->codes, (size_t)p->cap * p->M);
        p->ids = xrealloc(p->ids, (size_t)p->cap * sizeof(uint32_t));
    }
    pq_encode(p->codebooks, p->M, p->Ksub, p->dsub, vec, p->codes + (size_t)p->n * p->M);
    p->ids[p->n++] = id;
    return CO_OK;
}
int co_pq_add_batch(co_pq *p, const uint32_t *ids, float *rows, long n) {
    for (long i = 0; i < n; i++) {
        int rc = co_pq_add(p, ids[i], rows + (size_t)i * p->dim);
        if (rc) return rc;
    }
    return CO_OK;
}
int co_pq_remove(co_pq *p, uint32_t id) {
    int exists = 0;
    for (long i = 0; i < p->n; i++) if (p->ids[i] == id) { exists = 1; break; }
    if (!exists || idset_contains(&p->deleted, id)) return CO_ERR_NOT_FOUND;
    idset_add(&p->deleted, id);
    return CO_OK;
}
int co_pq_flush(co_pq *p) {
    if (p->deleted.n == 0) return CO_OK;
    long w = 0;
    for (long i = 0; i < p->n; i++) {
        if (idset_contains(&p->deleted, p->ids[i])) continue;
        if (w != i) { p->ids[w] = p->ids[i]; memmove(p->codes + (size_t)w * p->M, p->codes + (size_t)i * p->M, (size_t)p->M); }
        w++;
    }
    p->n = w;
    idset_clear(&p->deleted);
    return CO_OK;
}
long co_pq_size(const co_pq *p) { return p->n; }
const float *co_pq_codebooks(const co_pq *p) { return p->codebooks; }
const uint8_t *co_pq_codes(const co_pq *p) { return p->codes; }
const uint32_t *co_pq_ids(const co_pq *p) { return p->ids; }
void co_pq_encode(const co_pq *p, const float *vec, uint8_t *code) {
    pq_encode(p->codebooks, p->M, p->Ksub, p->dsub, vec, code);
}

/* pq_index_search.go:218-325 searchSingleQuery.  ADC ignores the metric: always sqrt(sum LUT). */
long co_pq_search(const co_pq *p, const float *query, long k_req, float threshold,
                  const uint32_t *filter_ids, long nfilter,
                  uint32_t *out_ids, float *out_scores) {
    if (!p->trained) return CO_ERR_NOT_TRAINED;
    if (p->n == 0) return 0;                                      /* :231-233 */
    int d = p->dim;
    float *q = xmalloc((size_t)d * sizeof(float));
    int rc = co_preprocess(p->metric, query, q, d);
    if (rc) { free(q); return rc; }
    float *lut = xmalloc((size_t)p->M * p->Ksub * sizeof(float));
    pq_build_lut(p->codebooks, p->M, p->Ksub, p->dsub, q, lut);
    idset filt; idset_from(&filt, filter_ids, nfilter);
    scored *res = xmalloc((size_t)p->n * sizeof(scored));
    long nres = 0;
    for (long i = 0; i < p->n; i++) {                             /* :277-306 */
        uint32_t id = p->ids[i];
        if (p->deleted.n && idset_contains(&p->deleted, id)) continue;
        if (filter_should_skip(&filt, nfilter, id)) continue;
        float dist = pq_adc(lut, p->M, p->Ksub, p->codes + (size_t)i * p->M);
        if (threshold > 0 && dist > threshold) continue;
        res[nres].score = dist; res[nres].pos = i; res[nres].id = id; nres++;
    }
    qsort(res, (size_t)nres, sizeof(scored), cmp_scored);
    long k = co_sanitize_k(k_req, nres);
    for (long i = 0; i < k; i++) { out_ids[i] = res[i].id; out_scores[i] = res[i].score; }
    free(res); idset_free(&filt); free(lut); free(q);
    return k;
}

/* ------------------------------------------------------------------------------------ */
/* ivfpq_index.go / ivfpq_index_search.go                                               */
/* ------------------------------------------------------------------------------------ */

struct co_ivfpq {
    int dim, metric, nlist, M, nbits, Ksub, dsub, trained;
    float *centroids;   /* nlist x dim */
    float *codebooks;   /* M x Ksub x dsub, trained on residuals */
    inv_list *lists;
    idset deleted;
};

static __thread long tl_ivfpq_scanned;
long co_ivfpq_last_scanned(void) { return tl_ivfpq_scanned; }

co_ivfpq *co_ivfpq_new(int dim, int metric, int nlist, int M, int nbits) {
    /* ivfpq_index.go:114-159 */
    if (dim <= 0 || nlist <= 0 || M <= 0 || dim % M != 0 || nbits <= 0 || nbits > 16) return NULL;
    if (metric < 0 || metric > 2) return NULL;
    co_ivfpq *x = calloc(1, sizeof(*x));
    x->dim = dim; x->metric = metric; x->nlist = nlist; x->M = M; x->nbits = nbits;
    x->Ksub = 1 << nbits; x->dsub = dim / M;
    x->lists = calloc((size_t)nlist, sizeof(inv_list));
    return x;
}
void co_ivfpq_free(co_ivfpq *x) {
    if (!x) return;
    for (int l = 0; l < x->nlist; l++) { free(x->lists[l].ids); free(x->lists[l].codes); }
    free(x->lists); free(x->centroids); free(x->codebooks); idset_free(&x->deleted); free(x);
}
/* ivfpq_index.go:180-259 Train: KMeans on raw vectors -> assign -> residuals -> per-m KMeansSubspace */
int co_ivfpq_train(co_ivfpq *x, const float *rows, long n) {
    if (n < (long)x->nlist * 10) return CO_ERR_TOO_FEW;              /* :185 */
    int d = x->dim;
    free(x->centroids); free(x->codebooks);
    x->centroids = xmalloc((size_t)x->nlist * d * sizeof(float));
    x->codebooks = xmalloc((size_t)x->M * x->Ksub * x->dsub * sizeof(float));
    int k = co_kmeans(rows, n, d, d, x->nlist, x->metric, 20, x->centroids, NULL);
    if (k != x->nlist) return CO_ERR_ARG;
    float *resid = xmalloc((size_t)n * d * sizeof(float));
    for (long i = 0; i < n; i++) {
        const float *v = rows + (size_t)i * d;
        int a = co_nearest_centroid(v, x->centroids, x->nlist, d, x->metric);   /* :213-216 */
        const float *c = x->centroids + (size_t)a * d;
        for (int j = 0; j < d; j++) resid[(size_t)i * d + j] = v[j] - c[j];     /* :219-227 */
    }
    int rc = pq_train_codebooks(resid, n, d, x->M, x->Ksub, x->dsub, x->codebooks);
    free(resid);
    if (rc) return rc;
    x->trained = 1;
    return CO_OK;
}
int co_ivfpq_set_trained(co_ivfpq *x, const float *centroids, const float *codebooks) {
    size_t cs = (size_t)x->nlist * x->dim * sizeof(float);
    size_t bs = (size_t)x->M * x->Ksub * x->dsub * sizeof(float);
    free(x->centroids); free(x->codebooks);
    x->centroids = xmalloc(cs); memcpy(x->centroids, centroids, cs);
    x->codebooks = xmalloc(bs); memcpy(x->codebooks, codebooks, bs);
    x->trained = 1;
    return CO_OK;
}
/* ivfpq_index.go:279-319 Add: preprocess in place, nearest centroid, residual, encodeResidual */
int co_ivfpq_add(co_ivfpq *x, uint32_t id, float *vec) {
    if (!x->trained) return CO_ERR_NOT_TRAINED;
    int d = x->dim;
    int rc = co_preprocess(x->metric, vec, vec, d);
    if (rc) return rc;
    int l = co_nearest_centroid(vec, x->centroids, x->nlist, d, x->metric);
    const float *c = x->centroids + (size_t)l * d;
    float *resid = xmalloc((size_t)d * sizeof(float));
    for (int j = 0; j < d; j++) resid[j] = vec[j] - c[j];
    uint8_t *code = xmalloc((size_t)x->M);
    pq_encode(x->codebooks, x->M, x->Ksub, x->dsub, resid, code);
    list_push(&x->lists[l], id, NULL, 0, code, x->M);
    free(code); free(resid);
    return CO_OK;
}
int co_ivfpq_add_batch(co_ivfpq *x, const uint32_t *ids, float *rows, long n) {
    for (long i = 0; i < n; i++) {
        int rc = co_ivfpq_add(x, ids[i], rows + (size_t)i * x->dim);
        if (rc) return rc;
    }
    return CO_OK;
}
/* Restore path (IVFPQIndex.ReadFrom, ivfpq_index.go:700+): compressed vectors return to the lists they were
 * stored in, in stored order, without being re-encoded -- what a search then scans is exactly the saved state. */
int co_ivfpq_load_codes(co_ivfpq *x, const uint32_t *ids, const uint8_t *codes, const int *list_of, long n) {
    if (!x->trained) return CO_ERR_NOT_TRAINED;
    for (long i = 0; i < n; i++) {
        if (list_of[i] < 0 || list_of[i] >= x->nlist) return CO_ERR_ARG;
        list_push(&x->lists[list_of[i]], ids[i], NULL, 0, codes + (size_t)i * x->M, x->M);
    }
    return CO_OK;
}
int co_ivfpq_remove(co_ivfpq *x, uint32_t id) {
    if (!lists_contain(x->lists, x->nlist, id)) return CO_ERR_NOT_FOUND;
    if (idset_contains(&x->deleted, id)) return CO_ERR_NOT_FOUND;
    idset_add(&x->deleted, id);
    return CO_OK;
}
int co_ivfpq_flush(co_ivfpq *x) {
    if (x->deleted.n == 0) return CO_OK;
    for (int l = 0; l < x->nlist; l++) {
        inv_list *L = &x->lists[l];
        long w = 0;
        for (long i = 0; i < L->n; i++) {
            if (idset_contains(&x->deleted, L->ids[i])) continue;
            if (w != i) { L->ids[w] = L->ids[i]; memmove(L->codes + (size_t)w * x->M, L->codes + (size_t)i * x->M, (size_t)x->M); }
            w++;
        }
        L->n = w;
    }
    idset_clear(&x->deleted);
    return CO_OK;
}
int co_ivfpq_default_nprobes(const co_ivfpq *x) { return (int)sqrt((double)x->nlist); }  /* ivfpq_index.go:446 */
const float *co_ivfpq_centroids(const co_ivfpq *x) { return x->centroids; }
const float *co_ivfpq_codebooks(const co_ivfpq *x) { return x->codebooks; }
long co_ivfpq_list_len(const co_ivfpq *x, int l) { return x->lists[l].n; }
void co_ivfpq_list_get(const co_ivfpq *x, int l, uint32_t *ids, uint8_t *codes) {
    const inv_list *L = &x->lists[l];
    if (ids) memcpy(ids, L->ids, (size_t)L->n * sizeof(uint32_t));
    if (codes) memcpy(codes, L->codes, (size_t)L->n * x->M);
}

/* ivfpq_index_search.go:231-346 searchSingleQuery (+ :350-390 helpers) */
long co_ivfpq_search(const co_ivfpq *x, const float *query, long k_req, int nprobes, float threshold,
                     const uint32_t *filter_ids, long nfilter,
                     uint32_t *out_ids, float *out_scores) {
    if (!x->trained) return CO_ERR_NOT_TRAINED;
    int d = x->dim;
    if (nprobes <= 0 || nprobes > x->nlist) nprobes = x->nlist;       /* :246-249 */
    float *q = xmalloc((size_t)d * sizeof(float));
    int rc = co_preprocess(x->metric, query, q, d);
    if (rc) { free(q); return rc; }
    scored *cd = coarse_sort(q, x->centroids, x->nlist, d, x->metric);
    idset filt; idset_from(&filt, filter_ids, nfilter);
    long total = 0;
    for (int i = 0; i < nprobes; i++) total += x->lists[cd[i].id].n;
    scored *res = xmalloc((size_t)(total ? total : 1) * sizeof(scored));
    float *qr = xmalloc((size_t)d * sizeof(float));
    float *lut = xmalloc((size_t)x->M * x->Ksub * sizeof(float));
    long nres = 0, pos = 0;
    for (int i = 0; i < nprobes; i++) {                               /* :282-323 */
        int l = (int)cd[i].id;
        const float *c = x->centroids + (size_t)l * d;
        for (int j = 0; j < d; j++) qr[j] = q[j] - c[j];              /* :287-290 residual */
        pq_build_lut(x->codebooks, x->M, x->Ksub, x->dsub, qr, lut);  /* :293 new LUT per probe */
        const inv_list *L = &x->lists[l];
        for (long j = 0; j < L->n; j++, pos++) {
            uint32_t id = L->ids[j];
            if (x->deleted.n && idset_contains(&x->deleted, id)) continue;
            if (filter_should_skip(&filt, nfilter, id)) continue;
            float dist = pq_adc(lut, x->M, x->Ksub, L->codes + (size_t)j * x->M);
            if (threshold > 0 && dist > threshold) continue;
            res[nres].score = dist; res[nres].pos = pos; res[nres].id = id; nres++;
        }
    }
    tl_ivfpq_scanned = pos;
    qsort(res, (size_t)nres, sizeof(scored), cmp_scored);
    long k = co_sanitize_k(k_req, nres);
    for (long i = 0; i < k; i++) { out_ids[i] = res[i].id; out_scores[i] = res[i].score; }
    free(lut); free(qr); free(res); idset_free(&filt); free(cd); free(q);
    return k;
}

/* ------------------------------------------------------------------------------------ */
/* hnsw_index.go / hnsw_index_search.go                                                 */
/* ------------------------------------------------------------------------------------ */

typedef struct { uint32_t id; float distance; } cand;   /* hnsw_index_search.go:363-366 */

typedef struct { uint32_t *v; int n, cap; } edge_list;

typedef struct {
    uint32_t id;
    int level;
    edge_list *edges;   /* level+1 lists (hnsw_index.go:50-61) */
} hnode;

struct co_hnsw {
    int dim, metric, M, efc, efs;
    int max_level;          /* -1 when empty */
    uint32_t entry;
    long n, cap;
    hnode *nodes;           /* slot order == insertion order */
    float *rows;            /* n x dim */
    /* id -> slot map (stands in for map[uint32]*hnswNode) */
    long map_cap; long *map_slot; uint32_t *map_key;
    idset deleted;
};

static __thread long tl_hnsw_evals, tl_hnsw_expansions;
long co_hnsw_last_dist_evals(void) { return tl_hnsw_evals; }
long co_hnsw_last_expansions(void) { return tl_hnsw_expansions; }

static long hmap_find(const co_hnsw *h, uint32_t id) {
    if (!h->map_cap) return -1;
    long m = h->map_cap - 1;
    long p = (long)((id * 2654435761u) & (uint32_t)m);
    while (h->map_slot[p] >= 0) {
        if (h->map_key[p] == id) return h->map_slot[p];
        p = (p + 1) & m;
    }
    return -1;
}
static void hmap_insert_raw(co_hnsw *h, uint32_t id, long slot) {
    long m = h->map_cap - 1;
    long p = (long)((id * 2654435761u) & (uint32_t)m);
    while (h->map_slot[p] >= 0 && h->map_key[p] != id) p = (p + 1) & m;
    h->map_slot[p] = slot; h->map_key[p] = id;
}
static void hmap_insert(co_hnsw *h, uint32_t id, long slot) {
    if ((h->n + 1) * 2 > h->map_cap) {
        long ncap = h->map_cap ? h->map_cap * 2 : 1024;
        long *os = h->map_slot; uint32_t *ok = h->map_key; long oc = h->map_cap;
        h->map_cap = ncap;
        h->map_slot = xmalloc((size_t)ncap * sizeof(long));
        h->map_key = xmalloc((size_t)ncap * sizeof(uint32_t));
        for (long i = 0; i < ncap; i++) h->map_slot[i] = -1;
        for (long i = 0; i < oc; i++) if (os[i] >= 0) hmap_insert_raw(h, ok[i], os[i]);
        free(os); free(ok);
    }
    hmap_insert_raw(h, id, slot);
}

co_hnsw *co_hnsw_new(int dim, int metric, int m, int efc, int efs) {
    /* hnsw_index.go:172-208 */
    if (dim <= 0 || metric < 0 || metric > 2) return NULL;
    if (m <= 0) m = 16;
    if (efc <= 0) efc = 200;
    if (efs <= 0) efs = efc;
    co_hnsw *h = calloc(1, sizeof(*h));
    h->dim = dim; h->metric = metric; h->M = m; h->efc = efc; h->efs = efs;
    h->max_level = -1; h->entry = 0;
    return h;
}
void co_hnsw_free(co_hnsw *h) {
    if (!h) return;
    for (long i = 0; i < h->n; i++) {
        for (int l = 0; l <= h->nodes[i].level; l++) free(h->nodes[i].edges[l].v);
        free(h->nodes[i].edges);
    }
    free(h->nodes); free(h->rows); free(h->map_slot); free(h->map_key);
    idset_free(&h->deleted); free(h);
}
static void edges_push(edge_list *e, uint32_t id) {
    if (e->n == e->cap) { e->cap = e->cap ? e->cap * 2 : 8; e->v = xrealloc(e->v, (size_t)e->cap * sizeof(uint32_t)); }
    e->v[e->n++] = id;
}
static const float *hrow(const co_hnsw *h, long slot) { return h->rows + (size_t)slot * h->dim; }

/* Go container/heap on []candidate: Push = append + up; Pop = swap(0,n-1) + down(0,n-1) + drop last.
 * less(i,j) is strict: '<' for the min-heap, '>' for the max-heap (hnsw_index_search.go:371-444). */
typedef struct { cand *a; long n, cap; int is_max; } gheap;
static int gh_less(const gheap *h, long i, long j) {
    return h->is_max ? (h->a[i].distance > h->a[j].distance) : (h->a[i].distance < h->a[j].distance);
}
static void gh_swap(gheap *h, long i, long j) { cand t = h->a[i]; h->a[i] = h->a[j]; h->a[j] = t; }
static void gh_push(gheap *h, cand c) {
    if (h->n == h->cap) { h->cap = h->cap ? h->cap * 2 : 64; h->a = xrealloc(h->a, (size_t)h->cap * sizeof(cand)); }
    h->a[h->n++] = c;
    long j = h->n - 1;
    for (;;) {                               /* heap.up */
        long i = (j - 1) / 2;
        if (i == j || !gh_less(h, j, i)) break;
        gh_swap(h, i, j); j = i;
    }
}
static cand gh_pop(gheap *h) {
    long n = h->n - 1;
    gh_swap(h, 0, n);
    long i = 0;
    for (;;) {                               /* heap.down(0, n) */
        long j1 = 2 * i + 1;
        if (j1 >= n || j1 < 0) break;
        long j = j1;
        long j2 = j1 + 1;
        if (j2 < n && gh_less(h, j2, j1)) j = j2;
        if (!gh_less(h, j, i)) break;
        gh_swap(h, i, j); i = j;
    }
    h->n = n;
    return h->a[n];
}

/* hnsw_index.go:565-629 searchLayer.  Returns malloc'd candidates ascending by distance. */
static cand *hnsw_search_layer(const co_hnsw *h, const float *q, uint32_t entry, int ef, int layer,
                               long *n_out, uint8_t *visited /* n bytes, zeroed; restored on exit */) {
    gheap cands = {0}, result = {0};
    result.is_max = 1;
    long *touched = xmalloc((size_t)(h->n + 1) * sizeof(long));
    long ntouched = 0;
    long eslot = hmap_find(h, entry);
    if (!idset_contains(&h->deleted, entry)) {                       /* :580-585 */
        float d = co_distance(h->metric, q, hrow(h, eslot), h->dim);
        tl_hnsw_evals++;
        cand c = { entry, d };
        gh_push(&cands, c); gh_push(&result, c);
    }
    if (eslot >= 0) { visited[eslot] = 1; touched[ntouched++] = eslot; }
    while (cands.n > 0) {                                            /* :588 */
        cand cur = gh_pop(&cands);
        if (result.n >= ef && cur.distance > result.a[0].distance) break;   /* :592-594 */
        long cs = hmap_find(h, cur.id);
        const hnode *node = &h->nodes[cs];
        tl_hnsw_expansions++;
        if (layer < node->level + 1) {                               /* :597 */
            const edge_list *E = &node->edges[layer];
            for (int e = 0; e < E->n; e++) {
                uint32_t nid = E->v[e];
                if (h->deleted.n && idset_contains(&h->deleted, nid)) continue;  /* :600-602 */
                long ns = hmap_find(h, nid);
                if (ns < 0) continue;   /* cannot happen with non-zero unique IDs */
                if (!visited[ns]) {                                  /* :604-605 */
                    visited[ns] = 1; touched[ntouched++] = ns;
                    float d = co_distance(h->metric, q, hrow(h, ns), h->dim);
                    tl_hnsw_evals++;
                    if (result.n < ef || d < result.a[0].distance) { /* :609 */
                        cand c = { nid, d };
                        gh_push(&cands, c); gh_push(&result, c);
                        if (result.n > ef) (void)gh_pop(&result);    /* :613-615 */
                    }
                }
            }
        }
    }
    long n = result.n;
    cand *out = xmalloc((size_t)(n ? n : 1) * sizeof(cand));
    for (long i = n - 1; i >= 0; i--) out[i] = gh_pop(&result);      /* :622-626 */
    for (long i = 0; i < ntouched; i++) visited[touched[i]] = 0;
    free(touched); free(cands.a); free(result.a);
    *n_out = n;
    return out;
}

static int cmp_cand_stable(const void *a, const void *b) {
    /* candidates carry their original index in a parallel struct; see sort_cands_stable */
    const scored *x = a, *y = b;
    return cmp_scored(x, y);
}
/* sort.Slice(candidates, distance <) -- stable stand-in */
static void sort_cands_stable(cand *c, long n) {
    scored *s = xmalloc((size_t)(n ? n : 1) * sizeof(scored));
    for (long i = 0; i < n; i++) { s[i].score = c[i].distance; s[i].pos = i; s[i].id = c[i].id; }
    qsort(s, (size_t)n, sizeof(scored), cmp_cand_stable);
    for (long i = 0; i < n; i++) { c[i].id = s[i].id; c[i].distance = s[i].score; }
    free(s);
}

/* hnsw_index.go:675-694 pruneConnections.  `idx.nodes[nid] == nil` skips the node being
 * inserted: insertNode runs BEFORE idx.nodes[id] = node (hnsw_index.go:282-284). */
static void hnsw_prune(co_hnsw *h, long slot, int layer, int M) {
    hnode *node = &h->nodes[slot];
    edge_list *E = &node->edges[layer];
    cand *cl = xmalloc((size_t)(E->n ? E->n : 1) * sizeof(cand));
    long n = 0;
    for (int e = 0; e < E->n; e++) {
        long ns = hmap_find(h, E->v[e]);
        if (ns < 0) continue;
        cl[n].id = E->v[e];
        cl[n].distance = co_distance(h->metric, hrow(h, slot), hrow(h, ns), h->dim);
        n++;
    }
    sort_cands_stable(cl, n);
    long keep = n < M ? n : M;
    E->n = 0;
    for (long i = 0; i < keep; i++) edges_push(E, cl[i].id);
    free(cl);
}

/* hnsw_index.go:228-288 Add + :493-552 insertNode */
int co_hnsw_add(co_hnsw *h, uint32_t id, float *vec, int level) {
    if (id == 0) return CO_ERR_UNSUPPORTED;   /* reference quirk (d), SURVEY.md section 2.1 */
    if (level < 0 || level > 16) return CO_ERR_ARG;
    int d = h->dim;
    int rc = co_preprocess(h->metric, vec, vec, d);              /* :237 */
    if (rc) return rc;
    if (hmap_find(h, id) >= 0) return CO_ERR_ARG;  /* duplicate IDs: map overwrite, unsupported */
    if (h->n == h->cap) {
        h->cap = h->cap ? h->cap * 2 : 256;
        h->nodes = xrealloc(h->nodes, (size_t)h->cap * sizeof(hnode));
        h->rows = xrealloc(h->rows, (size_t)h->cap * d * sizeof(float));
    }
    long slot = h->n;
    memcpy(h->rows + (size_t)slot * d, vec, (size_t)d * sizeof(float));
    hnode *node = &h->nodes[slot];
    node->id = id; node->level = level;
    node->edges = calloc((size_t)level + 1, sizeof(edge_list));
    if (level > h->max_level) h->max_level = level;              /* :266-268: before insertNode */

    if (h->entry == 0 && h->n == 0) {                            /* :273-278: first node only */
        h->entry = id;
        hmap_insert(h, id, slot); h->n++;
        return CO_OK;
    }

    /* insertNode: the new node is NOT yet in the map */
    const float *nv = hrow(h, slot);
    uint32_t curr = h->entry;
    float curr_dist = co_distance(h->metric, nv, hrow(h, hmap_find(h, curr)), d);
    for (int lc = h->max_level; lc > level; lc--) {              /* :498-521 */
        int changed = 1;
        while (changed) {
            changed = 0;
            const hnode *cn = &h->nodes[hmap_find(h, curr)];
            if (lc < cn->level + 1) {
                const edge_list *E = &cn->edges[lc];
                for (int e = 0; e < E->n; e++) {
                    uint32_t nid = E->v[e];
                    if (h->deleted.n && idset_contains(&h->deleted, nid)) continue;
                    long ns = hmap_find(h, nid);
                    if (ns < 0) continue;
                    float dd = co_distance(h->metric, nv, hrow(h, ns), d);
                    if (dd < curr_dist) { curr_dist = dd; curr = nid; changed = 1; }
                }
            }
        }
    }
    uint8_t *visited = calloc((size_t)h->n + 1, 1);
    for (int lc = level; lc >= 0; lc--) {                        /* :524-551 */
        long nc = 0;
        cand *cands = hnsw_search_layer(h, nv, curr, h->efc, lc, &nc, visited);
        int M = h->M;
        if (lc == 0) M *= 2;
        /* selectNeighbors (hnsw_index.go:637-656): all if <= M, else sort (in place!) + first M */
        long nsel = nc;
        if (nc > M) { sort_cands_stable(cands, nc); nsel = M; }
        for (long s = 0; s < nsel; s++) {
            uint32_t nid = cands[s].id;
            edges_push(&h->nodes[slot].edges[lc], nid);
            long ns = hmap_find(h, nid);
            hnode *nb = &h->nodes[ns];
            if (lc <= nb->level) {
                edges_push(&nb->edges[lc], id);
                if (nb->edges[lc].n > M) hnsw_prune(h, ns, lc, M);
            }
        }
        if (nc > 0) curr = cands[0].id;                          /* :548-550 */
        free(cands);
    }
    free(visited);
    hmap_insert(h, id, slot); h->n++;                            /* :284 */
    return CO_OK;
}
int co_hnsw_add_batch(co_hnsw *h, const uint32_t *ids, float *rows, const int *levels, long n) {
    for (long i = 0; i < n; i++) {
        int rc = co_hnsw_add(h, ids[i], rows + (size_t)i * h->dim, levels[i]);
        if (rc) return rc;
    }
    return CO_OK;
}
/* Restore path (HNSWIndex.ReadFrom, hnsw_index.go:898-1096): nodes with their stored (already preprocessed)
 * vectors, levels and per-layer edge lists, entry point and max level as saved.  edge_off has one entry per
 * (slot, layer <= level) pair in that order, plus the end. */
int co_hnsw_load_graph(co_hnsw *h, long n, const uint32_t *ids, const float *rows, const int *levels,
                       const long long *edge_off, const uint32_t *edge_ids, uint32_t entry, int max_level) {
    if (h->n != 0) return CO_ERR_ARG;
    int d = h->dim;
    h->cap = n > 0 ? n : 1;
    h->nodes = xrealloc(h->nodes, (size_t)h->cap * sizeof(hnode));
    h->rows = xrealloc(h->rows, (size_t)h->cap * d * sizeof(float));
    memcpy(h->rows, rows, (size_t)n * d * sizeof(float));
    long pair = 0;
    for (long s = 0; s < n; s++) {
        if (levels[s] < 0 || levels[s] > 16 || ids[s] == 0) return CO_ERR_ARG;
        hnode *node = &h->nodes[s];
        node->id = ids[s]; node->level = levels[s];
        node->edges = calloc((size_t)levels[s] + 1, sizeof(edge_list));
        for (int l = 0; l <= levels[s]; l++, pair++)
            for (long long e = edge_off[pair]; e < edge_off[pair + 1]; e++) edges_push(&node->edges[l], edge_ids[e]);
        hmap_insert(h, ids[s], s);
        h->n++;
    }
    h->entry = entry; h->max_level = max_level;
    return CO_OK;
}
/* hnsw_index.go:300-318 Remove (soft) */
int co_hnsw_remove(co_hnsw *h, uint32_t id) {
    if (hmap_find(h, id) < 0) return CO_ERR_NOT_FOUND;
    if (idset_contains(&h->deleted, id)) return CO_ERR_NOT_FOUND;
    idset_add(&h->deleted, id);
    return CO_OK;
}
long co_hnsw_size(const co_hnsw *h) { return h->n; }
int co_hnsw_max_level(const co_hnsw *h) { return h->max_level; }
uint32_t co_hnsw_entry_point(const co_hnsw *h) { return h->entry; }
int co_hnsw_ef_search(const co_hnsw *h) { return h->efs; }
void co_hnsw_export_nodes(const co_hnsw *h, uint32_t *ids, int *levels, float *rows) {
    for (long i = 0; i < h->n; i++) {
        if (ids) ids[i] = h->nodes[i].id;
        if (levels) levels[i] = h->nodes[i].level;
    }
    if (rows) memcpy(rows, h->rows, (size_t)h->n * h->dim * sizeof(float));
}
int co_hnsw_edges(const co_hnsw *h, long slot, int layer, uint32_t *out_ids) {
    const hnode *nd = &h->nodes[slot];
    if (layer > nd->level) return 0;
    const edge_list *E = &nd->edges[layer];
    if (out_ids) memcpy(out_ids, E->v, (size_t)E->n * sizeof(uint32_t));
    return E->n;
}

/* hnsw_index_search.go:248-354 searchSingleQuery */
long co_hnsw_search(const co_hnsw *h, const float *query, long k_req, int ef_search, float threshold,
                    const uint32_t *filter_ids, long nfilter,
                    uint32_t *out_ids, float *out_scores) {
    tl_hnsw_evals = 0; tl_hnsw_expansions = 0;
    if (h->n == 0 || h->max_level == -1) return 0;                   /* :258-260 */
    int d = h->dim;
    float *q = xmalloc((size_t)d * sizeof(float));
    int rc = co_preprocess(h->metric, query, q, d);
    if (rc) { free(q); return rc; }
    uint32_t curr = h->entry;
    float curr_dist = co_distance(h->metric, q, hrow(h, hmap_find(h, curr)), d);
    tl_hnsw_evals++;
    for (int lc = h->max_level; lc > 0; lc--) {                      /* :274-296 */
        int changed = 1;
        while (changed) {
            changed = 0;
            const hnode *cn = &h->nodes[hmap_find(h, curr)];
            if (lc < cn->level + 1) {
                const edge_list *E = &cn->edges[lc];
                for (int e = 0; e < E->n; e++) {
                    uint32_t nid = E->v[e];
                    if (h->deleted.n && idset_contains(&h->deleted, nid)) continue;
                    long ns = hmap_find(h, nid);
                    if (ns < 0) continue;
                    float dd = co_distance(h->metric, q, hrow(h, ns), d);
                    tl_hnsw_evals++;
                    if (dd < curr_dist) { curr_dist = dd; curr = nid; changed = 1; }
                }
            }
        }
    }
    int ef = ef_search;
    if (ef <= 0) ef = h->efs;                                        /* :302-305 */
    uint8_t *visited = calloc((size_t)h->n + 1, 1);
    long nc = 0;
    cand *cands = hnsw_search_layer(h, q, curr, ef, 0, &nc, visited);
    idset filt; idset_from(&filt, filter_ids, nfilter);
    scored *res = xmalloc((size_t)(nc ? nc : 1) * sizeof(scored));
    long nres = 0;
    for (long i = 0; i < nc; i++) {                                  /* :321-335: post-filter */
        if (filter_should_skip(&filt, nfilter, cands[i].id)) continue;
        if (threshold > 0 && cands[i].distance > threshold) continue;
        res[nres].score = cands[i].distance; res[nres].pos = i; res[nres].id = cands[i].id; nres++;
    }
    qsort(res, (size_t)nres, sizeof(scored), cmp_scored);            /* :338-340 */
    long k = co_sanitize_k(k_req, nres);
    for (long i = 0; i < k; i++) { out_ids[i] = res[i].id; out_scores[i] = res[i].score; }
    free(res); idset_free(&filt); free(cands); free(visited); free(q);
    return k;
}
