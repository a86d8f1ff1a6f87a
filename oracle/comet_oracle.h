/*
 * comet_oracle.h -- CPU restatement of wizenheimer/comet's vector distance + ANN search path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity oracle: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (comet_b200/, include/comet_b200.h, libcomet_b200.so) never links, imports or calls it.
 *
 * Every function cites the reference file:line it restates (paths relative to the reference
 * repo root).  The reference is pure Go and there is no Go toolchain in this image, so the
 * oracle cannot be checked against the running reference; it is pinned instead against every
 * known-answer test the reference's own *_test.go files hold for this path
 * (tests/test_oracle_kat.py lists them one by one).  For PQ / IVFPQ / HNSW the reference's
 * tests pin only counts, error cases and "exact match ranks first" -- numeric parity for
 * those three is therefore "pinned by restatement only" (see DESIGN.md, section Oracle).
 *
 * Arithmetic model: IEEE-754 binary32, one accumulator, strictly sequential over the
 * dimension, multiply and add rounded separately (what gc emits on amd64 with the default
 * GOAMD64=v1).  co_set_fma(1) switches every `sum += a*b` to a fused multiply-add, which is
 * what gc emits on arm64 (the Go spec allows either).
 *
 * Tie order: the reference sorts with sort.Slice (unstable pdqsort; insertion sort -- hence
 * stable -- for n <= 12).  The oracle sorts stably, i.e. by (score, scan order), which is
 * exactly the reference's result whenever no two scores are bit-equal, and for n <= 12.
 */
#ifndef COMET_ORACLE_H
#define COMET_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { CO_L2 = 0, CO_L2SQ = 1, CO_COSINE = 2 };           /* distance.go:21-38 */
enum { CO_AGG_SUM = 0, CO_AGG_MAX = 1, CO_AGG_MEAN = 2 }; /* aggregation.go:20-40 */

/* error codes (negative returns) */
enum {
    CO_OK = 0,
    CO_ERR_ZERO_VECTOR = -1,   /* distance.go:9-12 ErrZeroVector */
    CO_ERR_DIM = -2,           /* "... dimension mismatch" */
    CO_ERR_NOT_TRAINED = -3,   /* "index must be trained ..." */
    CO_ERR_TOO_FEW = -4,       /* "need at least ..." */
    CO_ERR_ARG = -5,
    CO_ERR_NOT_FOUND = -6,
    CO_ERR_UNSUPPORTED = -7
};

void co_set_fma(int on);
int  co_get_fma(void);
void co_set_threads(int n);          /* only co_*_search_batch use threads (one query per thread) */

/* ---- distance.go ---- */
float co_distance(int metric, const float *a, const float *b, int d);
int   co_normalize(const float *in, float *out, int d);   /* in may equal out */
int   co_preprocess(int metric, const float *in, float *out, int d);
float co_norm(const float *v, int d);

/* ---- aggregation.go / limiter.go ---- */
long  co_sanitize_k(long k, long max_results);
long  co_aggregate(int kind, const uint32_t *ids, const float *scores, long n,
                   uint32_t *out_ids, float *out_scores);
long  co_autocut(const float *y, long n, int cutoff);

/* ---- clustering.go ---- */
/* vectors: n x d row-major (stride ld floats). centroids out: k x d. assign out: n ints. returns k used */
int   co_kmeans(const float *vectors, long n, int d, long ld, int k, int metric, int max_iter,
                float *centroids, int *assign);
int   co_nearest_centroid(const float *v, const float *centroids, int k, int d, int metric);

/* ---- flat_index.go / flat_index_search.go ---- */
typedef struct co_flat co_flat;
co_flat *co_flat_new(int dim, int metric);
void  co_flat_free(co_flat *);
int   co_flat_add(co_flat *, uint32_t id, float *vec /* preprocessed IN PLACE, F7 */);
int   co_flat_add_batch(co_flat *, const uint32_t *ids, float *rows, long n);
int   co_flat_remove(co_flat *, uint32_t id);
int   co_flat_flush(co_flat *);
long  co_flat_size(const co_flat *);
const float *co_flat_rows(const co_flat *);   /* n x dim, preprocessed */
const uint32_t *co_flat_ids(const co_flat *);
/* searchSingleQuery.  Returns number of results (<= k_eff) or a negative error. out_pos may be NULL. */
long  co_flat_search(const co_flat *, const float *query, long k, float threshold,
                     const uint32_t *filter_ids, long nfilter,
                     uint32_t *out_ids, float *out_scores, long *out_pos);
/* nq independent searchSingleQuery calls (threads across queries); out arrays nq x kcap; counts nq */
int   co_flat_search_batch(const co_flat *, const float *queries, long nq, long k, float threshold,
                           long kcap, uint32_t *out_ids, float *out_scores, long *counts);

/* ---- ivf_index.go / ivf_index_search.go ---- */
typedef struct co_ivf co_ivf;
co_ivf *co_ivf_new(int dim, int nlist, int metric);
void  co_ivf_free(co_ivf *);
int   co_ivf_train(co_ivf *, const float *rows, long n);
int   co_ivf_set_centroids(co_ivf *, const float *centroids);
int   co_ivf_add(co_ivf *, uint32_t id, float *vec);
int   co_ivf_add_batch(co_ivf *, const uint32_t *ids, float *rows, long n);
int   co_ivf_remove(co_ivf *, uint32_t id);
int   co_ivf_flush(co_ivf *);
int   co_ivf_default_nprobes(const co_ivf *);
const float *co_ivf_centroids(const co_ivf *);
long  co_ivf_list_len(const co_ivf *, int list);
/* copies list contents: ids[len], rows[len*dim] (either may be NULL) */
void  co_ivf_list_get(const co_ivf *, int list, uint32_t *ids, float *rows);
long  co_ivf_search(const co_ivf *, const float *query, long k, int nprobes, float threshold,
                    const uint32_t *filter_ids, long nfilter,
                    uint32_t *out_ids, float *out_scores);

/* ---- pq_index.go / pq_index_search.go ---- */
typedef struct co_pq co_pq;
co_pq *co_pq_new(int dim, int metric, int M, int nbits);
void  co_pq_free(co_pq *);
int   co_pq_train(co_pq *, const float *rows, long n);
int   co_pq_set_codebooks(co_pq *, const float *codebooks /* M x Ksub x dsub */);
int   co_pq_add(co_pq *, uint32_t id, float *vec);
int   co_pq_add_batch(co_pq *, const uint32_t *ids, float *rows, long n);
int   co_pq_remove(co_pq *, uint32_t id);
int   co_pq_flush(co_pq *);
long  co_pq_size(const co_pq *);
const float *co_pq_codebooks(const co_pq *);
const uint8_t *co_pq_codes(const co_pq *);   /* n x M */
const uint32_t *co_pq_ids(const co_pq *);
void  co_pq_encode(const co_pq *, const float *vec, uint8_t *code);
long  co_pq_search(const co_pq *, const float *query, long k, float threshold,
                   const uint32_t *filter_ids, long nfilter,
                   uint32_t *out_ids, float *out_scores);

/* ---- ivfpq_index.go / ivfpq_index_search.go ---- */
typedef struct co_ivfpq co_ivfpq;
co_ivfpq *co_ivfpq_new(int dim, int metric, int nlist, int M, int nbits);
void  co_ivfpq_free(co_ivfpq *);
int   co_ivfpq_train(co_ivfpq *, const float *rows, long n);
int   co_ivfpq_set_trained(co_ivfpq *, const float *centroids, const float *codebooks);
int   co_ivfpq_add(co_ivfpq *, uint32_t id, float *vec);
int   co_ivfpq_add_batch(co_ivfpq *, const uint32_t *ids, float *rows, long n);
int   co_ivfpq_load_codes(co_ivfpq *, const uint32_t *ids, const uint8_t *codes, const int *list_of, long n);  /* ReadFrom */
int   co_ivfpq_remove(co_ivfpq *, uint32_t id);
int   co_ivfpq_flush(co_ivfpq *);
int   co_ivfpq_default_nprobes(const co_ivfpq *);
const float *co_ivfpq_centroids(const co_ivfpq *);
const float *co_ivfpq_codebooks(const co_ivfpq *);
long  co_ivfpq_list_len(const co_ivfpq *, int list);
void  co_ivfpq_list_get(const co_ivfpq *, int list, uint32_t *ids, uint8_t *codes);
long  co_ivfpq_search(const co_ivfpq *, const float *query, long k, int nprobes, float threshold,
                      const uint32_t *filter_ids, long nfilter,
                      uint32_t *out_ids, float *out_scores);
/* work counters of the last co_ivfpq_search on this thread: codes scanned */
long  co_ivfpq_last_scanned(void);

/* ---- hnsw_index.go / hnsw_index_search.go ---- */
typedef struct co_hnsw co_hnsw;
co_hnsw *co_hnsw_new(int dim, int metric, int m, int ef_construction, int ef_search);
void  co_hnsw_free(co_hnsw *);
/* level = what randomLevel() drew (hnsw_index.go:474-484); the caller owns the RNG (F8). id != 0. */
int   co_hnsw_add(co_hnsw *, uint32_t id, float *vec, int level);
int   co_hnsw_add_batch(co_hnsw *, const uint32_t *ids, float *rows, const int *levels, long n);
int   co_hnsw_load_graph(co_hnsw *, long n, const uint32_t *ids, const float *rows, const int *levels,
                         const long long *edge_off, const uint32_t *edge_ids, uint32_t entry, int max_level);  /* ReadFrom */
int   co_hnsw_remove(co_hnsw *, uint32_t id);
long  co_hnsw_size(const co_hnsw *);
int   co_hnsw_max_level(const co_hnsw *);
uint32_t co_hnsw_entry_point(const co_hnsw *);
int   co_hnsw_ef_search(const co_hnsw *);
/* export in insertion (slot) order: ids[n], levels[n], rows[n*dim] (any may be NULL) */
void  co_hnsw_export_nodes(const co_hnsw *, uint32_t *ids, int *levels, float *rows);
/* number of edges of slot at layer (0 if layer > level) and copy of neighbour IDs */
int   co_hnsw_edges(const co_hnsw *, long slot, int layer, uint32_t *out_ids);
long  co_hnsw_search(const co_hnsw *, const float *query, long k, int ef_search, float threshold,
                     const uint32_t *filter_ids, long nfilter,
                     uint32_t *out_ids, float *out_scores);
/* work counters of the last co_hnsw_search on this thread */
long  co_hnsw_last_dist_evals(void);
long  co_hnsw_last_expansions(void);

#ifdef __cplusplus
}
#endif
#endif
