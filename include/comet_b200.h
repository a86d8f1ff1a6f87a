/*
 * comet_b200.h -- C ABI of libcomet_b200.so: the B200 (sm_100a) device layer that replaces the
 * arithmetic loops of wizenheimer/comet's vector search path.  This header is the drop-in
 * boundary: every entry point below is what a cgo binding in comet's own Go package would call
 * in place of the reference function it cites (file:line, relative to the reference root).
 * INTEGRATION.md shows the Go side.
 *
 * Conventions
 *   - every function returns an int status, CM_OK == 0; cm_last_error() returns a thread-local
 *     message for the last non-zero status on the calling thread;
 *   - no function keeps a caller pointer after it returns; outputs are caller-allocated;
 *   - `*_search` entry points are re-entrant (the Go side holds idx.mu.RLock, e.g.
 *     flat_index_search.go:222); mutators need external exclusion (the Go side holds idx.mu.Lock,
 *     e.g. flat_index.go:171);
 *   - host-pointer entry points copy host<->device themselves; `*_device` variants take device
 *     pointers and a cudaStream_t (as void*) and enqueue without synchronising;
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     CM_ERR_CUDA.
 *
 * Result order: ascending by (score, scan position) -- the reference's sort.Slice result up to
 * groups of bit-equal scores (it is an unstable sort; for n <= 12 it is insertion sort and equals
 * this order exactly).  Scores are bit-identical to the reference's sequential float32 loops.
 */
#ifndef COMET_B200_H
#define COMET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* distance.go:21-38 DistanceKind "l2" | "l2_squared" | "cosine" */
enum { CM_L2 = 0, CM_L2SQ = 1, CM_COSINE = 2 };

enum {
    CM_OK = 0,
    CM_ERR_INVALID_ARG = 1,
    CM_ERR_DIM_MISMATCH = 2, /* "query dimension mismatch: expected %d, got %d" flat_index_search.go:227 */
    CM_ERR_ZERO_VECTOR = 3,  /* distance.go:9-12 ErrZeroVector */
    CM_ERR_NOT_TRAINED = 4,  /* "index must be trained before searching" ivf_index_search.go:223 */
    CM_ERR_NOT_FOUND = 5,    /* "vector with ID %d not found" flat_index.go:236 */
    CM_ERR_CUDA = 6,
    CM_ERR_UNSUPPORTED = 7,
    CM_ERR_TOO_FEW = 8,      /* "need at least %d training vectors" ivf_index.go:211 */
    CM_ERR_BUFFER_TOO_SMALL = 9
};

/* which device pipeline a flat search uses */
enum {
    CM_PATH_AUTO = 0,   /* exact scan for small batches, tensor-core candidate pass + exact re-score for large */
    CM_PATH_EXACT = 1,  /* reference-order fp32 scan of every row */
    CM_PATH_TENSOR = 2  /* bf16 tcgen05 Q x X^T candidate pass, then reference-order re-score */
};

/* how `sum += a*b` rounds: separately (gc on amd64, the default) or fused (gc on arm64) */
enum { CM_ROUND_SEPARATE = 0, CM_ROUND_FMA = 1 };

typedef struct {
    int64_t k;                  /* WithK: <= 0 or > n means "all" (limiter.go:12-17 sanitizeK) */
    float threshold;            /* WithThreshold: filters only when > 0 (flat_index_search.go:269) */
    int32_t nprobes;            /* WithNProbes (IVF, IVFPQ): <= 0 or > nlist means nlist */
    int32_t ef_search;          /* WithEfSearch (HNSW): <= 0 means the index default */
    const uint32_t *filter_ids; /* WithDocumentIDs: NULL/0 means every document is eligible */
    int64_t nfilter;
    int32_t path;               /* CM_PATH_* (flat only) */
    int32_t reserved;
} cm_search_params;

typedef struct cm_flat cm_flat;
typedef struct cm_ivf cm_ivf;
typedef struct cm_pq cm_pq;
typedef struct cm_ivfpq cm_ivfpq;
typedef struct cm_hnsw cm_hnsw;
typedef struct cm_flat_batcher cm_flat_batcher;
typedef struct cm_flat_sharded cm_flat_sharded;
typedef struct cm_ivf_sharded cm_ivf_sharded;
typedef struct cm_pq_sharded cm_pq_sharded;
typedef struct cm_ivfpq_sharded cm_ivfpq_sharded;

/* ---- runtime ------------------------------------------------------------------------------ */
int cm_init(const int *device_ids, int n_devices); /* NULL/0: use the current device */
void cm_shutdown(void);
const char *cm_last_error(void);
int cm_device_count(void);
int cm_set_rounding(int mode);                     /* CM_ROUND_*; process-wide */
int cm_get_rounding(void);
const char *cm_version(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
int64_t cm_kernel_launches(void);
/* per-kernel-class device timing (CUDA events on the launching stream); off by default */
enum {
    CM_PROF_FLAT_SCAN = 0,   /* exact flat scan (flat_scan_kernel) */
    CM_PROF_FLAT_GEMM = 1,   /* tcgen05 bf16 candidate pass */
    CM_PROF_RESCORE = 2,     /* reference-order re-score of candidates */
    CM_PROF_SELECT = 3,      /* top-K merge / selection */
    CM_PROF_IVF_SCAN = 4,
    CM_PROF_PQ_SCAN = 5,
    CM_PROF_HNSW = 6,
    CM_PROF_COARSE = 7,
    CM_PROF_CLASSES = 8
};
int cm_profile_enable(int on);
int cm_profile_reset(void);
int cm_profile_get(int kernel_class, double *total_ms, int64_t *launches);
/* pinned host memory for callers that want async copies (bench.py e2e leg) */
int cm_host_alloc(void **ptr, size_t bytes);
int cm_host_free(void *ptr);

/* ---- distance.go --------------------------------------------------------------------------- */
/* Distance.Calculate for n independent pairs a[i], b[i] (distance.go:114-121, 158-165, 201-216) */
int cm_distance_pairs(int metric, const float *a, const float *b, int64_t n, int dim, float *out);
/* Distance.PreprocessInPlace on n rows (distance.go:244-264); first zero vector -> CM_ERR_ZERO_VECTOR
 * with *bad_row set (rows before it are normalised, like n successive calls) */
int cm_preprocess_rows(int metric, float *rows, int64_t n, int dim, int64_t *bad_row);

/* ---- flat_index.go / flat_index_search.go -------------------------------------------------- */
int cm_flat_create(int dim, int metric, cm_flat **out);           /* NewFlatIndex flat_index.go:118 */
int cm_flat_destroy(cm_flat *h);
int cm_flat_reserve(cm_flat *h, int64_t n_rows);
/* n successive FlatIndex.Add calls (flat_index.go:169-189).  `rows` is preprocessed IN PLACE
 * (cosine normalises the caller's buffer) unless writeback == 0. */
int cm_flat_add(cm_flat *h, const uint32_t *ids, float *rows, int64_t n, int writeback);
/* Restore STORED (already preprocessed) vectors as they are (appended; a unit vector is not normalised twice): the
 * decoded-state half of FlatIndex.ReadFrom for hosts that parse the stream themselves.  cm_flat_load below takes
 * the bytes. */
int cm_flat_load_rows(cm_flat *h, const uint32_t *ids, const float *rows, int64_t n);
/* rows already resident on the device (device pointer, row-major n x dim) */
int cm_flat_add_device(cm_flat *h, const uint32_t *ids_host, const float *rows_dev, int64_t n, void *stream);
int cm_flat_get_ids(const cm_flat *h, int64_t first, int64_t n, uint32_t *out);   /* node IDs by scan position */
int cm_flat_remove(cm_flat *h, uint32_t id);                      /* flat_index.go:219-250 */
int cm_flat_flush(cm_flat *h);                                    /* flat_index.go:266-299 */
int64_t cm_flat_size(const cm_flat *h);                           /* len(idx.vectors), deleted included */
int cm_flat_dim(const cm_flat *h);
int cm_flat_metric(const cm_flat *h);
/* lookupNodeVectors (flat_index_search.go:171-196): stored (preprocessed) vector of a live node */
int cm_flat_get_vector(const cm_flat *h, uint32_t id, float *out);
/* stored vectors by scan position (VectorResult.Node.Vector(), index_search.go:84-90) */
int cm_flat_get_rows(const cm_flat *h, const int64_t *positions, int64_t n, float *out);
/* nq independent searchSingleQuery calls (flat_index_search.go:221-294) in one device batch.
 * out_ids/out_scores are nq x out_stride, out_counts[nq]; out_stride >= sanitizeK(k, n).
 * out_pos (optional, nq x out_stride) receives scan positions. */
int cm_flat_search(cm_flat *h, const float *queries, int64_t nq, int dim, const cm_search_params *p,
                   int64_t out_stride, uint32_t *out_ids, float *out_scores, int64_t *out_pos,
                   int64_t *out_counts);
/* Device-pointer variant: enqueues on `stream`, never synchronises, so it cannot do what the host entry point
 * does after its synchronise.  Two per-query conditions are therefore REPORTED in out_counts_dev instead:
 *   -1  the tensor path's candidate lists overflowed for this query (adversarial ties): its row of results is
 *       undefined and the caller must redo the query with p->path = CM_PATH_EXACT (cm_flat_search does);
 *   -2  cosine: the query is a zero vector (Distance.Preprocess fails with ErrZeroVector, distance.go:269-290):
 *       its row of results is undefined (cm_flat_search returns CM_ERR_ZERO_VECTOR). */
int cm_flat_search_device(cm_flat *h, const float *queries_dev, int64_t nq, int dim,
                          const cm_search_params *p, int64_t out_stride, uint32_t *out_ids_dev,
                          float *out_scores_dev, int64_t *out_pos_dev, int64_t *out_counts_dev,
                          void *stream);
/* Multi-GPU flat search: the corpus is row-sharded, shard r (rank r) holding rows [r*n/W, (r+1)*n/W) in
 * scan order.  Every rank searches its shard, the [nq][in_stride] result lists are all-gathered
 * (NCCL) into [W][nq][in_stride] device buffers, and this merges them into the global result in the
 * reference's order (score, scan position) == (score, shard, rank within the shard's list).
 * counts_dev: [W][nq] valid entries per list, or NULL when every list is full.  Replaces nothing in
 * the single-process reference; it is the path's one exchange step. */
int cm_merge_shards_device(const uint32_t *ids_dev, const float *scores_dev, const int64_t *counts_dev, int world,
                           int64_t nq, int64_t in_stride, int64_t k, int64_t out_stride, uint32_t *out_ids_dev,
                           float *out_scores_dev, int64_t *out_counts_dev, void *stream);
/* ---- row-sharded FlatIndex over the GPUs of one box, ONE host process ---------- */
/* The Go host is a single process, so the multi-GPU layout lives behind this ABI: NewFlatIndex on a box with W
 * GPUs binds cm_flat_sharded_create instead of cm_flat_create and nothing else changes for its callers.
 * Shard r lives on devices[r] (the same device may appear more than once) and holds up to rows_per_shard rows;
 * rows fill shard 0 first, then shard 1, ...: the reference's result order (score, scan position) is then
 * (score, shard, rank within the shard's list), so the per-shard top-K lists merge without touching a row again.
 * One search = queries to every shard over NVLink peer copies, cm_flat_search_device on every shard concurrently
 * (one stream per device, CUDA events for ordering), the [nq][K] lists back to devices[0], one merge kernel
 * (flat_index_search.go:277-291 on the union).  No NCCL: a single process owns the devices.
 * Searches on one handle are serialised (they share the gather buffers); mutators need external exclusion. */
int cm_flat_sharded_create(int dim, int metric, const int *devices, int n_devices, int64_t rows_per_shard,
                           cm_flat_sharded **out);
int cm_flat_sharded_destroy(cm_flat_sharded *h);
int cm_flat_sharded_shards(const cm_flat_sharded *h);
int64_t cm_flat_sharded_size(const cm_flat_sharded *h);                 /* len(idx.vectors) over all shards */
int cm_flat_sharded_shard_size(const cm_flat_sharded *h, int shard, int64_t *rows);
int cm_flat_sharded_reserve(cm_flat_sharded *h, int64_t n_rows);
/* n successive FlatIndex.Add calls (flat_index.go:169-189); rows written back like cm_flat_add */
int cm_flat_sharded_add(cm_flat_sharded *h, const uint32_t *ids, float *rows, int64_t n, int writeback);
/* rows already resident on the device of shard `shard` (filling several shards in parallel is allowed as long as
 * every shard before the last non-empty one is full before the first search) */
int cm_flat_sharded_add_device(cm_flat_sharded *h, int shard, const uint32_t *ids_host, const float *rows_dev,
                               int64_t n, void *stream);
int cm_flat_sharded_remove(cm_flat_sharded *h, uint32_t id);            /* flat_index.go:219-250 */
int cm_flat_sharded_flush(cm_flat_sharded *h);                          /* flat_index.go:266-299, shard by shard */
/* nq independent searchSingleQuery calls against the whole sharded corpus; same contract as cm_flat_search
 * (a query a shard's tensor path could not answer is redone on the exact path of every shard; a zero query under
 * cosine fails the call with CM_ERR_ZERO_VECTOR). */
int cm_flat_sharded_search(cm_flat_sharded *h, const float *queries, int64_t nq, int dim, const cm_search_params *p,
                           int64_t out_stride, uint32_t *out_ids, float *out_scores, int64_t *out_counts);
/* queries / outputs on devices[0], enqueued on `stream` (a stream of devices[0]); never synchronises; per-query
 * conditions are reported as counts -1 / -2 exactly like cm_flat_search_device */
int cm_flat_sharded_search_device(cm_flat_sharded *h, const float *queries_dev, int64_t nq, int dim,
                                  const cm_search_params *p, int64_t out_stride, uint32_t *out_ids_dev,
                                  float *out_scores_dev, int64_t *out_counts_dev, void *stream);
/* bytes that crossed NVLink in the last search (queries out + lists back; shards on devices[0] move nothing) */
int64_t cm_flat_sharded_last_exchange_bytes(const cm_flat_sharded *h);
/* device-side durations of the last search (CUDA events; waits for the search to finish): the slowest shard's
 * local search, the slowest shard's copy of its lists to devices[0], the merge kernel */
int cm_flat_sharded_last_timing(cm_flat_sharded *h, double *search_ms_max, double *gather_ms_max, double *merge_ms);
/* Cross-call dynamic batching.  The reference's callers issue one query per Execute()
 * (flat_index_search.go:109-165), often from many goroutines; cm_flat_batcher_search blocks its caller while a
 * worker coalesces the concurrent requests that share (k, threshold) into ONE device batch (at most max_batch
 * requests, the oldest waits at most max_wait_us) and hands each caller its own result.  Same results as a direct
 * cm_flat_search.  The batcher must be destroyed before the index. */
int cm_flat_batcher_create(cm_flat *index, int max_batch, int max_wait_us, cm_flat_batcher **out);
int cm_flat_batcher_destroy(cm_flat_batcher *b);
int cm_flat_batcher_search(cm_flat_batcher *b, const float *query, int dim, int64_t k, float threshold, int64_t out_stride,
                           uint32_t *out_ids, float *out_scores, int64_t *out_count);
int cm_flat_batcher_stats(cm_flat_batcher *b, int64_t *batches, int64_t *requests);
/* statistics of the last search on this handle (for bench.py / profiles): */
typedef struct {
    int32_t path_used;          /* CM_PATH_EXACT or CM_PATH_TENSOR */
    int32_t passes;             /* scan launches over the corpus */
    int64_t candidates;         /* tensor path: (query,row) pairs re-scored exactly */
    int64_t fallback_queries;   /* tensor path: queries redone by the exact scan (candidate overflow) */
    int64_t kernel_launches;
} cm_flat_stats;
int cm_flat_last_stats(const cm_flat *h, cm_flat_stats *out);
/* summed over the shards of the last sharded search */
int cm_flat_sharded_last_stats(const cm_flat_sharded *h, cm_flat_stats *out);

/* ---- ivf_index.go / ivf_index_search.go ----------------------------------------------------- */
int cm_ivf_create(int dim, int nlist, int metric, cm_ivf **out);     /* NewIVFIndex ivf_index.go:147 */
int cm_ivf_destroy(cm_ivf *h);
/* the result of IVFIndex.Train (ivf_index.go:205-246): nlist x dim centroids, stored as given */
int cm_ivf_set_centroids(cm_ivf *h, const float *centroids);
/* IVFIndex.Train on the device: the reference's deterministic KMeans (clustering.go:119-239), bit for bit */
int cm_ivf_train(cm_ivf *h, const float *rows, int64_t n);
int cm_ivf_get_centroids(const cm_ivf *h, float *out);                /* nlist x dim */
int cm_ivf_trained(const cm_ivf *h);
int64_t cm_ivf_size(const cm_ivf *h);
int cm_ivf_default_nprobes(const cm_ivf *h);                        /* ivf_index.go:410 int(sqrt(nlist)) */
/* n successive IVFIndex.Add calls (ivf_index.go:258-283): PreprocessInPlace (rows written back unless
 * writeback == 0), FindNearestCentroidIndex (clustering.go:252-272), append to the list.
 * out_lists (optional, n entries) receives the list each row went to. */
int cm_ivf_add(cm_ivf *h, const uint32_t *ids, float *rows, int64_t n, int writeback, int32_t *out_lists);
/* vectors of probed lists scanned by the last search on this handle (all queries): the measured
 * factor of the scan's algorithmic bytes */
int64_t cm_ivf_last_scanned(const cm_ivf *h);
/* IVFIndex.ReadFrom (ivf_index.go:611-785): stored vectors + the list of each, in list-append order */
int cm_ivf_load_lists(cm_ivf *h, const uint32_t *ids, const float *rows, const int32_t *list_of, int64_t n);
int cm_ivf_remove(cm_ivf *h, uint32_t id);                          /* ivf_index.go:296-330 */
int cm_ivf_flush(cm_ivf *h);                                        /* ivf_index.go:342-390 */
int cm_ivf_get_rows(const cm_ivf *h, const int64_t *positions, int64_t n, float *out);
int cm_ivf_get_ids(const cm_ivf *h, int64_t first, int64_t n, uint32_t *out);   /* node IDs by store position */
/* nq independent searchSingleQuery calls (ivf_index_search.go:217-322); p->nprobes as WithNProbes.
 * out_stride >= min(k, B) where B = the most candidates nprobes lists can hold (k <= 0: B itself). */
int cm_ivf_search(cm_ivf *h, const float *queries, int64_t nq, int dim, const cm_search_params *p,
                  int64_t out_stride, uint32_t *out_ids, float *out_scores, int64_t *out_pos,
                  int64_t *out_counts);
int cm_ivf_search_device(cm_ivf *h, const float *queries_dev, int64_t nq, int dim, const cm_search_params *p,
                         int64_t out_stride, uint32_t *out_ids_dev, float *out_scores_dev,
                         int64_t *out_pos_dev, int64_t *out_counts_dev, void *stream);

/* ---- IVF list shards over the GPUs of one box, ONE host process ------------------------------------ */
/* NewIVFIndex on a box with W GPUs: every shard (devices[r]) holds ALL centroids -- the coarse step is replicated -- and
 * the inverted lists it owns (list l starts on shard l mod W; cm_ivf_sharded_rebalance reassigns greedily by length).
 * A search scans, on every shard concurrently, the probed lists that shard holds; each shard returns its top-K together
 * with every winner's number in the reference's append loop over ALL probed lists (ivf_index_search.go:252-268), computed
 * from the replicated global list lengths; devices[0] merges by (score, that number): bit for bit the single index's
 * answer, ties included.  Queries are read and lists written through NVLink peer mappings (no NCCL: one process).
 * Searches on one handle are serialised; mutators need external exclusion.  k <= 0 ("all") is limited to
 * W x k <= 28K entries per query. */
int cm_ivf_sharded_create(int dim, int nlist, int metric, const int *devices, int n_devices, cm_ivf_sharded **out);
int cm_ivf_sharded_destroy(cm_ivf_sharded *h);
int cm_ivf_sharded_shards(const cm_ivf_sharded *h);
int cm_ivf_sharded_set_centroids(cm_ivf_sharded *h, const float *centroids);      /* result of IVFIndex.Train, to every shard */
int cm_ivf_sharded_train(cm_ivf_sharded *h, const float *rows, int64_t n);        /* IVFIndex.Train on devices[0] */
int cm_ivf_sharded_get_centroids(const cm_ivf_sharded *h, float *out);
int cm_ivf_sharded_trained(const cm_ivf_sharded *h);
int64_t cm_ivf_sharded_size(const cm_ivf_sharded *h);
int cm_ivf_sharded_shard_size(const cm_ivf_sharded *h, int shard, int64_t *rows);
int cm_ivf_sharded_owner(const cm_ivf_sharded *h, int list);                      /* shard holding the list */
int cm_ivf_sharded_default_nprobes(const cm_ivf_sharded *h);
/* n successive IVFIndex.Add calls (ivf_index.go:258-283): preprocessing and list assignment once on devices[0], then
 * every stored vector goes to the shard that owns its list */
int cm_ivf_sharded_add(cm_ivf_sharded *h, const uint32_t *ids, float *rows, int64_t n, int writeback, int32_t *out_lists);
int cm_ivf_sharded_remove(cm_ivf_sharded *h, uint32_t id);
int cm_ivf_sharded_flush(cm_ivf_sharded *h);
/* greedy-by-length list assignment (longest list first onto the lightest shard); lists that change owner move whole */
int cm_ivf_sharded_rebalance(cm_ivf_sharded *h);
int cm_ivf_sharded_search(cm_ivf_sharded *h, const float *queries, int64_t nq, int dim, const cm_search_params *p,
                          int64_t out_stride, uint32_t *out_ids, float *out_scores, int64_t *out_counts);
/* queries / outputs on devices[0]; never synchronises; a zero query under cosine is not detected on this entry point */
int cm_ivf_sharded_search_device(cm_ivf_sharded *h, const float *queries_dev, int64_t nq, int dim, const cm_search_params *p,
                                 int64_t out_stride, uint32_t *out_ids_dev, float *out_scores_dev, int64_t *out_counts_dev,
                                 void *stream);
/* vectors each shard scanned in the last host search, all queries together: per_shard[W] (balance of the assignment) */
int cm_ivf_sharded_last_scanned(const cm_ivf_sharded *h, int64_t *per_shard);

/* ---- pq_index.go / pq_index_search.go ------------------------------------------------------- */
int cm_pq_create(int dim, int metric, int M, int nbits, cm_pq **out);   /* NewPQIndex pq_index.go:135 */
int cm_pq_destroy(cm_pq *h);
/* the result of PQIndex.Train (pq_index.go:193-247): M x Ksub x dsub codebook floats */
int cm_pq_set_codebooks(cm_pq *h, const float *codebooks);
/* PQIndex.Train on the device: per sub-space KMeansSubspace on the RAW vectors (pq_index.go:210-243) */
int cm_pq_train(cm_pq *h, const float *rows, int64_t n);
int cm_pq_get_codebooks(const cm_pq *h, float *out);
int cm_pq_trained(const cm_pq *h);
int64_t cm_pq_size(const cm_pq *h);
/* n successive PQIndex.Add calls (pq_index.go:262-292): PreprocessInPlace + encode (pq_index.go:439-473) */
int cm_pq_add(cm_pq *h, const uint32_t *ids, float *rows, int64_t n, int writeback);
int cm_pq_get_codes(const cm_pq *h, int64_t first, int64_t n, uint8_t *out);   /* idx.codes[first:first+n] */
int cm_pq_get_ids(const cm_pq *h, int64_t first, int64_t n, uint32_t *out);     /* node IDs by store position */
int cm_pq_load_codes(cm_pq *h, const uint32_t *ids, const uint8_t *codes, int64_t n);   /* PQIndex.ReadFrom pq_index.go:652-846 */
int cm_pq_remove(cm_pq *h, uint32_t id);
int cm_pq_flush(cm_pq *h);
/* nq independent searchSingleQuery calls (pq_index_search.go:218-325) */
int cm_pq_search(cm_pq *h, const float *queries, int64_t nq, int dim, const cm_search_params *p, int64_t out_stride,
                 uint32_t *out_ids, float *out_scores, int64_t *out_pos, int64_t *out_counts);
int cm_pq_search_device(cm_pq *h, const float *queries_dev, int64_t nq, int dim, const cm_search_params *p,
                        int64_t out_stride, uint32_t *out_ids_dev, float *out_scores_dev, int64_t *out_pos_dev,
                        int64_t *out_counts_dev, void *stream);

/* ---- PQIndex row shards over the GPUs of one box, ONE host process ---------------------------------- */
/* PQ shards by rows exactly like flat (the candidate number of a PQ result IS its store position, so the per-shard lists
 * merge by (score, shard, rank) into the single index's order -- ADC scores tie constantly); the codebooks are replicated.
 * Same exchange as cm_flat_sharded_*: queries read and lists written through NVLink peer mappings, one merge kernel. */
int cm_pq_sharded_create(int dim, int metric, int M, int nbits, const int *devices, int n_devices, int64_t rows_per_shard,
                         cm_pq_sharded **out);
int cm_pq_sharded_destroy(cm_pq_sharded *h);
int cm_pq_sharded_shards(const cm_pq_sharded *h);
int64_t cm_pq_sharded_size(const cm_pq_sharded *h);
int cm_pq_sharded_trained(const cm_pq_sharded *h);
int cm_pq_sharded_train(cm_pq_sharded *h, const float *rows, int64_t n);            /* PQIndex.Train on devices[0], then replicated */
int cm_pq_sharded_set_codebooks(cm_pq_sharded *h, const float *codebooks);
int cm_pq_sharded_get_codebooks(const cm_pq_sharded *h, float *out);
int cm_pq_sharded_add(cm_pq_sharded *h, const uint32_t *ids, float *rows, int64_t n, int writeback);
int cm_pq_sharded_remove(cm_pq_sharded *h, uint32_t id);
int cm_pq_sharded_flush(cm_pq_sharded *h);
int cm_pq_sharded_search(cm_pq_sharded *h, const float *queries, int64_t nq, int dim, const cm_search_params *p, int64_t out_stride,
                         uint32_t *out_ids, float *out_scores, int64_t *out_counts);
int cm_pq_sharded_search_device(cm_pq_sharded *h, const float *queries_dev, int64_t nq, int dim, const cm_search_params *p,
                                int64_t out_stride, uint32_t *out_ids_dev, float *out_scores_dev, int64_t *out_counts_dev, void *stream);

/* ---- ivfpq_index.go / ivfpq_index_search.go ------------------------------------------------- */
int cm_ivfpq_create(int dim, int metric, int nlist, int M, int nbits, cm_ivfpq **out);   /* NewIVFPQIndex ivfpq_index.go:114 */
int cm_ivfpq_destroy(cm_ivfpq *h);
/* the result of IVFPQIndex.Train (ivfpq_index.go:180-259): nlist x dim centroids + residual codebooks */
int cm_ivfpq_set_trained(cm_ivfpq *h, const float *centroids, const float *codebooks);
/* IVFPQIndex.Train on the device: KMeans, assignment, residuals, per sub-space KMeansSubspace */
int cm_ivfpq_train(cm_ivfpq *h, const float *rows, int64_t n);
int cm_ivfpq_get_trained(const cm_ivfpq *h, float *centroids, float *codebooks);   /* either may be NULL */
int cm_ivfpq_trained(const cm_ivfpq *h);
int64_t cm_ivfpq_size(const cm_ivfpq *h);
int cm_ivfpq_default_nprobes(const cm_ivfpq *h);                        /* ivfpq_index.go:446 */
/* n successive IVFPQIndex.Add calls (ivfpq_index.go:279-319): preprocess, nearest centroid, residual, encode */
int cm_ivfpq_add(cm_ivfpq *h, const uint32_t *ids, float *rows, int64_t n, int writeback, int32_t *out_lists);
int cm_ivfpq_get_codes(const cm_ivfpq *h, int64_t first, int64_t n, uint8_t *out);   /* codes in arrival order */
int cm_ivfpq_get_ids(const cm_ivfpq *h, int64_t first, int64_t n, uint32_t *out);
int cm_ivfpq_load_codes(cm_ivfpq *h, const uint32_t *ids, const uint8_t *codes, const int32_t *list_of, int64_t n);
int64_t cm_ivfpq_last_scanned(const cm_ivfpq *h);                    /* codes scanned by the last search */
int cm_ivfpq_remove(cm_ivfpq *h, uint32_t id);
int cm_ivfpq_flush(cm_ivfpq *h);
/* nq independent searchSingleQuery calls (ivfpq_index_search.go:231-390) */
int cm_ivfpq_search(cm_ivfpq *h, const float *queries, int64_t nq, int dim, const cm_search_params *p,
                    int64_t out_stride, uint32_t *out_ids, float *out_scores, int64_t *out_pos,
                    int64_t *out_counts);
int cm_ivfpq_search_device(cm_ivfpq *h, const float *queries_dev, int64_t nq, int dim, const cm_search_params *p,
                           int64_t out_stride, uint32_t *out_ids_dev, float *out_scores_dev,
                           int64_t *out_pos_dev, int64_t *out_counts_dev, void *stream);

/* ---- IVFPQ list shards over the GPUs of one box, ONE host process ------------------------------------ */
/* Same layout and merge as cm_ivf_sharded_* (global candidate numbers from the replicated list lengths,
 * ivfpq_index_search.go:263-322); centroids and residual codebooks are replicated, Add encodes once on devices[0] and
 * routes (id, code) to the shard that owns the list. */
int cm_ivfpq_sharded_create(int dim, int metric, int nlist, int M, int nbits, const int *devices, int n_devices,
                            cm_ivfpq_sharded **out);
int cm_ivfpq_sharded_destroy(cm_ivfpq_sharded *h);
int cm_ivfpq_sharded_shards(const cm_ivfpq_sharded *h);
int cm_ivfpq_sharded_set_trained(cm_ivfpq_sharded *h, const float *centroids, const float *codebooks);
int cm_ivfpq_sharded_train(cm_ivfpq_sharded *h, const float *rows, int64_t n);
int cm_ivfpq_sharded_get_trained(const cm_ivfpq_sharded *h, float *centroids, float *codebooks);
int cm_ivfpq_sharded_trained(const cm_ivfpq_sharded *h);
int64_t cm_ivfpq_sharded_size(const cm_ivfpq_sharded *h);
int cm_ivfpq_sharded_shard_size(const cm_ivfpq_sharded *h, int shard, int64_t *rows);
int cm_ivfpq_sharded_owner(const cm_ivfpq_sharded *h, int list);
int cm_ivfpq_sharded_default_nprobes(const cm_ivfpq_sharded *h);
int cm_ivfpq_sharded_add(cm_ivfpq_sharded *h, const uint32_t *ids, float *rows, int64_t n, int writeback, int32_t *out_lists);
int cm_ivfpq_sharded_remove(cm_ivfpq_sharded *h, uint32_t id);
int cm_ivfpq_sharded_flush(cm_ivfpq_sharded *h);
int cm_ivfpq_sharded_rebalance(cm_ivfpq_sharded *h);
int cm_ivfpq_sharded_search(cm_ivfpq_sharded *h, const float *queries, int64_t nq, int dim, const cm_search_params *p,
                            int64_t out_stride, uint32_t *out_ids, float *out_scores, int64_t *out_counts);
int cm_ivfpq_sharded_search_device(cm_ivfpq_sharded *h, const float *queries_dev, int64_t nq, int dim, const cm_search_params *p,
                                   int64_t out_stride, uint32_t *out_ids_dev, float *out_scores_dev, int64_t *out_counts_dev,
                                   void *stream);
int cm_ivfpq_sharded_last_scanned(const cm_ivfpq_sharded *h, int64_t *per_shard);

/* ---- hnsw_index.go / hnsw_index_search.go --------------------------------------------------- */
int cm_hnsw_create(int dim, int metric, int m, int ef_construction, int ef_search, cm_hnsw **out);   /* NewHNSWIndex hnsw_index.go:172 */
int cm_hnsw_destroy(cm_hnsw *h);
int64_t cm_hnsw_size(const cm_hnsw *h);
int cm_hnsw_ef_search(const cm_hnsw *h);
/* Upload a graph built by HNSWIndex.Add / insertNode (hnsw_index.go:228-288, 493-552).  rows are the STORED
 * (already preprocessed) vectors, n x dim; slots = insertion order; edge_off has sum(levels[i]+1)+1 entries,
 * pairs ordered by (slot, layer 0..level); edge_ids are neighbour node IDs in the reference's edge order. */
int cm_hnsw_load_graph(cm_hnsw *h, int64_t n, const uint32_t *ids, const float *rows, const int32_t *levels,
                       const int64_t *edge_off, const uint32_t *edge_ids, uint32_t entry_id, int max_level);
/* n successive HNSWIndex.Add calls ON THE DEVICE (hnsw_index.go:228-288 + insertNode / selectNeighbors /
 * pruneConnections, :493-552, 637-694): PreprocessInPlace (rows written back unless writeback == 0), then the
 * insertion with the caller's level draws (randomLevel, :474-484, stays with the caller: the reference draws
 * from an unseeded global RNG).  IDs must be non-zero and new.  The graph is bit-identical to a sequential
 * replay of the reference with the same levels. */
int cm_hnsw_add(cm_hnsw *h, const uint32_t *ids, float *rows, const int32_t *levels, int64_t n, int writeback);
int cm_hnsw_max_level(const cm_hnsw *h);
/* the graph back to the host (HNSWIndex.WriteTo): levels[n]; edge_off[sum(levels+1)+1] and edge_ids
 * [cm_hnsw_edge_count()] in the layout cm_hnsw_load_graph takes */
int64_t cm_hnsw_edge_count(const cm_hnsw *h);
int cm_hnsw_export_graph(const cm_hnsw *h, int32_t *levels, int64_t *edge_off, uint32_t *edge_ids, uint32_t *entry_id,
                         int *max_level);
/* node IDs and STORED vectors by slot (insertion order); either output may be NULL */
int cm_hnsw_get_nodes(const cm_hnsw *h, int64_t first, int64_t n, uint32_t *ids_out, float *rows_out);
int cm_hnsw_remove(cm_hnsw *h, uint32_t id);                         /* soft delete, hnsw_index.go:300-330 */
/* HNSWIndex.Flush (hnsw_index.go:348-430): live nodes drop their edges to deleted nodes, a deleted entry point is
 * replaced (a live node at maxLevel, else one of the highest level left -- maxLevel follows -- else the index is
 * empty), deleted nodes are freed, the deleted set is cleared.  The reference picks the new entry point by walking a
 * Go map, i.e. at random among the eligible nodes; this picks the earliest inserted eligible node. */
int cm_hnsw_flush(cm_hnsw *h);
/* nq independent searchSingleQuery calls (hnsw_index_search.go:248-354 + searchLayer hnsw_index.go:565-629);
 * p->ef_search as WithEfSearch.  out_stride >= min(k, ef, n).  out_work (optional, nq x 2): distance
 * evaluations and node expansions per query (the algorithmic-bytes figure of the roofline). */
int cm_hnsw_search(cm_hnsw *h, const float *queries, int64_t nq, int dim, const cm_search_params *p, int64_t out_stride,
                   uint32_t *out_ids, float *out_scores, int64_t *out_pos, int64_t *out_counts, int64_t *out_work);
int cm_hnsw_search_device(cm_hnsw *h, const float *queries_dev, int64_t nq, int dim, const cm_search_params *p,
                          int64_t out_stride, uint32_t *out_ids_dev, float *out_scores_dev, int64_t *out_pos_dev,
                          int64_t *out_counts_dev, int64_t *work_dev, void *stream);

/* ---- wire formats: WriteTo / ReadFrom of the five index types -------------------------------- */
/* Byte for byte the reference's little-endian streams: magic ("FLAT" flat_index.go:366-470, "IVFX"
 * ivf_index.go:468-610, "PQIX" pq_index.go:509-650, "IVPQ" ivfpq_index.go:544-700, "HNSW" hnsw_index.go:734-896),
 * version 1, parameters, payload, roaring blob of the deleted set.
 *   cm_*_save   = WriteTo: FLUSHES the index first (as the reference does), then serialises into `buf`.  buf == NULL
 *                 is a size query (*bytes = exact size; the flush has happened).  cap too small ->
 *                 CM_ERR_BUFFER_TOO_SMALL with *bytes = needed.
 *   cm_*_load   = ReadFrom on a pre-constructed index with matching parameters (dimension, distance kind, nlist, M,
 *                 Nbits, m, ef...: mismatches fail with the reference's messages).  The stream is decoded and validated
 *                 completely before the index state is REPLACED.  *consumed = bytes read.  A non-empty deleted set in
 *                 the stream (the reference never writes one) is honoured: those IDs come back soft-deleted.
 *   cm_*_save_file / cm_*_load_file: the same on a file; a path ending in ".gz" is written gzip-compressed (the LSM
 *                 layer's vector_%06d.bin.gz segments, storage_provider.go:163-166, storage.go:693-760); loading
 *                 detects gzip by itself.
 * HNSW: the reference writes nodes in Go map order (random); cm_hnsw_save writes them in insertion order. */
int cm_flat_save(cm_flat *h, uint8_t *buf, int64_t cap, int64_t *bytes);
int cm_flat_load(cm_flat *h, const uint8_t *buf, int64_t len, int64_t *consumed);
int cm_flat_save_file(cm_flat *h, const char *path);
int cm_flat_load_file(cm_flat *h, const char *path);
int cm_ivf_save(cm_ivf *h, uint8_t *buf, int64_t cap, int64_t *bytes);
int cm_ivf_load(cm_ivf *h, const uint8_t *buf, int64_t len, int64_t *consumed);
int cm_ivf_save_file(cm_ivf *h, const char *path);
int cm_ivf_load_file(cm_ivf *h, const char *path);
int cm_pq_save(cm_pq *h, uint8_t *buf, int64_t cap, int64_t *bytes);
int cm_pq_load(cm_pq *h, const uint8_t *buf, int64_t len, int64_t *consumed);
int cm_pq_save_file(cm_pq *h, const char *path);
int cm_pq_load_file(cm_pq *h, const char *path);
int cm_ivfpq_save(cm_ivfpq *h, uint8_t *buf, int64_t cap, int64_t *bytes);
int cm_ivfpq_load(cm_ivfpq *h, const uint8_t *buf, int64_t len, int64_t *consumed);
int cm_ivfpq_save_file(cm_ivfpq *h, const char *path);
int cm_ivfpq_load_file(cm_ivfpq *h, const char *path);
int cm_hnsw_save(cm_hnsw *h, uint8_t *buf, int64_t cap, int64_t *bytes);
int cm_hnsw_load(cm_hnsw *h, const uint8_t *buf, int64_t len, int64_t *consumed);
int cm_hnsw_save_file(cm_hnsw *h, const char *path);
int cm_hnsw_load_file(cm_hnsw *h, const char *path);
/* the roaring portable-format decoder behind cm_*_load, exposed for its tests: IDs of `blob` into out_ids[cap] */
int cm_debug_decode_roaring(const uint8_t *blob, int64_t len, uint32_t *out_ids, int64_t cap, int64_t *count);

#ifdef __cplusplus
}
#endif
#endif
