// kmeans.cuh -- device k-means helpers (kmeans.cu): the reference's deterministic clustering.go, bit for bit.
#pragma once

#include "flat_index.cuh"

namespace cm {

// KMeansSubspace (clustering.go:96-118): L2^2 k-means on columns [off, off+d) of x ([n][ldx], device);
// centroids out: [k][d] contiguous (device).
int kmeans_subspace(const float *x, int64_t n, int64_t ldx, int off, int d, int k, int max_iter, float *cent_out, cudaStream_t st);
// KMeans (clustering.go:60-94) on full rows ([n_pad][cent.ld], zero padded, n_pad multiple of 8) with the
// index metric; the k centroids become the rows of `cent` (raw).  final_assign (optional, [n]) receives
// FindNearestCentroidIndex of every row against the FINAL centroids.
int kmeans_full(FlatIndex &cent, const float *rows, int64_t n, int k, int max_iter, long long *final_assign, cudaStream_t st);
int launch_residuals(const float *rows, int64_t n, int d, int ld, const float *cent, const long long *assign, float *out, cudaStream_t st);
int upload_training_rows(const float *rows_host, int64_t n, int dim, int ld, float **out, cudaStream_t st);

}  // namespace cm
