#!/usr/bin/env python
"""Writes reference_kats.json: the known answers the reference's own tests state for the hot path (file:line
relative to the reference root).  These are literals copied from the tests' expectations."""
import json
import os

KATS = {
    "distance": [
        {"src": "distance_test.go:87-145", "metric": "l2", "a": [0, 0], "b": [3, 4], "want": 5.0, "tol": 1e-6},
        {"src": "distance_test.go:87-145", "metric": "l2", "a": [-1, -2], "b": [1, 2], "want": 4.472136, "tol": 1e-6},
        {"src": "distance_test.go:87-145", "metric": "l2", "a": [1, 2, 3], "b": [1, 2, 3], "want": 0.0, "tol": 1e-6},
        {"src": "distance_test.go:214-266", "metric": "l2_squared", "a": [0, 0], "b": [3, 4], "want": 25.0, "tol": 1e-6},
        {"src": "distance_test.go:214-266", "metric": "l2_squared", "a": [-1, -2], "b": [1, 2], "want": 20.0, "tol": 1e-6},
        {"src": "distance_test.go:214-266", "metric": "l2_squared", "a": [1, 0], "b": [0, 0], "want": 1.0, "tol": 1e-6},
        {"src": "distance_test.go:335-387", "metric": "cosine", "a": [1, 0], "b": [1, 0], "want": 0.0, "tol": 1e-6},
        {"src": "distance_test.go:335-387", "metric": "cosine", "a": [1, 0], "b": [0, 1], "want": 1.0, "tol": 1e-6},
        {"src": "distance_test.go:335-387", "metric": "cosine", "a": [1, 0], "b": [-1, 0], "want": 2.0, "tol": 1e-6},
        {"src": "distance_test.go:335-387", "metric": "cosine", "a": [0.707107, 0.707107], "b": [1, 0], "want": 0.292893, "tol": 1e-6},
    ],
    "normalize": [
        {"src": "distance_test.go:417-491", "in": [3, 4], "want": [0.6, 0.8], "tol": 1e-6},
        {"src": "distance_test.go:417-491", "in": [0, 0, 0], "error": "ErrZeroVector"},
    ],
    # Norm / Normalize / NormalizeInPlace helpers (distance.go:312-428).  For a ZERO vector the helpers return the
    # vector unchanged, while cosine.Preprocess -- the call on the search path -- returns ErrZeroVector (above).
    "norm": [
        {"src": "distance_test.go:533-579", "in": [3, 4], "want": 5.0}, {"src": "distance_test.go:533-579", "in": [1, 0, 0], "want": 1.0},
        {"src": "distance_test.go:533-579", "in": [0, 0, 0], "want": 0.0}, {"src": "distance_test.go:533-579", "in": [-3, -4], "want": 5.0},
        {"src": "distance_test.go:533-579", "in": [7], "want": 7.0}, {"src": "distance_test.go:533-579", "in": [1, 1, 1, 1], "want": 2.0},
    ],
    "normalize_helper": [
        {"src": "distance_test.go:656-722", "in": [3, 4], "want": [0.6, 0.8]}, {"src": "distance_test.go:656-722", "in": [1, 0, 0], "want": [1, 0, 0]},
        {"src": "distance_test.go:656-722", "in": [-3, -4], "want": [-0.6, -0.8]},
        {"src": "distance_test.go:656-722", "in": [1, 1, 1, 1], "want": [0.5, 0.5, 0.5, 0.5]},
        {"src": "distance_test.go:724-781", "in": [2, 2, 2, 2], "want": [0.5, 0.5, 0.5, 0.5]},
    ],
    # TestHighDimensionalVectors (distance_test.go:786-816): a[i] = i % 10, b[i] = (i + 1) % 10, dim 768; every metric finite
    "high_dimensional": {"src": "distance_test.go:786-816", "dim": 768},
    # TestCalculateBatchConsistency (distance_test.go:886-925): per-query Calculate == batch, target all zeros
    "batch_consistency": {"src": "distance_test.go:886-925", "queries": [[1, 2, 3], [4, 5, 6], [7, 8, 9]], "target": [0, 0, 0]},
    # TestDistanceKindConstants (distance_test.go:870-884)
    "kind_constants": {"src": "distance_test.go:870-884", "l2": "l2", "l2_squared": "l2_squared", "cosine": "cosine"},
    "sanitize_k": [
        {"src": "limiter_test.go:7-73", "k": 0, "max": 5, "want": 5}, {"src": "limiter_test.go:7-73", "k": -1, "max": 5, "want": 5},
        {"src": "limiter_test.go:7-73", "k": 3, "max": 5, "want": 3}, {"src": "limiter_test.go:7-73", "k": 10, "max": 5, "want": 5},
    ],
    "aggregation": [
        {"src": "aggregation_test.go:7-115", "kind": "sum", "scores_of_node_1": [0.1, 0.15, 0.05], "want_f32_of": 0.3},
        {"src": "aggregation_test.go:7-115", "kind": "max", "scores_of_node_1": [0.1, 0.5, 0.05], "want": 0.5},
        {"src": "aggregation_test.go:7-115", "kind": "mean", "scores_of_node_1": [0.1, 0.3, 0.2], "want_f32_of": 0.2},
    ],
    "kmeans": [
        {"src": "clustering_test.go:262-301", "points": [[0, 0], [2, 2], [10, 10], [12, 12]], "k": 2,
         "want_centroids": [[1, 1], [11, 11]], "tol": 0.01},
    ],
    "flat_search": [
        {"src": "flat_index_search_test.go:10-48", "metric": "l2", "rows": {"1": [1, 0, 0], "2": [0, 1, 0], "3": [0, 0, 1], "4": [1, 1, 0]},
         "query": [1, 0, 0], "k": 2, "want_first_id": 1, "want_first_score": 0.0},
        {"src": "flat_index_search_test.go:490-536", "metric": "l2", "rows": {"1": [5, 0, 0], "2": [1, 0, 0], "3": [10, 0, 0], "4": [3, 0, 0]},
         "query": [0, 0, 0], "k": 4, "want_ids": [2, 4, 1, 3], "want_scores": [1, 3, 5, 10]},
    ],
}

if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kats.json")
    with open(out, "w") as f:
        json.dump(KATS, f, indent=1)
    print(out)
