"""CPU restatements for the pca / neighbors half of the path.  TEST INFRASTRUCTURE ONLY.

Parity status: UNPINNED.  The reference delegates these steps to un-vendored third-party packages
(scanpy >= 1.10 -> scikit-learn TruncatedSVD(arpack); scanpy neighbors -> umap-learn 0.5 ``fuzzy_simplicial_set``;
leidenalg) and its own tests assert nothing about their results.  scikit-learn is installed here and
is used directly by the tests; umap-learn is not, so its published routine (umap_.py: ``smooth_knn_dist`` /
``compute_membership_strengths`` / fuzzy union with set_op_mix_ratio = 1) is restated below in float32 numpy.
"""

from __future__ import annotations

import numpy as np
import scipy.sparse as sp

SMOOTH_K_TOLERANCE = 1e-5
MIN_K_DIST_SCALE = 1e-3


def smooth_knn_dist(distances: np.ndarray, k: float, n_iter: int = 64):
    """umap-learn ``smooth_knn_dist`` with local_connectivity = 1, bandwidth = 1 (float32 like numba's inference)."""
    distances = distances.astype(np.float32)
    target = np.float32(np.log2(k))
    n = distances.shape[0]
    rho = np.zeros(n, dtype=np.float32)
    result = np.zeros(n, dtype=np.float32)
    mean_distances = np.float32(np.mean(distances))
    for i in range(n):
        lo, hi, mid = np.float32(0.0), np.float32(np.inf), np.float32(1.0)
        ith = distances[i]
        nz = ith[ith > 0.0]
        if nz.shape[0] >= 1:
            rho[i] = nz[0]
        for _ in range(n_iter):
            psum = np.float32(0.0)
            for j in range(1, distances.shape[1]):
                d = ith[j] - rho[i]
                psum += np.exp(-(d / mid)) if d > 0 else np.float32(1.0)
            if np.fabs(psum - target) < SMOOTH_K_TOLERANCE:
                break
            if psum > target:
                hi = mid
                mid = (lo + hi) / np.float32(2.0)
            else:
                lo = mid
                mid = mid * np.float32(2) if hi == np.inf else (lo + hi) / np.float32(2.0)
        result[i] = mid
        if rho[i] > 0.0:
            mean_ith = np.float32(np.mean(ith))
            if result[i] < MIN_K_DIST_SCALE * mean_ith:
                result[i] = MIN_K_DIST_SCALE * mean_ith
        elif result[i] < MIN_K_DIST_SCALE * mean_distances:
            result[i] = MIN_K_DIST_SCALE * mean_distances
    return result, rho


def fuzzy_simplicial_set(knn_indices: np.ndarray, knn_dists: np.ndarray) -> sp.csr_matrix:
    """Connectivities ``A + A^T - A * A^T`` from kNN lists that contain the point itself in column 0."""
    n, k = knn_indices.shape
    sigmas, rhos = smooth_knn_dist(knn_dists, float(k))
    vals = np.zeros((n, k), dtype=np.float32)
    for i in range(n):
        for j in range(k):
            if knn_indices[i, j] == i:
                vals[i, j] = 0.0
            elif knn_dists[i, j] - rhos[i] <= 0.0 or sigmas[i] == 0.0:
                vals[i, j] = 1.0
            else:
                vals[i, j] = np.exp(-((np.float32(knn_dists[i, j]) - rhos[i]) / sigmas[i]))
    rows = np.repeat(np.arange(n), k)
    A = sp.coo_matrix((vals.ravel(), (rows, knn_indices.ravel())), shape=(n, n)).tocsr()
    A.eliminate_zeros()
    T = A.T.tocsr()
    P = A.multiply(T)
    return (A + T - P).tocsr()
