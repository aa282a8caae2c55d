"""TEST INFRASTRUCTURE ONLY (imported by tests/, never by the product): a sequential CPU restatement of the Leiden
algorithm (Traag, Waltman & van Eck, "From Louvain to Leiden: guaranteeing well-connected communities", Sci. Rep. 2019,
Algorithm 1 / supplementary pseudo-code) for the quality function leidenalg's ``RBConfigurationVertexPartition``
maximises:  Q = sum_c [ w_in(c) - gamma * K_c^2 / (2 * 2m) ]  (undirected, weights from the connectivities).

The reference reaches this algorithm through ``scanpy.tl.leiden`` -> ``leidenalg.find_partition``
(/root/reference/src/infercnvpy/tl/__init__.py:24-30); leidenalg / igraph are not installed in this image and the
reference's tests assert nothing about the clustering, so PARITY IS UNPINNED: this file follows the published
pseudo-code (queue-based fast local moving, refinement restricted to well-connected nodes and sub-communities,
aggregation on the refined partition with the non-refined partition as the initial assignment) with a seeded random
visiting order and the greedy (theta -> 0) choice in the refinement, and serves as the "CPU Leiden-equivalent" the GPU
result's quality and ARI are reported against.  Pure Python over CSR arrays: meant for graphs of a few thousand nodes.
"""

from __future__ import annotations

from collections import deque

import numpy as np
import scipy.sparse as sp


def quality(A: sp.csr_matrix, labels: np.ndarray, gamma: float = 1.0) -> float:
    """RB-configuration quality divided by 2m (== modularity for gamma = 1)."""
    A = sp.csr_matrix(A)
    k = np.asarray(A.sum(axis=1)).ravel()
    two_m = k.sum()
    rows = np.repeat(np.arange(A.shape[0]), np.diff(A.indptr))
    inside = A.data[labels[rows] == labels[A.indices]].sum()
    ctot = np.bincount(labels, weights=k)
    return float((inside - gamma * (ctot**2).sum() / two_m) / two_m)


def _move_nodes_fast(indptr, indices, data, k, comm, gamma, two_m, rng):
    n = len(k)
    ctot = np.bincount(comm, weights=k, minlength=n).astype(np.float64)
    order = rng.permutation(n)
    queue = deque(order.tolist())
    in_queue = np.ones(n, dtype=bool)
    moved = False
    while queue:
        v = queue.popleft()
        in_queue[v] = False
        cv = comm[v]
        wts: dict[int, float] = {}
        for e in range(indptr[v], indptr[v + 1]):
            u = indices[e]
            if u != v:
                wts[comm[u]] = wts.get(comm[u], 0.0) + data[e]
        kv = k[v]
        best_c, best = cv, wts.get(cv, 0.0) - gamma * kv * (ctot[cv] - kv) / two_m
        for c, wvc in wts.items():
            if c == cv:
                continue
            g = wvc - gamma * kv * ctot[c] / two_m
            if g > best + 1e-12:
                best, best_c = g, c
        if best_c != cv:
            ctot[cv] -= kv
            ctot[best_c] += kv
            comm[v] = best_c
            moved = True
            for e in range(indptr[v], indptr[v + 1]):
                u = indices[e]
                if comm[u] != best_c and not in_queue[u]:
                    queue.append(u)
                    in_queue[u] = True
    return moved


def _refine(indptr, indices, data, k, comm, gamma, two_m, rng):
    """Refined partition: inside every community, well-connected singletons merge greedily (theta -> 0)."""
    n = len(k)
    sub = np.arange(n)
    stot = k.astype(np.float64).copy()          # total degree per sub-community
    ssize = np.ones(n, dtype=np.int64)
    ctot = np.bincount(comm, weights=k, minlength=n)
    # weight of every sub-community to the rest of its community (for the well-connectedness test of the target)
    cut = np.zeros(n)
    for v in range(n):
        for e in range(indptr[v], indptr[v + 1]):
            u = indices[e]
            if u != v and comm[u] == comm[v]:
                cut[v] += data[e]
    for v in rng.permutation(n).tolist():
        if ssize[sub[v]] != 1:
            continue
        kv, Kc = k[v], ctot[comm[v]]
        if cut[sub[v]] < gamma * kv * (Kc - kv) / two_m:
            continue
        wts: dict[int, float] = {}
        for e in range(indptr[v], indptr[v + 1]):
            u = indices[e]
            if u != v and comm[u] == comm[v]:
                wts[sub[u]] = wts.get(sub[u], 0.0) + data[e]
        best_s, best = sub[v], 0.0
        for s_, wvs in wts.items():
            if s_ == sub[v]:
                continue
            if cut[s_] < gamma * stot[s_] * (Kc - stot[s_]) / two_m:
                continue  # target not well connected
            g = wvs - gamma * kv * stot[s_] / two_m
            if g > best + 1e-12:
                best, best_s = g, s_
        if best_s != sub[v]:
            old = sub[v]
            # cut of the merged sub-community: both cuts minus twice the weight between them
            cut[best_s] = cut[best_s] + cut[old] - 2.0 * wts[best_s]
            stot[best_s] += kv
            ssize[best_s] += 1
            stot[old], ssize[old], cut[old] = 0.0, 0, 0.0
            sub[v] = best_s
    return sub


def leiden(A, gamma: float = 1.0, seed: int = 0, max_levels: int = 50) -> np.ndarray:
    """Labels (0..k-1 by decreasing size) of the Leiden partition of the symmetric weighted graph ``A``."""
    A = sp.csr_matrix(A, dtype=np.float64)
    A.setdiag(0)
    A.eliminate_zeros()
    rng = np.random.default_rng(seed)
    n0 = A.shape[0]
    node_of = np.arange(n0)
    comm = np.arange(n0)
    for _ in range(max_levels):
        n = A.shape[0]
        indptr, indices, data = A.indptr, A.indices, A.data
        k = np.asarray(A.sum(axis=1)).ravel()  # the diagonal of an aggregated node already holds both directions of its inner edges
        two_m = k.sum()
        moved = _move_nodes_fast(indptr, indices, data, k, comm, gamma, two_m, rng)
        sub = _refine(indptr, indices, data, k, comm, gamma, two_m, rng)
        uniq, inv = np.unique(sub, return_inverse=True)
        nc = len(uniq)
        if nc == n and not moved:
            break
        first = np.full(nc, n)
        np.minimum.at(first, inv, np.arange(n))
        _, new_comm = np.unique(comm[first], return_inverse=True)
        node_of = inv[node_of]
        comm = new_comm
        if nc < n:
            S = sp.csr_matrix((np.ones(n), (np.arange(n), inv)), shape=(n, nc))
            A = sp.csr_matrix(S.T @ A @ S)
            A.sum_duplicates()
    lab = comm[node_of]
    sizes = np.bincount(lab)
    order = np.argsort(-sizes, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    return rank[lab]
