"""CPU oracle for prod_knn_sample (reference Model.py:75-106).

TEST INFRASTRUCTURE ONLY.  The neighbour search itself is oracle/knn_oracle.c
(float64 brute force restating scikit-learn's arithmetic and tie rule); this
module restates the sampler around it and is pinned by
tests/test_oracle_knn.py against tests/golden/knn.npz, i.e. against the
reference function run unmodified on top of scikit-learn 1.9.0.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "libknn_oracle.so")
        if not os.path.exists(path):
            build()
        lib = ctypes.CDLL(path)
        lib.knn_oracle_f64.restype = ctypes.c_int
        lib.knn_oracle_f64.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_int, ctypes.c_void_p, ctypes.c_long,
                                       ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        _LIB = lib
    return _LIB


def sklearn_route(width, k, n_fit):
    """sklearn/neighbors/_base.py:615-648 with algorithm='auto', euclidean."""
    return "brute" if (width > 15 or k >= n_fit // 2) else "kd_tree"


def knn(keys, queries, k, excluded=None, route="brute", threads=None):
    """k nearest keys per query (original key indices, nearest first, exact
    ties -> lowest index).  float32 inputs, float64 arithmetic."""
    keys = np.ascontiguousarray(keys, dtype=np.float32)
    queries = np.ascontiguousarray(queries, dtype=np.float32)
    n, w = keys.shape
    m = queries.shape[0]
    out = np.empty((m, k), dtype=np.int64)
    dist = np.empty((m, k), dtype=np.float64)
    exc = None
    if excluded is not None:
        exc = np.ascontiguousarray(excluded, dtype=np.uint8)
    lib = _lib()

    def run(lo, hi):        # ctypes drops the GIL: one thread per query slice
        return lib.knn_oracle_f64(keys.ctypes.data, n, w, queries[lo:hi].ctypes.data, hi - lo,
                                  exc.ctypes.data if exc is not None else None, k,
                                  1 if route == "brute" else 0, out[lo:hi].ctypes.data, dist[lo:hi].ctypes.data)
    nt = max(1, min(threads or (os.cpu_count() or 1), m))
    cuts = np.linspace(0, m, nt + 1).astype(int)
    if nt == 1:
        st = run(0, m)
    else:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(nt) as ex:
            st = max(ex.map(lambda ab: run(*ab), zip(cuts[:-1], cuts[1:])))
    if st == 1:
        raise MemoryError
    if st == 2:
        # sklearn/neighbors/_base.py:840-851
        raise ValueError("Expected n_neighbors <= n_samples_fit")
    return out, dist


def draw_ids(N, m):
    """np.random.choice(range(N), size=m, replace=False) (Model.py:81): same
    values and same global-RNG state as permutation(N)[:m]."""
    if m > N:
        raise ValueError("Cannot take a larger sample than population when 'replace=False'")
    return np.random.permutation(N)[:m]


def prod_knn_sample(X, Y, Z, batch_size, k_neighbor, radius=None, ids=None):
    """Returns (batch_x, batch_y, batch_z, ids, nbr_compacted).  ``radius`` is
    accepted and ignored, as in the reference."""
    X = np.asarray(X, dtype=np.float32)
    Y = np.asarray(Y, dtype=np.float32)
    Z = np.asarray(Z, dtype=np.float32)
    N = X.shape[0]
    m = batch_size // k_neighbor
    if ids is None:
        ids = draw_ids(N, m)
    excluded = np.zeros(N, dtype=np.uint8)
    excluded[ids] = 1
    route = sklearn_route(Z.shape[1], k_neighbor, N - m)
    nbr_orig, _ = knn(Z, Z[ids], k_neighbor, excluded, route)
    sorted_ids = np.sort(ids)
    nbr_comp = nbr_orig - np.searchsorted(sorted_ids, nbr_orig)      # index into the compacted pool
    bx = X[nbr_orig.reshape(-1)]
    by = Y[np.repeat(ids, k_neighbor)]
    bz = Z[np.repeat(ids, k_neighbor)]
    wmax = max(bx.shape[1], by.shape[1], bz.shape[1])
    bx, by, bz = (np.tile(b, (1, wmax // b.shape[1])) if b.shape[1] != wmax else b for b in (bx, by, bz))
    return bx, by, bz, ids, nbr_comp
