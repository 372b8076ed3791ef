"""ORACLE (test infrastructure).  ctypes front-end of knn_oracle.c + edge-list helpers.

Frozen rules, see knn_oracle.c header.  Reference call sites:
/root/reference/data_creator_2d.py:66,75-76 (sklearn 30-NN), :258 (radius_graph), :260 (knn_graph).
"""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libmmpde_oracle.so")
    src = os.path.join(_HERE, "knn_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared",
                               "-o", so, src, "-lm"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.mmpde_oracle_knn.restype = ctypes.c_int
        _LIB.mmpde_oracle_radius.restype = ctypes.c_int
    return _LIB


def _offsets(batch, n):
    """batch: sorted int64 [n] sample ids (or None = one sample) -> offsets [S+1]."""
    if batch is None:
        return np.array([0, n], dtype=np.int64)
    b = np.asarray(batch, dtype=np.int64)
    assert np.all(np.diff(b) >= 0), "batch vector must be sorted"
    S = int(b.max()) + 1 if b.size else 0
    counts = np.bincount(b, minlength=S)
    return np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)


def knn_indices(pts, qry, k, pts_batch=None, qry_batch=None, exclude_self=False, rule="f32"):
    """Ordered k nearest points of every query.  Returns (idx int64 [Q,k] global rows of pts, d2 [Q,k])."""
    pts = np.ascontiguousarray(np.asarray(pts, dtype=np.float32).reshape(-1, 2))
    qry = np.ascontiguousarray(np.asarray(qry, dtype=np.float32).reshape(-1, 2))
    po, qo = _offsets(pts_batch, len(pts)), _offsets(qry_batch, len(qry))
    assert len(po) == len(qo)
    idx = np.empty((len(qry), k), dtype=np.int64)
    d2 = np.empty((len(qry), k), dtype=np.float64)
    rc = _lib().mmpde_oracle_knn(
        pts.ctypes.data_as(ctypes.c_void_p), qry.ctypes.data_as(ctypes.c_void_p),
        po.ctypes.data_as(ctypes.c_void_p), qo.ctypes.data_as(ctypes.c_void_p),
        ctypes.c_int(len(po) - 1), ctypes.c_int(k), ctypes.c_int(int(exclude_self)),
        ctypes.c_int(0 if rule == "f32" else 1),
        idx.ctypes.data_as(ctypes.c_void_p), d2.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return idx, d2


def knn_graph(x, k, batch=None, loop=False):
    """torch_cluster.knn_graph restatement -> edge_index int64 [2,E]; row0 = neighbour (source j),
    row1 = centre (target i); grouped by ascending i, ascending distance inside a group."""
    assert not loop
    xn = x.detach().cpu().numpy()
    bn = None if batch is None else batch.detach().cpu().numpy()
    idx, _ = knn_indices(xn, xn, k, bn, bn, exclude_self=True, rule="f32")
    tgt = np.repeat(np.arange(len(xn), dtype=np.int64), k)
    src = idx.reshape(-1)
    keep = src >= 0
    return torch.from_numpy(np.stack([src[keep], tgt[keep]]))


def radius_graph(x, r, batch=None, loop=False, max_num_neighbors=32):
    assert not loop
    xn = np.ascontiguousarray(x.detach().cpu().numpy().astype(np.float32).reshape(-1, 2))
    bn = None if batch is None else batch.detach().cpu().numpy()
    off = _offsets(bn, len(xn))
    idx = np.empty((len(xn), max_num_neighbors), dtype=np.int64)
    rc = _lib().mmpde_oracle_radius(
        xn.ctypes.data_as(ctypes.c_void_p), off.ctypes.data_as(ctypes.c_void_p),
        ctypes.c_int(len(off) - 1), ctypes.c_float(float(r)), ctypes.c_int(max_num_neighbors),
        idx.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    tgt = np.repeat(np.arange(len(xn), dtype=np.int64), max_num_neighbors)
    src = idx.reshape(-1)
    keep = src >= 0
    return torch.from_numpy(np.stack([src[keep], tgt[keep]]))
