"""The frozen kNN rules of oracle/knn_oracle.c against independent implementations
(sklearn = the library the reference calls at data_creator_2d.py:66, scipy cKDTree, brute force)."""
import numpy as np
import pytest
import torch
from scipy.spatial import cKDTree
from sklearn.neighbors import NearestNeighbors

from oracle import knn


def _cloud(n, seed, jitter=True):
    rng = np.random.default_rng(seed)
    side = int(np.ceil(np.sqrt(n)))
    gx, gy = np.meshgrid(np.linspace(0, 1, side), np.linspace(0, 1, side), indexing="ij")
    p = np.stack([gx.ravel(), gy.ravel()], 1)[:n]
    if jitter:
        p = p + rng.uniform(-0.3, 0.3, p.shape) / (side - 1)
    return p.astype(np.float32)


@pytest.mark.parametrize("n,k", [(400, 30), (2521, 30), (5000, 35)])
def test_itp_rule_equals_sklearn_on_generic_clouds(n, k):
    pts, qry = _cloud(n, 1), _cloud(n, 2)
    idx, _ = knn.knn_indices(pts, qry, k, rule="f64")
    ref = NearestNeighbors(n_neighbors=k).fit(pts).kneighbors(qry)[1]
    assert np.array_equal(idx, ref)


def test_graph_rule_matches_kdtree_sets_and_order():
    pts = _cloud(3000, 3)
    ei = knn.knn_graph(torch.from_numpy(pts), 35)
    assert ei.shape == (2, 3000 * 35)
    assert torch.equal(ei[1], torch.arange(3000).repeat_interleave(35))
    d, ref = cKDTree(pts.astype(np.float64)).query(pts.astype(np.float64), k=36)
    src = ei[0].reshape(3000, 35).numpy()
    same = sum(set(src[i]) == set(ref[i, 1:]) for i in range(3000))
    assert same >= 2995           # fp32 vs fp64 distances may reorder a near-tie at the k-th place
    dd = np.sum((pts[src] - pts[:, None, :]) ** 2, -1)
    assert np.all(np.diff(dd, axis=1) >= -1e-9)


def test_graph_rule_batched_never_crosses_samples_and_ragged():
    a, b = _cloud(50, 4), _cloud(120, 5)
    x = torch.from_numpy(np.concatenate([a, b]))
    batch = torch.cat([torch.zeros(50, dtype=torch.long), torch.ones(120, dtype=torch.long)])
    ei = knn.knn_graph(x, 35, batch)
    assert torch.all(batch[ei[0]] == batch[ei[1]])
    assert torch.all(ei[0] != ei[1])
    assert torch.equal(torch.bincount(ei[1], minlength=170), torch.full((170,), 35))
    # sample smaller than k+1 -> every other node of the sample, degree n-1
    small = knn.knn_graph(torch.from_numpy(_cloud(10, 6)), 35)
    assert small.shape[1] == 10 * 9


def test_lattice_ties_resolved_by_lower_index():
    g = np.linspace(0, 1, 12, dtype=np.float32)
    pts = np.stack(np.meshgrid(g, g, indexing="ij"), -1).reshape(-1, 2)
    idx, d2 = knn.knn_indices(pts, pts, 35, exclude_self=True, rule="f32")
    assert np.all(np.diff(d2, axis=1) >= 0)
    ties = np.diff(d2, axis=1) == 0
    assert ties.any()
    assert np.all(np.diff(idx, axis=1)[ties] > 0)
    # tie-aware validity: every kept distance <= every dropped distance
    full = ((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1)
    np.fill_diagonal(full, np.inf)
    kth = np.sort(full, axis=1)[:, 34]
    assert np.allclose(d2[:, -1], kth, rtol=1e-5)


def test_duplicates_and_empty():
    pts = np.zeros((5, 2), np.float32)
    idx, d2 = knn.knn_indices(pts, pts, 3, exclude_self=True, rule="f32")
    assert np.array_equal(idx[0], [1, 2, 3]) and np.array_equal(idx[4], [0, 1, 2]) and np.all(d2 == 0)
    e = knn.knn_graph(torch.zeros(0, 2), 35)
    assert e.shape == (2, 0)


def test_radius_rule():
    pts = _cloud(300, 7)
    ei = knn.radius_graph(torch.from_numpy(pts), 0.2, max_num_neighbors=32)
    d = np.linalg.norm(pts[ei[0]] - pts[ei[1]], axis=1)
    assert np.all(d < 0.2 + 1e-6) and torch.all(ei[0] != ei[1])
    assert torch.bincount(ei[1], minlength=300).max() <= 32
