"""Property tests (hypothesis) of the CUDA path through the C ABI -- the cases a fixed fixture does not reach
(SURVEY.md section 4): k >= n, coincident points, ragged samples, targets without incoming edges,
and permutation equivariance of the message-passing layer.  Example counts are small: every example launches kernels."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

pytestmark = pytest.mark.gpu

from oracle import knn as oknn  # noqa: E402

COMMON = dict(deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)


def _dev():
    return torch.device("cuda:0")


def _off(sizes, dev):
    return torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int32, device=dev)


@st.composite
def clouds(draw):
    """Ragged batch of small 2-D clouds on a coarse coordinate grid: many exact ties and coincident points."""
    n_samples = draw(st.integers(1, 4))
    sizes = [draw(st.integers(1, 60)) for _ in range(n_samples)]
    levels = draw(st.sampled_from([3, 8, 1000]))            # 3 -> almost everything coincides / ties
    seed = draw(st.integers(0, 2 ** 16))
    rng = np.random.default_rng(seed)
    pts = (rng.integers(0, levels, size=(sum(sizes), 2)) / float(levels)).astype(np.float32)
    return sizes, pts


@settings(max_examples=25, **COMMON)
@given(clouds(), st.sampled_from([1, 5, 35, 64]), st.booleans())
def test_knn_any_k_ties_duplicates_ragged(cloud, k, graph_rule):
    """Bit-exact ordered neighbour lists for every k (also k >= sample size: -1 pads), both rules, with duplicates."""
    from mmpde_b200 import ops
    sizes, pts = cloud
    dev = _dev()
    batch = np.repeat(np.arange(len(sizes)), sizes)
    rule, excl = ("f32", True) if graph_rule else ("f64", False)
    ref, _ = oknn.knn_indices(pts, pts, k, batch, batch, exclude_self=excl, rule=rule)
    t = torch.from_numpy(pts).to(dev)
    off = _off(sizes, dev)
    got = ops.knn_indices(t, off, t, off, k, rule=0 if graph_rule else 1, exclude_self=excl).cpu().numpy()
    assert np.array_equal(got, ref)
    got = ops._knn_grid(t, off, t, off, k, 0 if graph_rule else 1, excl, (0.0, 0.0, 1.0, 1.0), max(sizes)).cpu().numpy()
    assert np.array_equal(got, ref)
    # pads: exactly max(0, k - available) per row, always at the end
    avail = np.repeat(np.array(sizes) - (1 if excl else 0), sizes)
    assert np.array_equal((got < 0).sum(1), np.maximum(0, k - avail))


def _layer_inputs(n, n_edges, seed, dev, isolate):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 128, generator=g)
    node4 = torch.rand(n, 4, generator=g)
    src = torch.randint(0, n, (n_edges,), generator=g)
    dst = torch.randint(0, n, (n_edges,), generator=g)
    if isolate:                                             # targets without any incoming edge (mean over nothing = 0)
        dst = dst[dst % 3 != 0]
        src = src[:dst.numel()]
    order = torch.argsort(dst, stable=True)
    ei = torch.stack((src[order], dst[order]))
    return x.to(dev), node4.to(dev), ei.to(dev)


@settings(max_examples=8, **COMMON)
@given(st.integers(16, 300), st.integers(0, 4000), st.integers(0, 1000), st.booleans())
def test_layer_matches_oracle_on_random_multigraphs(n, n_edges, seed, isolate):
    """GNN_Layer_FS_2D forward on arbitrary target-sorted multigraphs (repeated edges, self loops, empty targets, ragged
    in-degrees, E = 0) equals the oracle layer (reference formulation with the materialised [E,260] tensor)."""
    from mmpde_b200.gnn_2d import GNN_Layer_FS_2D
    from oracle import processor as oproc
    from tests.golden.common import fill_params
    dev = _dev()
    x, node4, ei = _layer_inputs(n, n_edges, seed, dev, isolate)
    layer = fill_params(GNN_Layer_FS_2D(128, 128, 128, 1, 1), 3).to(dev).train()
    olayer = fill_params(oproc.GNN_Layer_FS_2D(128, 128, 128, 1, 1), 3).train()
    out = layer(x, node4[:, 0:1], node4[:, 1:2], node4[:, 2:3], node4[:, 3:4], ei)
    xc, n4c = x.cpu(), node4.cpu()
    ref = olayer(xc, n4c[:, 0:1], n4c[:, 1:2], n4c[:, 2:3], n4c[:, 3:4], ei.cpu(), None)
    err = float((out.cpu() - ref).norm() / ref.norm().clamp_min(1e-30))
    assert err < 1e-3, err


@settings(max_examples=6, **COMMON)
@given(st.integers(40, 400), st.integers(0, 1000))
def test_layer_is_permutation_equivariant(n, seed):
    """Relabelling the nodes (and the edge list with them) permutes the layer's output rows and nothing else."""
    from mmpde_b200 import ops
    from mmpde_b200.gnn_2d import GNN_Layer_FS_2D
    from tests.golden.common import fill_params
    dev = _dev()
    x, node4, ei = _layer_inputs(n, 12 * n, seed, dev, isolate=False)
    layer = fill_params(GNN_Layer_FS_2D(128, 128, 128, 1, 1), 5).to(dev).train()
    out = layer(x, node4[:, 0:1], node4[:, 1:2], node4[:, 2:3], node4[:, 3:4], ei)
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(seed + 1)).to(dev)     # new id of node i = inv[i]
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(n, device=dev)
    ei_p = inv[ei]
    xp, n4p = x[perm], node4[perm]
    out_p = layer(xp, n4p[:, 0:1], n4p[:, 1:2], n4p[:, 2:3], n4p[:, 3:4], ei_p)
    err = float((out_p - out[perm]).norm() / out.norm())
    assert err < 2e-5, err
    assert ops.EdgeList.from_edge_index(ei_p, n).n_edges == ei.shape[1]
