"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol that
include/mmpde_b200.h declares, the ctypes table matches the header, host-side logic works on CPU
tensors, and the product path fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "mmpde_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = re.findall(r"\bint\s+(mmpde_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S)
    out = {}
    for name, args in decls:
        args = args.strip()
        out[name] = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
    return out


def _ensure_built():
    from mmpde_b200 import _cabi
    if not os.path.exists(_cabi.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _cabi


def test_library_exports_every_declared_symbol():
    _cabi = _ensure_built()
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    declared = _header_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in mmpde_b200.h but not exported"
    assert lib.mmpde_abi_version() == 1


def test_ctypes_table_matches_header():
    _cabi = _ensure_built()
    declared = _header_functions()
    assert set(declared) == set(_cabi.SIGNATURES)
    for name, n_args in declared.items():
        assert len(_cabi.SIGNATURES[name]) == n_args, name


def test_argument_errors_are_reported_before_any_device_work():
    """Entry points validate their arguments on the host and return MMPDE_EINVAL (-1) without touching a device, so a
    wrong call fails loudly even on a box without a GPU; the struct layouts of the grouped launches match the header."""
    _cabi = _ensure_built()
    lib = _cabi.lib()
    assert lib.mmpde_node_wgrad_grouped(None, -1, None) == -1
    assert lib.mmpde_node_wgrad_grouped(None, 2, None) == -1                      # tasks missing
    bad = (_cabi.WgradTask * 1)(_cabi.WgradTask(0x1000, 130, None, 0, None, None, 0, None, 0, None, 64))   # lda % 4 != 0
    assert lib.mmpde_node_wgrad_grouped(ctypes.addressof(bad), 1, None) == -1
    unpaired = (_cabi.WgradTask * 1)(_cabi.WgradTask(0x1000, 128, 0x2000, 128, None, None, 0, None, 0, None, 64))   # B without dW
    assert lib.mmpde_node_wgrad_grouped(ctypes.addressof(unpaired), 1, None) == -1
    assert lib.mmpde_knn_grid_multi(None, 1, None) == -1
    k0 = (_cabi.KnnTask * 1)(_cabi.KnnTask(0x1000, 0x1000, 0x1000, 0x1000, 1, 0, 10, 0.0, 0.0, 1.0, 4, 4, 0x1000, 0x1000, 0, 0, 0x1000))
    assert lib.mmpde_knn_grid_multi(ctypes.addressof(k0), 1, None) == -1          # k = 0
    k65 = (_cabi.KnnTask * 1)(_cabi.KnnTask(0x1000, 0x1000, 0x1000, 0x1000, 1, 65, 10, 0.0, 0.0, 1.0, 4, 4, 0x1000, 0x1000, 0, 0, 0x1000))
    assert lib.mmpde_knn_grid_multi(ctypes.addressof(k65), 1, None) == -1         # k > 64
    assert lib.mmpde_bn_exchange(0x1000, 16, 0x1000, 3, 2, 0x1000, None) == -1    # rank >= world
    assert lib.mmpde_bn_exchange(0x1000, 16, 0x1000, 0, 17, 0x1000, None) == -1   # world > 16
    assert lib.mmpde_bn_finalize(0x1000, 0, 10.0, 1e-5, 0.1, 0x1000, None, None, None) == -1     # n_rep < 1
    assert lib.mmpde_node_gemm(0x1000, 128, None, 0, 0x1000, 128, 1, None, 0, 0, None, None, None, 2, None, 0, None, 0,
                               0x1000, 128, 64, None) == -1                      # gate mode without the activation
    # struct sizes as the C compiler lays them out (header: mmpde_wgrad_task, mmpde_knn_task)
    assert ctypes.sizeof(_cabi.WgradTask) == 11 * 8
    assert ctypes.sizeof(_cabi.KnnTask) == 104
    from mmpde_b200 import ops
    assert ops.BN_REPLICAS == 16
    with open(os.path.join(ROOT, "include", "mmpde_b200.h")) as f:
        hdr = f.read()
    assert "#define MMPDE_BN_REPLICAS 16" in hdr and "#define MMPDE_BN_EXCHANGE_BYTES (1024 + 4 * 16 * 256 * 8)" in hdr


def test_argument_errors_of_the_round2_entry_points():
    """mmpde_dmm_displacement, the two halves of the BatchNorm exchange and the CTA cap reject bad arguments on the host."""
    lib = _ensure_built().lib()
    P = 0x1000
    assert lib.mmpde_dmm_displacement(P, P, P, 33, P, P, P, 512, 10, 5, P, None) == -1      # trunk width > 32
    assert lib.mmpde_dmm_displacement(P, P, P, 32, P, P, P, 510, 10, 5, P, None) == -1      # J not a multiple of 4
    assert lib.mmpde_dmm_displacement(P, P, P, 32, P, P, P, 512, 10, 0, P, None) == -1      # per_sample <= 0
    assert lib.mmpde_dmm_displacement(None, P, P, 32, P, P, P, 512, 10, 5, P, None) == -1   # missing points
    assert lib.mmpde_dmm_displacement(P, P, P, 32, P, P, P, 512, 0, 5, P, None) == 0        # nothing to do
    assert lib.mmpde_bn_exchange_wait(None, 0, 2, P, None) == -1
    assert lib.mmpde_bn_exchange_wait(P, 2, 2, P, None) == -1                                # rank >= world
    assert lib.mmpde_bn_bwd_reduce_post(P, 128, None, 0, 0, P, 128, None, 0, 64, P, P, P, P, None, 0, 2, None) == -1   # no peers
    assert lib.mmpde_bn_bwd_reduce_post(P, 128, None, 0, 0, P, 128, None, 0, 64, P, P, P, P, P, 0, 1, None) == -1      # world < 2
    assert lib.mmpde_set_persistent_ctas(-1) == -1
    assert lib.mmpde_set_persistent_ctas(74) == 0 and lib.mmpde_set_persistent_ctas(0) == 0


def test_branch_width_is_only_used_for_overlapped_training_steps():
    """train_helper_2d._branch_ctas: no cap without a mesh mover, on the CPU, or when the overlap is switched off."""
    _ensure_built()
    from mmpde_b200 import ops, train_helper_2d as th
    assert th._branch_ctas("cpu", object()) == (0, 0)
    assert th._branch_ctas("cpu", None) == (0, 0)
    with ops.persistent_ctas(0):                 # n = 0 is a no-op and needs no device
        pass


def test_no_cpu_fallback():
    from mmpde_b200 import ops, _cabi
    from mmpde_b200.gnn_2d import GNN_Layer_FS_2D
    layer = GNN_Layer_FS_2D(128, 128, 128, 1, 1)
    x = torch.randn(10, 128)
    s = torch.randn(10, 1)
    ei = torch.stack((torch.arange(1, 10), torch.zeros(9, dtype=torch.long)))
    with pytest.raises(_cabi.MMPDEError):
        layer(x, s, s, s, s, ei, None)
    with pytest.raises(_cabi.MMPDEError):
        ops.knn_indices(torch.rand(5, 2), torch.tensor([0, 5], dtype=torch.int32), torch.rand(5, 2),
                        torch.tensor([0, 5], dtype=torch.int32), 3, 0, True)


def test_state_dict_keys_match_oracle():
    from mmpde_b200.gnn_2d import MP_PDE_Solver_2D
    from mmpde_b200.interpolate import ItpNet
    from mmpde_b200.PDEs import burgers
    from mmpde_b200.mesh.dmm_model import DMM
    from oracle import processor, itp, pdes, dmm
    a = MP_PDE_Solver_2D(burgers()).state_dict()
    b = processor.MP_PDE_Solver_2D(pdes.burgers()).state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(a[k].shape == b[k].shape for k in a)
    assert repr(MP_PDE_Solver_2D(burgers(), hidden_layer=1)) == "GNN"
    for args in ((12, 12), (77, None)):
        a = ItpNet(*args, [128, 64], [128, 64], [1, 4, 16, 4, 1]).state_dict()
        b = itp.ItpNet(*args, [128, 64], [128, 64], [1, 4, 16, 4, 1]).state_dict()
        assert list(a.keys()) == list(b.keys()) and all(a[k].shape == b[k].shape for k in a)
    kw = dict(s=12, mode="array", branch_layer=7, trunk_layer=[2, 32, 512], out_layer=[1024, 512, 1])
    assert sorted(DMM(**kw).state_dict().keys()) == sorted(dmm.DMM(**kw).state_dict().keys())
    assert burgers().dt == pdes.burgers().dt == 1.0


def test_edge_list_from_edge_index_sorts_and_counts():
    from mmpde_b200.ops import EdgeList
    ei = torch.tensor([[1, 2, 0, 3, 0], [2, 0, 1, 0, 2]])
    e = EdgeList.from_edge_index(ei, 5)
    assert e.dst.tolist() == [0, 0, 1, 2, 2] and e.src.tolist() == [2, 3, 0, 1, 0]
    assert torch.allclose(e.inv_deg, torch.tensor([0.5, 1.0, 0.5, 1.0, 1.0]))
    assert e.src.dtype == torch.int32 and e.n_edges == 5
    nbr = torch.tensor([[1, 2], [0, -1], [-1, -1]], dtype=torch.int32)
    p = EdgeList.from_knn(nbr, has_pad=True)
    assert p.src.tolist() == [1, 2, 0] and p.dst.tolist() == [0, 0, 1]
    assert torch.allclose(p.inv_deg, torch.tensor([0.5, 1.0, 1.0]))
    assert torch.equal(p.edge_index(), torch.tensor([[1, 2, 0], [0, 0, 1]]))


def test_create_data_matches_oracle_slicing():
    from mmpde_b200.data_creator_2d import GraphCreator_FS_2D
    from mmpde_b200.PDEs import burgers
    from oracle import creator, pdes
    u = torch.randn(5, 31, 6, 6)
    steps = [1, 7, 30, 12]                      # shorter than the batch: zip truncation (appendix C.5)
    a = GraphCreator_FS_2D(burgers(), 35, "knn", 1, 31).create_data(u, steps)
    b = creator.GraphCreator_FS_2D(pdes.burgers(), 35, "knn", 1, 31).create_data(u, steps)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and a[0].shape == (4, 1, 6, 6)


def test_step_graph_host_logic_on_cpu():
    """StepGraph on a box without CUDA: the optimizer check is a host-side check, CPU inputs run eagerly (prepare +
    body every call, nothing recorded), release() is safe to call at any time."""
    from mmpde_b200.train_helper_2d import StepGraph
    w = torch.nn.Parameter(torch.zeros(3))
    with pytest.raises(ValueError):
        StepGraph.hyper(torch.optim.AdamW([w], lr=1e-3))
    sig_a = StepGraph.hyper(torch.optim.AdamW([w], lr=1e-3, capturable=True))
    sig_b = StepGraph.hyper(torch.optim.AdamW([w], lr=4e-4, capturable=True))
    assert sig_a != sig_b and sig_a == StepGraph.hyper(torch.optim.AdamW([w], lr=1e-3, capturable=True), None)
    sg = StepGraph(eager_steps=1)
    calls = []
    for i in range(4):
        out = sg.run(("train", (2, 3)), lambda x: (calls.append("body"), x * 2)[1], (torch.ones(2, 3) * i,),
                     prepare=lambda: calls.append("prepare"))
        assert torch.equal(out, torch.ones(2, 3) * 2 * i)
    assert calls == ["prepare", "body"] * 4 and sg.replays == 0 and len(sg._graphs) == 0
    sg.release()
