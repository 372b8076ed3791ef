"""ORACLE -- CPU restatement of the MM-PDE hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker or the timed CPU
baseline -- never from ``mm-pde_b200/`` (the product fails loudly without its CUDA library).

What it restates (all file:line relative to /root/reference):
  * knn.py        -- torch_cluster knn_graph / radius_graph and sklearn NearestNeighbors rules
                     (data_creator_2d.py:66,75-76,258,260; mesh/dmm_model.py:228)  [C: knn_oracle.c]
  * processor.py  -- GNN_Layer_FS_2D / MP_PDE_Solver_2D (gnn_2d.py:19-141) incl. PyG propagate,
                     torch_scatter mean and PyG BatchNorm semantics
  * itp.py        -- ItpNet (interpolate.py:5-98)
  * creator.py    -- GraphCreator_FS_2D (data_creator_2d.py:18-305)
  * loops.py      -- training_itp / training_loop_branch / test_timestep_losses
                     (train_helper_2d.py:9-200), criterion (mmpde.py:33-36)
  * dmm.py        -- DMM mesh mover (mesh/dmm_model.py:9-234), CPU-capable
  * pdes.py       -- burgers / cy constants (PDEs.py:20-67)

Parity status: the reference ships no tests, fixtures or golden vectors and its third-party
operators (PyG 2.0.3, torch_cluster 1.5.9, torch_scatter 2.0.9) cannot be installed here, so
parity is UNPINNED by the reference itself.  It is pinned as far as possible by
tests/golden/make_golden.py, which imports the reference's own unmodified gnn_2d.py,
interpolate.py, data_creator_2d.py, train_helper_2d.py and PDEs.py on top of thin shims of
the missing third-party calls and freezes their outputs as fixtures under tests/golden/.
"""
