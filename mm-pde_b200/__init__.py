"""mm-pde_b200 -- B200-native (sm_100a) implementation of MM-PDE's data-parallel hot path:
the MP-PDE message-passing processor and the moved-mesh <-> reference-mesh interpolation.

Host side = Python mirrors of the reference's module API (gnn_2d, interpolate, data_creator_2d,
train_helper_2d, mmpde, PDEs, mesh.dmm_model); device side = libmmpde_b200.so (hand-written CUDA,
C ABI in include/mmpde_b200.h) bound with ctypes in _cabi.py.  There is no CPU fallback: every
hot-path op raises if the CUDA library or a CUDA device is missing.
"""
__version__ = "0.1.0"
