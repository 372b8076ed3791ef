"""ctypes binding of libmmpde_b200.so (C ABI: include/mmpde_b200.h).

The product path has NO fallback: if the shared library is missing, fails to load, or a tensor is not
a CUDA tensor, the call raises.  Build with ``python -c 'import __graft_entry__ as g; g.build()'``
or ``make -C mm-pde_b200/csrc``.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmmpde_b200.so")
_lib = None

_p, _i, _l, _f, _d = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_double

class WgradTask(ctypes.Structure):
    """include/mmpde_b200.h: mmpde_wgrad_task."""
    _fields_ = [("A", _p), ("lda", _l), ("B", _p), ("ldb", _l), ("Bext", _p), ("dW", _p), ("ldw", _l),
                ("dWext", _p), ("ldwext", _l), ("dbias", _p), ("M", _l)]


class WimgTask(ctypes.Structure):
    """include/mmpde_b200.h: mmpde_wimg_task."""
    _fields_ = [("W", _p), ("w_ns", _l), ("w_ks", _l), ("scale", _f), ("image", _p)]


class KnnTask(ctypes.Structure):
    """include/mmpde_b200.h: mmpde_knn_task."""
    _fields_ = [("pts", _p), ("pts_off", _p), ("qry", _p), ("qry_off", _p), ("n_samples", ctypes.c_int32), ("k", ctypes.c_int32),
                ("n_queries", _l), ("x0", _f), ("y0", _f), ("inv_cell", _f), ("gx", ctypes.c_int32), ("gy", ctypes.c_int32),
                ("cell_start", _p), ("order", _p), ("rule", ctypes.c_int32), ("exclude_self", ctypes.c_int32), ("out_idx", _p)]


# name -> argtypes, exactly mirroring include/mmpde_b200.h
SIGNATURES = {
    "mmpde_abi_version": [],
    "mmpde_device_info": [_p, _p, _p],
    "mmpde_set_persistent_ctas": [_i],
    "mmpde_knn": [_p, _p, _p, _p, _i, _l, _i, _i, _i, _p, _p],
    "mmpde_knn_grid_build": [_p, _p, _i, _l, _f, _f, _f, _i, _i, _p, _p, _p, _p, _p],
    "mmpde_knn_grid": [_p, _p, _p, _p, _i, _l, _f, _f, _f, _i, _i, _p, _p, _i, _i, _i, _p, _p],
    "mmpde_knn_grid_multi": [_p, _i, _p],
    "mmpde_radius": [_p, _p, _i, _l, _f, _i, _p, _p],
    "mmpde_gemm": [_p, _l, _i, _p, _l, _i, _p, _l, _l, _i, _l, _p, _p, _l, _p, _i, _i, _i, _p],
    "mmpde_node_gemm": [_p, _l, _p, _l, _p, _l, _l, _p, _l, _l, _p, _p, _p, _i, _p, _l, _p, _l, _p, _l, _l, _p],
    "mmpde_weight_images": [_p, _i, _p],
    "mmpde_node_gemm_img": [_p, _l, _p, _l, _p, _p, _p, _p, _p, _i, _p, _l, _p, _l, _p, _l, _l, _p],
    "mmpde_node_wgrad": [_p, _l, _p, _l, _p, _p, _l, _p, _l, _p, _l, _p],
    "mmpde_node_wgrad_grouped": [_p, _i, _p],
    "mmpde_edge_fwd": [_p, _p, _p, _p, _l, _p, _p, _p, _l, _p, _p],
    "mmpde_edge_bwd": [_p, _p, _p, _p, _l, _p, _p, _p, _l, _p, _p, _p, _p],
    "mmpde_bn_stats": [_p, _l, _p, _l, _l, _p, _p],
    "mmpde_bn_finalize": [_p, _i, _d, _f, _f, _p, _p, _p, _p],
    "mmpde_bn_exchange": [_p, _i, _p, _i, _i, _p, _p],
    "mmpde_bn_exchange_set_timeout": [_d],
    "mmpde_bn_stats_fused": [_p, _l, _p, _l, _l, _p, _p, _d, _f, _f, _p, _p, _p, _p, _i, _i, _p],
    "mmpde_bn_bwd_reduce_fused": [_p, _l, _p, _l, _i, _p, _l, _p, _l, _l, _p, _p, _p, _p, _p, _p, _i, _i, _p],
    "mmpde_bn_bwd_reduce_post": [_p, _l, _p, _l, _i, _p, _l, _p, _l, _l, _p, _p, _p, _p, _p, _i, _i, _p],
    "mmpde_bn_exchange_wait": [_p, _i, _i, _p, _p],
    "mmpde_bn_apply": [_p, _l, _p, _l, _l, _p, _p, _p, _i, _p, _l, _p],
    "mmpde_bn_bwd_reduce": [_p, _l, _p, _l, _i, _p, _l, _p, _l, _l, _p, _p, _p],
    "mmpde_bn_bwd_apply": [_p, _l, _p, _l, _i, _p, _l, _p, _l, _l, _p, _p, _p, _d, _p, _l, _i, _p, _l, _p],
    "mmpde_relu_bwd": [_p, _l, _p, _l, _l, _p, _l, _p, _p],
    "mmpde_colsum": [_p, _l, _l, _i, _p, _p],
    "mmpde_decoder_fwd": [_p, _l, _l, _p, _f, _p, _p],
    "mmpde_decoder_fwd_acts": [_p, _l, _l, _p, _f, _p, _p, _p, _p],
    "mmpde_decoder_bwd": [_p, _l, _l, _p, _f, _p, _p, _l, _p, _p],
    "mmpde_outer_gate": [_p, _p, _p, _l, _p, _l, _l, _p],
    "mmpde_itp_fwd": [_p, _p, _p, _p, _l, _p, _p, _p],
    "mmpde_itp_bwd": [_p, _p, _p, _p, _l, _p, _p, _p, _p, _p],
    "mmpde_itp_fwd_tc": [_p, _p, _p, _p, _l, _p, _p, _p],
    "mmpde_itp_bwd_tc": [_p, _p, _p, _p, _l, _p, _p, _p, _p, _p, _p, _p, _p],
    "mmpde_rescut_fwd": [_p, _l, _i, _i, _p, _p, _p, _p],
    "mmpde_rescut_bwd": [_p, _l, _i, _i, _p, _p, _p, _p, _p, _p, _p],
    "mmpde_dmm_gnn_layer": [_p, _p, _p, _p, _l, _p, _p, _p],
    "mmpde_dmm_displacement": [_p, _p, _p, _i, _p, _p, _p, _i, _l, _l, _p, _p],
    "mmpde_rows_gather": [_p, _l, _p, _l, _i, _p, _p],
    "mmpde_rows_scatter_add": [_p, _p, _l, _i, _p, _l, _p],
    "mmpde_node4_linear": [_p, _p, _p, _p, _l, _l, _p],
    "mmpde_rows_dot": [_p, _l, _i, _p, _p, _l, _l, _i, _p],
}

launches = 0          # number of kernel-launching C-ABI calls made so far (bench.py reports the delta)


class MMPDEError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MMPDEError(f"{LIB_PATH} is missing: the CUDA extension is required (no CPU fallback). "
                             "Build it with `make -C mm-pde_b200/csrc`.")
        _lib = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(_lib, name)
            fn.argtypes = argtypes
            fn.restype = ctypes.c_int
    return _lib


def call(name, *args):
    """Invoke an entry point; raise on any non-zero status."""
    global launches
    rc = getattr(lib(), name)(*args)
    if rc != 0:
        kind = "argument error" if rc < 0 else "cudaError"
        raise MMPDEError(f"{name} failed: {kind} {rc}")
    launches += 1
    return rc
