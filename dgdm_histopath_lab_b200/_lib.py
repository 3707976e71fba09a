"""ctypes binding of libdgdm_b200.so (include/dgdm_b200.h). There is no CPU fallback: if the library is
missing or cannot be loaded every op raises."""
import ctypes as C
import os

from . import build as _build

_LIB = None

c_void_p, c_int, c_float, c_size_t = C.c_void_p, C.c_int, C.c_float, C.c_size_t


class GemmParams(C.Structure):
    _fields_ = [
        ("A", c_void_p), ("lda", c_int), ("M", c_int), ("K", c_int),
        ("W_hi", c_void_p), ("W_lo", c_void_p), ("ldw", c_int),
        ("N", c_int),
        ("C", c_void_p), ("ldc", c_int),
        ("bias", c_void_p),
        ("r_pre", c_void_p), ("ld_r_pre", c_int),
        ("r_post", c_void_p), ("ld_r_post", c_int),
        ("act_pre", c_int),
        ("ln1_g", c_void_p), ("ln1_b", c_void_p),
        ("ln2_g", c_void_p), ("ln2_b", c_void_p),
        ("ln_eps", c_float),
        ("act_post", c_int),
        ("dot_w", c_void_p), ("dot_b", c_float), ("dot_act", c_int), ("dot_out", c_void_p),
        ("a_scale", c_float),
        ("row_bias_index", c_void_p), ("row_bias", c_void_p), ("ld_row_bias", c_int),
        ("C_peers", c_void_p * 8), ("n_peers", c_int), ("peers_multicast", c_int),
    ]


# name -> (restype, argtypes); must list every symbol include/dgdm_b200.h declares (tests/test_abi.py checks)
SIGNATURES = {
    "dgdm_abi_version": (c_int, []),
    "dgdm_device_check": (c_int, [c_int]),
    "dgdm_gemm_f16x2": (c_int, [C.POINTER(GemmParams), c_void_p]),
    "dgdm_gemm_simt": (c_int, [C.POINTER(GemmParams), c_void_p, c_void_p]),
    "dgdm_split_weight_f16": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "dgdm_rowwise_epilogue": (c_int, [c_void_p, c_int, c_int, c_int, C.POINTER(GemmParams), c_void_p]),
    "dgdm_csr_build_work_ints": (c_size_t, [c_int, c_int]),
    "dgdm_csr_build": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dgdm_spmm_gcn": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_int,
                              c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "dgdm_edge_attr_aggregate": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "dgdm_topk_pool_work_ints": (c_size_t, [c_int]),
    "dgdm_topk_select": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dgdm_pool_gather": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p]),
    "dgdm_csr_pool": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_void_p, c_void_p]),
    "dgdm_unpool_add_act": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "dgdm_segment_softmax_scores": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int,
                                            c_void_p, c_void_p]),
    "dgdm_segment_minmax": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "dgdm_spatial_pe_add": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p]),
    "dgdm_spatial_attention_work_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "dgdm_spatial_attention": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int,
                                       c_int, c_int, c_int, c_int, c_float, c_int, c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    "dgdm_spatial_attention_weights": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p]),
    "dgdm_mha_dense": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                               C.c_longlong, C.c_longlong, C.c_longlong, C.c_longlong, c_void_p, c_void_p, c_void_p, c_int, c_float,
                               c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "dgdm_spatial_attention_backward": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_int,
                                                c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "dgdm_spatial_attention_backward_tc": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_int,
                                                   c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "dgdm_spatial_attention_lse": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int,
                                           c_int, c_int, c_int, c_int, c_float, c_int, c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "dgdm_attention_pool_work_floats": (c_size_t, [c_int, c_int, c_int, c_int]),
    "dgdm_attention_pool": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "dgdm_diffusion_add_noise": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                         c_void_p, c_void_p]),
    "dgdm_groupnorm_silu": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p, c_int, c_void_p]),
    "dgdm_ddpm_update": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_float, c_float, c_float, c_size_t, c_void_p, c_void_p]),
    "dgdm_validate_inputs": (c_int, [c_void_p, c_size_t, c_void_p, c_size_t, c_int, c_void_p, c_void_p, c_void_p]),
    "dgdm_knn_cosine_work_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "dgdm_knn_cosine": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                c_void_p]),
    "dgdm_row_sqnorm_f64": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "dgdm_knn_cosine_tc_work_bytes": (c_size_t, [c_int]),
    "dgdm_knn_cosine_tc": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_size_t, c_void_p]),
    "dgdm_gemm_f16x2_strided": (c_int, [C.POINTER(GemmParams), C.c_longlong, c_void_p]),
    "dgdm_row_normalize": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_int, c_void_p]),
    "dgdm_contrastive_accumulate": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_float, c_void_p, c_void_p, c_void_p,
                                            c_void_p]),
    "dgdm_knn2d_grid_work_bytes": (c_size_t, [c_int]),
    "dgdm_knn2d": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
}


class NativeLibraryError(RuntimeError):
    pass


def lib_path():
    return _build.LIB_PATH


def load():
    """Load (once) and return the ctypes library with prototypes set. Raises NativeLibraryError if absent."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise NativeLibraryError(
            f"{path} not found: build it with `python -m dgdm_histopath_lab_b200.build` (nvcc, sm_100a). "
            "There is no CPU or PyTorch fallback for the DGDM hot path.")
    try:
        lib = C.CDLL(path)
    except OSError as e:
        raise NativeLibraryError(f"cannot load {path}: {e}") from e
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _LIB = lib
    return lib
