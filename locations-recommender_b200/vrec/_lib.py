"""ctypes binding of libvrec.so (include/vrec.h).  There is no CPU fallback: if the
CUDA library is missing or no sm_100 device is present, this fails loudly."""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# VREC_LIB_PATH: an alternative build of the same library (A/B measurements of kernel variants)
LIB_PATH = os.environ.get("VREC_LIB_PATH") or os.path.join(_PKG, "libvrec.so")

OK, ENOENT, ENOMEM, ENODEV, EINVAL, ECUDA, ENCCL = 0, -2, -12, -19, -22, -100, -101

i64p = C.POINTER(C.c_int64)
i32p = C.POINTER(C.c_int32)
f64p = C.POINTER(C.c_double)
vp = C.c_void_p

# name -> (restype, argtypes); mirrors include/vrec.h one to one
SIGNATURES = {
    "vrec_init": (C.c_int, [C.c_int, C.POINTER(vp)]),
    "vrec_shutdown": (None, [vp]),
    "vrec_last_error": (C.c_char_p, []),
    "vrec_abi_version": (C.c_int, []),
    "vrec_stream": (vp, [vp]),
    "vrec_launch_count": (C.c_int64, [vp]),
    "vrec_synchronize": (C.c_int, [vp]),
    "vrec_comm_unique_id": (C.c_int, [vp]),
    "vrec_comm_init": (C.c_int, [vp, C.c_int, C.c_int, vp]),
    "vrec_comm_rank": (C.c_int, [vp]),
    "vrec_comm_world": (C.c_int, [vp]),
    "vrec_sg_load_partitioned": (C.c_int, [vp, C.c_int64, i64p, i64p, f64p, C.POINTER(vp)]),
    "vrec_knn_load": (C.c_int, [vp, C.c_int64, i64p, i64p, i32p, f64p, C.c_int32, i64p, i32p, f64p, C.c_int32,
                                C.c_int64, i64p, i64p, i64p, C.POINTER(vp)]),
    "vrec_knn_free": (None, [vp]),
    "vrec_knn_query": (C.c_int, [vp, i64p, C.c_int32, C.c_double, C.c_double, C.c_int32, i64p, C.c_int64,
                                 C.c_int32, i64p, f64p, i32p, i32p]),
    "vrec_knn_set_filter": (C.c_int, [vp, i64p, C.c_int64]),
    "vrec_knn_query_device": (C.c_int, [vp, vp, C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_int32,
                                        vp, vp, vp, vp]),
    "vrec_knn_neighbours": (C.c_int, [vp, C.c_int64, C.c_double, C.c_double, C.c_int32, i64p, f64p, C.c_int32,
                                      i32p]),
    "vrec_knn_estimates": (C.c_int, [vp, C.c_int64, C.c_double, C.c_double, C.c_int32, i64p, f64p, C.c_int64,
                                     i64p]),
    "vrec_knn_person_ids": (C.c_int, [vp, i64p]),
    "vrec_knn_similarities": (C.c_int, [vp, C.c_int64, C.c_double, C.c_double, C.c_int32, f64p]),
    "vrec_knn_set_option": (C.c_int, [vp, C.c_char_p, C.c_int64]),
    "vrec_knn_resident_bytes": (C.c_int64, [vp]),
    "vrec_knn_debug_last_neighbours": (C.c_int, [vp, C.c_int32, C.c_int32, i64p, f64p, i32p]),
    "vrec_knn_debug_stats": (C.c_int, [vp, C.POINTER(C.c_uint64)]),
    "vrec_knn_debug_tc_cycles": (C.c_int, [vp, C.POINTER(C.c_uint64)]),
    "vrec_knn_debug_tc_block_cycles": (C.c_int, [vp, C.POINTER(C.c_uint64), C.c_int]),
    "vrec_knn_debug_probe": (C.c_int, [vp, C.POINTER(C.c_uint64)]),
    "vrec_knn_last_dense_ms": (C.c_int, [vp, C.POINTER(C.c_double)]),
    "vrec_build_rating_vectors": (C.c_int, [vp, C.c_int64, i64p, i64p, i64p, C.c_int32, C.POINTER(C.c_int64),
                                            C.POINTER(C.c_int64), i64p, i64p, i32p, f64p, C.POINTER(C.c_int32)]),
    "vrec_build_place_visits": (C.c_int, [vp, C.c_int64, i64p, f64p, f64p, i64p, i64p, C.c_int64, i64p, f64p, f64p, i64p,
                                          i64p, C.c_int32, C.c_double, C.c_int64, C.POINTER(C.c_int64), i64p, i64p, i64p,
                                          i64p, i64p]),
    "vrec_build_edge_family": (C.c_int, [vp, C.c_int64, i64p, i64p, i64p, C.c_int32, C.c_double, C.c_int64,
                                         C.POINTER(C.c_int64), i64p, i64p, f64p]),
    "vrec_build_stochastic_graph": (C.c_int, [vp, C.c_int64, i64p, i64p, i64p, i64p, C.c_double, C.c_double, C.c_int64,
                                              C.POINTER(C.c_int64), i64p, i64p, f64p]),
    "vrec_sg_load": (C.c_int, [vp, C.c_int64, i64p, i64p, f64p, C.POINTER(vp)]),
    "vrec_sg_free": (None, [vp]),
    "vrec_sg_vertex_count": (C.c_int64, [vp]),
    "vrec_sg_edge_count": (C.c_int64, [vp]),
    "vrec_sg_vertex_ids": (C.c_int, [vp, i64p]),
    "vrec_sg_query": (C.c_int, [vp, i64p, C.c_int32, C.c_double, C.c_int32, i64p, C.c_int64, C.c_int32,
                                i64p, f64p, i32p, i32p, i32p, i32p]),
    "vrec_sg_stationary": (C.c_int, [vp, C.c_int64, C.c_double, C.c_int32, f64p, i32p, i32p, f64p]),
    "vrec_sg_iterate_device": (C.c_int, [vp, C.c_int32]),
    "vrec_sg_set_option": (C.c_int, [vp, C.c_char_p, C.c_int32]),
    "vrec_sg_batch_info": (C.c_int64, [vp, C.c_int32]),
    "vrec_sg_resident_bytes": (C.c_int64, [vp]),
    "vrec_sg_generate": (C.c_int, [vp, C.c_int64, C.c_int32, C.c_uint64, C.c_int32, C.c_int32, C.POINTER(vp)]),
    "vrec_host_sg_csr": (C.c_int, [C.c_int64, i64p, i64p, f64p, i64p, i64p, i32p, i32p, f64p]),
    "vrec_sg_export_csr": (C.c_int, [vp, i32p, i32p, f64p]),
    "vrec_sg_row_range": (C.c_int, [vp, i64p, i64p]),
    "vrec_sg_group_load": (C.c_int, [vp, C.c_int32, C.c_int64, i64p, i64p, f64p, C.POINTER(vp)]),
    "vrec_sg_group_stationary": (C.c_int, [C.POINTER(vp), C.c_int32, C.c_int64, C.c_double, C.c_int32, f64p, i32p,
                                           i32p, f64p]),
    "vrec_host_sg_partition": (C.c_int, [C.c_int64, i32p, C.c_int32, i64p]),
    "vrec_debug_tc_selftest": (C.c_int, [vp, f64p]),
}

_lib = None


def load() -> C.CDLL:
    """Loads libvrec.so; raises if it has not been built (no fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(libvrec has no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)       # AttributeError if the library lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error() -> str:
    return (load().vrec_last_error() or b"").decode("utf-8", "replace")
