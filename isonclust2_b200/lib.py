"""ctypes binding of libisonclust2_b200.so (the C ABI declared in include/isonclust2_b200.h).

There is no fallback: if the library is missing or a symbol the header declares is absent, load() raises.
"""
import ctypes as C
import os
import re

# one CUDA stream per batch in flight (icb_cluster_batches): give them their own hardware queues.  Read by the driver when
# the process initialises CUDA, so it is set at import time (a value the user exported wins).
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libisonclust2_b200.so")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "isonclust2_b200.h")

_lib = None
_p, _i64, _i32, _dbl = C.c_void_p, C.c_int64, C.c_int, C.c_double


def header_symbols():
    """Every function name include/isonclust2_b200.h declares."""
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(icb_[a-z0-9_]+)\s*\(", txt)))


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing — build it with `python __graft_entry__.py` (nvcc, sm_100a). "
                           "There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    missing = [s for s in header_symbols() if not hasattr(L, s)]
    if missing:
        raise RuntimeError(f"{LIB_PATH} does not export: {missing}")
    L.icb_device_count.restype = _i32
    L.icb_ctx_create.argtypes = [_i32, C.POINTER(_p)]
    L.icb_ctx_destroy.argtypes = [_p]
    L.icb_ctx_destroy.restype = None
    L.icb_ctx_sync.argtypes = [_p]
    L.icb_ctx_stream.argtypes = [_p]
    L.icb_ctx_stream.restype = _p
    L.icb_last_error.argtypes = [_p]
    L.icb_last_error.restype = C.c_char_p
    L.icb_strerror.argtypes = [_i32]
    L.icb_strerror.restype = C.c_char_p
    L.icb_ctx_timing.argtypes = [_p, _i32]
    L.icb_ctx_timing_reset.argtypes = [_p]
    L.icb_ctx_kernel_time.argtypes = [_p, C.c_char_p, C.POINTER(_i64), C.POINTER(_dbl)]
    L.icb_ctx_launch_count.argtypes = [_p]
    L.icb_ctx_launch_count.restype = _i64
    L.icb_debug_residency_start.argtypes = [_p, _i64]
    L.icb_debug_residency_dump.argtypes = [_p, _p, _i64, C.POINTER(_i64)]
    L.icb_readset_create.argtypes = [_p, _p, _p, _p, _i64, C.POINTER(_p)]
    L.icb_readset_destroy.argtypes = [_p]
    L.icb_readset_destroy.restype = None
    L.icb_readset_create_reserve.argtypes = [_p, _p, _p, _p, _i64, _i64, _i64, _i64, C.POINTER(_p)]
    L.icb_readset_append_from.argtypes = [_p, _p, _i64, _dbl, _dbl, _dbl, C.POINTER(_i64)]
    L.icb_readset_set_errors.argtypes = [_p, _p, _p, _p]
    L.icb_poa_graph_create.argtypes = [C.POINTER(_p)]
    L.icb_poa_graph_destroy.argtypes = [_p]
    L.icb_poa_graph_destroy.restype = None
    L.icb_poa_graph_load.argtypes = [_p, _i64, C.POINTER(_p)]
    L.icb_poa_graph_save.argtypes = [_p, _p, _i64, C.POINTER(_i64)]
    L.icb_poa_graph_sizes.argtypes = [_p, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]
    L.icb_poa_graph_clear.argtypes = [_p]
    L.icb_poa_align.argtypes = [_p, C.c_char_p, _i64, _i32, _p, _p, _i64, C.POINTER(_i64), C.POINTER(C.c_int32)]
    L.icb_poa_graph_add.argtypes = [_p, C.c_char_p, _i64, C.c_uint32, _i32]
    L.icb_poa_graph_consensus.argtypes = [_p, _p, _i64, C.POINTER(_i64)]
    L.icb_readset_n_reads.argtypes = [_p]
    L.icb_readset_n_reads.restype = _i64
    L.icb_readset_n_bases.argtypes = [_p]
    L.icb_readset_n_bases.restype = _i64
    L.icb_score_reads.argtypes = [_p, _i32, _p, _p]
    L.icb_scores_download.argtypes = [_p, _p, _p]
    L.icb_sketch_reads.argtypes = [_p, _i32, _i32, _p, _p]
    L.icb_sketch_sizes.argtypes = [_p, C.POINTER(_i64), C.POINTER(_i64)]
    L.icb_sketch_download.argtypes = [_p] * 10
    L.icb_align_pairs.argtypes = [_p, _p, _i64, _i32, _p]
    L.icb_clusterer_create.argtypes = [_p, _p, C.POINTER(_p)]
    L.icb_clusterer_destroy.argtypes = [_p]
    L.icb_clusterer_destroy.restype = None
    L.icb_cluster_reads.argtypes = [_p, _p, _i64, _p, _p]
    L.icb_clusterer_n_clusters.argtypes = [_p]
    L.icb_clusterer_n_clusters.restype = _i64
    L.icb_clusterer_founders.argtypes = [_p, _p]
    L.icb_clusterer_stats.argtypes = [_p, _p]
    L.icb_clusterer_seed.argtypes = [_p, _p, _i64]
    L.icb_clusterer_update_representative.argtypes = [_p, C.c_int32, C.c_int32]
    L.icb_cluster_batches.argtypes = [_p, _p, _p, _p, _i32, _i32, _p, _p, _p, _p]
    L.icb_cluster_tree.argtypes = [_p, _p, _p, _p, _i32, _i32, _i32, _p, _p, C.POINTER(_i64), _p]
    L.icb_init_qual_tab.argtypes = [_p, _i32]
    L.icb_init_qual_tab.restype = None
    L.icb_quality_filter.argtypes = [_p, _i64, _dbl, _p]
    L.icb_sort_by_score.argtypes = [_p, _i64, _p]
    # per-primitive face
    L.icb_hpc_compress.argtypes = [_p, _p, _p, _p, _i64, _p, _p, _p]
    L.icb_kmer_encode.argtypes = [_p, _p, _i64, _i32, _p, C.POINTER(_i64)]
    L.icb_kmer_minimizers.argtypes = [_p, _p, _i64, _i32, _i32, _p, C.POINTER(_i64)]
    L.icb_calc_error_rate.argtypes = [_p, _p, _p, _i64, _p, _p]
    L.icb_index_create.argtypes = [_p, C.POINTER(_p)]
    L.icb_index_destroy.argtypes = [_p]
    L.icb_index_destroy.restype = None
    L.icb_index_add.argtypes = [_p, C.c_uint32, _p, _i64]
    L.icb_index_update_cluster.argtypes = [_p, C.c_uint32, _p, _i64, _p, _i64]
    L.icb_index_query_hits.argtypes = [_p, _p, _i64, _p, _i64, _p, _i64, C.POINTER(_i64)]
    L.icb_mapped_length.argtypes = [_p, _p, _p, _i64, _i64, _i64, C.c_int32, C.POINTER(_i64)]
    L.icb_count_windows.argtypes = [_p, _p, _i64, _i32, _i32, C.POINTER(_i64)]
    L.icb_align_traceback.argtypes = [_p, _p, _i32, _p, _p, _i64]
    L.icb_load_p_shared.argtypes = [C.c_char_p, _i32, _i32, _p]
    _lib = L
    return L
