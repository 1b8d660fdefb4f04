"""ctypes binding of include/swgpu.h.  Fails loudly when the CUDA library is missing."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class SwParams(C.Structure):
    """sw_params, include/sw_types.h (mirrors TilerMetaParameters, reference process/Tiler.h:64-75)."""

    _fields_ = [
        ("sampling", C.c_int32),
        ("tiling", C.c_int32),
        ("spacing_at_root", C.c_float),
        ("max_depth", C.c_uint32),
        ("max_points_per_node", C.c_uint64),
        ("bounds_min", C.c_double * 3),
        ("bounds_max", C.c_double * 3),
        ("concurrency", C.c_uint32),
        ("reserved", C.c_uint32),
    ]


class SwLasTransform(C.Structure):
    """sw_las_transform, include/sw_types.h (reference io/LASFile.cpp:79-94, process/TilerProcess.cpp:552-559)."""

    _fields_ = [
        ("scale", C.c_double * 3),
        ("offset", C.c_double * 3),
        ("header_min", C.c_double * 3),
        ("header_max", C.c_double * 3),
        ("center", C.c_double * 3),
        ("shift_to_center", C.c_int32),
        ("reserved", C.c_int32),
    ]


class SwgpuStats(C.Structure):
    _fields_ = [
        ("n_points", C.c_uint64),
        ("n_output_ids", C.c_uint64),
        ("n_nodes", C.c_uint64),
        ("n_levels", C.c_uint32),
        ("n_reconstruct_levels", C.c_uint32),
        ("sweep_points", C.c_uint64),
        ("bytes_index", C.c_uint64),
        ("bytes_sort", C.c_uint64),
        ("bytes_gather", C.c_uint64),
        ("bytes_sample", C.c_uint64),
        ("ms_index", C.c_float),
        ("ms_sort", C.c_float),
        ("ms_gather", C.c_float),
        ("ms_sample", C.c_float),
        ("ms_total", C.c_float),
        ("kernel_launches", C.c_uint32),
        ("min_distance_rounds", C.c_uint32),
        ("bytes_traffic", C.c_uint64),
        ("sort_passes", C.c_uint32),
        ("sort_first_bit", C.c_uint32),
        ("sort_fallback", C.c_uint32),
        ("ms_sort_finish", C.c_float),
        ("sort_scan_steps", C.c_uint64),
        ("sort_moved", C.c_uint64),
    ]


# swgpu_allreduce_u32_fn: int (*)(void* ctx, uint32_t* device_counters, uint64_t count, void* cuda_stream)
ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p)
# swgpu_allgatherv_fn: int (*)(void* ctx, const void* send, uint64_t send_bytes, void** recv, uint64_t* recv_bytes, void* stream)
ALLGATHERV_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64),
                            C.c_void_p)
MAX_RANKS = 16
PREFIX_BINS = 262144

# every symbol include/swgpu.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("swgpu_create", C.c_int, [C.POINTER(SwParams), C.c_int, C.POINTER(C.c_void_p)]),
    ("swgpu_destroy", None, [C.c_void_p]),
    ("swgpu_last_error", C.c_char_p, [C.c_void_p]),
    ("swgpu_set_stream", C.c_int, [C.c_void_p, C.c_void_p]),
    ("swgpu_reserve", C.c_int, [C.c_void_p, C.c_uint64]),
    ("swgpu_index_batch", C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
    ("swgpu_index_batch_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
    ("swgpu_index_batch_las", C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(SwLasTransform)]),
    ("swgpu_index_batch_las_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(SwLasTransform)]),
    ("swgpu_get_positions", C.c_int, [C.c_void_p, C.c_void_p]),
    ("swgpu_get_payload_pnts", C.c_int, [C.c_void_p, C.c_void_p]),
    ("swgpu_get_payload_pnts_device", C.c_int, [C.c_void_p, C.c_void_p]),
    ("swgpu_get_payload_las", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    ("swgpu_get_payload_las_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    ("swgpu_finalize", C.c_int, [C.c_void_p]),
    ("swgpu_set_multi_batch", C.c_int, [C.c_void_p, C.c_int]),
    ("swgpu_set_deep_node_policy", C.c_int, [C.c_void_p, C.c_int]),
    ("swgpu_set_sort_mode", C.c_int, [C.c_void_p, C.c_int]),
    ("swgpu_result_size", C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    ("swgpu_get_nodes", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    ("swgpu_get_nodes_device_ids", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    ("swgpu_get_start_level", C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    ("swgpu_get_clamped_count", C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    ("swgpu_get_keys", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    ("swgpu_gather_attribute_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]),
    ("swgpu_morton_encode_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    ("swgpu_sort_keys_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    ("swgpu_prefix_histogram_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    ("swgpu_prefix_histogram_coarse_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p]),
    ("swgpu_estimate_start_level", C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_int32)]),
    ("swgpu_choose_splitters", C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]),
    ("swgpu_max_shard_levels", C.c_int, [C.c_void_p, C.POINTER(C.c_uint32)]),
    ("swgpu_partition_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32,
                                         C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("swgpu_partition_to_peers_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32,
                                                  C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("swgpu_set_shard", C.c_int, [C.c_void_p, C.c_uint32, C.c_int32, ALLREDUCE_FN, C.c_void_p, C.c_void_p]),
    ("swgpu_set_partition_attributes", C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]),
    ("swgpu_set_shard_faces", C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, ALLGATHERV_FN, C.c_void_p]),
    ("swgpu_multi_create", C.c_int, [C.POINTER(SwParams), C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p)]),
    ("swgpu_multi_destroy", None, [C.c_void_p]),
    ("swgpu_multi_last_error", C.c_char_p, [C.c_void_p]),
    ("swgpu_multi_index_batch", C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32]),
    ("swgpu_multi_finalize", C.c_int, [C.c_void_p]),
    ("swgpu_multi_result_size", C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    ("swgpu_multi_get_nodes", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    ("swgpu_multi_get_info", C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64),
                                       C.c_void_p]),
    ("swgpu_multi_rank_result_size", C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    ("swgpu_multi_get_rank_attributes", C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("swgpu_multi_set_min_distance_faces", C.c_int, [C.c_void_p, C.c_int]),
    ("swgpu_enable_timing", C.c_int, [C.c_void_p, C.c_int]),
    ("swgpu_get_stats", C.c_int, [C.c_void_p, C.POINTER(SwgpuStats)]),
]


def library_path() -> str:
    # SWGPU_LIB selects an alternative build of the SAME library (kernel tuning experiments)
    return os.environ.get("SWGPU_LIB") or os.path.join(HERE, "libswgpu.so")


def load_library():
    """Loads libswgpu.so and binds every declared symbol.  No fallback of any kind."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise ImportError(
            "schwarzwald_b200/libswgpu.so is missing: build it with ./build_native.sh "
            "(or __graft_entry__.build()).  There is no CPU fallback for the tiler kernels.")
    lib = C.CDLL(path)
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = restype
        fn.argtypes = argtypes
    _LIB = lib
    return lib
