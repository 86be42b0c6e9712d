"""ctypes loader for libwiggletools_b200.so (the C ABI in include/wiggletools_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C
wiggletools_b200/csrc``.  There is no fallback: a missing library or a missing
CUDA device raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# WB200_LIB: development override (kernel experiments load alternative builds of the same library)
LIB_PATH = os.environ.get("WB200_LIB") or os.path.join(HERE, "libwiggletools_b200.so")


class CTrack(C.Structure):
    _fields_ = [("start", C.c_void_p), ("end", C.c_void_p), ("value", C.c_void_p),
                ("n", C.c_int64), ("default_value", C.c_double), ("value32", C.c_void_p)]


class CSegments(C.Structure):
    _fields_ = [("start", C.c_void_p), ("end", C.c_void_p), ("value", C.c_void_p),
                ("n", C.c_int64), ("default_value", C.c_double), ("auc", C.c_double),
                ("n_ends", C.c_int64), ("end_index", C.c_void_p), ("end_value", C.c_void_p)]


class CStats(C.Structure):
    _fields_ = [("intervals", C.c_int64), ("segments", C.c_int64), ("windows", C.c_int64),
                ("slow_segments", C.c_int64), ("kernels", C.c_int32), ("fast_path", C.c_int32),
                ("ms_count", C.c_float), ("ms_emit", C.c_float), ("ms_total", C.c_float),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("limbs", C.c_int32), ("sweep_kernel", C.c_int32)]


class CDeviceShard(C.Structure):
    """wb200_device_shard (include/wiggletools_b200.h)."""
    _fields_ = [("d_start", C.c_void_p), ("d_end", C.c_void_p), ("d_value", C.c_void_p), ("h_off", C.c_void_p),
                ("pos_lo", C.c_int32), ("pos_hi", C.c_int32),
                ("d_out_start", C.c_void_p), ("d_out_end", C.c_void_p), ("d_out_value", C.c_void_p), ("out_cap", C.c_int64),
                ("n_out", C.c_int64), ("auc", C.c_double), ("slow_segments", C.c_int64),
                ("ms_cursor", C.c_float), ("ms_sweep", C.c_float), ("limbs", C.c_int32), ("sweep_kernel", C.c_int32)]


class CMap(C.Structure):
    _fields_ = [("op", C.c_int32), ("reserved", C.c_int32), ("arg", C.c_double)]


class CSinkState(C.Structure):
    _fields_ = [("records", C.c_int64), ("point_by_point", C.c_int32), ("reserved", C.c_int32)]


_lib = None


def load():
    """Load the shared library (once) and declare its signatures."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "wiggletools_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C wiggletools_b200/csrc`; there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.wb200_create.restype = C.c_void_p
    lib.wb200_create.argtypes = [C.c_int]
    lib.wb200_destroy.argtypes = [C.c_void_p]
    lib.wb200_last_error.restype = C.c_char_p
    lib.wb200_last_error.argtypes = [C.c_void_p]
    lib.wb200_last_stats.restype = C.POINTER(CStats)
    lib.wb200_last_stats.argtypes = [C.c_void_p]
    lib.wb200_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    lib.wb200_set_window.argtypes = [C.c_void_p, C.c_int]
    lib.wb200_set_breakpoint_output.argtypes = [C.c_void_p, C.c_int]
    lib.wb200_bind_host_thread.argtypes = [C.c_void_p]
    lib.wb200_host_alloc.restype = C.c_void_p
    lib.wb200_host_alloc.argtypes = [C.c_int64]
    lib.wb200_host_free.argtypes = [C.c_void_p]
    lib.wb200_reduce_host.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.wb200_rows_host.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
    lib.wb200_reduce_host_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.wb200_reduce_device.argtypes = [
        C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.wb200_reduce_device_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.wb200_pearson_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.wb200_pearson_pair_host.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.wb200_pearson_merge.argtypes = [C.c_void_p, C.c_void_p]
    lib.wb200_pearson_result.restype = C.c_double
    lib.wb200_pearson_result.argtypes = [C.c_void_p]
    lib.wb200_bw_open.restype = C.c_void_p
    lib.wb200_bw_open.argtypes = [C.c_char_p]
    lib.wb200_bw_close.argtypes = [C.c_void_p]
    lib.wb200_bw_last_error.restype = C.c_char_p
    lib.wb200_bw_last_error.argtypes = [C.c_void_p]
    lib.wb200_bw_nchrom.argtypes = [C.c_void_p]
    lib.wb200_bw_chrom_name.restype = C.c_char_p
    lib.wb200_bw_chrom_name.argtypes = [C.c_void_p, C.c_int]
    lib.wb200_bw_chrom_length.restype = C.c_int64
    lib.wb200_bw_chrom_length.argtypes = [C.c_void_p, C.c_int]
    lib.wb200_bw_set_stretch.argtypes = [C.c_void_p, C.c_int]
    lib.wb200_bw_decode_chrom.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.wb200_debug_phases.argtypes = [C.c_void_p, C.c_void_p]
    lib.wb200_debug_format_value.argtypes = [C.c_double, C.c_char_p]
    lib.wb200_set_input_maps.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.wb200_map_default.restype = C.c_double
    lib.wb200_map_default.argtypes = [C.c_int, C.c_double, C.c_double]
    lib.wb200_format_device.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                        C.c_void_p, C.c_void_p, C.c_void_p]
    lib.wb200_format_host.argtypes = lib.wb200_format_device.argtypes
    lib.wb200_reduce_format_host.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_char_p,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.wb200_apply_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_int64,
                                       C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    lib.wb200_apply_host.argtypes = lib.wb200_apply_device.argtypes
    lib.wb200_apply_last.argtypes = [C.c_void_p, C.c_double, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    lib.wb200_debug_tdist_device.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.wb200_debug_window_smem_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.wb200_debug_window_smem_bytes.restype = C.c_longlong
    lib.wb200_multi_create.restype = C.c_void_p
    lib.wb200_multi_create.argtypes = [C.c_int, C.c_void_p]
    lib.wb200_multi_destroy.argtypes = [C.c_void_p]
    lib.wb200_multi_last_error.restype = C.c_char_p
    lib.wb200_multi_last_error.argtypes = [C.c_void_p]
    lib.wb200_multi_ndev.argtypes = [C.c_void_p]
    lib.wb200_multi_owner.argtypes = [C.c_void_p, C.c_int]
    lib.wb200_multi_reduce_host.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.wb200_multi_pearson_host.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.wb200_tdist_Q.restype = C.c_double
    lib.wb200_tdist_Q.argtypes = [C.c_double, C.c_double]
    _lib = lib
    return lib


# every symbol include/wiggletools_b200.h declares
EXPORTS = [
    "wb200_create", "wb200_destroy", "wb200_last_error", "wb200_last_stats", "wb200_set_stream",
    "wb200_set_window", "wb200_set_breakpoint_output", "wb200_bind_host_thread", "wb200_host_alloc", "wb200_host_free", "wb200_reduce_host", "wb200_rows_host", "wb200_reduce_host_batch",
    "wb200_reduce_device", "wb200_reduce_device_batch", "wb200_pearson_host", "wb200_pearson_pair_host", "wb200_pearson_merge", "wb200_pearson_result",
    "wb200_tdist_Q", "wb200_bw_open", "wb200_bw_close", "wb200_bw_last_error", "wb200_bw_nchrom", "wb200_bw_chrom_name",
    "wb200_bw_chrom_length", "wb200_bw_set_stretch", "wb200_bw_decode_chrom", "wb200_debug_phases", "wb200_debug_format_value",
    "wb200_format_device", "wb200_format_host", "wb200_reduce_format_host", "wb200_set_input_maps", "wb200_map_default",
    "wb200_apply_device", "wb200_apply_host", "wb200_apply_last", "wb200_debug_tdist_device", "wb200_debug_window_smem_bytes",
    "wb200_multi_create", "wb200_multi_destroy", "wb200_multi_last_error", "wb200_multi_ndev", "wb200_multi_owner",
    "wb200_multi_reduce_host", "wb200_multi_pearson_host",
]
