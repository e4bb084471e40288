"""Loads libfemto_b200.so (the C ABI of include/femto_b200.h) and declares its prototypes.

There is deliberately no fallback: if the shared library is missing the import fails loudly and
tells the user how to build it.  The library itself has no CPU query path either.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfemto_b200.so")

i32, i64, u16, u8 = C.c_int32, C.c_int64, C.c_uint16, C.c_uint8
P = C.POINTER
vp = C.c_void_p


class FmInfo(C.Structure):
    _fields_ = [
        ("total_length", i64), ("num_documents", i64), ("num_blocks", i64),
        ("block_size", i32), ("bucket_size", i32), ("mark_period", i32), ("chunk_size", i32),
        ("first_row", i64), ("end_row", i64), ("hbm_bytes", i64), ("rank_block_bytes", i64),
        ("device", i32), ("max_code_len", i32), ("rank_block_size", i32), ("levels_per_block", i32),
    ]


# name -> (restype, argtypes); exactly the symbols declared in include/femto_b200.h
PROTOTYPES = {
    "fm_open": (C.c_int, [C.c_char_p, C.c_int, P(vp)]),
    "fm_open_shard": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, P(vp)]),
    "fm_close": (None, [vp]),
    "fm_info": (C.c_int, [vp, P(FmInfo)]),
    "fm_last_error": (C.c_char_p, []),
    "fm_count": (C.c_int, [vp, C.c_int, P(C.c_int), P(P(u16)), P(i64), P(i64)]),
    "fm_count_flat": (C.c_int, [vp, i64, P(i32), P(u16), P(i64), P(i64), P(i64)]),
    "fm_count_bytes": (C.c_int, [vp, i64, P(i32), P(u8), P(i64), P(i64), P(i64)]),
    "fm_count_device": (C.c_int, [vp, i64, vp, vp, vp, vp, vp, vp]),
    "fm_count_shard_step": (C.c_int, [vp, i64, vp, vp, vp, vp, vp, C.c_int, vp]),
    "fm_locate_shard_step": (C.c_int, [vp, i64, vp, vp, C.c_int, vp]),
    "fm_take_status": (C.c_int, [vp, vp, P(C.c_int)]),
    "fm_mesh_create": (C.c_int, [vp, C.c_int, C.c_int, i64, C.c_int, P(vp)]),
    "fm_mesh_destroy": (None, [vp]),
    "fm_mesh_export": (C.c_int, [vp, vp, i64]),
    "fm_mesh_connect": (C.c_int, [vp, vp, i64]),
    "fm_mesh_connect_local": (C.c_int, [vp, P(vp)]),
    "fm_mesh_set_limits": (C.c_int, [vp, C.c_int, C.c_double]),
    "fm_mesh_count": (C.c_int, [vp, vp, vp, vp, C.c_int, i64, i64, vp, vp, vp]),
    "fm_mesh_locate_rows": (C.c_int, [vp, i64, vp, vp, vp]),
    "fm_mesh_finish": (C.c_int, [vp, vp, P(C.c_int), P(C.c_uint64)]),
    "fm_last_transfer": (C.c_int, [vp, P(i64), P(i64)]),
    "fm_doc_name": (C.c_int, [vp, i64, vp, i64, P(i64)]),
    "fm_range_documents": (C.c_int, [vp, i64, i64, P(i64), i64, P(i64)]),
    "fm_chunk_documents": (C.c_int, [vp, i64, P(i64), P(i64), P(i64), i64, P(i64)]),
    "fm_locate": (C.c_int, [vp, C.c_int, P(C.c_int), P(P(u16)), C.c_int, P(C.c_int), P(P(i64))]),
    "fm_locate_flat": (C.c_int, [vp, i64, P(i32), P(u16), P(i64), C.c_int, P(i32), P(i64), P(i64), i64]),
    "fm_locate_range": (C.c_int, [vp, i64, i64, P(i64)]),
    "fm_locate_rows": (C.c_int, [vp, i64, P(i64), P(i64)]),
    "fm_locate_rows_device": (C.c_int, [vp, i64, vp, vp, vp]),
    "fm_back_step": (C.c_int, [vp, i64, P(i64), P(i32), P(i64), P(i64)]),
    "fm_occ": (C.c_int, [vp, i64, P(u16), P(i64), P(i64)]),
    "fm_backward_step": (C.c_int, [vp, i64, P(i64), P(i64), P(u16), P(i64), P(i64)]),
    "fm_doc_info": (C.c_int, [vp, i64, P(i64), P(i64)]),
    "fm_resolve": (C.c_int, [vp, i64, P(i64), P(i64), P(i64)]),
    "fm_extract": (C.c_int, [vp, i64, P(u16), i64, P(i64)]),
    "fm_extract_batch": (C.c_int, [vp, i64, P(i64), P(u16), i64, P(i64)]),
    "fm_generic_request": (C.c_int, [vp, C.c_char_p, P(C.c_char_p)]),
    "fm_host_alloc": (vp, [C.c_size_t]),
    "fm_host_free": (None, [vp]),
    "fm_kernel_launches": (i64, [vp]),
    "fm_set_lanes_per_query": (C.c_int, [vp, C.c_int]),
    "fm_set_count_schedule": (C.c_int, [vp, C.c_int, C.c_int]),
    "fm_set_default_block_bytes": (C.c_int, [C.c_int]),
    "fm_set_default_levels_per_block": (C.c_int, [C.c_int]),
    "fm_count_stats": (C.c_int, [vp, i64, P(i32), P(u16), P(i64), P(C.c_uint64)]),
    "fm_walk_stats": (C.c_int, [vp, i64, P(i64), P(C.c_uint64)]),
    "fm_probe_random_reads": (C.c_int, [vp, C.c_int, C.c_int, P(i64), P(C.c_double)]),
    "fm_builder_create": (C.c_int, [C.c_char_p, i64, i64, P(i64), i32, i32, i32, i32, C.c_int, P(vp)]),
    "fm_builder_append": (C.c_int, [vp, i64, P(u16), P(i64)]),
    "fm_builder_set_doc_info": (C.c_int, [vp, i64, vp, i64]),
    "fm_builder_finish": (C.c_int, [vp]),
    "fm_builder_abort": (None, [vp]),
    "fm_builder_create_range": (C.c_int, [C.c_char_p, i64, i64, P(i64), i32, i32, i32, i32, C.c_int, i64, i64, P(vp)]),
    "fm_builder_finish_range": (C.c_int, [vp, P(i64), P(i64)]),
    "fm_builder_write_header": (C.c_int, [C.c_char_p, i64, i64, P(i64), i32, i32, i32, i32, P(i64), P(i64), P(vp),
                                          P(i64)]),
    "fm_flatten": (C.c_int, [C.c_char_p, C.c_char_p]),
    "fm_suffix_sort_host": (C.c_int, [P(u16), i64, P(i64)]),
}

# loader/builder inspection hooks used by the CPU test-suite (not part of the public header)
DEBUG_PROTOTYPES = {
    "fm_debug_bseq_encode": (C.c_int, [C.c_char_p, i64, C.c_int, P(vp), P(i64)]),
    "fm_debug_bseq_expand": (C.c_int, [C.c_char_p, i64, C.c_char_p, i64, P(i64)]),
    "fm_debug_free": (None, [vp]),
    "fm_debug_image_open": (vp, [C.c_char_p, C.c_int, C.c_int, C.c_int, P(C.c_int)]),
    "fm_debug_image_close": (None, [vp]),
    "fm_debug_image_stats": (None, [vp, P(i64)]),
    "fm_debug_image_occ": (i64, [vp, C.c_int, i64]),
    "fm_debug_image_back_step": (C.c_int, [vp, i64, P(i32), P(i64), P(i64)]),
}

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (or `make -C femto_b200/csrc`). "
            "femto_b200 has no pure-Python or CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for table in (PROTOTYPES, DEBUG_PROTOTYPES):
        for name, (res, args) in table.items():
            fn = getattr(lib, name)  # AttributeError here == a declared symbol is not exported
            fn.restype = res
            fn.argtypes = args
    _lib = lib
    return lib
