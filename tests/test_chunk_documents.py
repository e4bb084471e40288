"""Document chunks: the per-chunk document lists femto stores in every bucket
(block_get_chunk, src/main/index.c:2147-2197; results encoding src/main/results.c:133-152,356-371),
read back by the loader (fmb::chunk_documents; C ABI fm_chunk_documents) -- host-side format work,
checked on CPU against (a) brute force: the documents of SA[first..last] from the oracle and (b) the
live reference's block_chunk_request where oracle/_ref exists."""
import ctypes as C

import numpy as np
import pytest

from oracle.bindings import Oracle, Reference, have_reference
from test_loader_image import Image

CASES = ["gen400_big_buckets", "gen400_small_buckets", "gen400_small_blocks", "gen13_small_blocks", "single_symbol",
         "multi_doc_mixed", "acgt_64k", "english_100k"]


def image_chunk(im, row, cap=1 << 16):
    fn = im.lib.fm_debug_image_chunk
    fn.restype = C.c_int64
    fn.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
    a, b = C.c_int64(), C.c_int64()
    docs = np.zeros(cap, dtype=np.int64)
    n = fn(im.h, row, C.byref(a), C.byref(b), docs.ctypes.data, cap)
    return n, a.value, b.value, docs[:max(n, 0)]


@pytest.mark.parametrize("name", CASES)
def test_chunks_match_brute_force_and_reference(name, built_indexes, corpora):
    path = built_indexes[name]
    params = corpora[name][1]
    im = Image(path, levels=4)
    try:
        with Oracle(path) as o:
            info = o.header_info()
            n, bs, cs = info["total_length"], info["block_size"], params["chunk_size"]
            sa = o.locate_range(0, n - 1)
            doc_of = np.array([o.resolve(int(x))[0] for x in sa], dtype=np.int64)
        ref = Reference(path) if have_reference() else None
        rows = sorted(set(list(range(0, n, max(1, cs // 2))) + [n - 1]))
        for row in rows:
            got_n, first, last, docs = image_chunk(im, row)
            row0 = (row // bs) * bs
            want_first = row0 + ((row - row0) // cs) * cs
            want_last = min(want_first + cs - 1, min(row0 + bs, n) - 1)
            assert (first, last) == (want_first, want_last) and first <= row <= last
            want = np.unique(doc_of[first:last + 1])
            assert got_n == len(want) and (docs == want).all(), row
            if ref is not None:
                rf, rl, rdocs = ref.chunk_documents(row)
                assert (rf, rl) == (first, last) and (rdocs == docs).all(), row
        if ref is not None:
            ref.close()
    finally:
        im.close()


def test_index_without_chunks_reports_missing(built_indexes):
    im = Image(built_indexes["skewed_deep"], levels=4)       # built with chunk_size = 0
    try:
        n, _, _, _ = image_chunk(im, 5)
        assert n == -8                                       # FM_ERR_MISSING (include/femto_b200.h)
    finally:
        im.close()
