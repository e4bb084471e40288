"""The BWT-range shard map (fm_open_shard / the mesh kernels / femto_b200.sharded.shard_of_block): contiguous
block ranges, balanced by rows, identical in the C++ loader and the Python drivers."""
import ctypes as C
import os

import numpy as np
import pytest

from femto_b200 import _lib, sharded


def c_shard_of_block(b, bs, n, g):
    lib = _lib.load()
    fn = lib.fm_debug_shard_of_block
    fn.restype = C.c_int
    fn.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_int]
    return fn(b, bs, n, g)


@pytest.mark.parametrize("n,bs", [((1 << 32) + 1, 1 << 27), (401, 16), (1467, 512), (100001, 65536), ((1 << 37) + 12345, 1 << 27),
                                  (17179885568, 1 << 27), (5, 16)])
@pytest.mark.parametrize("g", [1, 2, 3, 4, 8, 16])
def test_shard_map_is_contiguous_balanced_and_the_same_everywhere(n, bs, g):
    nb = (n + bs - 1) // bs
    step = max(1, nb // 3000)
    blocks = list(range(0, nb, step)) + [nb - 1]
    owners = [sharded.shard_of_block(b, bs, n, g) for b in blocks]
    assert owners == [c_shard_of_block(b, bs, n, g) for b in blocks]
    assert owners == sorted(owners) and 0 <= owners[0] and owners[-1] <= g - 1       # contiguous ranges, in order
    if step == 1 and nb >= 2 * g:
        rows = np.zeros(g, dtype=np.int64)
        for b, o in zip(blocks[:-1], owners[:-1]):
            rows[o] += min(bs, n - b * bs)
        # no shard holds more than one block above the even share
        assert rows.max() <= n / g + bs


def test_4gib_index_over_8_gpus_gets_four_full_blocks_each():
    n, bs = (1 << 32) + 1, 1 << 27
    owners = [sharded.shard_of_block(b, bs, n, 8) for b in range(33)]
    assert [owners.count(r) for r in range(8)] == [4, 4, 4, 4, 4, 4, 4, 5]      # the 33rd block holds one row


def test_loader_shards_partition_the_rows(built_indexes):
    """The loader's residency ranges (fm_debug_image_open on a shard) tile [0, n) for several shard counts."""
    lib = _lib.load()
    lib.fm_debug_image_open.restype = C.c_void_p
    from oracle.bindings import Oracle
    for name in ("gen400_small_blocks", "multi_doc_mixed", "acgt_64k"):
        path = built_indexes[name]
        with Oracle(path) as o:
            total = o.header_info()["total_length"]
        for g in (1, 2, 3, 5):
            prev_end = 0
            for r in range(g):
                err = C.c_int(0)
                h = lib.fm_debug_image_open(os.fsencode(path), r, g, 0, C.byref(err))
                assert h, err.value
                st = (C.c_int64 * 16)()
                lib.fm_debug_image_stats(C.c_void_p(h), st)
                lib.fm_debug_image_close(C.c_void_p(h))
                first_row, end_row = int(st[5]), int(st[6])     # fm_debug_image_stats: [5] first_row, [6] end_row
                assert first_row == prev_end or first_row == end_row     # an empty shard may sit anywhere
                if end_row > first_row:
                    prev_end = end_row
            assert prev_end == total
