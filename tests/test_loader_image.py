"""The load-time decode of the on-disk format into the HBM rank image, verified on the CPU.

fm_debug_image_* (femto_b200/csrc/fm_debug.cc) walk the HOST copy of the image -- rank blocks,
node records, folded Occ bases, mark bit-vectors, SA samples -- with plain loops; they are
compared with the oracle for every row x symbol on the small corpora and on samples of the larger
ones.  This isolates "the image is right" from "the kernel reads it right" (tests -m gpu).
"""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from femto_b200 import _lib
from oracle.bindings import Oracle


class Image:
    def __init__(self, path, shard=0, nshards=1, block_bytes=128, levels=1):
        self.lib = _lib.load()
        err = C.c_int()
        assert self.lib.fm_set_default_block_bytes(block_bytes) == 0
        assert self.lib.fm_set_default_levels_per_block(levels) == 0
        try:
            self.h = self.lib.fm_debug_image_open(os.fsencode(path), shard, nshards, 4, C.byref(err))
        finally:
            self.lib.fm_set_default_block_bytes(0)
            self.lib.fm_set_default_levels_per_block(0)
        assert self.h, f"image open failed err={err.value}"

    def stats(self):
        out = (C.c_int64 * 8)()
        self.lib.fm_debug_image_stats(self.h, out)
        keys = ["rank_blocks", "wtree_blocks", "nodes", "buckets", "markvals", "first_row", "end_row", "max_code_len"]
        return dict(zip(keys, list(out)))

    def occ(self, ch, row):
        return self.lib.fm_debug_image_occ(self.h, ch, row)

    def back_step(self, row):
        ch, nxt, off = C.c_int32(), C.c_int64(), C.c_int64()
        rc = self.lib.fm_debug_image_back_step(self.h, row, C.byref(ch), C.byref(nxt), C.byref(off))
        assert rc == 0, rc
        return ch.value, nxt.value, off.value

    def close(self):
        self.lib.fm_debug_image_close(self.h)


@pytest.mark.parametrize("name", ["two_docs", "gen400_big_buckets", "gen400_small_buckets", "gen400_small_blocks",
                                  "gen13_small_blocks", "gen3", "single_symbol", "multi_doc_mixed"])
def test_image_matches_oracle_exhaustive(name, built_indexes):
    path = built_indexes[name]
    im = Image(path)
    with Oracle(path) as o:
        n = o.header_info()["total_length"]
        assert im.stats()["first_row"] == 0 and im.stats()["end_row"] == n
        step = 1 if n <= 500 else 5
        for row in range(0, n, step):
            assert im.back_step(row) == o.back_step(row), row
            for ch in range(261):
                assert im.occ(ch, row) == o.occ(ch, row)[0], (ch, row)
    im.close()


@pytest.mark.parametrize("block_bytes", [64, 32])
@pytest.mark.parametrize("name", ["two_docs", "gen400_small_blocks", "multi_doc_mixed"])
def test_image_small_rank_blocks_exhaustive(name, block_bytes, built_indexes):
    """The 64- and 32-byte rank block layouts decode to the same Occ / LF / marks."""
    path = built_indexes[name]
    im = Image(path, block_bytes=block_bytes)
    with Oracle(path) as o:
        n = o.header_info()["total_length"]
        for row in range(0, n, 1 if n <= 500 else 3):
            assert im.back_step(row) == o.back_step(row), row
            for ch in range(0, 261, 2):
                assert im.occ(ch, row) == o.occ(ch, row)[0], (ch, row)
    im.close()


@pytest.mark.parametrize("block_bytes", [128, 64])
@pytest.mark.parametrize("name", ["two_docs", "gen400_small_blocks", "gen400_big_buckets", "gen13_small_blocks", "gen3",
                                  "single_symbol", "multi_doc_mixed"])
def test_paired_level_image_exhaustive(name, block_bytes, built_indexes):
    """The paired-level layout (two wavelet-tree levels per block, fm_image.hpp) decodes to the same
    Occ / LF / marks as the reference for every row and symbol."""
    path = built_indexes[name]
    im = Image(path, block_bytes=block_bytes, levels=2)
    with Oracle(path) as o:
        n = o.header_info()["total_length"]
        for row in range(0, n, 1 if n <= 500 else 3):
            assert im.back_step(row) == o.back_step(row), row
            for ch in range(261):
                assert im.occ(ch, row) == o.occ(ch, row)[0], (ch, row)
    im.close()


@pytest.mark.parametrize("name", ["two_docs", "gen400_small_blocks", "gen400_big_buckets", "gen400_small_buckets",
                                  "gen13_small_blocks", "gen3", "single_symbol", "multi_doc_mixed"])
def test_quad_level_image_exhaustive(name, built_indexes):
    """The quad-level layout (four wavelet-tree levels per 128-byte block, fm_image.hpp) decodes to
    the same Occ / LF / marks as the reference for every row and symbol."""
    path = built_indexes[name]
    im = Image(path, block_bytes=128, levels=4)
    with Oracle(path) as o:
        n = o.header_info()["total_length"]
        for row in range(0, n, 1 if n <= 500 else 3):
            assert im.back_step(row) == o.back_step(row), row
            for ch in range(261):
                assert im.occ(ch, row) == o.occ(ch, row)[0], (ch, row)
    im.close()


def test_paired_level_layout_needs_wide_blocks(built_indexes):
    lib = _lib.load()
    err = C.c_int()
    lib.fm_set_default_block_bytes(32)
    lib.fm_set_default_levels_per_block(2)
    try:
        assert not lib.fm_debug_image_open(os.fsencode(built_indexes["two_docs"]), 0, 1, 1, C.byref(err))
        assert err.value != 0
    finally:
        lib.fm_set_default_block_bytes(0)
        lib.fm_set_default_levels_per_block(0)


@pytest.mark.parametrize("layout", [(128, 1), (64, 1), (32, 1), (128, 2), (64, 2), (128, 4)])
@pytest.mark.parametrize("name", ["acgt_64k", "bytes_200k", "skewed_deep", "english_100k"])
def test_image_matches_oracle_sampled(name, layout, built_indexes):
    path = built_indexes[name]
    im = Image(path, block_bytes=layout[0], levels=layout[1])
    rng = np.random.default_rng(9)
    with Oracle(path) as o:
        n = o.header_info()["total_length"]
        rows = np.concatenate([rng.integers(0, n, 600), [0, n - 1]])
        for row in rows:
            assert im.back_step(int(row)) == o.back_step(int(row))
        for row, ch in zip(rng.integers(0, n, 4000), rng.integers(0, 261, 4000)):
            assert im.occ(int(ch), int(row)) == o.occ(int(ch), int(row))[0]
    st = im.stats()
    if name == "skewed_deep":
        assert st["max_code_len"] > 8          # deep Huffman tree really exercised
    im.close()


def test_image_of_reference_built_golden_indexes():
    for case in sorted(os.listdir(GOLDEN_DIR)):
        idx = os.path.join(GOLDEN_DIR, case, "index")
        if not os.path.isdir(idx):
            continue
        im = Image(idx)
        with Oracle(idx) as o:
            n = o.header_info()["total_length"]
            for row in range(n):
                assert im.back_step(row) == o.back_step(row)
                for ch in range(0, 261, 3):
                    assert im.occ(ch, row) == o.occ(ch, row)[0]
        im.close()


def test_sharded_images_cover_the_index(built_indexes):
    """BWT row-range sharding by data block: shards are disjoint, contiguous and agree with the oracle."""
    path = built_indexes["acgt_64k"]            # 2 data blocks of 32768 rows
    with Oracle(path) as o:
        n = o.header_info()["total_length"]
        nblocks = o.header_info()["nblocks"]
        assert nblocks >= 2
        covered = 0
        for shard in range(2):
            im = Image(path, shard, 2)
            st = im.stats()
            assert st["first_row"] == covered
            covered = st["end_row"]
            rng = np.random.default_rng(shard)
            for row in rng.integers(st["first_row"], st["end_row"], 200):
                assert im.back_step(int(row)) == o.back_step(int(row))
                assert im.occ(7, int(row)) == o.occ(7, int(row))[0]
            if st["end_row"] < n:
                assert im.occ(7, st["end_row"]) == -1        # rows of other shards are not resident
            if st["first_row"] > 0:
                assert im.occ(7, st["first_row"] - 1) == -1
            im.close()
        assert covered == n


def test_image_rejects_corrupt_index(built_indexes, tmp_path):
    import shutil
    src = built_indexes["two_docs"]
    bad = str(tmp_path / "bad")
    shutil.copytree(src, bad)
    data = bytearray(open(os.path.join(bad, "01"), "rb").read())
    data[0] ^= 0xFF                                  # break the data block magic
    open(os.path.join(bad, "01"), "wb").write(bytes(data))
    lib = _lib.load()
    err = C.c_int()
    assert not lib.fm_debug_image_open(os.fsencode(bad), 0, 1, 1, C.byref(err))
    assert err.value == 4                            # ERR_FORMAT, as read_block_header (index.c:1360)
    err = C.c_int()
    assert not lib.fm_debug_image_open(os.fsencode(str(tmp_path / "missing")), 0, 1, 1, C.byref(err))
    assert err.value == 2                            # ERR_IO
