"""The chunk plan of the streamed host-buffer count (femto_b200/csrc/fm_stream_plan.hpp), checked on
CPU through the debug export fm_debug_stream_plan (not part of the public header).  The properties
are what keeps a kernel that runs AHEAD of its input copies correct: every pattern is delivered
exactly once and in order, and no 128-byte line of plen / offs / symbols is shared by two chunks."""
import ctypes as C

import numpy as np
import pytest

from femto_b200 import _lib


def plan(plen, offs, flat_len):
    lib = _lib.load()
    fn = lib.fm_debug_stream_plan
    fn.restype = C.c_int64
    fn.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
    out = np.zeros((256, 5), dtype=np.int64)
    k = fn(len(plen), plen.ctypes.data, offs.ctypes.data, flat_len, out.ctypes.data, 256)
    return k, out[:max(k, 0)]


def batch(n, lengths, seed):
    rng = np.random.default_rng(seed)
    plen = rng.choice(np.array(lengths, dtype=np.int32), n).astype(np.int32)
    offs = np.zeros(n, dtype=np.int64)
    offs[1:] = np.cumsum(plen[:-1], dtype=np.int64)
    return plen, offs, int(plen.sum())


def test_small_batches_are_not_streamed():
    plen, offs, flat_len = batch(131071, [32], 1)
    assert plan(plen, offs, flat_len)[0] == -1


@pytest.mark.parametrize("n,lengths", [(131072, [32]), (1 << 20, [32]), (300077, [1, 2, 3, 5, 8, 13, 21, 34]),
                                        (200003, [0, 0, 7, 255]), (999999, [31])])
def test_plan_tiles_the_batch_on_line_boundaries(n, lengths):
    plen, offs, flat_len = batch(n, lengths, 7)
    k, p = plan(plen, offs, flat_len)
    assert k > 2
    mid = (n - n // 4) & ~31
    # two kernels, in order; the second takes the last quarter
    assert list(p[:, 0]) == sorted(p[:, 0]) and set(p[:, 0]) == {0, 1}
    assert p[p[:, 0] == 0][-1, 2] == mid == p[p[:, 0] == 1][0, 1]
    # patterns: a tiling of [0, n), every inner boundary a multiple of 32 (128 bytes of plen)
    assert p[0, 1] == 0 and p[-1, 2] == n and (p[1:, 1] == p[:-1, 2]).all()
    assert (p[1:, 1] % 32 == 0).all()
    # chunk sizes grow from 8 Ki to 128 Ki patterns
    sizes = p[:, 2] - p[:, 1]
    assert sizes[0] == 8192 and sizes.max() <= 131072
    # symbols: a tiling of [0, flat_len), inner cuts on 128-byte lines, never before the end of the
    # chunk's last pattern (so a chunk's mark is written only after all of its symbols)
    assert p[0, 3] == 0 and p[-1, 4] == flat_len and (p[1:, 3] == p[:-1, 4]).all()
    assert (p[:-1, 4] % 64 == 0).all() or flat_len in p[:-1, 4]
    ends = offs[p[:, 2] - 1] + plen[p[:, 2] - 1]
    assert (p[:, 4] >= ends).all()
    # ... and less than one line beyond it
    assert (p[:-1, 4] - ends[:-1] < 64).all()


def plan_bytes(plen, offs, flat_len):
    lib = _lib.load()
    fn = lib.fm_debug_stream_plan2
    fn.restype = C.c_int64
    fn.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64]
    out = np.zeros((256, 5), dtype=np.int64)
    k = fn(len(plen), plen.ctypes.data, offs.ctypes.data, flat_len, 1, out.ctypes.data, 256)
    return k, out[:max(k, 0)]


@pytest.mark.parametrize("n,lengths", [(1 << 20, [32]), (300077, [1, 2, 3, 5, 8, 13, 21, 34]), (200003, [0, 0, 7, 255])])
def test_plan_for_byte_patterns_cuts_on_128_symbol_lines(n, lengths):
    """fm_count_bytes: one byte per symbol, so a 128-byte line holds 128 symbols."""
    plen, offs, flat_len = batch(n, lengths, 9)
    k, p = plan_bytes(plen, offs, flat_len)
    assert k > 2
    assert p[0, 3] == 0 and p[-1, 4] == flat_len and (p[1:, 3] == p[:-1, 4]).all()
    assert (p[:-1, 4] % 128 == 0).all() or flat_len in p[:-1, 4]
    ends = offs[p[:, 2] - 1] + plen[p[:, 2] - 1]
    assert (p[:, 4] >= ends).all() and (p[:-1, 4] - ends[:-1] < 128).all()
