"""femto_b200 -- B200-native FM-index query engine behind femto's count/locate/extract API.

This package is a thin ctypes mirror of the C ABI in ``include/femto_b200.h`` (the product is the
shared library ``libfemto_b200.so`` built from ``femto_b200/csrc``).  Names follow the reference:
``Index.count`` == ``parallel_count`` (reference src/main/femto.c:275), ``Index.locate`` ==
``parallel_locate`` (:331), ``Index.locate_range`` == ``parallel_locate_range`` (:481),
``Index.extract`` == the extract-document query (src/main/server.c:6364).

Symbols are ``alpha_t`` values: ``5 + byte`` for text bytes (src/main/index_types.h:64-68).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import FmInfo

CHARACTER_OFFSET = 5
FM_ERR_FULL = 10          # include/femto_b200.h fm_err_t (ERR_FULL, utils/error.h)
ALPHA_SIZE = 261
ESCAPE_CODE_SEOF = 2

ERR_NAMES = {0: "OK", 1: "MEM", 2: "IO", 3: "PARAM", 4: "FORMAT", 5: "BZ_DATA", 6: "INVALID", 7: "PTHREADS",
             8: "MISSING", 9: "CANCELED", 10: "FULL", 11: "OVERWORKED", 12: "UNKNOWN"}


class FemtoError(RuntimeError):
    def __init__(self, code: int, where: str, msg: str = ""):
        self.code = code
        super().__init__(f"{where}: ERR_{ERR_NAMES.get(code, code)} {msg}".strip())


def _check(rc: int, where: str) -> None:
    if rc:
        msg = _lib.load().fm_last_error()
        raise FemtoError(rc, where, msg.decode(errors="replace") if msg else "")


def _ptr(a: np.ndarray, t):
    return a.ctypes.data_as(C.POINTER(t))


def bytes_to_alpha(b: bytes) -> np.ndarray:
    """strtoalpha (src/main/index_types.h:85-97)."""
    return np.frombuffer(b, dtype=np.uint8).astype(np.uint16) + CHARACTER_OFFSET


def flatten_patterns(pats: Sequence[np.ndarray]) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    plen = np.array([len(p) for p in pats], dtype=np.int32)
    offs = np.zeros(len(pats), dtype=np.int64)
    if len(pats) > 1:
        offs[1:] = np.cumsum(plen[:-1], dtype=np.int64)
    total = int(plen.sum()) if len(pats) else 0
    flat = (np.concatenate([np.asarray(p, dtype=np.uint16) for p in pats]) if total
            else np.zeros(1, dtype=np.uint16))
    return plen, np.ascontiguousarray(flat), offs


class Index:
    """An index resident in one GPU's HBM (``fm_open`` ... ``fm_close``)."""

    def __init__(self, path: str, device: int = 0, shard: int = 0, nshards: int = 1):
        self.lib = _lib.load()
        h = C.c_void_p()
        if nshards == 1:
            rc = self.lib.fm_open(os.fsencode(path), device, C.byref(h))
        else:
            rc = self.lib.fm_open_shard(os.fsencode(path), device, shard, nshards, C.byref(h))
        _check(rc, "fm_open")
        self.h = h
        info = FmInfo()
        _check(self.lib.fm_info(self.h, C.byref(info)), "fm_info")
        self.info = info

    # -- lifecycle -------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "h", None):
            self.lib.fm_close(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def total_length(self) -> int:
        return int(self.info.total_length)

    @property
    def num_documents(self) -> int:
        return int(self.info.num_documents)

    def set_lanes_per_query(self, lanes: int) -> None:
        _check(self.lib.fm_set_lanes_per_query(self.h, lanes), "fm_set_lanes_per_query")

    def set_count_schedule(self, merged: bool, lanes: int) -> None:
        _check(self.lib.fm_set_count_schedule(self.h, int(merged), lanes), "fm_set_count_schedule")

    def kernel_launches(self) -> int:
        return int(self.lib.fm_kernel_launches(self.h))

    # -- count -----------------------------------------------------------------------------
    def count_flat(self, plen: np.ndarray, flat: np.ndarray, offs: np.ndarray,
                   first: Optional[np.ndarray] = None, last: Optional[np.ndarray] = None):
        n = len(plen)
        if first is None:
            first = np.empty(max(n, 1), dtype=np.int64)
        if last is None:
            last = np.empty(max(n, 1), dtype=np.int64)
        rc = self.lib.fm_count_flat(self.h, n, _ptr(plen, C.c_int32), _ptr(flat, C.c_uint16), _ptr(offs, C.c_int64),
                                    _ptr(first, C.c_int64), _ptr(last, C.c_int64))
        _check(rc, "fm_count_flat")
        return first[:n], last[:n]

    def count_bytes(self, plen: np.ndarray, text: np.ndarray, offs: np.ndarray,
                    first: Optional[np.ndarray] = None, last: Optional[np.ndarray] = None):
        """fm_count_bytes: patterns as raw text bytes (uint8), symbols = CHARACTER_OFFSET + byte."""
        n = len(plen)
        if first is None:
            first = np.empty(max(n, 1), dtype=np.int64)
        if last is None:
            last = np.empty(max(n, 1), dtype=np.int64)
        rc = self.lib.fm_count_bytes(self.h, n, _ptr(plen, C.c_int32), _ptr(text, C.c_uint8), _ptr(offs, C.c_int64),
                                     _ptr(first, C.c_int64), _ptr(last, C.c_int64))
        _check(rc, "fm_count_bytes")
        return first[:n], last[:n]

    def count(self, pats: Sequence[np.ndarray]):
        """[first,last] BWT row range per pattern, through the reference-shaped pointer-array call."""
        n = len(pats)
        arrs = [np.ascontiguousarray(p, dtype=np.uint16) for p in pats]
        plen = (C.c_int * max(n, 1))(*[len(a) for a in arrs])
        ptrs = (C.POINTER(C.c_uint16) * max(n, 1))(*[_ptr(a, C.c_uint16) for a in arrs])
        first = np.empty(max(n, 1), dtype=np.int64)
        last = np.empty(max(n, 1), dtype=np.int64)
        _check(self.lib.fm_count(self.h, n, plen, ptrs, _ptr(first, C.c_int64), _ptr(last, C.c_int64)), "fm_count")
        return first[:n], last[:n]

    def count_stats(self, plen: np.ndarray, flat: np.ndarray, offs: np.ndarray) -> dict:
        """Counters of one instrumented count launch (see fm_count_stats)."""
        st = (C.c_uint64 * 8)()
        _check(self.lib.fm_count_stats(self.h, len(plen), _ptr(plen, C.c_int32), _ptr(flat, C.c_uint16),
                                       _ptr(offs, C.c_int64), st), "fm_count_stats")
        return {"block_reads": int(st[0]), "distinct_block_reads": int(st[1]), "occ_evals": int(st[2]),
                "steps": int(st[3]), "group_slots": int(st[4]), "groups_active": int(st[5])}

    def walk_stats(self, rows: np.ndarray) -> dict:
        """Counters of one instrumented locate walk launch over `rows` (see fm_walk_stats)."""
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        st = (C.c_uint64 * 4)()
        _check(self.lib.fm_walk_stats(self.h, len(rows), _ptr(rows, C.c_int64), st), "fm_walk_stats")
        return {"lf_steps": int(st[0]), "wtree_blocks": int(st[1]), "mark_blocks": int(st[2]), "sa_samples": int(st[3])}

    def probe_random_reads(self, bytes_per_access: int, steps: int = 400) -> dict:
        """Dependent random reads over the resident rank blocks (fm_probe_random_reads): the
        access-rate ceiling of this GPU's memory system for the count kernel's access shape."""
        acc, ms = C.c_int64(), C.c_double()
        _check(self.lib.fm_probe_random_reads(self.h, bytes_per_access, steps, C.byref(acc), C.byref(ms)),
               "fm_probe_random_reads")
        return {"accesses": int(acc.value), "ms": float(ms.value),
                "accesses_per_s": acc.value / (ms.value / 1e3) if ms.value > 0 else 0.0,
                "gb_per_s": acc.value * bytes_per_access / (ms.value / 1e3) / 1e9 if ms.value > 0 else 0.0}

    def count_device(self, npats: int, d_plen: int, d_flat: int, d_offs: int, d_first: int, d_last: int,
                     stream: int = 0) -> None:
        """Device-pointer form (integers are raw device addresses, e.g. ``tensor.data_ptr()``)."""
        _check(self.lib.fm_count_device(self.h, npats, d_plen, d_flat, d_offs, d_first, d_last, stream),
               "fm_count_device")

    # -- locate ----------------------------------------------------------------------------
    def locate(self, pats: Sequence[np.ndarray], max_occs: int) -> List[np.ndarray]:
        n = len(pats)
        arrs = [np.ascontiguousarray(p, dtype=np.uint16) for p in pats]
        plen = (C.c_int * max(n, 1))(*[len(a) for a in arrs])
        ptrs = (C.POINTER(C.c_uint16) * max(n, 1))(*[_ptr(a, C.c_uint16) for a in arrs])
        noccs = (C.c_int * max(n, 1))()
        outs = (C.POINTER(C.c_int64) * max(n, 1))()
        _check(self.lib.fm_locate(self.h, n, plen, ptrs, int(max_occs), noccs, outs), "fm_locate")
        res = []
        libc = C.CDLL(None)
        libc.free.argtypes = [C.c_void_p]
        for i in range(n):
            k = noccs[i]
            if k > 0:
                res.append(np.ctypeslib.as_array(outs[i], shape=(k,)).copy())
                libc.free(C.cast(outs[i], C.c_void_p))
            else:
                res.append(np.zeros(0, dtype=np.int64))
        return res

    def locate_flat(self, plen, flat, offs, max_occs: int, cap: int):
        n = len(plen)
        noccs = np.zeros(max(n, 1), dtype=np.int32)
        start = np.zeros(max(n, 1), dtype=np.int64)
        out = np.zeros(max(cap, 1), dtype=np.int64)
        rc = self.lib.fm_locate_flat(self.h, n, _ptr(plen, C.c_int32), _ptr(flat, C.c_uint16), _ptr(offs, C.c_int64),
                                     int(max_occs), _ptr(noccs, C.c_int32), _ptr(start, C.c_int64),
                                     _ptr(out, C.c_int64), cap)
        _check(rc, "fm_locate_flat")
        return noccs[:n], start[:n], out

    def locate_range(self, first: int, last: int) -> np.ndarray:
        n = max(last - first + 1, 0)
        out = np.zeros(max(n, 1), dtype=np.int64)
        _check(self.lib.fm_locate_range(self.h, first, last, _ptr(out, C.c_int64)), "fm_locate_range")
        return out[:n]

    def locate_rows(self, rows: np.ndarray) -> np.ndarray:
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        out = np.zeros(max(len(rows), 1), dtype=np.int64)
        _check(self.lib.fm_locate_rows(self.h, len(rows), _ptr(rows, C.c_int64), _ptr(out, C.c_int64)),
               "fm_locate_rows")
        return out[:len(rows)]

    # -- leaf interface --------------------------------------------------------------------
    def back_step(self, rows: np.ndarray):
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        n = len(rows)
        ch = np.zeros(max(n, 1), dtype=np.int32)
        nxt = np.zeros(max(n, 1), dtype=np.int64)
        off = np.zeros(max(n, 1), dtype=np.int64)
        _check(self.lib.fm_back_step(self.h, n, _ptr(rows, C.c_int64), _ptr(ch, C.c_int32), _ptr(nxt, C.c_int64),
                                     _ptr(off, C.c_int64)), "fm_back_step")
        return ch[:n], nxt[:n], off[:n]

    def occ(self, ch: np.ndarray, rows: np.ndarray) -> np.ndarray:
        ch = np.ascontiguousarray(ch, dtype=np.uint16)
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        out = np.zeros(max(len(rows), 1), dtype=np.int64)
        _check(self.lib.fm_occ(self.h, len(rows), _ptr(ch, C.c_uint16), _ptr(rows, C.c_int64), _ptr(out, C.c_int64)),
               "fm_occ")
        return out[:len(rows)]

    def backward_step(self, first: np.ndarray, last: np.ndarray, ch: np.ndarray):
        """One backward-search step per (range, symbol): backward_search_query (server.c:948)."""
        first = np.ascontiguousarray(first, dtype=np.int64)
        last = np.ascontiguousarray(last, dtype=np.int64)
        ch = np.ascontiguousarray(ch, dtype=np.uint16)
        n = len(first)
        nf = np.zeros(max(n, 1), dtype=np.int64)
        nl = np.zeros(max(n, 1), dtype=np.int64)
        _check(self.lib.fm_backward_step(self.h, n, _ptr(first, C.c_int64), _ptr(last, C.c_int64),
                                         _ptr(ch, C.c_uint16), _ptr(nf, C.c_int64), _ptr(nl, C.c_int64)),
               "fm_backward_step")
        return nf[:n], nl[:n]

    # -- documents -------------------------------------------------------------------------
    def doc_info(self, doc: int) -> Tuple[int, int]:
        a, b = C.c_int64(), C.c_int64()
        _check(self.lib.fm_doc_info(self.h, doc, C.byref(a), C.byref(b)), "fm_doc_info")
        return a.value, b.value

    def resolve(self, offsets: np.ndarray):
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets)
        doc = np.zeros(max(n, 1), dtype=np.int64)
        off = np.zeros(max(n, 1), dtype=np.int64)
        _check(self.lib.fm_resolve(self.h, n, _ptr(offsets, C.c_int64), _ptr(doc, C.c_int64), _ptr(off, C.c_int64)),
               "fm_resolve")
        return doc[:n], off[:n]

    def doc_name(self, doc: int) -> bytes:
        """The info bytes stored with the document at build time (document_info, index.c:1767-1784)."""
        n = C.c_int64()
        rc = self.lib.fm_doc_name(self.h, doc, None, 0, C.byref(n))
        if rc not in (0, FM_ERR_FULL):
            _check(rc, "fm_doc_name")
        buf = C.create_string_buffer(max(n.value, 1))
        _check(self.lib.fm_doc_name(self.h, doc, buf, n.value, C.byref(n)), "fm_doc_name")
        return buf.raw[:n.value]

    def range_documents(self, first: int, last: int) -> np.ndarray:
        """Ascending document numbers holding the suffixes of rows first..last (range_to_results, documents)."""
        cap = max(min(last - first + 1, int(self.info.num_documents)), 1)
        docs = np.zeros(cap, dtype=np.int64)
        n = C.c_int64()
        _check(self.lib.fm_range_documents(self.h, first, last, _ptr(docs, C.c_int64), cap, C.byref(n)),
               "fm_range_documents")
        return docs[:n.value]

    def chunk_documents(self, row: int):
        """(first row, last row, ascending documents) of the document chunk holding `row` (block_chunk_request)."""
        a, b, n = C.c_int64(), C.c_int64(), C.c_int64()
        cap = max(int(self.info.num_documents), 1)
        docs = np.zeros(cap, dtype=np.int64)
        _check(self.lib.fm_chunk_documents(self.h, row, C.byref(a), C.byref(b), _ptr(docs, C.c_int64), cap, C.byref(n)),
               "fm_chunk_documents")
        return a.value, b.value, docs[:n.value]

    def extract(self, doc: int) -> np.ndarray:
        ln, _ = self.doc_info(doc)
        out = np.zeros(max(ln, 1), dtype=np.uint16)
        n = C.c_int64()
        _check(self.lib.fm_extract(self.h, doc, _ptr(out, C.c_uint16), ln, C.byref(n)), "fm_extract")
        return out[:n.value]


    def generic_request(self, request: str) -> str:
        """femto's generic request interface (femto.h:75-149) for the string_rows* requests."""
        resp = C.c_char_p()
        _check(self.lib.fm_generic_request(self.h, request.encode(), C.byref(resp)), "fm_generic_request")
        try:
            return resp.value.decode()
        finally:
            libc = C.CDLL(None)
            libc.free.argtypes = [C.c_void_p]
            libc.free(C.cast(resp, C.c_void_p))

    def extract_batch(self, docs: Sequence[int]) -> List[np.ndarray]:
        """Several documents in one launch (fm_extract_batch)."""
        d = np.ascontiguousarray(docs, dtype=np.int64)
        start = np.zeros(len(d) + 1, dtype=np.int64)
        total = sum(self.doc_info(int(x))[0] - 1 for x in d)
        out = np.zeros(max(total, 1), dtype=np.uint16)
        _check(self.lib.fm_extract_batch(self.h, len(d), _ptr(d, C.c_int64), _ptr(out, C.c_uint16), total,
                                         _ptr(start, C.c_int64)), "fm_extract_batch")
        return [out[start[k]:start[k + 1]] for k in range(len(d))]


# ---------------------------------------------------------------------------------------------
# index construction (host side)

def prepare_text(docs: Sequence[bytes]) -> Tuple[np.ndarray, np.ndarray]:
    """The reference's prepared text for documents without headers: each document's bytes as
    5+byte followed by one SEOF symbol (append_file_mem, src/main/bwt_prepare.c:231-311).
    Returns (text uint16, doc_ends int64)."""
    parts = []
    ends = []
    total = 0
    for d in docs:
        a = np.empty(len(d) + 1, dtype=np.uint16)
        a[:-1] = np.frombuffer(d, dtype=np.uint8).astype(np.uint16) + CHARACTER_OFFSET
        a[-1] = ESCAPE_CODE_SEOF
        parts.append(a)
        total += len(a)
        ends.append(total)
    return np.concatenate(parts), np.array(ends, dtype=np.int64)


def suffix_sort_host(text: np.ndarray) -> np.ndarray:
    lib = _lib.load()
    text = np.ascontiguousarray(text, dtype=np.uint16)
    sa = np.empty(len(text), dtype=np.int64)
    _check(lib.fm_suffix_sort_host(_ptr(text, C.c_uint16), len(text), _ptr(sa, C.c_int64)), "fm_suffix_sort_host")
    return sa


def bwt_from_sa(text: np.ndarray, sa: np.ndarray) -> np.ndarray:
    """L[r] = text[sa[r]-1], SEOF for sa[r]==0 (get_L_char_from_offsets, src/main/bwt_qsufsort.c:63-84)."""
    L = text[sa - 1].copy()  # sa==0 wraps to the last symbol, which is the final document's SEOF
    return L


class IndexBuilder:
    """Streams BWT rows into femto's on-disk format (``fm_builder_*``).

    With ``first_block`` / ``range_blocks`` it is a RANGE builder (``fm_builder_create_range``): it writes
    only those data blocks, is fed their rows, and ends with ``finish_range()``; ``write_index_header``
    completes the index once all ranges are written."""

    def __init__(self, out_dir: str, doc_ends: np.ndarray, block_size: int = 128 << 20, bucket_size: int = 1 << 20,
                 chunk_size: int = 2048, mark_period: int = 20, nthreads: int = 0,
                 first_block: Optional[int] = None, range_blocks: int = 0):
        self.lib = _lib.load()
        doc_ends = np.ascontiguousarray(doc_ends, dtype=np.int64)
        b = C.c_void_p()
        self.ndocs = len(doc_ends)
        self.range_blocks = range_blocks if first_block is not None else None
        if first_block is None:
            _check(self.lib.fm_builder_create(os.fsencode(out_dir), int(doc_ends[-1]), len(doc_ends),
                                              _ptr(doc_ends, C.c_int64), block_size, bucket_size, chunk_size,
                                              mark_period, nthreads, C.byref(b)), "fm_builder_create")
        else:
            _check(self.lib.fm_builder_create_range(os.fsencode(out_dir), int(doc_ends[-1]), len(doc_ends),
                                                    _ptr(doc_ends, C.c_int64), block_size, bucket_size, chunk_size,
                                                    mark_period, nthreads, first_block, range_blocks, C.byref(b)),
                   "fm_builder_create_range")
        self.b = b

    def finish_range(self) -> Tuple[np.ndarray, np.ndarray]:
        """-> (block_counts [range_blocks, 261], eof_rows [ndocs], -1 where the row is not in this range)."""
        assert self.range_blocks is not None, "not a range builder"
        counts = np.zeros((self.range_blocks, ALPHA_SIZE), dtype=np.int64)
        eof = np.full(self.ndocs, -1, dtype=np.int64)
        b, self.b = self.b, None
        _check(self.lib.fm_builder_finish_range(b, _ptr(counts, C.c_int64), _ptr(eof, C.c_int64)),
               "fm_builder_finish_range")
        return counts, eof

    def set_doc_info(self, doc: int, info: bytes) -> None:
        _check(self.lib.fm_builder_set_doc_info(self.b, doc, info, len(info)), "fm_builder_set_doc_info")

    def append(self, L: np.ndarray, sa: np.ndarray) -> None:
        L = np.ascontiguousarray(L, dtype=np.uint16)
        sa = np.ascontiguousarray(sa, dtype=np.int64)
        assert len(L) == len(sa)
        _check(self.lib.fm_builder_append(self.b, len(L), _ptr(L, C.c_uint16), _ptr(sa, C.c_int64)),
               "fm_builder_append")

    def finish(self) -> None:
        b, self.b = self.b, None
        _check(self.lib.fm_builder_finish(b), "fm_builder_finish")

    def abort(self) -> None:
        if self.b:
            self.lib.fm_builder_abort(self.b)
            self.b = None

    def __del__(self):
        try:
            self.abort()
        except Exception:
            pass


def write_index_header(out_dir: str, doc_ends: np.ndarray, block_counts: np.ndarray, eof_rows: np.ndarray,
                       block_size: int = 128 << 20, bucket_size: int = 1 << 20, chunk_size: int = 2048,
                       mark_period: int = 20, doc_infos: Optional[Sequence[Optional[bytes]]] = None) -> None:
    """Header block of an index whose data blocks were written by range builders (``fm_builder_write_header``):
    block_counts [nblocks, 261] in block order, eof_rows [ndocs] merged over the ranges."""
    lib = _lib.load()
    doc_ends = np.ascontiguousarray(doc_ends, dtype=np.int64)
    block_counts = np.ascontiguousarray(block_counts, dtype=np.int64)
    eof_rows = np.ascontiguousarray(eof_rows, dtype=np.int64)
    total = int(doc_ends[-1])
    nblocks = (total + block_size - 1) // block_size
    if block_counts.shape != (nblocks, ALPHA_SIZE) or eof_rows.shape != (len(doc_ends),):
        raise ValueError("write_index_header: block_counts must be [nblocks, 261] and eof_rows [ndocs]")
    info_p = info_n = None
    keep = []
    if doc_infos is not None:
        info_p = (C.c_void_p * len(doc_ends))()
        info_n = (C.c_int64 * len(doc_ends))()
        for i, info in enumerate(doc_infos):
            if info is None:
                info_p[i], info_n[i] = None, -1
            else:
                buf = C.create_string_buffer(bytes(info), len(info))
                keep.append(buf)
                info_p[i], info_n[i] = C.cast(buf, C.c_void_p), len(info)
    _check(lib.fm_builder_write_header(os.fsencode(out_dir), total, len(doc_ends), _ptr(doc_ends, C.c_int64),
                                       block_size, bucket_size, chunk_size, mark_period,
                                       _ptr(block_counts, C.c_int64), _ptr(eof_rows, C.c_int64), info_p, info_n),
           "fm_builder_write_header")


def build_index_host(docs: Sequence[bytes], out_dir: str, doc_infos: Optional[Sequence[bytes]] = None,
                     **params) -> None:
    """Small-corpus builder: host suffix sort + the format-identical emitter."""
    text, ends = prepare_text(docs)
    sa = suffix_sort_host(text)
    L = bwt_from_sa(text, sa)
    b = IndexBuilder(out_dir, ends, **params)
    if doc_infos is not None:
        for i, info in enumerate(doc_infos):
            b.set_doc_info(i, info)
    b.append(L, sa)
    b.finish()


def flatten(index_dir: str, out_file: str) -> None:
    _check(_lib.load().fm_flatten(os.fsencode(index_dir), os.fsencode(out_file)), "fm_flatten")
