"""ctypes bindings for the two CPU checkers -- TEST INFRASTRUCTURE ONLY.

* ``Oracle``    -> oracle/libfm_oracle.so  (our plain-C restatement, fm_oracle.c)
* ``Reference`` -> oracle/_ref/libfemto_ref.so (the unmodified reference + ref_shim.c)

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` /
``--impl reference`` legs may import this module; nothing under femto_b200/ does.
Both classes expose the same methods so tests can diff them call for call.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libfm_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libfemto_ref.so")

CHARACTER_OFFSET = 5  # reference src/main/index_types.h:64
ALPHA_SIZE = 261


def build(ref: bool = True, quiet: bool = True) -> None:
    """Compile the checkers (oracle always; oracle/_ref only where /root/reference exists)."""
    targets = ["oracle"] + (["ref"] if ref else [])
    subprocess.run(["make", "-s", "-C", HERE, "-j8"] + targets, check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def have_reference() -> bool:
    return os.path.exists(REF_SO)


def flatten_patterns(pats: Sequence[np.ndarray]) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """list of uint16 arrays -> (plen int32, flat uint16, offs int64)"""
    plen = np.array([len(p) for p in pats], dtype=np.int32)
    offs = np.zeros(len(pats), dtype=np.int64)
    if len(pats):
        offs[1:] = np.cumsum(plen[:-1], dtype=np.int64)
    flat = (np.concatenate([np.asarray(p, dtype=np.uint16) for p in pats])
            if len(pats) and plen.sum() else np.zeros(1, dtype=np.uint16))
    return plen, np.ascontiguousarray(flat), offs


def bytes_to_alpha(b: bytes) -> np.ndarray:
    return np.frombuffer(b, dtype=np.uint8).astype(np.uint16) + CHARACTER_OFFSET


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(C.POINTER(t))


class _Base:
    prefix = ""
    lib: C.CDLL

    def _fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def __init__(self, path: str):
        raise NotImplementedError

    # ---- common API over the handle ----
    def header_info(self) -> dict:
        out = np.zeros(7, dtype=np.int64)
        rc = self._fn("header_info")(self.h, _p(out, C.c_int64))
        if rc:
            raise RuntimeError(f"header_info rc={rc}")
        keys = ["nblocks", "total_length", "ndocs", "block_size", "bucket_size", "mark_period", "chunk_size"]
        return dict(zip(keys, (int(x) for x in out)))

    def C(self, ch: int) -> int:
        out = C.c_int64()
        rc = self._fn("C")(self.h, ch, C.byref(out))
        if rc:
            raise RuntimeError(f"C rc={rc}")
        return out.value

    def occ(self, ch: int, row: int) -> Tuple[int, int]:
        a, b = C.c_int64(), C.c_int64()
        rc = self._fn("occ")(self.h, ch, C.c_int64(row), C.byref(a), C.byref(b))
        if rc:
            raise RuntimeError(f"occ rc={rc}")
        return a.value, b.value

    def back_step(self, row: int) -> Tuple[int, int, int]:
        ch, nxt, off = C.c_int(), C.c_int64(), C.c_int64()
        rc = self._fn("back_step")(self.h, C.c_int64(row), C.byref(ch), C.byref(nxt), C.byref(off))
        if rc:
            raise RuntimeError(f"back_step rc={rc}")
        return ch.value, nxt.value, off.value

    def count(self, pats: Sequence[np.ndarray]) -> Tuple[np.ndarray, np.ndarray]:
        plen, flat, offs = flatten_patterns(pats)
        return self.count_flat(plen, flat, offs)

    def count_flat(self, plen, flat, offs) -> Tuple[np.ndarray, np.ndarray]:
        n = len(plen)
        first = np.zeros(max(n, 1), dtype=np.int64)
        last = np.zeros(max(n, 1), dtype=np.int64)
        rc = self._fn("count")(self.h, n, _p(plen, C.c_int32), _p(flat, C.c_uint16), _p(offs, C.c_int64),
                               _p(first, C.c_int64), _p(last, C.c_int64))
        if rc:
            raise RuntimeError(f"count rc={rc}")
        return first[:n], last[:n]

    def locate(self, pats: Sequence[np.ndarray], max_occs: int, cap: Optional[int] = None):
        plen, flat, offs = flatten_patterns(pats)
        n = len(plen)
        if cap is None:
            f, l = self.count_flat(plen, flat, offs)
            cap = int(np.maximum(l - f + 1, 0).clip(max=max_occs + 1).sum()) + 1
        noccs = np.zeros(max(n, 1), dtype=np.int32)
        start = np.zeros(max(n, 1), dtype=np.int64)
        out = np.zeros(max(cap, 1), dtype=np.int64)
        rc = self._fn("locate")(self.h, n, _p(plen, C.c_int32), _p(flat, C.c_uint16), _p(offs, C.c_int64),
                                int(max_occs), _p(noccs, C.c_int32), _p(start, C.c_int64),
                                _p(out, C.c_int64), C.c_int64(cap))
        if rc:
            raise RuntimeError(f"locate rc={rc}")
        return [out[start[i]:start[i] + noccs[i]].copy() for i in range(n)]

    def locate_range(self, first: int, last: int) -> np.ndarray:
        out = np.zeros(max(last - first + 1, 1), dtype=np.int64)
        rc = self._fn("locate_range")(self.h, C.c_int64(first), C.c_int64(last), _p(out, C.c_int64))
        if rc:
            raise RuntimeError(f"locate_range rc={rc}")
        return out[:max(last - first + 1, 0)]

    def chunk_documents(self, row: int, cap: int = 1 << 16):
        """(first row, last row, ascending documents) of the chunk holding `row` -- the live reference only
        (block_chunk_request); the plain-C oracle does not restate chunks."""
        a, b, n = C.c_int64(), C.c_int64(), C.c_int64()
        docs = np.zeros(cap, dtype=np.int64)
        rc = self._fn("chunk_documents")(self.h, C.c_int64(row), C.byref(a), C.byref(b), _p(docs, C.c_int64),
                                         C.c_int64(cap), C.byref(n))
        if rc:
            raise RuntimeError(f"chunk_documents rc={rc}")
        return a.value, b.value, docs[:n.value]

    def doc_info(self, doc: int) -> Tuple[int, int]:
        a, b = C.c_int64(), C.c_int64()
        rc = self._fn("doc_info")(self.h, C.c_int64(doc), C.byref(a), C.byref(b))
        if rc:
            raise RuntimeError(f"doc_info rc={rc}")
        return a.value, b.value

    def resolve(self, offset: int) -> Tuple[int, int]:
        a, b = C.c_int64(), C.c_int64()
        rc = self._fn("resolve")(self.h, C.c_int64(offset), C.byref(a), C.byref(b))
        if rc:
            raise RuntimeError(f"resolve rc={rc}")
        return a.value, b.value

    def close(self):
        if getattr(self, "h", None):
            self._fn("close")(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Oracle(_Base):
    prefix = "fmo_"
    _lib = None

    @classmethod
    def load(cls):
        if cls._lib is None:
            if not os.path.exists(ORACLE_SO):
                build(ref=False)
            lib = C.CDLL(ORACLE_SO)
            lib.fmo_open.restype = C.c_void_p
            lib.fmo_open.argtypes = [C.c_char_p, C.POINTER(C.c_int)]
            for name in ("close", "header_info", "C", "occ", "back_step", "count", "locate", "locate_range",
                         "doc_info", "doc_name", "resolve", "extract", "counters", "reset_counters"):
                getattr(lib, "fmo_" + name).argtypes = None
            lib.fmo_close.argtypes = [C.c_void_p]
            lib.fmo_close.restype = None
            cls._lib = lib
        return cls._lib

    def __init__(self, path: str):
        self.lib = self.load()
        err = C.c_int()
        self.h = self.lib.fmo_open(path.encode(), C.byref(err))
        if not self.h:
            raise RuntimeError(f"fmo_open({path}) failed: err={err.value}")
        self.h = C.c_void_p(self.h)

    def extract(self, doc: int) -> np.ndarray:
        ln, _ = self.doc_info(doc)
        out = np.zeros(max(ln, 1), dtype=np.uint16)
        n = C.c_int64()
        rc = self.lib.fmo_extract(self.h, C.c_int64(doc), _p(out, C.c_uint16), C.c_int64(ln), C.byref(n))
        if rc:
            raise RuntimeError(f"extract rc={rc}")
        return out[:n.value]

    def counters(self) -> dict:
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        self.lib.fmo_counters(self.h, C.byref(a), C.byref(b), C.byref(c))
        return {"bytes": a.value, "occ_calls": b.value, "levels": c.value}

    def reset_counters(self):
        self.lib.fmo_reset_counters(self.h)

    @classmethod
    def bseq_rank(cls, zdata: bytes, index1: int) -> Tuple[int, int, int]:
        lib = cls.load()
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        lib.fmo_bseq_rank(zdata, index1, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value


class Reference(_Base):
    prefix = "ref_"
    _lib = None

    @classmethod
    def load(cls):
        if cls._lib is None:
            if not os.path.exists(REF_SO):
                raise RuntimeError("oracle/_ref/libfemto_ref.so is missing (run `make -C oracle ref` "
                                   "where /root/reference is mounted)")
            lib = C.CDLL(REF_SO)
            lib.ref_open.restype = C.c_void_p
            lib.ref_open.argtypes = [C.c_char_p]
            lib.ref_close.argtypes = [C.c_void_p]
            lib.ref_close.restype = None
            lib.ref_free.argtypes = [C.c_void_p]
            lib.ref_free.restype = None
            cls._lib = lib
        return cls._lib

    def __init__(self, path: str):
        self.lib = self.load()
        self.h = self.lib.ref_open(path.encode())
        if not self.h:
            raise RuntimeError(f"ref_open({path}) failed")
        self.h = C.c_void_p(self.h)

    def generic_request(self, index_path: str, request: str) -> str:
        """femto_create_generic_request .. femto_response_for_generic_request (femto.h:75-149)."""
        resp = C.c_char_p()
        self.lib.ref_generic_request.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_char_p)]
        rc = self.lib.ref_generic_request(self.h, index_path.encode(), request.encode(), C.byref(resp))
        if rc:
            raise RuntimeError(f"generic request failed rc={rc}")
        try:
            return resp.value.decode()
        finally:
            self.lib.ref_free(C.cast(resp, C.c_void_p))

    @classmethod
    def build_index(cls, docs: Sequence[bytes], index_path: str, scratch_dir: str, block_size: int = 0,
                    bucket_size: int = 0, chunk_size: int = -1, mark_period: int = -1) -> None:
        """Build an index with the reference's own in-memory builder (qsufsort + index_documents)."""
        lib = cls.load()
        os.makedirs(index_path, exist_ok=True)
        n = len(docs)
        lens = (C.c_int64 * n)(*[len(d) for d in docs])
        bufs = [C.create_string_buffer(d, max(len(d), 1)) for d in docs]
        ptrs = (C.c_char_p * n)(*[C.cast(b, C.c_char_p) for b in bufs])
        rc = lib.ref_build_index(n, lens, ptrs, index_path.encode(), scratch_dir.encode(),
                                 int(block_size), int(bucket_size), int(chunk_size), int(mark_period))
        if rc:
            raise RuntimeError(f"ref_build_index rc={rc}")

    @classmethod
    def bseq_construct(cls, bits: np.ndarray, force_type: int = 0) -> bytes:
        """bits: array of 0/1; returns the reference's encoded bseq (wtree.c:364)."""
        lib = cls.load()
        packed = np.packbits(np.asarray(bits, dtype=np.uint8))  # MSB-first, as the reference reads it
        z = C.c_void_p()
        zlen = C.c_int()
        rc = lib.ref_bseq_construct(int(len(bits)), packed.ctypes.data_as(C.c_char_p), int(force_type),
                                    C.byref(z), C.byref(zlen))
        if rc:
            raise RuntimeError(f"ref_bseq_construct rc={rc}")
        out = C.string_at(z, zlen.value)
        lib.ref_free(z)
        return out

    @classmethod
    def bseq_rank(cls, zdata: bytes, index1: int) -> Tuple[int, int, int]:
        lib = cls.load()
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        lib.ref_bseq_rank(zdata, index1, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value
