"""Index build on several GPUs: every rank writes the data blocks it will serve ("next" row f-1 of
SURVEY.md section 8, for corpora larger than one GPU -- BASELINE configs[4]).

The reference's partition unit is the data block (src/main/index.h:83-100): a block is a
self-contained file holding block_size consecutive BWT rows, and the header block only needs every
block's symbol counts and the rows of the document ends (constructor_construct_header,
src/main/construct.c:407-460).  So the build splits BY BWT ROW RANGE, with the same block -> rank map
the query side uses (``shard_of_block``): rank r produces rows [first_block * block_size, ...) of its
own blocks and nothing else.

What makes that cheap on B200s: the TEXT is replicated (one byte per symbol, 128 GiB of corpus in
180 GB of HBM), so a rank can sort any set of suffixes with local gathers -- there is no exchange
step in the sort at all.  The ranks agree on the row range of every first-symbol (and, where
needed, longer-prefix) bucket from histograms each computes by itself; a rank sorts the buckets
that overlap its rows (those at its borders, and any bucket larger than a batch, are split by the next
symbol first, so that what it sorts beyond its own rows stays small) and slices.  The only communication is one gather of
block_counts / eof_rows (KBs) before rank 0 writes the header.

Text beyond ~150 GiB would not fit replicated; it would be sharded and the key gathers of
``_sort_batch`` would become all-to-all exchanges of (position, key) pairs.  Not built.

Tested on CPU tensors with gloo, world size 2 and 3, byte for byte against the one-process builder
(tests/test_build_dist.py); on GPUs it has run once, 2 ranks on a 64 MiB corpus
(profiles/r02_bench_sharded_mode_small_n2.json) -- not at the sizes it is meant for.
"""
from __future__ import annotations

import time
from typing import Callable, Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ALPHA_SIZE, CHARACTER_OFFSET, ESCAPE_CODE_SEOF, IndexBuilder, write_index_header
from .build_gpu import PAD, SYM_BITS, SYMS_PER_KEY, _sort_batch
from .sharded import shard_of_block


class ByteText:
    """Prepared text kept as ONE BYTE per position plus the sorted positions of the document ends.

    symbol(p) = SEOF where p is a document's last position, else byte[p] + CHARACTER_OFFSET
    (src/main/bwt_prepare.c:231-311).  Half the memory of the int16 text of ``prepare_text_gpu``;
    ``build_gpu._pack_keys`` and the functions below accept either."""

    def __init__(self, data: torch.Tensor, doc_ends: np.ndarray):
        assert data.dtype == torch.uint8
        self.n = int(doc_ends[-1])
        assert data.numel() >= self.n + PAD, "text must carry PAD bytes behind its end"
        self.data = data
        self.device = data.device
        self.doc_ends = np.asarray(doc_ends, dtype=np.int64)
        self.seof = torch.from_numpy(self.doc_ends - 1).to(data.device)
        lens = np.diff(np.concatenate([[0], self.doc_ends]))
        # with every document (SEOF included) at least SYMS_PER_KEY long, a key holds at most one SEOF
        self.sparse_seof = bool(lens.min() >= SYMS_PER_KEY)

    @classmethod
    def from_docs(cls, docs: Sequence[torch.Tensor]) -> "ByteText":
        n = sum(int(d.numel()) + 1 for d in docs)
        data = torch.zeros(n + PAD, dtype=torch.uint8, device=docs[0].device)
        ends, pos = [], 0
        for d in docs:
            m = int(d.numel())
            data[pos: pos + m] = d
            pos += m + 1
            ends.append(pos)
        return cls(data, np.array(ends, dtype=np.int64))

    def is_seof(self, pos: torch.Tensor) -> torch.Tensor:
        i = torch.searchsorted(self.seof, pos)
        i = torch.clamp(i, max=self.seof.numel() - 1)
        return self.seof[i] == pos

    def symbols(self, pos: torch.Tensor) -> torch.Tensor:
        """int64 symbols at arbitrary positions < n + PAD (0 behind the end of the text)."""
        s = self.data[pos].long() + CHARACTER_OFFSET
        s = torch.where(pos < self.n, s, torch.zeros_like(s))
        return torch.where(self.is_seof(pos), torch.full_like(s, ESCAPE_CODE_SEOF), s)

    def slice_symbols(self, lo: int, hi: int) -> torch.Tensor:
        """int16 symbols of positions [lo, hi)."""
        s = self.data[lo:hi].to(torch.int16) + CHARACTER_OFFSET
        if hi > self.n:
            s[max(0, self.n - lo):] = 0
        a, b = np.searchsorted(self.doc_ends - 1, [lo, hi])
        if b > a:
            s[self.seof[a:b] - lo] = ESCAPE_CODE_SEOF
        return s

    def pack_keys(self, pos: torch.Tensor, depth: int) -> torch.Tensor:
        """SYMS_PER_KEY symbols from pos + depth on, SYM_BITS each, first symbol highest."""
        p0 = pos + depth
        key = torch.zeros_like(pos)
        if self.sparse_seof:
            # one search per key: the first document end at or behind p0
            i = torch.clamp(torch.searchsorted(self.seof, p0), max=self.seof.numel() - 1)
            q = self.seof[i]
            for k in range(SYMS_PER_KEY):
                p = p0 + k
                s = self.data[p].long() + CHARACTER_OFFSET
                s = torch.where(p < self.n, s, torch.zeros_like(s))
                s = torch.where(p == q, torch.full_like(s, ESCAPE_CODE_SEOF), s)
                key = (key << SYM_BITS) | s
            return key
        for k in range(SYMS_PER_KEY):
            key = (key << SYM_BITS) | self.symbols(p0 + k)
        return key


# --------------------------------------------------------------------------------------------
# which suffixes a row range needs

def _pair_histogram(T, n: int, step: int = 1 << 27) -> np.ndarray:
    """h[a, b] = number of positions p < n with symbol a at p and symbol b at p + 1 (0 behind the text)."""
    h = torch.zeros(512 * 512, dtype=torch.int64, device=T.device)
    for s in range(0, n, step):
        e = min(n, s + step)
        t = T.slice_symbols(s, e + 1) if isinstance(T, ByteText) else T[s:e + 1]
        h += torch.bincount(t[: e - s].long() * 512 + t[1:].long(), minlength=512 * 512)
    return h.cpu().numpy().reshape(512, 512)


MAX_PREFIX = 6   # symbols a bucket's prefix may have; a bucket still too large at that depth is sorted whole


def _symbols(T, lo: int, hi: int) -> torch.Tensor:
    return T.slice_symbols(lo, hi) if isinstance(T, ByteText) else T[lo:hi]


def _prefix_mask(t: torch.Tensor, length: int, prefix: Tuple[int, ...]) -> Optional[torch.Tensor]:
    """mask[p] = the symbols t[p .. p + len(prefix)) equal `prefix`, for p < length (None: empty prefix)."""
    m = None
    for j, c in enumerate(prefix):
        mj = t[j: j + length] == c
        m = mj if m is None else m & mj
    return m


def _next_histogram(T, n: int, prefix: Tuple[int, ...], step: int = 1 << 28) -> np.ndarray:
    """Histogram of the symbol that follows `prefix`, over the positions of the text where it occurs."""
    k = len(prefix)
    h = torch.zeros(512, dtype=torch.int64, device=T.device)
    for s in range(0, n, step):
        e = min(n, s + step)
        t = _symbols(T, s, e + k)
        nxt = t[k: k + e - s]
        m = _prefix_mask(t, e - s, prefix)
        h += torch.bincount((nxt if m is None else nxt[m]).long(), minlength=512)
    return h.cpu().numpy()


def _select(T, n: int, prefix: Tuple[int, ...], lo: int, hi: int, step: int = 1 << 28) -> torch.Tensor:
    """Positions where `prefix` occurs followed by a symbol in [lo, hi], ascending."""
    k = len(prefix)
    parts = []
    for s in range(0, n, step):
        e = min(n, s + step)
        t = _symbols(T, s, e + k)
        nxt = t[k: k + e - s]
        m = (nxt >= lo) & (nxt <= hi)
        pm = _prefix_mask(t, e - s, prefix)
        if pm is not None:
            m &= pm
        parts.append(torch.nonzero(m).squeeze(1) + s)
    return torch.cat(parts) if len(parts) > 1 else parts[0]


Job = Tuple[Tuple[int, ...], int, int, int, int]


def plan_jobs(count_next: Callable[[Tuple[int, ...]], np.ndarray], batch: int, row_lo: int, row_hi: int,
              max_prefix: int = MAX_PREFIX) -> List[Job]:
    """Sort jobs that together cover the BWT rows [row_lo, row_hi), in row order.

    A job = (prefix, lo, hi, row_start, count): the suffixes that begin with `prefix` followed by a symbol
    in [lo, hi] occupy the rows [row_start, row_start + count).  count_next(prefix) = histogram of the
    symbol following `prefix` in the text (prefix () = the symbols themselves).  A bucket is split by
    its next symbol while it holds more than `batch` suffixes, or more than batch / 8 and sticks out of the
    row range (so that a rank sorts little beyond its own rows); neighbouring small buckets of one
    prefix are merged into one job of at most `batch`.  Every rank derives the same buckets from the same
    histograms, whatever its row range."""
    jobs: List[Job] = []

    def visit(prefix: Tuple[int, ...], start: int, hist: np.ndarray) -> None:
        run: List[int] = []
        run_start = run_total = 0

        def flush() -> None:
            nonlocal run, run_total
            if run:
                jobs.append((prefix, run[0], run[-1], run_start, run_total))
            run, run_total = [], 0

        pos = start
        for c in (int(x) for x in np.nonzero(hist)[0]):
            cnt = int(hist[c])
            lo, hi = pos, pos + cnt
            pos = hi
            if hi <= row_lo or lo >= row_hi:
                flush()
                continue
            inside = lo >= row_lo and hi <= row_hi
            too_big = cnt > batch or (not inside and cnt > batch // 8)
            if too_big and len(prefix) + 1 < max_prefix and c != 0:   # (0 = behind the text: nothing follows)
                flush()
                visit(prefix + (c,), lo, count_next(prefix + (c,)))
                continue
            if run and run_total + cnt > batch:
                flush()
            if not run:
                run_start = lo
            run.append(c)
            run_total += cnt
        flush()

    visit((), 0, count_next(()))
    return jobs


def suffix_batches_range(T, n: int, row_lo: int, row_hi: int, batch: int = 1 << 27) -> Iterator[torch.Tensor]:
    """The suffix array entries of rows [row_lo, row_hi), in order, in pieces (int64, on T's device).
    T: ByteText, or the int16 prepared text of ``prepare_text_gpu``."""
    if row_hi <= row_lo:
        return
    pairs = _pair_histogram(T, n)       # one pass answers every prefix of length 0 and 1

    def count_next(prefix: Tuple[int, ...]) -> np.ndarray:
        if len(prefix) == 0:
            return pairs.sum(axis=1)
        if len(prefix) == 1:
            return pairs[prefix[0]]
        return _next_histogram(T, n, prefix)

    for prefix, lo, hi, start, count in plan_jobs(count_next, batch, row_lo, row_hi):
        pos = _select(T, n, prefix, lo, hi)
        assert pos.numel() == count, "bucket plan and text disagree"
        sa = _sort_batch(T, n, pos)
        a, b = max(row_lo, start) - start, min(row_hi, start + count) - start
        yield sa[a:b] if (a, b) != (0, count) else sa


# --------------------------------------------------------------------------------------------
# the build

def blocks_of_rank(nblocks: int, block_size: int, total_length: int, rank: int, world: int) -> Tuple[int, int]:
    """[first, first + count) = the data blocks rank `rank` serves (and therefore builds)."""
    mine = [b for b in range(nblocks) if shard_of_block(b, block_size, total_length, world) == rank]
    if not mine:
        return 0, 0
    assert mine == list(range(mine[0], mine[-1] + 1))
    return mine[0], len(mine)


def build_rank_blocks(T, doc_ends: np.ndarray, out_dir: str, rank: int, world: int,
                      block_size: int = 128 << 20, bucket_size: int = 1 << 20, chunk_size: int = 2048,
                      mark_period: int = 20, nthreads: int = 0, batch: int = 1 << 27, host_chunk: int = 1 << 26,
                      log: Optional[Callable[[str], None]] = None):
    """Rank `rank`'s share of the build, no communication: sorts the suffixes of its rows and writes
    its data blocks into out_dir.  -> ((first_block, block_counts, eof_rows), timings)."""
    t0 = time.time()
    n = int(doc_ends[-1])
    nblocks = (n + block_size - 1) // block_size
    first, count = blocks_of_rank(nblocks, block_size, n, rank, world)
    row_lo, row_hi = min(n, first * block_size), min(n, (first + count) * block_size)
    builder = IndexBuilder(out_dir, doc_ends, block_size=block_size, bucket_size=bucket_size, chunk_size=chunk_size,
                           mark_period=mark_period, nthreads=nthreads, first_block=first, range_blocks=count)
    on_gpu = T.device.type == "cuda"
    host_chunk = max(1, min(host_chunk, row_hi - row_lo))
    pin_L = torch.empty(host_chunk, dtype=torch.int16)
    pin_S = torch.empty(host_chunk, dtype=torch.int64)
    if on_gpu:
        pin_L, pin_S = pin_L.pin_memory(), pin_S.pin_memory()
    rows = 0
    t_sort = t_emit = 0.0
    t1 = time.time()
    for sa in suffix_batches_range(T, n, row_lo, row_hi, batch):
        prev = torch.where(sa == 0, torch.full_like(sa, n - 1), sa - 1)
        L = T.symbols(prev).to(torch.int16) if isinstance(T, ByteText) else T[prev]
        if on_gpu:
            torch.cuda.synchronize()
        t2 = time.time()
        t_sort += t2 - t1
        for s in range(0, sa.numel(), host_chunk):
            e = min(sa.numel(), s + host_chunk)
            pin_L[: e - s].copy_(L[s:e])
            pin_S[: e - s].copy_(sa[s:e])
            if on_gpu:
                torch.cuda.synchronize()
            builder.append(pin_L[: e - s].numpy().view(np.uint16), pin_S[: e - s].numpy())
        rows += sa.numel()
        del sa, L, prev
        t1 = time.time()
        t_emit += t1 - t2
        if log:
            log(f"  build[{rank}]: {rows}/{row_hi - row_lo} rows  sort {t_sort:.1f}s emit {t_emit:.1f}s")
    assert rows == row_hi - row_lo
    counts, eof = builder.finish_range()
    return (first, counts, eof), {"rows": rows, "first_block": first, "blocks": count, "sort_s": t_sort,
                                  "emit_s": t_emit, "total_s": time.time() - t0}


def write_header_from_parts(out_dir: str, doc_ends: np.ndarray, parts, block_size: int = 128 << 20,
                            bucket_size: int = 1 << 20, chunk_size: int = 2048, mark_period: int = 20,
                            doc_infos=None) -> None:
    """Header block from every rank's (first_block, block_counts, eof_rows)."""
    n = int(doc_ends[-1])
    nblocks = (n + block_size - 1) // block_size
    all_counts = np.zeros((nblocks, ALPHA_SIZE), dtype=np.int64)
    all_eof = np.full(len(doc_ends), -1, dtype=np.int64)
    covered = 0
    for f, c, e in parts:
        all_counts[f: f + len(c)] = c
        covered += len(c)
        all_eof = np.maximum(all_eof, e)
    if covered != nblocks:
        raise ValueError(f"the ranks wrote {covered} data blocks of {nblocks}")
    write_index_header(out_dir, doc_ends, all_counts, all_eof, block_size=block_size, bucket_size=bucket_size,
                       chunk_size=chunk_size, mark_period=mark_period, doc_infos=doc_infos)


def build_index_distributed(T, doc_ends: np.ndarray, out_dir: str, rank: int, world: int, group=None,
                            log: Optional[Callable[[str], None]] = None, **params) -> Dict[str, float]:
    """Collective over `group` (world ranks, all on one file system): rank r sorts and writes its data
    blocks of the index of text T (replicated on every rank) into out_dir; rank 0 adds the header.
    params: block_size, bucket_size, chunk_size, mark_period, nthreads, batch, host_chunk.
    world == 1 needs no process group."""
    import torch.distributed as dist

    t0 = time.time()
    part, stats = build_rank_blocks(T, doc_ends, out_dir, rank, world, log=log, **params)
    if world > 1:
        parts: List = [None] * world
        dist.all_gather_object(parts, part, group=group)
    else:
        parts = [part]
    if rank == 0:
        fmt = {k: v for k, v in params.items() if k in ("block_size", "bucket_size", "chunk_size", "mark_period")}
        write_header_from_parts(out_dir, doc_ends, parts, **fmt)
    if world > 1:
        dist.barrier(group=group)
    stats["total_s"] = time.time() - t0
    return stats
