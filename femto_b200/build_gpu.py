"""GPU-side suffix sorting and synthetic corpora for large indexes (index build = "next" row f-1).

The query engine needs an index in femto's on-disk format.  The reference builds one at ~1 MB/s
(SURVEY.md section 6), i.e. more than an hour for the 4 GiB headline corpus, so the build is done
here: suffix array on the GPU with torch ops (plumbing, not the hot path), then the host emitter
``IndexBuilder`` (femto_b200/csrc/fm_builder.cc) writes blocks that are byte-identical to what the
reference's constructor would write for the same rows (tests/test_builder_format.py).

Suffix order = plain lexicographic order of the prepared text's suffixes, a suffix that is a proper
prefix of another first -- the order of the reference's builders (bwt_qsufsort.c / dcx_cc).

Method: suffixes are handled in batches of consecutive first symbols (<= ``batch`` suffixes each);
inside a batch every suffix gets a 63-bit key packing its first 7 symbols (9 bits each), the batch
is sorted by key, and groups of equal keys are refined with the next 7 symbols until no ties
remain.  Random and natural-language-like corpora resolve in one or a few rounds; highly
repetitive corpora should use the host sorter instead (``femto_b200.suffix_sort_host``).
"""
from __future__ import annotations

import time
from typing import Callable, Dict, Iterator, List, Optional, Tuple

import numpy as np
import torch

from . import CHARACTER_OFFSET, ESCAPE_CODE_SEOF, IndexBuilder

SYM_BITS = 9
SYMS_PER_KEY = 7
PAD = 64


# --------------------------------------------------------------------------------------------
# deterministic synthetic corpora (counter-based, identical on any device / torch version)

def _splitmix64(x: torch.Tensor) -> torch.Tensor:
    """splitmix64 finaliser on int64 tensors (wrapping arithmetic, logical shifts emulated)."""
    def lsr(v, k):
        return (v >> k) & ((1 << (64 - k)) - 1)
    x = x + (-7046029254386353131)            # 0x9E3779B97F4A7C15
    x = (x ^ lsr(x, 30)) * (-4658895280553007687)   # 0xBF58476D1CE4E5B9
    x = (x ^ lsr(x, 27)) * (-7723592293110705685)   # 0x94D049BB133111EB
    return x ^ lsr(x, 31)


def synthetic_bytes(n: int, seed: int, device, alphabet: Optional[bytes] = None, chunk: int = 1 << 27) -> torch.Tensor:
    """n i.i.d. uniform symbols as a uint8 tensor: byte k of splitmix64(seed*2^40 + i) for the 8
    bytes of word i; with ``alphabet`` the byte is reduced modulo len(alphabet) (power of two)."""
    out = torch.empty(n, dtype=torch.uint8, device=device)
    nwords = (n + 7) // 8
    lut = None
    if alphabet is not None:
        assert len(alphabet) & (len(alphabet) - 1) == 0, "alphabet size must be a power of two"
        lut = torch.tensor(list(alphabet), dtype=torch.uint8, device=device)
    for w0 in range(0, nwords, chunk):
        w1 = min(nwords, w0 + chunk)
        idx = torch.arange(w0, w1, dtype=torch.int64, device=device) + (int(seed) << 40)
        b = _splitmix64(idx).view(torch.uint8)
        if lut is not None:
            b = lut[(b & (len(alphabet) - 1)).long()]
        lo, hi = w0 * 8, min(n, w1 * 8)
        out[lo:hi] = b[: hi - lo]
    return out


def synthetic_bytes_numpy(n: int, seed: int, alphabet: Optional[bytes] = None) -> np.ndarray:
    """CPU twin of synthetic_bytes (same bytes), for tests."""
    nwords = (n + 7) // 8
    x = (np.arange(nwords, dtype=np.uint64) + (np.uint64(seed) << np.uint64(40)))
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    b = x.view(np.uint8)[:n]
    if alphabet is not None:
        b = np.frombuffer(alphabet, dtype=np.uint8)[b & (len(alphabet) - 1)]
    return b


# --------------------------------------------------------------------------------------------
# prepared text

def prepare_text_gpu(docs: List[torch.Tensor]) -> Tuple[torch.Tensor, np.ndarray]:
    """uint8 document tensors -> (int16 prepared text padded with PAD zeros, doc_ends).
    Layout as the reference's prepared text without headers: bytes as 5+byte, one SEOF per
    document (src/main/bwt_prepare.c:231-311)."""
    n = sum(int(d.numel()) + 1 for d in docs)
    dev = docs[0].device
    T = torch.zeros(n + PAD, dtype=torch.int16, device=dev)
    ends = []
    pos = 0
    for d in docs:
        m = int(d.numel())
        step = 1 << 28
        for s in range(0, m, step):
            e = min(m, s + step)
            T[pos + s: pos + e] = d[s:e].to(torch.int16) + CHARACTER_OFFSET
        T[pos + m] = ESCAPE_CODE_SEOF
        pos += m + 1
        ends.append(pos)
    return T, np.array(ends, dtype=np.int64)


def english_vocabulary(vocab: int, seed: int) -> Tuple[np.ndarray, np.ndarray]:
    """A fixed vocabulary for synthetic_english: `vocab` lower-case words, frequent ones short
    (rank r has 2 + min(8, floor(log2(r+1))) letters, drawn from splitmix64).  Returns (letters
    concatenated, word offsets [vocab + 1]).  Host side, numpy, device independent."""
    ranks = np.arange(vocab, dtype=np.int64)
    wlen = 2 + np.minimum(8, np.floor(np.log2(ranks + 1.0)).astype(np.int64))
    woff = np.zeros(vocab + 1, dtype=np.int64)
    woff[1:] = np.cumsum(wlen)
    total = int(woff[-1])
    # letters with English-like frequencies (per mille, a..z), by inverse CDF on 16-bit draws
    permille = np.array([82, 15, 28, 43, 127, 22, 20, 61, 70, 2, 8, 40, 24, 67, 75, 19, 1, 60, 63, 91, 28, 10,
                         24, 2, 20, 1], dtype=np.float64)
    cdf = np.cumsum(permille) / permille.sum()
    raw = synthetic_bytes_numpy(2 * total, seed * 7919 + 13).astype(np.uint32)
    u = (raw[0::2] * 256 + raw[1::2]) / 65536.0
    letters = (97 + np.minimum(np.searchsorted(cdf, u, side="right"), 25)).astype(np.uint8)
    return letters, woff


def synthetic_english(n: int, seed: int, device, vocab: int = 50000, words_per_chunk: int = 1 << 22) -> torch.Tensor:
    """n bytes of English-like text (BASELINE configs[3]): words of a fixed `vocab`-word vocabulary
    drawn with Zipf(s=1) rank frequencies and separated by single spaces.  Counter-based and integer /
    exact-float only, so the bytes are identical on any device and torch version.  Byte entropy is
    ~4.2 bits, so the index gets Huffman codes of 3..15 bits (deep wavelet trees: three quad rounds for
    the rare letters), and repeats long enough to need several tie-refinement rounds in the sorter."""
    letters_h, woff_h = english_vocabulary(vocab, seed)
    ranks = np.arange(1, vocab + 1, dtype=np.float64)
    cdf_h = np.cumsum(1.0 / ranks)
    cdf_h /= cdf_h[-1]
    letters = torch.from_numpy(letters_h).to(device)
    woff = torch.from_numpy(woff_h).to(device)
    cdf = torch.from_numpy(cdf_h).to(device)
    out = torch.empty(n, dtype=torch.uint8, device=device)
    pos, w0 = 0, 0
    while pos < n:
        idx = torch.arange(w0, w0 + words_per_chunk, dtype=torch.int64, device=device) + ((int(seed) + 1) << 44)
        x = _splitmix64(idx)
        u = ((x >> 11) & ((1 << 53) - 1)).to(torch.float64) * (2.0 ** -53)   # exact in float64
        wid = torch.clamp(torch.searchsorted(cdf, u, right=True), max=vocab - 1)
        wl = woff[wid + 1] - woff[wid]
        ends = torch.cumsum(wl + 1, 0)                      # each word is followed by one space
        total = int(ends[-1])
        take = min(total, n - pos)
        p = torch.arange(take, dtype=torch.int64, device=device)
        w = torch.searchsorted(ends, p, right=True)         # word that byte p belongs to
        k = p - (ends[w] - wl[w] - 1)                       # position inside "word + space"
        inside = k < wl[w]
        src = woff[wid[w]] + torch.where(inside, k, torch.zeros_like(k))
        out[pos:pos + take] = torch.where(inside, letters[src], torch.full_like(letters[src], 32))
        pos += take
        w0 += words_per_chunk
    return out


# --------------------------------------------------------------------------------------------
# suffix sorting

def _pack_keys(T, pos: torch.Tensor, depth: int) -> torch.Tensor:
    if not isinstance(T, torch.Tensor):  # build_dist.ByteText: one byte per position + document ends
        return T.pack_keys(pos, depth)
    key = torch.zeros_like(pos)
    for k in range(SYMS_PER_KEY):
        key = (key << SYM_BITS) | T[pos + (depth + k)].long()
    return key


def _positions_with_first_symbol(T: torch.Tensor, n: int, lo: int, hi: int, second: Optional[Tuple[int, int]],
                                 chunk: int = 1 << 29) -> torch.Tensor:
    parts = []
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        t = T[s:e]
        m = (t >= lo) & (t <= hi)
        if second is not None:
            t1 = T[s + 1: e + 1]
            m &= (t1 >= second[0]) & (t1 <= second[1])
        parts.append(torch.nonzero(m).squeeze(1) + s)
    return torch.cat(parts) if len(parts) > 1 else parts[0]


def _sort_batch(T: torch.Tensor, n: int, pos: torch.Tensor, max_rounds: int = 100000) -> torch.Tensor:
    """Sort the suffixes starting at `pos` (all share nothing in particular) -> positions in suffix order."""
    if pos.numel() <= 1:
        return pos
    key = _pack_keys(T, pos, 0)
    key, order = torch.sort(key)
    pos = pos[order]
    del order
    # group id of every element = index of the first element of its run of equal keys
    depth = SYMS_PER_KEY
    neq = torch.ones(pos.numel(), dtype=torch.bool, device=pos.device)
    neq[1:] = key[1:] != key[:-1]
    del key
    rounds = 0
    while True:
        tied = ~neq
        tied[:-1] |= ~neq[1:]              # element i is tied if it equals its predecessor or successor
        idx = torch.nonzero(tied).squeeze(1)
        if idx.numel() == 0:
            return pos
        rounds += 1
        if rounds > max_rounds or depth >= n + SYMS_PER_KEY:
            raise RuntimeError("suffix sort did not converge (highly repetitive text): use suffix_sort_host")
        group = torch.cumsum(neq.long(), 0)[idx]           # run id of each tied element
        p = pos[idx]
        k2 = _pack_keys(T, torch.clamp(p, max=n + PAD - depth - SYMS_PER_KEY - 1), depth)
        k2 = torch.where(p + depth < n, k2, torch.zeros_like(k2))   # past the end: smallest
        # sort tied elements by (group, key): two stable passes
        k2s, o1 = torch.sort(k2, stable=True)
        g1 = group[o1]
        g2, o2 = torch.sort(g1, stable=True)
        perm = o1[o2]
        k2s = k2s[o2]
        pos[idx] = p[perm]
        new_neq = torch.ones(idx.numel(), dtype=torch.bool, device=pos.device)
        new_neq[1:] = (g2[1:] != g2[:-1]) | (k2s[1:] != k2s[:-1])
        neq[idx] = new_neq
        depth += SYMS_PER_KEY


def suffix_batches(T: torch.Tensor, n: int, batch: int = 1 << 28) -> Iterator[torch.Tensor]:
    """Yields the suffix array in order, in pieces (int64 position tensors on T's device)."""
    hist = torch.zeros(512, dtype=torch.int64, device=T.device)
    step = 1 << 29
    for s in range(0, n, step):
        hist += torch.bincount(T[s:min(n, s + step)].long(), minlength=512)
    hist = hist.cpu().numpy()
    syms = [int(c) for c in np.nonzero(hist)[0]]
    i = 0
    while i < len(syms):
        c = syms[i]
        if hist[c] > batch:
            # one first symbol is too frequent: split it by the second symbol
            h2 = torch.zeros(512, dtype=torch.int64, device=T.device)
            for s in range(0, n, step):
                e = min(n, s + step)
                sel = T[s:e] == c
                h2 += torch.bincount(T[s + 1:e + 1][sel].long(), minlength=512)
            h2 = h2.cpu().numpy()
            seconds = [int(x) for x in np.nonzero(h2)[0]]
            j = 0
            while j < len(seconds):
                lo2 = seconds[j]
                tot = int(h2[lo2])
                k = j + 1
                while k < len(seconds) and tot + int(h2[seconds[k]]) <= batch:
                    tot += int(h2[seconds[k]])
                    k += 1
                pos = _positions_with_first_symbol(T, n, c, c, (lo2, seconds[k - 1]))
                yield _sort_batch(T, n, pos)
                j = k
            i += 1
            continue
        tot = int(hist[c])
        k = i + 1
        while k < len(syms) and tot + int(hist[syms[k]]) <= batch:
            tot += int(hist[syms[k]])
            k += 1
        pos = _positions_with_first_symbol(T, n, c, syms[k - 1], None)
        yield _sort_batch(T, n, pos)
        i = k


def suffix_array_gpu(T: torch.Tensor, n: int, batch: int = 1 << 28) -> torch.Tensor:
    return torch.cat(list(suffix_batches(T, n, batch)))


# --------------------------------------------------------------------------------------------
# end-to-end: documents on the GPU -> index directory

def build_index_gpu(docs: List[torch.Tensor], out_dir: str, block_size: int = 128 << 20, bucket_size: int = 1 << 20,
                    chunk_size: int = 2048, mark_period: int = 20, nthreads: int = 0, batch: int = 1 << 28,
                    host_chunk: int = 1 << 26, log: Optional[Callable[[str], None]] = None) -> Dict[str, float]:
    """Suffix-sort on the GPU, stream (L, SA) rows to the host emitter.  Returns timing info."""
    t0 = time.time()
    T, ends = prepare_text_gpu(docs)
    n = int(ends[-1])
    builder = IndexBuilder(out_dir, ends, block_size=block_size, bucket_size=bucket_size, chunk_size=chunk_size,
                           mark_period=mark_period, nthreads=nthreads)
    on_gpu = T.is_cuda
    host_chunk = min(host_chunk, n)
    pin_L = torch.empty(host_chunk, dtype=torch.int16)
    pin_S = torch.empty(host_chunk, dtype=torch.int64)
    if on_gpu:
        pin_L, pin_S = pin_L.pin_memory(), pin_S.pin_memory()

    def sync():
        if on_gpu:
            torch.cuda.synchronize()

    t_sort = t_emit = 0.0
    rows = 0
    t1 = time.time()
    for sa in suffix_batches(T, n, batch):
        prev = torch.where(sa == 0, torch.full_like(sa, n - 1), sa - 1)
        L = T[prev]
        sync()
        t2 = time.time()
        t_sort += t2 - t1
        for s in range(0, sa.numel(), host_chunk):
            e = min(sa.numel(), s + host_chunk)
            pin_L[: e - s].copy_(L[s:e])
            pin_S[: e - s].copy_(sa[s:e])
            sync()
            builder.append(pin_L[: e - s].numpy().view(np.uint16), pin_S[: e - s].numpy())
        rows += sa.numel()
        del sa, L, prev
        t1 = time.time()
        t_emit += t1 - t2
        if log:
            log(f"  build: {rows}/{n} rows  sort {t_sort:.1f}s emit {t_emit:.1f}s")
    assert rows == n
    t2 = time.time()
    builder.finish()
    t_emit += time.time() - t2
    del T
    if on_gpu:
        torch.cuda.empty_cache()
    return {"rows": n, "sort_s": t_sort, "emit_s": t_emit, "total_s": time.time() - t0}
