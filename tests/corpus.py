"""Deterministic synthetic corpora and pattern sets shared by the CPU and GPU tests."""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

CHARACTER_OFFSET = 5


def generate_text(n: int) -> bytes:
    """Same idea as the reference's test generator (src/main/index_test_funcs.c:316-337):
    a counter written in base 6 over 'a'..'f' -- many repeats, interesting to search."""
    out = bytearray(b"x" * n)
    i, j = 0, 0
    while i < n:
        k = j
        while k > 0 and i < n:
            out[i] = ord("a") + (k - 1) % 6
            i += 1
            k //= 6
        j += 1
    return bytes(out)


def random_bytes(n: int, seed: int) -> bytes:
    return np.random.default_rng(seed).integers(0, 256, n, dtype=np.uint8).tobytes()


def random_acgt(n: int, seed: int) -> bytes:
    return np.random.default_rng(seed).choice(np.frombuffer(b"ACGT", dtype=np.uint8), n).tobytes()


def skewed_text(n: int, seed: int, nsym: int = 60) -> bytes:
    """Geometric symbol distribution: Huffman codes far deeper than 8 levels."""
    rng = np.random.default_rng(seed)
    p = 0.5 ** np.arange(1, nsym + 1)
    p /= p.sum()
    return (rng.choice(nsym, n, p=p).astype(np.uint8) + 33).tobytes()


def english_like(n: int, seed: int) -> bytes:
    rng = np.random.default_rng(seed)
    vocab = [bytes(rng.integers(97, 123, rng.integers(2, 9), dtype=np.uint8)) for _ in range(300)]
    ranks = np.arange(1, len(vocab) + 1)
    p = (1.0 / ranks) / (1.0 / ranks).sum()
    words = rng.choice(len(vocab), n // 4 + 8, p=p)
    text = b" ".join(vocab[w] for w in words)
    return text[:n]


def all_bytes_doc() -> bytes:
    """Every byte value incl. NUL and 0x01..0x04, as the reference's CLI test (src/test/test.pl:52-55)."""
    return bytes(range(256)) + bytes(range(255, -1, -1))


def to_alpha(b: bytes) -> np.ndarray:
    return np.frombuffer(b, dtype=np.uint8).astype(np.uint16) + CHARACTER_OFFSET


def sample_patterns(docs: Sequence[bytes], n: int, lengths: Sequence[int], seed: int,
                    random_fraction: float = 0.25) -> List[np.ndarray]:
    """Patterns sampled from the documents (count >= 1) mixed with uniformly random ones."""
    rng = np.random.default_rng(seed)
    pats: List[np.ndarray] = []
    docs = [d for d in docs if len(d) > 0]
    for i in range(n):
        m = int(lengths[i % len(lengths)])
        if rng.random() < random_fraction or not docs:
            pats.append(rng.integers(0, 256, m, dtype=np.uint8).astype(np.uint16) + CHARACTER_OFFSET)
            continue
        d = docs[int(rng.integers(0, len(docs)))]
        if len(d) < m:
            pats.append(to_alpha(d))
            continue
        s = int(rng.integers(0, len(d) - m + 1))
        pats.append(to_alpha(d[s:s + m]))
    return pats


def brute_count(docs: Sequence[bytes], pat: np.ndarray) -> int:
    """Occurrences of pat inside documents (patterns never contain SEOF, so no match spans two)."""
    if len(pat) == 0:
        return sum(len(d) + 1 for d in docs)
    if (pat < CHARACTER_OFFSET).any():
        return -1
    p = bytes((pat - CHARACTER_OFFSET).astype(np.uint8))
    total = 0
    for d in docs:
        start = 0
        while True:
            k = d.find(p, start)
            if k < 0:
                break
            total += 1
            start = k + 1
    return total


def brute_locate(docs: Sequence[bytes], pat: np.ndarray) -> List[int]:
    """Global text offsets (documents laid end to end, one SEOF after each)."""
    p = bytes((pat - CHARACTER_OFFSET).astype(np.uint8))
    out = []
    base = 0
    for d in docs:
        start = 0
        while True:
            k = d.find(p, start)
            if k < 0:
                break
            out.append(base + k)
            start = k + 1
        base += len(d) + 1
    return sorted(out)


# The corpora every suite runs on: name -> (documents, builder parameters)
def standard_corpora():
    two = [b"test_one;", b"test_two_fun;"]
    return {
        # the reference's golden fixture (src/main/index_test.c:514-533)
        "two_docs": (two, dict(mark_period=100)),
        # the reference's non-default parameter sets (src/main/index_test_funcs.c:46-86)
        "gen400_big_buckets": ([generate_text(400)], dict(block_size=10000, bucket_size=1000, chunk_size=1000)),
        "gen400_small_buckets": ([generate_text(400)], dict(block_size=10000, bucket_size=4, chunk_size=2)),
        "gen400_small_blocks": ([generate_text(400)], dict(block_size=16, bucket_size=4, chunk_size=8 // 2)),
        "gen13_small_blocks": ([generate_text(13)], dict(block_size=16, bucket_size=4, chunk_size=4)),
        "gen3": ([generate_text(3)], dict()),
        "single_symbol": ([b"a" * 300], dict(block_size=256, bucket_size=64, chunk_size=32, mark_period=7)),
        "multi_doc_mixed": ([b"", b"a", all_bytes_doc(), generate_text(50), random_acgt(3000, 7), b"\x00\x01\x02\x03\x04"],
                            dict(block_size=2048, bucket_size=512, chunk_size=128, mark_period=5)),
        "acgt_64k": ([random_acgt(30000, 1), random_acgt(35000, 2)],
                     dict(block_size=32768, bucket_size=4096, chunk_size=1024)),
        "bytes_200k": ([random_bytes(200000, 3)], dict(block_size=131072, bucket_size=65536, chunk_size=2048)),
        "skewed_deep": ([skewed_text(60000, 4)], dict(block_size=65536, bucket_size=16384, chunk_size=0)),
        "english_100k": ([english_like(50000, 5), english_like(50000, 6)],
                         dict(block_size=65536, bucket_size=8192, chunk_size=2048, mark_period=20)),
    }
