"""femto_b200/build_dist.py -- the index build split over ranks by BWT row range -- on CPU tensors.

  * ByteText (one byte per position + document ends) yields the symbols and sort keys of the int16
    prepared text, also with documents shorter than a key (several SEOF inside one key);
  * suffix_batches_range(row_lo, row_hi) = that slice of the whole suffix array, for both text forms,
    whatever the batch size (first-symbol groups, second-symbol splits, borders inside a bucket);
  * range builders (fm_builder_create_range / _finish_range / _write_header) write, together, files
    byte-identical to the one-process builder's -- for 1, 2, 3, 5 and more-ranks-than-blocks splits;
  * the real thing: 2 and 3 gloo processes building one index side by side.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import corpus
import femto_b200 as fb
from femto_b200 import build_dist, build_gpu
from femto_b200.build_dist import ByteText


def _t(d: bytes) -> torch.Tensor:
    return torch.frombuffer(bytearray(d), dtype=torch.uint8) if d else torch.zeros(0, dtype=torch.uint8)


def _docs_cases():
    return {
        "random_bytes": [corpus.random_bytes(20000, 1)],
        "acgt_two_docs": [corpus.random_acgt(9000, 2), corpus.random_acgt(7000, 3)],
        "english_docs": [corpus.english_like(6000, 4), corpus.english_like(5000, 5), corpus.english_like(4000, 6)],
        "repetitive": [b"abcabcabc" * 300, b"abcabc" * 100, b"a" * 500],
        "tiny": [b"", b"a", b"ba", b"", b"abcdefgh", b"b"],
    }


@pytest.mark.parametrize("name", list(_docs_cases()))
def test_byte_text_equals_prepared_text(name):
    docs = _docs_cases()[name]
    T, ends = build_gpu.prepare_text_gpu([_t(d) for d in docs])
    B = ByteText.from_docs([_t(d) for d in docs])
    n = int(ends[-1])
    assert B.n == n and (B.doc_ends == ends).all()
    assert B.sparse_seof == (name != "tiny")
    assert (B.slice_symbols(0, n + build_gpu.PAD) == T).all()
    for lo, hi in [(0, 1), (n - 1, n + 3), (n // 3, n // 2), (n, n + 5)]:
        assert (B.slice_symbols(lo, hi) == T[lo:hi]).all()
    rng = np.random.default_rng(0)
    pos = torch.from_numpy(np.concatenate([rng.integers(0, n, 500), np.arange(max(0, n - 12), n), B.doc_ends - 1,
                                           np.maximum(B.doc_ends - 4, 0)]).astype(np.int64))
    assert (B.symbols(pos) == T[pos].long()).all()
    for depth in (0, 7, 21):
        p = torch.clamp(pos, max=n + build_gpu.PAD - depth - build_gpu.SYMS_PER_KEY - 1)
        key_t = build_gpu._pack_keys(T, p, depth)
        assert (build_gpu._pack_keys(B, p, depth) == key_t).all()


@pytest.mark.parametrize("name", list(_docs_cases()))
@pytest.mark.parametrize("batch", [1 << 28, 2500, 300])
def test_row_range_of_the_suffix_array(name, batch):
    docs = _docs_cases()[name]
    T, ends = build_gpu.prepare_text_gpu([_t(d) for d in docs])
    B = ByteText.from_docs([_t(d) for d in docs])
    n = int(ends[-1])
    text, _ = fb.prepare_text(docs)
    full = fb.suffix_sort_host(text)
    cuts = sorted({0, 1, n // 7, n // 3, n // 2, n - 1, n})
    for text_form in (T, B):
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            got = list(build_dist.suffix_batches_range(text_form, n, lo, hi, batch=batch))
            got = torch.cat(got).numpy() if got else np.zeros(0, np.int64)
            assert (got == full[lo:hi]).all(), (lo, hi)
        assert list(build_dist.suffix_batches_range(text_form, n, 5, 5, batch=batch)) == []


def test_plan_sorts_little_beyond_the_rows_it_needs():
    """A rank whose rows end inside a first-symbol bucket sorts only the finer buckets that overlap."""
    docs = [corpus.random_bytes(200000, 3)]
    B = ByteText.from_docs([_t(d) for d in docs])
    n = B.n
    pairs = build_dist._pair_histogram(B, n)
    assert pairs.sum() == n and pairs[fb.ESCAPE_CODE_SEOF, 0] == 1
    assert (build_dist._next_histogram(B, n, ()) == pairs.sum(axis=1)).all()
    assert (build_dist._next_histogram(B, n, (70,)) == pairs[70]).all()
    h3 = build_dist._next_histogram(B, n, (70, 71))
    assert h3.sum() == pairs[70, 71] and h3.sum() > 0
    calls = []

    def count_next(prefix):
        calls.append(prefix)
        return pairs.sum(axis=1) if not prefix else pairs[prefix[0]] if len(prefix) == 1 else \
            build_dist._next_histogram(B, n, prefix)

    lo, hi = n // 4 + 13, n // 2 + 7
    jobs = build_dist.plan_jobs(count_next, 20000, lo, hi)
    sorted_rows = sum(j[4] for j in jobs)
    assert hi - lo <= sorted_rows <= (hi - lo) + 2 * (20000 // 8)          # at most a small bucket per border
    assert all(j[4] <= 20000 for j in jobs)
    assert jobs[0][3] <= lo and jobs[-1][3] + jobs[-1][4] >= hi
    for a, b in zip(jobs[:-1], jobs[1:]):
        assert a[3] + a[4] == b[3]                                          # contiguous in row order
    assert len(calls) <= 4                                                  # the root and the border buckets only


def test_plan_splits_big_buckets_as_deep_as_needed():
    """English-like text: ' ' and ' t' hold far more suffixes than a batch; they are split by longer prefixes."""
    text = build_gpu.synthetic_english(400000, 9, "cpu", vocab=500, words_per_chunk=1 << 14)
    B = ByteText.from_docs([text])
    n = B.n
    full = fb.suffix_sort_host(fb.prepare_text([bytes(text.numpy())])[0])
    pairs = build_dist._pair_histogram(B, n)
    deepest = 0

    def count_next(prefix):
        nonlocal deepest
        deepest = max(deepest, len(prefix))
        return pairs.sum(axis=1) if not prefix else pairs[prefix[0]] if len(prefix) == 1 else \
            build_dist._next_histogram(B, n, prefix)

    batch = 3000
    jobs = build_dist.plan_jobs(count_next, batch, 0, n)
    assert deepest >= 3 and sum(j[4] for j in jobs) == n
    big = [j for j in jobs if j[4] > batch]
    assert all(len(j[0]) == build_dist.MAX_PREFIX - 1 and j[1] == j[2] for j in big)   # only at the depth limit
    got = torch.cat(list(build_dist.suffix_batches_range(B, n, n // 3, n // 2, batch=batch))).numpy()
    assert (got == full[n // 3: n // 2]).all()


PARAMS = dict(block_size=4096, bucket_size=1024, chunk_size=256, mark_period=20)


def _same_files(a: str, b: str):
    assert sorted(os.listdir(a)) == sorted(os.listdir(b))
    for f in sorted(os.listdir(a)):
        assert open(os.path.join(a, f), "rb").read() == open(os.path.join(b, f), "rb").read(), f


@pytest.mark.parametrize("world", [1, 2, 3, 5, 9])
@pytest.mark.parametrize("chunk_size", [256, 0])
def test_range_builders_write_the_one_process_index(tmp_path, world, chunk_size):
    docs = [corpus.random_bytes(9000, 5), corpus.english_like(8000, 6), b"x", corpus.random_acgt(3500, 7)]
    params = dict(PARAMS, chunk_size=chunk_size)
    a, b = str(tmp_path / "host"), str(tmp_path / f"ranks{world}")
    infos = [b"first", None, b"", b"last one"]
    fb.build_index_host(docs, a, doc_infos=[i if i is not None else b"doc1" for i in infos], **params)
    B = ByteText.from_docs([_t(d) for d in docs])
    parts = []
    nblocks = (B.n + params["block_size"] - 1) // params["block_size"]
    assert nblocks == 6
    for rank in reversed(range(world)):      # any order: the blocks are independent files
        part, stats = build_dist.build_rank_blocks(B, B.doc_ends, b, rank, world, batch=3000, host_chunk=2500, **params)
        parts.append(part)
        assert stats["rows"] == min(B.n, (part[0] + stats["blocks"]) * params["block_size"]) - min(
            B.n, part[0] * params["block_size"])
    assert sum(len(p[1]) for p in parts) == nblocks
    build_dist.write_header_from_parts(b, B.doc_ends, parts, doc_infos=infos, **params)
    _same_files(a, b)


def test_range_builder_argument_checks(tmp_path):
    ends = np.array([10000], dtype=np.int64)
    with pytest.raises(fb.FemtoError):
        fb.IndexBuilder(str(tmp_path / "x"), ends, first_block=2, range_blocks=2, **PARAMS)      # 3 blocks only
    b = fb.IndexBuilder(str(tmp_path / "x"), ends, first_block=1, range_blocks=1, **PARAMS)
    L = np.full(4096, 70, dtype=np.uint16)
    with pytest.raises(fb.FemtoError):
        b.finish_range()                                                                           # no rows yet
    b = fb.IndexBuilder(str(tmp_path / "x"), ends, first_block=1, range_blocks=1, **PARAMS)
    b.append(L, np.arange(4096, dtype=np.int64))
    with pytest.raises(fb.FemtoError):
        b.append(L[:1], np.zeros(1, dtype=np.int64))                                               # beyond its range
    b.abort()
    b = fb.IndexBuilder(str(tmp_path / "x"), ends, first_block=1, range_blocks=1, **PARAMS)
    with pytest.raises(fb.FemtoError):
        b.finish()                                                                                 # wrong kind of finish
    counts = np.zeros((3, fb.ALPHA_SIZE), dtype=np.int64)
    with pytest.raises(fb.FemtoError):                                                             # counts do not add up
        fb.write_index_header(str(tmp_path / "x"), ends, counts, np.zeros(1, dtype=np.int64), **PARAMS)
    with pytest.raises(ValueError):
        fb.write_index_header(str(tmp_path / "x"), ends, counts[:2], np.zeros(1, dtype=np.int64), **PARAMS)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _docs_for_ranks():
    return [corpus.random_bytes(14000, 11), corpus.english_like(9000, 12)]


def _build_worker(rank, world, port, out_dir, use_bytes):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        docs = [_t(d) for d in _docs_for_ranks()]   # every rank makes the same text (synthetic corpora are seeded)
        if use_bytes:
            T = ByteText.from_docs(docs)
            ends = T.doc_ends
        else:
            T, ends = build_gpu.prepare_text_gpu(docs)
        stats = build_dist.build_index_distributed(T, ends, out_dir, rank, world, batch=4000, nthreads=2, **PARAMS)
        assert stats["blocks"] >= 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,use_bytes", [(2, True), (3, False)])
def test_gloo_ranks_build_one_index(tmp_path, world, use_bytes):
    a, b = str(tmp_path / "host"), str(tmp_path / "dist")
    fb.build_index_host(_docs_for_ranks(), a, **PARAMS)
    mp.spawn(_build_worker, args=(world, _free_port(), b, use_bytes), nprocs=world, join=True)
    _same_files(a, b)


@pytest.mark.parametrize("kind,corpus_mib,doc_mib,piece_mib", [("bytes", 1, 1, 16384), ("acgt", 1, 1, 16384),
                                                               ("english", 2, 1, 16384), ("english", 7, 3, 2),
                                                               ("english", 5, 1, 2), ("english", 3, 4, 2)])
def test_bench_sharded_text_is_the_single_gpu_corpus(kind, corpus_mib, doc_mib, piece_mib):
    """bench.py --parallelism sharded builds from ByteText; the symbols must be those of the text the
    single-GPU path (ensure_index -> build_index_gpu) indexes under the same cache name -- also when the
    English-like corpus is several generator streams and documents straddle them -- and the sampled
    patterns must come from inside documents."""
    import argparse
    import bench
    args = argparse.Namespace(kind=kind, corpus_mib=corpus_mib, seed=2, doc_mib=doc_mib, chunk_size=2048,
                              english_piece_mib=piece_mib, npats=64, plen=32, patterns="text")
    B, sample = bench.sharded_text(args, torch.device("cpu"))
    plain = bench.corpus_tensor(args, torch.device("cpu"))
    assert plain.numel() == corpus_mib << 20
    docs = bench.corpus_docs(args, plain)
    T, ends = build_gpu.prepare_text_gpu(docs)
    assert (B.doc_ends == ends).all() and len(ends) == (1 if kind != "english" else -(-corpus_mib // doc_mib))
    assert (B.slice_symbols(0, B.n + build_gpu.PAD) == T).all()
    if kind == "english" and piece_mib < corpus_mib:      # stream k is seeded seed + k
        second = build_gpu.synthetic_english(1000, 3, "cpu")
        assert (plain[piece_mib << 20: (piece_mib << 20) + 1000] == second).all()
    pats = sample(0, 1)
    assert pats.shape == (64, 32) and not (pats == sample(1, 1)).all()
    whole = bytes(plain.numpy())
    for p in pats[:16]:
        raw = bytes((p - 5).to(torch.uint8).numpy())
        at = whole.find(raw)
        assert at >= 0
        if kind == "english":                               # inside one document
            assert any(whole.find(raw, d * (doc_mib << 20), (d + 1) * (doc_mib << 20)) >= 0
                       for d in range(len(docs)))
