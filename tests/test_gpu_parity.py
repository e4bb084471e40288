"""Parity tests proper: the CUDA path (through the C ABI) against the oracle on the same index and
the same seeded inputs.  Bit-exact: every first/last row, every located offset (in row order),
every L symbol, LF target and SA sample.  Run on the B200 box with `-m gpu`."""
import json
import os

import ctypes as C

import numpy as np
import pytest

import corpus
import femto_b200 as fb
from conftest import GOLDEN_DIR
from oracle.bindings import Oracle, Reference, have_reference

pytestmark = pytest.mark.gpu

ALL = ["two_docs", "gen400_big_buckets", "gen400_small_buckets", "gen400_small_blocks", "gen13_small_blocks",
       "gen3", "single_symbol", "multi_doc_mixed", "acgt_64k", "bytes_200k", "skewed_deep", "english_100k"]


def open_with_layout(built_indexes, block_bytes, levels):
    """Open every test index with `levels` wavelet-tree levels per rank block of block_bytes."""
    from femto_b200 import _lib
    lib = _lib.load()
    assert lib.fm_set_default_block_bytes(block_bytes) == 0
    assert lib.fm_set_default_levels_per_block(levels) == 0
    try:
        opened = {name: fb.Index(path, device=0) for name, path in built_indexes.items()}
    finally:
        lib.fm_set_default_block_bytes(0)
        lib.fm_set_default_levels_per_block(0)
    for ix in opened.values():
        assert ix.info.rank_block_size == block_bytes and ix.info.levels_per_block == levels
    return opened


@pytest.fixture(scope="module")
def gpu_indexes(built_indexes):
    """One wavelet-tree level per 128-byte rank block (the layout the schedule matrix below covers)."""
    opened = open_with_layout(built_indexes, 128, 1)
    yield opened
    for ix in opened.values():
        ix.close()


def edge_patterns():
    return [np.zeros(0, dtype=np.uint16),                      # plen == 0 -> [0, n-1]
            np.array([2], dtype=np.uint16),                    # the SEOF symbol itself
            np.array([1], dtype=np.uint16), np.array([3, 4], dtype=np.uint16),   # unused escape codes
            np.array([260], dtype=np.uint16), np.array([260, 260], dtype=np.uint16),
            np.array([5], dtype=np.uint16), np.array([5, 5, 5], dtype=np.uint16)]


@pytest.fixture(scope="module")
def gpu_indexes_small_blocks(built_indexes):
    """The same indexes loaded with 64- and 32-byte rank blocks."""
    opened = {}
    for bb in (64, 32):
        for name, ix in open_with_layout(built_indexes, bb, 1).items():
            opened[(name, bb)] = ix
    yield opened
    for ix in opened.values():
        ix.close()


@pytest.mark.parametrize("cfg", [(64, 1, 1), (64, 1, 2), (64, 1, 4), (64, 0, 2), (64, 0, 4),
                                 (32, 1, 1), (32, 1, 2), (32, 0, 2)])
@pytest.mark.parametrize("name", ALL)
def test_count_small_rank_blocks(name, cfg, gpu_indexes_small_blocks, built_indexes, corpora):
    bb, merged, lanes = cfg
    docs, _ = corpora[name]
    ix = gpu_indexes_small_blocks[(name, bb)]
    ix.set_count_schedule(merged, lanes)
    pats = corpus.sample_patterns(docs, 1200, [1, 2, 3, 4, 5, 6, 8, 12, 16, 24, 32, 64], seed=33) + edge_patterns()
    with Oracle(built_indexes[name]) as o:
        of, ol = o.count(pats)
    f, l = ix.count(pats)
    assert (f == of).all() and (l == ol).all()


@pytest.mark.parametrize("bb", [64, 32])
@pytest.mark.parametrize("name", ["multi_doc_mixed", "skewed_deep", "english_100k", "gen400_small_blocks"])
def test_locate_extract_small_rank_blocks(name, bb, gpu_indexes_small_blocks, built_indexes, corpora):
    docs, _ = corpora[name]
    ix = gpu_indexes_small_blocks[(name, bb)]
    pats = corpus.sample_patterns(docs, 200, [1, 2, 3, 4, 6, 8, 16], seed=43)
    with Oracle(built_indexes[name]) as o:
        want = o.locate(pats, 50)
        n = o.header_info()["total_length"]
        rows = np.arange(n) if n <= 3000 else np.random.default_rng(1).integers(0, n, 2000)
        ch, nxt, off = ix.back_step(rows)
        for i, r in enumerate(rows):
            assert (int(ch[i]), int(nxt[i]), int(off[i])) == o.back_step(int(r))
    got = ix.locate(pats, 50)
    assert all((a == b).all() for a, b in zip(got, want))
    for d in range(len(docs)):
        assert bytes((ix.extract(d) - fb.CHARACTER_OFFSET).astype(np.uint8)) == docs[d]


@pytest.mark.parametrize("sched", [(1, 4), (1, 2), (1, 8), (0, 4), (0, 8)])
@pytest.mark.parametrize("name", ALL)
def test_count_matches_oracle(name, sched, gpu_indexes, built_indexes, corpora):
    docs, _ = corpora[name]
    ix = gpu_indexes[name]
    ix.set_count_schedule(*sched)
    pats = corpus.sample_patterns(docs, 1500, [1, 2, 3, 4, 5, 6, 8, 12, 16, 24, 32, 64], seed=31) + edge_patterns()
    with Oracle(built_indexes[name]) as o:
        of, ol = o.count(pats)
    f, l = ix.count(pats)                                    # reference-shaped pointer-array call
    assert (f == of).all() and (l == ol).all()
    plen, flat, offs = fb.flatten_patterns(pats)
    f2, l2 = ix.count_flat(plen, flat, offs)                 # flat host-buffer call
    assert (f2 == of).all() and (l2 == ol).all()
    for i in range(0, len(pats), 97):
        c = corpus.brute_count(docs, pats[i])
        if c >= 0:
            assert max(l[i] - f[i] + 1, 0) == c
    ix.set_count_schedule(1, 4)


@pytest.mark.parametrize("name", ALL)
def test_occ_and_back_step_match_oracle(name, gpu_indexes, built_indexes):
    ix = gpu_indexes[name]
    with Oracle(built_indexes[name]) as o:
        n = o.header_info()["total_length"]
        rng = np.random.default_rng(4)
        if n <= 500:
            rows = np.repeat(np.arange(n), 261)
            chs = np.tile(np.arange(261), n)
        else:
            rows = rng.integers(0, n, 5000)
            chs = rng.integers(0, 261, 5000)
        got = ix.occ(chs, rows)
        want = np.array([o.occ(int(c), int(r))[0] for c, r in zip(chs, rows)], dtype=np.int64)
        assert (got == want).all()
        srows = np.arange(n) if n <= 3000 else np.concatenate([rng.integers(0, n, 3000), [0, n - 1]])
        ch, nxt, off = ix.back_step(srows)
        for i, r in enumerate(srows):
            assert (int(ch[i]), int(nxt[i]), int(off[i])) == o.back_step(int(r)), r


@pytest.mark.parametrize("lanes", [4, 8])
@pytest.mark.parametrize("name", ALL)
def test_locate_matches_oracle(name, lanes, gpu_indexes, built_indexes, corpora):
    docs, _ = corpora[name]
    ix = gpu_indexes[name]
    ix.set_lanes_per_query(lanes)
    pats = corpus.sample_patterns(docs, 300, [1, 2, 3, 4, 6, 8, 16, 32], seed=41) + edge_patterns()[1:]
    with Oracle(built_indexes[name]) as o:
        n = o.header_info()["total_length"]
        for max_occs in (1, 2, 7, 100000):                   # incl. the '>' clip quirk (server.c:4411)
            want = o.locate(pats, max_occs)
            got = ix.locate(pats, max_occs)
            assert all((a == b).all() for a, b in zip(got, want)), max_occs
        if n <= 70000:
            assert (ix.locate_range(0, n - 1) == o.locate_range(0, n - 1)).all()
        else:
            assert (ix.locate_range(1000, 9000) == o.locate_range(1000, 9000)).all()
    # every located offset really is an occurrence
    got = ix.locate(pats[:60], 50)
    for p, offs_ in zip(pats[:60], got):
        truth = set(corpus.brute_locate(docs, p)) if (p >= 5).all() and len(p) else None
        if truth is not None:
            assert set(offs_.tolist()) <= truth
    ix.set_lanes_per_query(4)


@pytest.mark.parametrize("name", ALL)
def test_extract_and_doc_tables(name, gpu_indexes, built_indexes, corpora):
    docs, _ = corpora[name]
    ix = gpu_indexes[name]
    with Oracle(built_indexes[name]) as o:
        n = o.header_info()["total_length"]
        for d in range(len(docs)):
            assert ix.doc_info(d) == o.doc_info(d)
            got = bytes((ix.extract(d) - fb.CHARACTER_OFFSET).astype(np.uint8))
            assert got == docs[d]
        offs_ = np.arange(0, n, max(1, n // 200), dtype=np.int64)
        doc, doff = ix.resolve(offs_)
        for i, off in enumerate(offs_):
            assert (int(doc[i]), int(doff[i])) == o.resolve(int(off))


@pytest.mark.parametrize("name", ["multi_doc_mixed", "english_100k", "bytes_200k"])
def test_backward_step_api_reproduces_count(name, gpu_indexes, built_indexes, corpora):
    """fm_backward_step (the reference's backward_search_query, one step per call) iterated over a
    pattern must walk through exactly the ranges of the oracle's backward search."""
    docs, _ = corpora[name]
    ix = gpu_indexes[name]
    pats = [p for p in corpus.sample_patterns(docs, 400, [6], seed=61) if len(p) == 6]
    with Oracle(built_indexes[name]) as o:
        first = np.array([o.C(int(p[-1])) for p in pats], dtype=np.int64)
        last = np.array([o.C(int(p[-1]) + 1) - 1 for p in pats], dtype=np.int64)
        alive = np.ones(len(pats), dtype=bool)
        for k in range(4, -1, -1):
            alive &= first <= last
            idx = np.nonzero(alive)[0]
            if len(idx) == 0:
                break
            ch = np.array([pats[i][k] for i in idx], dtype=np.uint16)
            nf, nl = ix.backward_step(first[idx], last[idx], ch)
            first[idx], last[idx] = nf, nl
        of, ol = o.count(pats)
    assert (first == of).all() and (last == ol).all()


def test_golden_reference_outputs():
    """Committed answers of the unmodified reference on indexes built by the reference."""
    for case in sorted(os.listdir(GOLDEN_DIR)):
        base = os.path.join(GOLDEN_DIR, case)
        if not os.path.exists(os.path.join(base, "expected.json")):
            continue
        exp = json.load(open(os.path.join(base, "expected.json")))
        with fb.Index(os.path.join(base, "index")) as ix:
            pats = [np.array(p, dtype=np.uint16) for p in exp["patterns"]]
            f, l = ix.count(pats)
            assert f.tolist() == exp["first"] and l.tolist() == exp["last"]
            assert [x.tolist() for x in ix.locate(pats, exp["max_occs"])] == exp["locate"]
            n = exp["total_length"]
            ch, nxt, off = ix.back_step(np.arange(n))
            assert [[int(a), int(b), int(c)] for a, b, c in zip(ch, nxt, off)] == exp["back_step"]
            assert ix.locate_range(0, n - 1).tolist() == exp["sa"]
            occ = np.array(exp["occ_samples"])
            assert ix.occ(occ[:, 0], occ[:, 1]).tolist() == occ[:, 2].tolist()
            # femto's generic requests (femto.h:75-149): the response text, character for character
            for req, want in exp["generic_requests"].items():
                assert ix.generic_request(req) == want, req
            with pytest.raises(fb.FemtoError):
                ix.generic_request("no_such_request 1 2")
            with pytest.raises(fb.FemtoError):
                ix.generic_request("string_rows 1 x")


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref did not travel")
def test_live_reference_on_gpu_box(gpu_indexes, built_indexes, corpora):
    """The unmodified reference, run on this box, against the GPU on the same index."""
    for name in ("multi_doc_mixed", "english_100k", "bytes_200k"):
        docs, _ = corpora[name]
        pats = corpus.sample_patterns(docs, 500, [2, 4, 8, 16, 32], seed=51)
        with Reference(built_indexes[name]) as r:
            rf, rl = r.count(pats)
            rloc = r.locate(pats[:100], 20)
        f, l = gpu_indexes[name].count(pats)
        assert (f == rf).all() and (l == rl).all()
        assert all((a == b).all() for a, b in zip(gpu_indexes[name].locate(pats[:100], 20), rloc))


def test_flattened_container_and_reopen(built_indexes, corpora, tmp_path):
    docs, _ = corpora["acgt_64k"]
    flat = str(tmp_path / "acgt.femto")
    fb.flatten(built_indexes["acgt_64k"], flat)
    pats = corpus.sample_patterns(docs, 400, [4, 8, 12], seed=61)
    with fb.Index(built_indexes["acgt_64k"]) as a, fb.Index(flat) as b:
        fa, la = a.count(pats)
        fbb, lb = b.count(pats)
        assert (fa == fbb).all() and (la == lb).all()
        assert a.kernel_launches() > 0


def test_device_pointer_api_with_torch(built_indexes, corpora):
    """fm_count_device: inputs and outputs stay in HBM (what bench.py's `value` times)."""
    import torch
    docs, _ = corpora["english_100k"]
    pats = corpus.sample_patterns(docs, 5000, [8, 16, 32], seed=71)
    plen, flat, offs = fb.flatten_patterns(pats)
    with fb.Index(built_indexes["english_100k"]) as ix, Oracle(built_indexes["english_100k"]) as o:
        d_plen = torch.from_numpy(plen).cuda()
        d_flat = torch.from_numpy(flat.view(np.int16)).cuda()
        d_offs = torch.from_numpy(offs).cuda()
        d_first = torch.empty(len(pats), dtype=torch.int64, device="cuda")
        d_last = torch.empty(len(pats), dtype=torch.int64, device="cuda")
        s = torch.cuda.current_stream().cuda_stream
        ix.count_device(len(pats), d_plen.data_ptr(), d_flat.data_ptr(), d_offs.data_ptr(), d_first.data_ptr(),
                        d_last.data_ptr(), s)
        torch.cuda.synchronize()
        of, ol = o.count(pats)
        assert (d_first.cpu().numpy() == of).all() and (d_last.cpu().numpy() == ol).all()


def test_mixed_length_zipf_batch(tmp_path):
    """BASELINE configs[3] in miniature: lengths Zipf-distributed over [8,256] on an English-like
    multi-document corpus (deep Huffman trees, patterns dying at different steps, groups of a warp
    finishing at different times)."""
    docs = [corpus.english_like(200000, 200 + d) for d in range(5)]
    path = str(tmp_path / "english_1m")
    fb.build_index_host(docs, path, block_size=1 << 18, bucket_size=1 << 14, chunk_size=2048)
    rng = np.random.default_rng(17)
    ranks = np.arange(8, 257)
    p = (1.0 / (ranks - 7)) / (1.0 / (ranks - 7)).sum()
    lengths = rng.choice(ranks, 20000, p=p).tolist()
    pats = corpus.sample_patterns(docs, 20000, lengths, seed=18, random_fraction=0.2)
    # mutate a symbol in a quarter of the patterns so that they die somewhere in the middle
    for k in range(0, len(pats), 4):
        q = pats[k].copy()
        q[int(rng.integers(0, len(q)))] = 5 + int(rng.integers(97, 123))
        pats[k] = q
    with fb.Index(path) as ix, Oracle(path) as o:
        for lanes in (1, 2, 1):                              # the default image, both quad schedules
            ix.set_count_schedule(1, lanes)
            f, l = ix.count(pats)
            sub = list(range(0, len(pats), 7))
            of, ol = o.count([pats[i] for i in sub])
            assert (f[sub] == of).all() and (l[sub] == ol).all()
        cnt = np.maximum(l - f + 1, 0)
        assert (cnt[1::4] >= 1).sum() > 3000                 # text-sampled ones are found
        loc = ix.locate(pats[:2000], 30)
        oloc = o.locate(pats[:2000:5], 30)
        assert all((a == b).all() for a, b in zip(loc[::5], oloc))


def test_properties_at_scale(tmp_path):
    """Size-independent properties on a corpus too large for the scalar oracle to sweep:
    count == occurrences found by locate, every located offset matches the text, LF is a
    permutation walk that reproduces the document (extract), ranges nest when a pattern grows."""
    docs = [corpus.random_bytes(1 << 20, 101), corpus.random_acgt(1 << 20, 102)]
    path = str(tmp_path / "scale")
    fb.build_index_host(docs, path, block_size=1 << 20, bucket_size=1 << 16, chunk_size=0)
    text = b"".join(d + b"\x00" for d in docs)    # offsets only; SEOF positions never match a pattern
    pats = corpus.sample_patterns(docs, 20000, [3, 4, 6, 8, 12, 20, 32], seed=81, random_fraction=0.1)
    with fb.Index(path) as ix:
        f, l = ix.count(pats)
        cnt = np.maximum(l - f + 1, 0)
        loc = ix.locate(pats[:3000], 64)
        for p, c, offs_ in zip(pats[:3000], cnt[:3000], loc):
            assert len(offs_) == min(int(c), 64) or (int(c) == 65 and len(offs_) == 65)
            pb = bytes((p - 5).astype(np.uint8))
            for off in offs_:
                assert text[off:off + len(pb)] == pb
            assert len(set(offs_.tolist())) == len(offs_)
        # nesting: the range of a suffix of the pattern contains ... (backward search narrows)
        shorter = [p[1:] for p in pats[:2000]]
        fs, ls = ix.count(shorter)
        assert (np.maximum(ls - fs + 1, 0) >= cnt[:2000]).all()
        # SA is a permutation on a sampled range, and offsets agree with the suffix order
        sa = ix.locate_range(5000, 9000)
        assert len(set(sa.tolist())) == len(sa)
        for d in range(2):
            got = bytes((ix.extract(d) - 5).astype(np.uint8))
            assert got == docs[d]


# ---- paired- and quad-level wavelet blocks (2 / 4 tree levels per HBM read; fm_image.hpp) ------------
@pytest.fixture(scope="module")
def gpu_indexes_multilevel(built_indexes):
    opened = {}
    for bb, levels in ((128, 2), (64, 2), (128, 4)):
        for name, ix in open_with_layout(built_indexes, bb, levels).items():
            opened[(name, bb, levels)] = ix
    yield opened
    for ix in opened.values():
        ix.close()


@pytest.mark.parametrize("cfg", [(128, 2, 4), (128, 2, 2), (128, 2, 1), (64, 2, 2), (64, 2, 1), (128, 4, 2), (128, 4, 1)])
@pytest.mark.parametrize("name", ALL)
def test_count_multilevel_blocks(name, cfg, gpu_indexes_multilevel, built_indexes, corpora):
    bb, levels, lanes = cfg
    docs, _ = corpora[name]
    ix = gpu_indexes_multilevel[(name, bb, levels)]
    ix.set_count_schedule(1, lanes)
    pats = corpus.sample_patterns(docs, 1500, [1, 2, 3, 4, 5, 6, 8, 12, 16, 24, 32, 64], seed=35) + edge_patterns()
    with Oracle(built_indexes[name]) as o:
        of, ol = o.count(pats)
    f, l = ix.count(pats)
    assert (f == of).all() and (l == ol).all()
    plen, flat, offs = fb.flatten_patterns(pats)
    f2, l2 = ix.count_flat(plen, flat, offs)
    assert (f2 == of).all() and (l2 == ol).all()


@pytest.mark.parametrize("cfg", [(128, 2, 4), (128, 2, 2), (64, 2, 2), (64, 2, 1), (128, 4, 2)])
@pytest.mark.parametrize("name", ALL)
def test_walks_multilevel_blocks(name, cfg, gpu_indexes_multilevel, built_indexes, corpora):
    """occ / back_step / locate / extract over paired- and quad-level blocks, every lane configuration."""
    bb, levels, lanes = cfg
    docs, _ = corpora[name]
    ix = gpu_indexes_multilevel[(name, bb, levels)]
    ix.set_lanes_per_query(lanes)
    pats = corpus.sample_patterns(docs, 200, [1, 2, 3, 4, 6, 8, 16], seed=45) + edge_patterns()[1:]
    with Oracle(built_indexes[name]) as o:
        n = o.header_info()["total_length"]
        rng = np.random.default_rng(6)
        if n <= 500:
            rows = np.repeat(np.arange(n), 261)
            chs = np.tile(np.arange(261), n)
        else:
            rows = rng.integers(0, n, 5000)
            chs = rng.integers(0, 261, 5000)
        got = ix.occ(chs, rows)
        want = np.array([o.occ(int(c), int(r))[0] for c, r in zip(chs, rows)], dtype=np.int64)
        assert (got == want).all()
        srows = np.arange(n) if n <= 3000 else np.concatenate([rng.integers(0, n, 2000), [0, n - 1]])
        ch, nxt, off = ix.back_step(srows)
        for i, r in enumerate(srows):
            assert (int(ch[i]), int(nxt[i]), int(off[i])) == o.back_step(int(r)), r
        want_loc = o.locate(pats, 50)
    got_loc = ix.locate(pats, 50)
    assert all((a == b).all() for a, b in zip(got_loc, want_loc))
    for d in range(len(docs)):
        assert bytes((ix.extract(d) - fb.CHARACTER_OFFSET).astype(np.uint8)) == docs[d]
    ix.set_lanes_per_query(4)


def test_default_layout_and_probe(built_indexes):
    """fm_open without tuning calls builds a multi-level image; the random-read probe runs."""
    ix = fb.Index(built_indexes["bytes_200k"], device=0)
    assert (ix.info.levels_per_block, ix.info.rank_block_size) in ((2, 64), (4, 128))
    for b in (32, 64, 128):
        r = ix.probe_random_reads(b, steps=50)
        assert r["accesses"] > 0 and r["ms"] > 0
    with pytest.raises(Exception):
        ix.probe_random_reads(48, steps=10)
    ix.close()


# ---- streamed host-buffer batches (count kernel launched while the patterns are still arriving) ----
@pytest.mark.parametrize("uniform", [True, False])
def test_streamed_batch_equals_plain_batch(uniform, built_indexes, corpora):
    """Batches of >= 128 Ki ordered patterns take the streamed path of fm_count_flat (two kernels gated
    by arrival counters, uniform-length batches without plen / offs transfers).  Results must equal the
    plain path (same patterns addressed out of order, which disables streaming) and the oracle."""
    name = "english_100k"
    docs, _ = corpora[name]
    n = 300000 + 77                                           # several chunks per half, ragged tail
    lengths = [12] if uniform else [1, 2, 3, 5, 8, 13, 21, 34]
    base = corpus.sample_patterns(docs, 3000, lengths, seed=61, random_fraction=0.2)
    rng = np.random.default_rng(62)
    pick = rng.integers(0, len(base), n)
    pats = [base[i] for i in pick]
    plen, flat, offs = fb.flatten_patterns(pats)
    with fb.Index(built_indexes[name], device=0) as ix, Oracle(built_indexes[name]) as o:
        of, ol = o.count(base)
        for _ in range(2):                                    # twice: device buffers are reused between calls
            f, l = ix.count_flat(plen, flat, offs)
            assert (f == of[pick]).all() and (l == ol[pick]).all()
        # the same patterns, flat buffer reversed pattern by pattern: not "in order" -> plain path
        order = np.arange(n)[::-1]
        rflat = np.concatenate([pats[i] for i in order]).astype(np.uint16)
        roffs = np.zeros(n, dtype=np.int64)
        roffs[order] = np.concatenate([[0], np.cumsum(plen[order][:-1])])
        f2, l2 = ix.count_flat(plen, rflat, roffs)
        assert (f2 == f).all() and (l2 == l).all()
        # last == NULL: counts in first (parallel_count, femto.c:313-318), streamed as well
        cnt = np.empty(n, dtype=np.int64)
        rc = ix.lib.fm_count_flat(ix.h, n, fb._ptr(plen, C.c_int32), fb._ptr(flat, C.c_uint16), fb._ptr(offs, C.c_int64),
                                  fb._ptr(cnt, C.c_int64), None)
        assert rc == 0 and (cnt == l - f + 1).all()


# ---- documents: info bytes and the documents of a row range ----------------
def test_doc_names_and_range_documents(tmp_path):
    """fm_doc_name returns the bytes stored at build time (document_info, index.c:1767-1784);
    fm_range_documents = the ascending set of documents of SA[first..last] (range_to_results for
    documents, server.c:4549-4889), checked against the oracle's locate + its document table."""
    docs = [corpus.english_like(3000, 300 + d) for d in range(7)] + [b"", b"zzzz the end"]
    names = [f"file://corpus/doc_{d:03d}.txt".encode() for d in range(len(docs))]
    names[3] = b""                                           # a document without info bytes
    path = str(tmp_path / "named")
    fb.build_index_host(docs, path, doc_infos=names, block_size=4096, bucket_size=512, chunk_size=256)
    with fb.Index(path) as ix, Oracle(path) as o:
        assert [ix.doc_name(d) for d in range(len(docs))] == names
        n = o.header_info()["total_length"]
        sa = o.locate_range(0, n - 1)
        doc_of = np.array([o.resolve(int(x))[0] for x in sa], dtype=np.int64)
        for first, last in [(0, n - 1), (0, 0), (n - 1, n - 1), (17, 16), (100, 180), (2000, 2600), (n - 40, n - 1)]:
            want = np.unique(doc_of[first:last + 1]) if last >= first else np.zeros(0, dtype=np.int64)
            got = ix.range_documents(first, last)
            assert (got == want).all() and len(got) == len(want), (first, last)
        # a pattern's documents: count -> range -> documents
        pats = corpus.sample_patterns(docs, 40, [2, 3, 5], seed=5)
        f, l = ix.count(pats)
        for p, a, b in zip(pats, f, l):
            pb = bytes((p - 5).astype(np.uint8))
            want = [d for d, text in enumerate(docs) if pb in text]
            assert ix.range_documents(int(a), int(b)).tolist() == want


# ---- BASELINE configs[3] at a size the reference still opens quickly ------------------------------
def test_mixed_length_zipf_on_64mib_english_like_corpus(tmp_path):
    """Mixed-length batch (8-256 symbols, Zipf) on an English-like multi-document corpus of 64 MiB,
    index built by the GPU pipeline with DEFAULT parameters (1 Mi-row buckets: Huffman trees of
    depth 3-15, three quad rounds for the rare symbols).  A sample is checked against the oracle
    (and the live reference where it travelled); the whole batch through count == locate sizes and
    located offsets matching the text."""
    import torch
    from femto_b200 import build_gpu
    ndocs, per = 16, 4 << 20
    docs = [corpus.english_like(per, 900 + d) for d in range(ndocs)]
    path = str(tmp_path / "english_64m")
    dev = torch.device("cuda", 0)
    t = build_gpu.build_index_gpu([torch.frombuffer(bytearray(d), dtype=torch.uint8).to(dev) for d in docs], path)
    assert t["rows"] == ndocs * (per + 1)
    rng = np.random.default_rng(23)
    ranks = np.arange(8, 257)
    p = (1.0 / (ranks - 7)) / (1.0 / (ranks - 7)).sum()
    lengths = rng.choice(ranks, 200000, p=p).tolist()
    pats = corpus.sample_patterns(docs, 200000, lengths, seed=24, random_fraction=0.1)
    for k in range(0, len(pats), 5):                        # a fifth dies somewhere in the middle
        q = pats[k].copy()
        q[int(rng.integers(0, len(q)))] = 5 + int(rng.integers(65, 91))
        pats[k] = q
    plen, flat, offs = fb.flatten_patterns(pats)
    with fb.Index(path) as ix:
        f, l = ix.count_flat(plen, flat, offs)              # streamed path, ragged lengths
        sub = list(range(0, len(pats), 997))
        with Oracle(path) as o:
            of, ol = o.count([pats[i] for i in sub])
        assert (f[sub] == of).all() and (l[sub] == ol).all()
        if have_reference():
            with Reference(path) as r:
                rf, rl = r.count([pats[i] for i in sub[:60]])
            assert (f[sub[:60]] == rf).all() and (l[sub[:60]] == rl).all()
        cnt = np.maximum(l - f + 1, 0)
        assert (cnt[1::5] >= 1).mean() > 0.85               # unmutated and (90 %) text-sampled: found
        text = b"".join(d + b"\x00" for d in docs)
        k0 = [i for i in range(1, 4000, 5)]
        loc = ix.locate([pats[i] for i in k0], 20)
        for i, offs_ in zip(k0, loc):
            assert len(offs_) == min(int(cnt[i]), 20) or (int(cnt[i]) == 21 and len(offs_) == 21)
            pb = bytes((pats[i] - 5).astype(np.uint8))
            for off in offs_:
                assert text[off:off + len(pb)] == pb


def test_streamed_batch_lazy_validation_falls_back(built_indexes, corpora):
    """Large batches are validated chunk by chunk behind the running kernel, on the shape their first
    and last pattern claim.  Batches that break the claim half-way must be caught before anything is
    read out of bounds and still give the reference's answers (or the reference's error)."""
    name = "english_100k"
    docs, _ = corpora[name]
    n = 200000
    base = corpus.sample_patterns(docs, 2000, [12], seed=71, random_fraction=0.2)
    rng = np.random.default_rng(72)
    pick = rng.integers(0, len(base), n)
    pats = [base[i] for i in pick]
    with fb.Index(built_indexes[name], device=0) as ix, Oracle(built_indexes[name]) as o:
        of, ol = o.count(base)
        want_f, want_l = of[pick].copy(), ol[pick].copy()
        # (a) claims "all of length 12" (first two patterns, total size) but two patterns in the middle are 11 and 13
        a_pats = list(pats)
        a_pats[n // 2] = a_pats[n // 2][:11]
        a_pats[n // 2 + 1] = np.concatenate([a_pats[n // 2 + 1], a_pats[n // 2 + 1][:1]])
        plen, flat, offs = fb.flatten_patterns(a_pats)
        assert int(plen.sum()) == 12 * n
        f, l = ix.count_flat(plen, flat, offs)
        af, al = o.count([a_pats[n // 2], a_pats[n // 2 + 1]])
        want_a_f, want_a_l = want_f.copy(), want_l.copy()
        want_a_f[n // 2:n // 2 + 2], want_a_l[n // 2:n // 2 + 2] = af, al
        assert (f == want_a_f).all() and (l == want_a_l).all()
        # (b) claims "densely packed" but one unused symbol sits between two patterns in the middle
        plen, flat, offs = fb.flatten_patterns([pats[0][:11]] + pats[1:])   # not uniform: the dense claim is tried
        cut = int(offs[n // 3])
        flat_b = np.concatenate([flat[:cut], np.array([77], dtype=np.uint16), flat[cut:]])
        offs_b = offs.copy()
        offs_b[n // 3:] += 1
        f, l = ix.count_flat(plen, flat_b, offs_b)
        f0, l0 = o.count([pats[0][:11]])
        assert f[0] == f0[0] and l[0] == l0[0] and (f[1:] == want_f[1:]).all() and (l[1:] == want_l[1:]).all()
        # (c) a negative length in the middle: the reference's parameter error, nothing else
        plen_c = plen.copy()
        plen_c[n - 1000] = -3
        with pytest.raises(Exception):
            ix.count_flat(plen_c, flat_b, offs_b)
        f, l = ix.count_flat(plen, flat_b, offs_b)            # the handle is fine afterwards
        assert (f[1:] == want_f[1:]).all()


def test_pointer_array_call_gathers_behind_the_kernel(built_indexes, corpora):
    """fm_count with one pointer per pattern (parallel_count's prototype): the patterns are gathered
    into pinned staging by background threads while the streamed kernel already runs.  Ragged lengths
    incl. empty patterns; must equal the flat call and the oracle."""
    name = "english_100k"
    docs, _ = corpora[name]
    base = corpus.sample_patterns(docs, 3000, [1, 2, 3, 5, 8, 13, 21, 34, 70], seed=81, random_fraction=0.2)
    base += [np.zeros(0, dtype=np.uint16)] * 40
    rng = np.random.default_rng(82)
    n = 180000 + 13
    pick = rng.integers(0, len(base), n)
    pats = [base[i] for i in pick]
    plen, flat, offs = fb.flatten_patterns(pats)
    with fb.Index(built_indexes[name], device=0) as ix, Oracle(built_indexes[name]) as o:
        of, ol = o.count(base)
        for _ in range(2):
            f, l = ix.count(pats)
            assert (f == of[pick]).all() and (l == ol[pick]).all()
        f2, l2 = ix.count_flat(plen, flat, offs)
        assert (f2 == f).all() and (l2 == l).all()


def test_streamed_batch_back_pointing_last_pattern(built_indexes, corpora):
    """A dense prefix whose LAST pattern points back to offset 0: the shape claimed from the first and
    last pattern is a 12-symbol buffer, so every chunk must fail the lazy check (each pattern has to
    end inside the claimed buffer) before the kernel touches symbols that were never copied; the call
    then validates the whole batch and answers it the plain way."""
    name = "english_100k"
    docs, _ = corpora[name]
    n = 150000
    base = corpus.sample_patterns(docs, 1000, [12], seed=81, random_fraction=0.2)
    rng = np.random.default_rng(82)
    pick = rng.integers(0, len(base), n)
    pats = [base[i] for i in pick]
    plen, flat, offs = fb.flatten_patterns(pats)
    offs = offs.copy()
    offs[n - 1] = 0                                           # -> claim_len = 12
    with fb.Index(built_indexes[name], device=0) as ix, Oracle(built_indexes[name]) as o:
        of, ol = o.count(base)
        want_f, want_l = of[pick].copy(), ol[pick].copy()
        want_f[n - 1], want_l[n - 1] = want_f[0], want_l[0]   # the last pattern now IS the first one
        f, l = ix.count_flat(plen, flat, offs)
        assert (f == want_f).all() and (l == want_l).all()
        f, l = ix.count_flat(plen, flat, offs)                # and the handle still streams fine afterwards
        assert (f == want_f).all() and (l == want_l).all()


@pytest.mark.parametrize("uniform", [True, False])
def test_count_bytes_equals_count_flat(uniform, built_indexes, corpora):
    """fm_count_bytes (raw text bytes, widened by the kernel as it reads) against fm_count_flat on the
    same patterns as alpha_t symbols and against the oracle: a small batch (plain path) and a batch of
    >= 128 Ki patterns (streamed path: chunk cuts at 128 bytes = 128 symbols)."""
    name = "english_100k"
    docs, _ = corpora[name]
    lengths = [12] if uniform else [1, 2, 3, 5, 8, 13, 21, 34]
    base = corpus.sample_patterns(docs, 3000, lengths, seed=91, random_fraction=0.2)
    base = [p for p in base if len(p)]                                        # bytes cannot say "empty" differently
    with fb.Index(built_indexes[name], device=0) as ix, Oracle(built_indexes[name]) as o:
        of, ol = o.count(base)
        for n in (len(base), 200000 + 33):
            pick = np.arange(n) if n == len(base) else np.random.default_rng(92).integers(0, len(base), n)
            pats = [base[i] for i in pick]
            plen, flat, offs = fb.flatten_patterns(pats)
            text = (flat[:int(plen.sum())] - fb.CHARACTER_OFFSET).astype(np.uint8)
            f, l = ix.count_bytes(plen, text, offs)
            assert (f == of[pick]).all() and (l == ol[pick]).all()
            f2, l2 = ix.count_flat(plen, flat, offs)
            assert (f2 == f).all() and (l2 == l).all()


def test_count_bytes_other_layouts_widen_on_the_host(built_indexes, corpora):
    name = "acgt_64k"
    docs, _ = corpora[name]
    base = [p for p in corpus.sample_patterns(docs, 500, [2, 7, 16], seed=93) if len(p)]
    plen, flat, offs = fb.flatten_patterns(base)
    text = (flat[:int(plen.sum())] - fb.CHARACTER_OFFSET).astype(np.uint8)
    opened = open_with_layout({name: built_indexes[name]}, 128, 1)
    try:
        with Oracle(built_indexes[name]) as o:
            of, ol = o.count(base)
        f, l = opened[name].count_bytes(plen, text, offs)
        assert (f == of).all() and (l == ol).all()
    finally:
        for ix in opened.values():
            ix.close()


def test_locate_negative_max_occs_gives_no_rows(built_indexes, corpora):
    """max_occs < 0: the reference's clip leaves last < first, i.e. no rows (server.c:4411-4415)."""
    name = "acgt_64k"
    docs, _ = corpora[name]
    pats = corpus.sample_patterns(docs, 50, [3, 5], seed=94, random_fraction=0.0)
    plen, flat, offs = fb.flatten_patterns(pats)
    with fb.Index(built_indexes[name], device=0) as ix:
        noccs, start, out = ix.locate_flat(plen, flat, offs, -5, 16)
        assert (noccs == 0).all()


def test_extract_batch_and_chunkless_range_documents(tmp_path):
    """fm_extract_batch: several documents in one launch equal the documents (and fm_extract one by one);
    fm_range_documents on an index built WITHOUT document chunks locates every row and gives the same sets."""
    docs = [corpus.english_like(4000, 500 + d) for d in range(9)] + [b"", b"x", corpus.random_acgt(2500, 9)]
    for chunk_size in (128, 0):
        path = str(tmp_path / f"idx{chunk_size}")
        fb.build_index_host(docs, path, block_size=4096, bucket_size=1024, chunk_size=chunk_size)
        with fb.Index(path) as ix, Oracle(path) as o:
            order = [5, 0, 11, 9, 10, 3, 3, 8]
            got = ix.extract_batch(order)
            for d, sym in zip(order, got):
                assert bytes((sym - fb.CHARACTER_OFFSET).astype(np.uint8)) == docs[d]
                assert (sym == ix.extract(d)).all()
            n = o.header_info()["total_length"]
            sa = o.locate_range(0, n - 1)
            doc_of = np.array([o.resolve(int(x))[0] for x in sa], dtype=np.int64)
            for first, last in [(0, n - 1), (5, 5), (100, 700), (n - 300, n - 1)]:
                assert (ix.range_documents(first, last) == np.unique(doc_of[first:last + 1])).all()
