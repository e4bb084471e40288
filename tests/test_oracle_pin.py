"""Pins the oracle (oracle/fm_oracle.c) before anything trusts it.

Three anchors:
  1. the reference's own golden values (src/main/index_test.c:595-725) on the two-document index;
  2. committed fixtures under tests/golden/ that were produced by the unmodified reference
     (tests/golden/make_golden.py) -- these work where /root/reference does not exist;
  3. the live reference (oracle/_ref/libfemto_ref.so), call for call, on every standard corpus:
     exhaustive Occ for every row x symbol on the small ones, count, locate, LF walk.
"""
import json
import os

import numpy as np
import pytest

import corpus
from conftest import GOLDEN_DIR
from oracle.bindings import Oracle, Reference, bytes_to_alpha, have_reference

OFF = 5


def test_reference_golden_values_two_docs(built_indexes):
    """src/main/index_test.c:595-725 (test_construct) replayed through the oracle."""
    with Oracle(built_indexes["two_docs"]) as o:
        n = o.header_info()["total_length"]
        assert n == 24
        assert o.back_step(7)[0] == OFF + ord("n")                 # L[7] == 'n'
        assert o.occ(OFF + ord("n"), 19)[1] == 2                   # Occ('n',19) == 2
        assert o.occ(OFF + ord("e"), 1)[1] == 0                    # Occ('e',1) == 0
        assert o.back_step(8)[0] == OFF + ord("t")                 # L[8] == 't'
        assert o.occ(OFF + ord("t"), 8)[1] == 3                    #   and Occ == 3
        assert o.occ(OFF + ord("t"), 19)[1] == 4                   # Occ('t',19) == 4
        assert o.back_step(19)[2] == 0                             #   row 19 marked with offset 0
        assert o.back_step(21)[2] == -1                            # row 21 unmarked
        assert o.back_step(20)[2] == 10                            # row 20: start of 2nd document
        assert o.resolve(10) == (1, 0)                             # resolves to (doc 1, offset 0)
        assert o.doc_info(1)[0] == 24 - 10
        assert o.C(255) == n                                       # C of an unused high symbol == n
        assert o.C(OFF + ord("e")) == 7
        assert o.C(OFF + ord("t")) == 17


def _golden_cases():
    if not os.path.isdir(GOLDEN_DIR):
        return []
    return sorted(d for d in os.listdir(GOLDEN_DIR) if os.path.exists(os.path.join(GOLDEN_DIR, d, "expected.json")))


@pytest.mark.parametrize("case", _golden_cases())
def test_oracle_against_committed_reference_outputs(case):
    base = os.path.join(GOLDEN_DIR, case)
    exp = json.load(open(os.path.join(base, "expected.json")))
    with Oracle(os.path.join(base, "index")) as o:
        info = o.header_info()
        assert info["total_length"] == exp["total_length"]
        pats = [np.array(p, dtype=np.uint16) for p in exp["patterns"]]
        f, l = o.count(pats)
        assert f.tolist() == exp["first"] and l.tolist() == exp["last"]
        loc = o.locate(pats, exp["max_occs"])
        assert [x.tolist() for x in loc] == exp["locate"]
        steps = [list(o.back_step(r)) for r in range(info["total_length"])]
        assert steps == exp["back_step"]
        for ch, row, val in exp["occ_samples"]:
            assert o.occ(ch, row)[0] == val
        assert o.locate_range(0, info["total_length"] - 1).tolist() == exp["sa"]


needs_ref = pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built (no /root/reference here)")


@needs_ref
@pytest.mark.parametrize("name", ["two_docs", "gen400_big_buckets", "gen400_small_buckets", "gen400_small_blocks",
                                  "gen13_small_blocks", "gen3", "single_symbol", "multi_doc_mixed"])
def test_oracle_vs_reference_exhaustive(name, built_indexes, corpora):
    docs, _ = corpora[name]
    path = built_indexes[name]
    with Oracle(path) as o, Reference(path) as r:
        assert o.header_info() == r.header_info()
        n = o.header_info()["total_length"]
        for ch in list(range(261)) + [261]:
            assert o.C(ch) == r.C(ch)
        step = 1 if n <= 500 else 7
        for row in range(0, n, step):
            assert o.back_step(row) == r.back_step(row), row
            for ch in range(261):
                assert o.occ(ch, row) == r.occ(ch, row), (ch, row)
        for d in range(len(docs)):
            assert o.doc_info(d) == r.doc_info(d)
            got = bytes((o.extract(d) - OFF).astype(np.uint8))
            assert got == docs[d]
        for off in range(0, n, max(1, n // 50)):
            assert o.resolve(off) == r.resolve(off)
        assert (o.locate_range(0, n - 1) == r.locate_range(0, n - 1)).all()


@needs_ref
@pytest.mark.parametrize("name", ["acgt_64k", "bytes_200k", "skewed_deep", "english_100k", "multi_doc_mixed"])
def test_oracle_vs_reference_queries(name, built_indexes, corpora):
    docs, _ = corpora[name]
    path = built_indexes[name]
    pats = corpus.sample_patterns(docs, 400, [1, 2, 3, 4, 6, 8, 12, 16, 24, 32], seed=21)
    pats += [np.zeros(0, dtype=np.uint16), np.array([2], dtype=np.uint16), np.array([260, 260], dtype=np.uint16)]
    with Oracle(path) as o, Reference(path) as r:
        of, ol = o.count(pats)
        rf, rl = r.count(pats)
        assert (of == rf).all() and (ol == rl).all()
        for i, p in enumerate(pats):
            c = corpus.brute_count(docs, p)
            if c >= 0:
                assert max(ol[i] - of[i] + 1, 0) == c
        sub = pats[:120]
        for max_occs in (1, 3, 1000):
            ol_ = o.locate(sub, max_occs)
            rl_ = r.locate(sub, max_occs)
            assert all((a == b).all() for a, b in zip(ol_, rl_))
        n = o.header_info()["total_length"]
        rng = np.random.default_rng(5)
        for row in rng.integers(0, n, 300):
            assert o.back_step(int(row)) == r.back_step(int(row))
        for row, ch in zip(rng.integers(0, n, 1500), rng.integers(0, 261, 1500)):
            assert o.occ(int(ch), int(row)) == r.occ(int(ch), int(row))


@needs_ref
def test_bseq_rank_vs_reference_all_positions():
    """As src/main/wtree_test.c:440-582: every position, three segment modes."""
    rng = np.random.default_rng(3)
    inputs = [rng.integers(0, 2, 3000), np.zeros(2000, dtype=np.uint8), np.ones(2000, dtype=np.uint8),
              np.unpackbits(np.full(200, 0x55, dtype=np.uint8)), np.unpackbits(np.full(200, 0x11, dtype=np.uint8)),
              (rng.random(20000) < 0.02).astype(np.uint8)]
    for bits in inputs:
        for force in (0, 1, -1):
            z = Reference.bseq_construct(bits, force)
            step = 1 if len(bits) <= 3000 else 17
            for idx in range(1, len(bits) + 1, step):
                assert Oracle.bseq_rank(z, idx) == Reference.bseq_rank(z, idx), (len(bits), force, idx)


def test_oracle_counts_match_brute_force(built_indexes, corpora):
    for name in ("multi_doc_mixed", "english_100k", "acgt_64k"):
        docs, _ = corpora[name]
        pats = corpus.sample_patterns(docs, 150, [1, 2, 3, 5, 8, 13], seed=2)
        with Oracle(built_indexes[name]) as o:
            f, l = o.count(pats)
            loc = o.locate(pats, 10 ** 6)
            for i, p in enumerate(pats):
                assert max(l[i] - f[i] + 1, 0) == corpus.brute_count(docs, p)
                assert sorted(loc[i].tolist()) == corpus.brute_locate(docs, p)
