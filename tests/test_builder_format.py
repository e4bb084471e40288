"""The index emitter (femto_b200/csrc/fm_builder.cc) must write the reference's format byte for byte.

Anchors:
  * bseq level: our encoder vs the reference's bseq_construct_forcetype (src/main/wtree.c:364) on
    the inputs of src/main/wtree_test.c:440-582 (fixed strings, random, all-0, all-1, 0x55, 0x11)
    in all three segment modes, and our decoder round-trips them;
  * index level: our builder vs (a) the reference's in-memory builder (qsufsort + index_documents)
    and (b) the reference's real builder femto_index (difference-cover sort, document chunks),
    compared file by file;
  * committed golden indexes (built by the reference) are reproduced from their documents.
"""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

import corpus
import femto_b200 as fb
from conftest import GOLDEN_DIR
from femto_b200 import _lib
from oracle.bindings import REF_SO, Reference, have_reference

needs_ref = pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built (no /root/reference here)")
FEMTO_INDEX = os.path.join(os.path.dirname(REF_SO), "femto_index")


def our_encode(bits, force=0):
    lib = _lib.load()
    packed = np.packbits(np.asarray(bits, dtype=np.uint8)).tobytes()
    out, n = C.c_void_p(), C.c_int64()
    assert lib.fm_debug_bseq_encode(packed, len(bits), force, C.byref(out), C.byref(n)) == 0
    z = C.string_at(out, n.value)
    lib.fm_debug_free(out)
    return z


def our_expand(z, nbits):
    lib = _lib.load()
    buf = C.create_string_buffer((nbits + 7) // 8 + 8)
    n = C.c_int64()
    assert lib.fm_debug_bseq_expand(z, len(z), buf, nbits, C.byref(n)) == 0
    assert n.value == nbits
    return np.unpackbits(np.frombuffer(buf.raw[:(nbits + 7) // 8], dtype=np.uint8))[:nbits]


def bseq_inputs():
    rng = np.random.default_rng(1)
    cases = []
    for n in (1, 2, 7, 8, 63, 64, 65, 510, 511, 512, 513, 1023, 5000):
        cases += [rng.integers(0, 2, n), np.zeros(n, np.uint8), np.ones(n, np.uint8),
                  np.tile([0, 1], n)[:n], np.tile([0, 0, 0, 1], n)[:n]]
    cases += [np.unpackbits(np.full(300, 0x55, np.uint8)), np.unpackbits(np.full(300, 0x11, np.uint8)),
              rng.integers(0, 2, 131072)]
    for p in (0.01, 0.1, 0.5, 0.9, 0.999):
        cases.append((rng.random(40000) < p).astype(np.uint8))
    for _ in range(10):  # alternating compressible / incompressible stretches
        parts = []
        for _ in range(int(rng.integers(2, 25))):
            if rng.random() < 0.5:
                parts.append(rng.integers(0, 2, int(rng.integers(1, 3000))))
            else:
                parts.append(np.full(int(rng.integers(1, 5000)), int(rng.integers(0, 2))))
        cases.append(np.concatenate(parts).astype(np.uint8))
    return cases


@needs_ref
def test_bseq_encoder_matches_reference_bytes():
    for i, bits in enumerate(bseq_inputs()):
        for force in (0, 1, -1):
            assert our_encode(bits, force) == Reference.bseq_construct(bits, force), (i, len(bits), force)


def test_bseq_decoder_round_trips():
    for bits in bseq_inputs():
        for force in (0, 1, -1):
            z = our_encode(bits, force)
            assert (our_expand(z, len(bits)) == np.asarray(bits, dtype=np.uint8)).all()


def _same_files(a, b):
    names = sorted(n for n in os.listdir(a) if n != "_femto_index")
    assert names == sorted(n for n in os.listdir(b) if n != "_femto_index")
    for n in names:
        x, y = open(os.path.join(a, n), "rb").read(), open(os.path.join(b, n), "rb").read()
        if x != y:
            k = next((i for i in range(min(len(x), len(y))) if x[i] != y[i]), min(len(x), len(y)))
            raise AssertionError(f"block file {n} differs at byte {k} (sizes {len(x)} vs {len(y)})")


@needs_ref
@pytest.mark.parametrize("name", ["two_docs", "gen400_big_buckets", "gen400_small_buckets", "gen400_small_blocks",
                                  "gen13_small_blocks", "gen3", "single_symbol", "multi_doc_mixed", "acgt_64k",
                                  "skewed_deep", "english_100k"])
def test_builder_matches_reference_inmemory_builder(name, corpora, tmp_path):
    docs, params = corpora[name]
    ref_dir, our_dir = str(tmp_path / "ref"), str(tmp_path / "ours")
    p = {k: v for k, v in params.items() if k != "chunk_size"}
    Reference.build_index(docs, ref_dir, str(tmp_path), **p)
    fb.build_index_host(docs, our_dir, chunk_size=0, **p)   # the in-memory reference path writes no chunks
    _same_files(ref_dir, our_dir)


@needs_ref
@pytest.mark.skipif(not os.path.exists(FEMTO_INDEX), reason="femto_index not built")
@pytest.mark.parametrize("params", [
    dict(block_size=16384, bucket_size=4096, chunk_size=1024, mark_period=20),
    dict(block_size=8192, bucket_size=1024, chunk_size=64, mark_period=3),
    dict(),
])
def test_builder_matches_femto_index(params, tmp_path):
    docs = [b"test_one;", b"test_two_fun;", corpus.random_bytes(5000, 5), corpus.random_acgt(20000, 6),
            corpus.english_like(9000, 7), corpus.all_bytes_doc()]
    files = []
    for i, d in enumerate(docs):
        f = tmp_path / f"doc{i}.bin"
        f.write_bytes(d)
        files.append(str(f))
    ref_dir, our_dir, scratch = str(tmp_path / "ref"), str(tmp_path / "ours"), tmp_path / "scratch"
    scratch.mkdir()
    cmd = [FEMTO_INDEX, "--tmp", str(scratch), "--outdir", ref_dir, "--no-enable-core"]
    if params:
        cmd += ["--param", ",".join(f"{k}={v}" for k, v in params.items())]
    subprocess.run(cmd + files, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    # femto_index stores the file path as document info
    fb.build_index_host(docs, our_dir, doc_infos=[f.encode() for f in files], **params)
    _same_files(ref_dir, our_dir)


def test_builder_reproduces_committed_reference_indexes(tmp_path):
    cases = [d for d in sorted(os.listdir(GOLDEN_DIR)) if os.path.exists(os.path.join(GOLDEN_DIR, d, "expected.json"))]
    assert cases
    for case in cases:
        exp = json.load(open(os.path.join(GOLDEN_DIR, case, "expected.json")))
        docs = [bytes.fromhex(h) for h in exp["docs_hex"]]
        p = {k: v for k, v in exp["params"].items() if k != "chunk_size"}
        out = str(tmp_path / case)
        fb.build_index_host(docs, out, chunk_size=0, **p)
        _same_files(os.path.join(GOLDEN_DIR, case, "index"), out)


def test_flatten_round_trip(built_indexes, tmp_path):
    from oracle.bindings import Oracle
    src = built_indexes["multi_doc_mixed"]
    flat = str(tmp_path / "index.femto")
    fb.flatten(src, flat)
    with Oracle(src) as a, Oracle(flat) as b:
        assert a.header_info() == b.header_info()
        n = a.header_info()["total_length"]
        assert (a.locate_range(0, n - 1) == b.locate_range(0, n - 1)).all()
    if have_reference():
        with Reference(flat) as r, Oracle(src) as a:
            n = a.header_info()["total_length"]
            assert (r.locate_range(0, n - 1) == a.locate_range(0, n - 1)).all()


def test_builder_rejects_bad_input(tmp_path):
    with pytest.raises(fb.FemtoError):
        fb.IndexBuilder(str(tmp_path / "x"), np.array([5, 3]))          # unordered document ends
    with pytest.raises(fb.FemtoError):
        fb.IndexBuilder(str(tmp_path / "y"), np.array([10]), block_size=100, bucket_size=30)
    b = fb.IndexBuilder(str(tmp_path / "z"), np.array([4]), block_size=16, bucket_size=4, chunk_size=0)
    with pytest.raises(fb.FemtoError):
        b.finish()                                                      # no rows appended
