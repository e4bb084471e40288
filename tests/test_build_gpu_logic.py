"""femto_b200/build_gpu.py (GPU suffix sort + streaming emit) exercised on CPU tensors: the same
torch code path, small inputs.  Checks the suffix order against the host sorter, the batch
splitting (by first and by second symbol), multi-round tie refinement on repetitive text, the
deterministic corpus generator, and that the streamed index equals the host-built one."""
import os

import numpy as np
import pytest
import torch

import corpus
import femto_b200 as fb
from femto_b200 import build_gpu


def _docs_cases():
    return {
        "random_bytes": [corpus.random_bytes(20000, 1)],
        "acgt_two_docs": [corpus.random_acgt(9000, 2), corpus.random_acgt(7000, 3)],
        "english": [corpus.english_like(15000, 4)],
        "repetitive": [b"abcabcabc" * 300, b"abcabc" * 100, b"a" * 500],
        "tiny": [b"", b"a", b"ba"],
    }


@pytest.mark.parametrize("name", list(_docs_cases()))
@pytest.mark.parametrize("batch", [1 << 28, 3000, 400])
def test_suffix_array_matches_host_sorter(name, batch):
    docs = _docs_cases()[name]
    T, ends = build_gpu.prepare_text_gpu([torch.frombuffer(bytearray(d), dtype=torch.uint8) if d else
                                          torch.zeros(0, dtype=torch.uint8) for d in docs])
    n = int(ends[-1])
    text, ends2 = fb.prepare_text(docs)
    assert (ends == ends2).all() and (T[:n].numpy().astype(np.uint16) == text).all()
    sa = build_gpu.suffix_array_gpu(T, n, batch=batch).numpy()
    assert (sa == fb.suffix_sort_host(text)).all()


def test_synthetic_generator_is_deterministic_and_device_independent():
    a = build_gpu.synthetic_bytes(100003, 7, "cpu").numpy()
    b = build_gpu.synthetic_bytes_numpy(100003, 7)
    assert (a == b).all()
    assert abs(a.mean() - 127.5) < 2 and len(np.unique(a)) == 256
    c = build_gpu.synthetic_bytes(5000, 7, "cpu", alphabet=b"ACGT").numpy()
    assert set(np.unique(c).tolist()) == set(b"ACGT")
    assert (c == build_gpu.synthetic_bytes_numpy(5000, 7, alphabet=b"ACGT")).all()
    assert (build_gpu.synthetic_bytes(64, 8, "cpu").numpy() != a[:64]).any()


def test_streamed_build_equals_host_build(tmp_path):
    docs = [corpus.random_bytes(30000, 5), corpus.english_like(8000, 6)]
    params = dict(block_size=16384, bucket_size=4096, chunk_size=1024, mark_period=20)
    a, b = str(tmp_path / "host"), str(tmp_path / "stream")
    fb.build_index_host(docs, a, **params)
    tdocs = [torch.frombuffer(bytearray(d), dtype=torch.uint8) for d in docs]
    info = build_gpu.build_index_gpu(tdocs, b, batch=5000, host_chunk=7000, **params)
    assert info["rows"] == sum(len(d) + 1 for d in docs)
    for f in sorted(os.listdir(a)):
        assert open(os.path.join(a, f), "rb").read() == open(os.path.join(b, f), "rb").read(), f


def test_english_like_generator_properties():
    """synthetic_english (the corpus of BASELINE configs[3]; prepared for the 16 GiB run of round 2):
    deterministic, independent of the chunking, words from the fixed vocabulary with Zipf-like
    frequencies, a skewed byte distribution (deep Huffman codes)."""
    a = build_gpu.synthetic_english(300001, 5, "cpu", vocab=2000, words_per_chunk=1 << 14).numpy()
    b = build_gpu.synthetic_english(300001, 5, "cpu", vocab=2000, words_per_chunk=1 << 14).numpy()
    assert len(a) == 300001 and (a == b).all()
    c = build_gpu.synthetic_english(100000, 5, "cpu", vocab=2000, words_per_chunk=1 << 14).numpy()
    assert (c == a[:100000]).all()                           # a prefix of the longer text
    assert (build_gpu.synthetic_english(1000, 6, "cpu", vocab=2000).numpy() != a[:1000]).any()
    letters, woff = build_gpu.english_vocabulary(2000, 5)
    vocab = {bytes(letters[woff[i]:woff[i + 1]]) for i in range(2000)}
    words = bytes(a).split(b" ")
    assert all(w in vocab for w in words[:-1])               # the last word may be cut
    assert set(np.unique(a).tolist()) <= set(range(97, 123)) | {32}
    # Zipf: the most frequent word is rank 0, and frequencies fall steeply
    from collections import Counter
    cnt = Counter(words[:-1])
    top = cnt.most_common(3)
    assert top[0][0] == bytes(letters[woff[0]:woff[1]]) and top[0][1] > 2.5 * top[2][1] * 0.5
    # skewed bytes: space is by far the most frequent, some letters are rare
    freq = np.bincount(a, minlength=256) / len(a)
    assert freq[32] > 0.08 and freq[freq > 0].min() < 0.005 and freq[ord("e")] > 0.06


def test_english_like_corpus_sorts_and_indexes(tmp_path):
    """The GPU pipeline (on CPU tensors here) handles the English-like corpus: repeats force several
    tie-refinement rounds; the streamed index equals the host-built one."""
    text = build_gpu.synthetic_english(60000, 9, "cpu", vocab=500, words_per_chunk=1 << 12)
    docs = [bytes(text[:25000].numpy()), bytes(text[25000:].numpy())]
    params = dict(block_size=16384, bucket_size=4096, chunk_size=1024, mark_period=20)
    a, b = str(tmp_path / "host"), str(tmp_path / "stream")
    fb.build_index_host(docs, a, **params)
    build_gpu.build_index_gpu([torch.frombuffer(bytearray(d), dtype=torch.uint8) for d in docs], b, batch=9000, **params)
    for f in sorted(os.listdir(a)):
        assert open(os.path.join(a, f), "rb").read() == open(os.path.join(b, f), "rb").read(), f
