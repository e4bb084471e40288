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
