"""The multi-rank index build (femto_b200/build_dist.py) on CUDA tensors: two ranks' shares built one after
the other on this GPU must give the files of the host builder, and the index must answer like the oracle.
(Runs last: the file name sorts behind the other GPU tests.)"""
import os

import numpy as np
import pytest
import torch

import corpus
import femto_b200 as fb
from femto_b200 import build_dist
from oracle.bindings import Oracle

pytestmark = pytest.mark.gpu


def test_two_ranks_shares_on_the_gpu_equal_the_host_build(tmp_path):
    docs = [corpus.random_bytes(60000, 21), corpus.english_like(30000, 22), b"q"]
    params = dict(block_size=16384, bucket_size=4096, chunk_size=1024, mark_period=20)
    a, b = str(tmp_path / "host"), str(tmp_path / "ranks")
    fb.build_index_host(docs, a, **params)
    dev = torch.device("cuda", 0)
    B = build_dist.ByteText.from_docs([torch.frombuffer(bytearray(d), dtype=torch.uint8).to(dev) for d in docs])
    parts = [build_dist.build_rank_blocks(B, B.doc_ends, b, rank, 2, batch=400, host_chunk=15000, **params)[0]
             for rank in range(2)]
    build_dist.write_header_from_parts(b, B.doc_ends, parts, **params)
    for f in sorted(os.listdir(a)):
        assert open(os.path.join(a, f), "rb").read() == open(os.path.join(b, f), "rb").read(), f
    pats = corpus.sample_patterns(docs, 300, [1, 3, 8, 20], seed=5, random_fraction=0.2)
    with fb.Index(b) as ix, Oracle(b) as o:
        f, l = ix.count(pats)
        of, ol = o.count(pats)
        assert (f == of).all() and (l == ol).all()
