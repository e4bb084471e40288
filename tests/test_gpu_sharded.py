"""Range-sharded count on real GPUs: each rank holds a BWT row range (fm_open_shard), pattern states
are routed between ranks with NCCL all-to-all (femto_b200/sharded.py), results must equal the
oracle's on the whole index.  Needs >= 2 GPUs (skipped on a single-GPU box); the routing logic itself
is covered on CPU by tests/test_sharded_routing.py."""
import os
import socket

import numpy as np
import pytest
import torch

import corpus
import femto_b200 as fb

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, index_path, pats, out_dir):
    import torch.distributed as dist
    from femto_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    ix = fb.Index(index_path, device=rank, shard=rank, nshards=world)
    plen, flat, offs = fb.flatten_patterns(pats)
    d_plen = torch.from_numpy(plen).to(dev)
    d_flat = torch.from_numpy(flat.view(np.int16)).to(dev)
    d_offs = torch.from_numpy(offs).to(dev)
    npat = len(pats)
    lo, hi = npat * rank // world, npat * (rank + 1) // world
    step = sharded.cuda_step_fn(ix, d_plen, d_flat, d_offs, world)
    first, last, rounds = sharded.sharded_count(step, lo, hi, rank, world, dev)
    torch.cuda.synchronize()
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), first=first.cpu().numpy(), last=last.cpu().numpy(),
             rounds=rounds, first_row=ix.info.first_row, end_row=ix.info.end_row)
    ix.close()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
@pytest.mark.parametrize("name", ["acgt_64k", "english_100k", "gen400_small_blocks"])
def test_sharded_count_on_gpus(name, built_indexes, corpora, tmp_path):
    import torch.multiprocessing as mp
    from oracle.bindings import Oracle
    docs, _ = corpora[name]
    path = built_indexes[name]
    world = 2
    pats = corpus.sample_patterns(docs, 3000, [1, 2, 3, 5, 8, 12, 20, 32], seed=93)
    pats += [np.zeros(0, dtype=np.uint16), np.array([2], dtype=np.uint16)]
    mp.spawn(_worker, args=(world, _free_port(), path, pats, str(tmp_path)), nprocs=world, join=True)
    with Oracle(path) as o:
        of, ol = o.count(pats)
        n = o.header_info()["total_length"]
    parts = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    assert (np.concatenate([p["first"] for p in parts]) == of).all()
    assert (np.concatenate([p["last"] for p in parts]) == ol).all()
    assert int(parts[0]["first_row"]) == 0 and int(parts[-1]["end_row"]) == n
    assert int(parts[0]["end_row"]) == int(parts[1]["first_row"])
    assert int(parts[0]["rounds"]) >= 1


def _locate_worker(rank, world, port, index_path, pats, max_occs, out_dir):
    import torch.distributed as dist
    from femto_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    ix = fb.Index(index_path, device=rank, shard=rank, nshards=world)
    plen, flat, offs = fb.flatten_patterns(pats)
    d_plen = torch.from_numpy(plen).to(dev)
    d_flat = torch.from_numpy(flat.view(np.int16)).to(dev)
    d_offs = torch.from_numpy(offs).to(dev)
    npat = len(pats)
    lo, hi = npat * rank // world, npat * (rank + 1) // world
    step = sharded.cuda_step_fn(ix, d_plen, d_flat, d_offs, world)
    walk = sharded.cuda_walk_fn(ix, world)
    cnt, offsets, r1, r2 = sharded.sharded_locate(step, walk, lo, hi, max_occs, rank, world, dev)
    torch.cuda.synchronize()
    np.savez(os.path.join(out_dir, f"l{rank}.npz"), cnt=cnt.cpu().numpy(), offsets=offsets.cpu().numpy(), r1=r1, r2=r2)
    ix.close()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
@pytest.mark.parametrize("name", ["acgt_64k", "english_100k", "gen400_small_blocks"])
def test_sharded_locate_on_gpus(name, built_indexes, corpora, tmp_path):
    """parallel_locate over two BWT-range shards: walk states routed by NCCL all-to-all."""
    import torch.multiprocessing as mp
    from oracle.bindings import Oracle
    docs, _ = corpora[name]
    path = built_indexes[name]
    world, max_occs = 2, 25
    pats = corpus.sample_patterns(docs, 1500, [2, 3, 5, 8, 12, 20], seed=95)
    pats += [np.zeros(0, dtype=np.uint16)]
    mp.spawn(_locate_worker, args=(world, _free_port(), path, pats, max_occs, str(tmp_path)), nprocs=world, join=True)
    with Oracle(path) as o:
        want = o.locate(pats, max_occs)
    parts = [np.load(tmp_path / f"l{r}.npz") for r in range(world)]
    cnt = np.concatenate([p["cnt"] for p in parts])
    offsets = np.concatenate([p["offsets"] for p in parts])
    ends = np.cumsum(cnt)
    for k, w in enumerate(want):
        got = offsets[ends[k] - cnt[k]:ends[k]]
        assert len(got) == len(w) and (got == w).all(), k


def test_chunk_documents_through_the_c_abi(built_indexes, corpora):
    """fm_chunk_documents on an opened index (host-side decode of the stored chunk lists; the decode
    itself is pinned on CPU against the live reference in tests/test_chunk_documents.py)."""
    from oracle.bindings import Oracle
    name = "multi_doc_mixed"
    path = built_indexes[name]
    cs = corpora[name][1]["chunk_size"]
    with fb.Index(path, device=0) as ix, Oracle(path) as o:
        n = o.header_info()["total_length"]
        sa = o.locate_range(0, n - 1)
        doc_of = np.array([o.resolve(int(x))[0] for x in sa], dtype=np.int64)
        for row in list(range(0, n, max(1, cs // 2)))[:40] + [n - 1]:
            first, last, docs = ix.chunk_documents(row)
            assert first <= row <= last
            assert (docs == np.unique(doc_of[first:last + 1])).all()
