"""Device-initiated sharded count (fm_mesh_*, femto_b200/csrc/fm_mesh.cu) against the oracle.

The single-GPU tests open SEVERAL BWT-range shards of one index on device 0 in this process and wire
their meshes together directly (fm_mesh_connect_local): the persistent kernels of all shards run
side by side on the one GPU (each bounded to a share of its SMs) and exchange pattern states through
the same inbox rings, tags and termination counters as on separate GPUs.  The multi-GPU test runs one
process per GPU with the inboxes mapped through CUDA IPC."""
import os
import socket

import numpy as np
import pytest
import torch

import corpus
import femto_b200 as fb
from femto_b200 import sharded

pytestmark = pytest.mark.gpu


def _run_local(path, pats, nshards, window, cap_log2, max_ctas, batches=1, want_last=True):
    """Counts `pats` split evenly over nshards ranks living in this process; returns (first, last, stats)."""
    dev = torch.device("cuda", 0)
    ixs = [fb.Index(path, device=0, shard=r, nshards=nshards) for r in range(nshards)]
    meshes = [sharded.Mesh(ix, r, nshards, window=window, cap_log2=cap_log2, connect=False) for r, ix in enumerate(ixs)]
    sharded.Mesh.connect_local(meshes)
    for m in meshes:
        m.set_limits(max_ctas=max_ctas, timeout_seconds=5.0)
    plen, flat, offs = fb.flatten_patterns(pats)
    uniform = int(plen[0]) if len(plen) and plen[0] > 0 and (plen == plen[0]).all() else 0
    d_plen = torch.from_numpy(plen).to(dev)
    d_flat = torch.from_numpy(flat.view(np.int16)).to(dev)
    d_offs = torch.from_numpy(offs).to(dev)
    n = len(pats)
    bounds = [n * r // nshards for r in range(nshards + 1)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(nshards)]
    torch.cuda.synchronize()
    out = None
    for _ in range(batches):
        firsts = [torch.full((max(bounds[r + 1] - bounds[r], 1),), -7, dtype=torch.int64, device=dev) for r in range(nshards)]
        lasts = [torch.full_like(f, -7) for f in firsts]
        torch.cuda.synchronize()
        for r, m in enumerate(meshes):
            m.launch_count(d_plen, d_flat, d_offs, uniform, bounds[r], bounds[r + 1] - bounds[r], firsts[r],
                           lasts[r] if want_last else None, stream=streams[r].cuda_stream)
        stats = [m.finish(stream=streams[r].cuda_stream) for r, m in enumerate(meshes)]
        first = np.concatenate([f.cpu().numpy()[:bounds[r + 1] - bounds[r]] for r, f in enumerate(firsts)])
        last = np.concatenate([l.cpu().numpy()[:bounds[r + 1] - bounds[r]] for r, l in enumerate(lasts)])
        out = (first, last, stats)
    for m in meshes:
        m.close()
    for ix in ixs:
        ix.close()
    return out


@pytest.mark.parametrize("name,nshards", [("acgt_64k", 2), ("english_100k", 2), ("gen400_small_blocks", 2),
                                          ("gen400_small_blocks", 3), ("multi_doc_mixed", 3), ("bytes_200k", 2),
                                          ("single_symbol", 2)])
def test_mesh_count_shards_on_one_gpu(name, nshards, built_indexes, corpora):
    from oracle.bindings import Oracle
    docs, _ = corpora[name]
    path = built_indexes[name]
    pats = corpus.sample_patterns(docs, 4000, [1, 2, 3, 5, 8, 12, 20, 32], seed=193)
    pats += [np.zeros(0, dtype=np.uint16), np.array([2], dtype=np.uint16), np.array([1], dtype=np.uint16),
             np.array([3, 4], dtype=np.uint16), np.array([260, 260], dtype=np.uint16), np.array([5, 5, 5], dtype=np.uint16)]
    with Oracle(path) as o:
        of, ol = o.count(pats)
    first, last, stats = _run_local(path, pats, nshards, window=0, cap_log2=0, max_ctas=64)
    assert (first == of).all() and (last == ol).all()
    assert sum(s["injected"] for s in stats) == len(pats)
    assert sum(s["sent"] for s in stats) == sum(s["received"] for s in stats)
    if name != "single_symbol":
        assert sum(s["sent"] for s in stats) > 0          # states did travel between the shards


def test_mesh_small_ring_wraps_and_window_throttles(built_indexes, corpora):
    """A 16384-slot ring and 64 patterns in flight per rank: the ring wraps many times within a batch and
    the injection window is what keeps it from overrunning; three batches in a row reuse it (epochs)."""
    from oracle.bindings import Oracle
    name = "english_100k"
    docs, _ = corpora[name]
    path = built_indexes[name]
    pats = corpus.sample_patterns(docs, 20000, [16], seed=7, random_fraction=0.05)   # equal lengths: uniform_len path
    with Oracle(path) as o:
        of, ol = o.count(pats)
    first, last, stats = _run_local(path, pats, 2, window=64, cap_log2=14, max_ctas=8, batches=3)
    assert (first == of).all() and (last == ol).all()
    assert sum(s["sent"] for s in stats) > 16384 * 4


def test_mesh_counts_only(built_indexes, corpora):
    """d_last == NULL: first receives the number of occurrences (parallel_count, femto.c:313-318)."""
    from oracle.bindings import Oracle
    name = "acgt_64k"
    docs, _ = corpora[name]
    path = built_indexes[name]
    pats = corpus.sample_patterns(docs, 3000, [3, 6, 11], seed=5)
    with Oracle(path) as o:
        of, ol = o.count(pats)
    first, _, _ = _run_local(path, pats, 2, window=0, cap_log2=0, max_ctas=64, want_last=False)
    assert (first == ol - of + 1).all()


def test_mesh_rejects_a_ring_too_small_for_its_window(built_indexes):
    path = built_indexes["acgt_64k"]
    ix = fb.Index(path, device=0, shard=0, nshards=1)
    m = sharded.Mesh(ix, 0, 1, window=1 << 20, cap_log2=12, connect=False)
    d = torch.zeros(8, dtype=torch.int64, device="cuda:0")
    with pytest.raises(fb.FemtoError):
        m.launch_count(d, d, d, 4, 0, 1, d, d)
    m.close()
    ix.close()


@pytest.mark.parametrize("name,nshards", [("acgt_64k", 2), ("english_100k", 2), ("multi_doc_mixed", 3),
                                          ("gen400_small_blocks", 3)])
def test_mesh_locate_rows_on_one_gpu(name, nshards, built_indexes, corpora):
    """SA[row] for every row of the index, each shard walking its share of the rows; a walk hops to the
    shard owning LF(row) until it meets a mark (fm_mesh_locate_rows)."""
    from oracle.bindings import Oracle
    path = built_indexes[name]
    dev = torch.device("cuda", 0)
    with Oracle(path) as o:
        n = o.header_info()["total_length"]
        want = o.locate_range(0, n - 1)
    ixs = [fb.Index(path, device=0, shard=r, nshards=nshards) for r in range(nshards)]
    meshes = [sharded.Mesh(ix, r, nshards, connect=False) for r, ix in enumerate(ixs)]
    sharded.Mesh.connect_local(meshes)
    for m in meshes:
        m.set_limits(max_ctas=64, timeout_seconds=5.0)
    rng = np.random.default_rng(5)
    perm = rng.permutation(n)                       # every rank gets rows from everywhere
    bounds = [n * r // nshards for r in range(nshards + 1)]
    rows = [torch.from_numpy(perm[bounds[r]:bounds[r + 1]].astype(np.int64)).to(dev) for r in range(nshards)]
    outs = [torch.full((max(len(x), 1),), -9, dtype=torch.int64, device=dev) for x in rows]
    streams = [torch.cuda.Stream(device=dev) for _ in range(nshards)]
    torch.cuda.synchronize()
    for r, m in enumerate(meshes):
        m.launch_locate_rows(rows[r], outs[r], stream=streams[r].cuda_stream)
    stats = [m.finish(stream=streams[r].cuda_stream) for r, m in enumerate(meshes)]
    got = np.empty(n, dtype=np.int64)
    for r in range(nshards):
        got[perm[bounds[r]:bounds[r + 1]]] = outs[r].cpu().numpy()[:bounds[r + 1] - bounds[r]]
    assert (got == want).all()
    assert sum(s["sent"] for s in stats) == sum(s["received"] for s in stats) > 0
    for m in meshes:
        m.close()
    for ix in ixs:
        ix.close()


# ---- one process per GPU, inboxes mapped through CUDA IPC -------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, index_path, pats, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    ix = fb.Index(index_path, device=rank, shard=rank, nshards=world)
    mesh = sharded.Mesh(ix, rank, world)
    mesh.set_limits(timeout_seconds=10.0)
    n = len(pats)
    lo, hi = n * rank // world, n * (rank + 1) // world
    plen, flat, offs = fb.flatten_patterns(pats[lo:hi])
    # every rank contributes ITS patterns; the batch is replicated with NCCL all-gathers
    plen_all, flat_all, offs_all, pid_lo = sharded.gather_ragged_batch(
        torch.from_numpy(plen).to(dev), torch.from_numpy(flat.view(np.int16)[:int(plen.sum())]).to(dev), world)
    assert pid_lo == lo
    first = torch.empty(max(hi - lo, 1), dtype=torch.int64, device=dev)
    last = torch.empty_like(first)
    for _ in range(2):
        mesh.launch_count(plen_all, flat_all, offs_all, 0, pid_lo, hi - lo, first, last)
        stats = mesh.finish()
        dist.barrier()
    cnt, offs_out = sharded.mesh_locate(mesh, first[:hi - lo], last[:hi - lo], 25)
    np.savez(os.path.join(out_dir, f"m{rank}.npz"), first=first.cpu().numpy()[:hi - lo], last=last.cpu().numpy()[:hi - lo],
             sent=stats["sent"], cnt=cnt.cpu().numpy(), offsets=offs_out.cpu().numpy())
    mesh.close()
    ix.close()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
@pytest.mark.parametrize("name", ["acgt_64k", "english_100k"])
def test_mesh_count_across_gpus(name, built_indexes, corpora, tmp_path):
    import torch.multiprocessing as mp
    from oracle.bindings import Oracle
    docs, _ = corpora[name]
    path = built_indexes[name]
    world = min(torch.cuda.device_count(), 2)
    pats = corpus.sample_patterns(docs, 6000, [1, 2, 3, 5, 8, 12, 20, 32], seed=93)
    pats += [np.zeros(0, dtype=np.uint16)]
    mp.spawn(_worker, args=(world, _free_port(), path, pats, str(tmp_path)), nprocs=world, join=True)
    with Oracle(path) as o:
        of, ol = o.count(pats)
    parts = [np.load(tmp_path / f"m{r}.npz") for r in range(world)]
    assert (np.concatenate([p["first"] for p in parts]) == of).all()
    assert (np.concatenate([p["last"] for p in parts]) == ol).all()
    assert sum(int(p["sent"]) for p in parts) > 0
    with Oracle(path) as o:
        want = o.locate(pats, 25)
    cnt = np.concatenate([p["cnt"] for p in parts])
    offsets = np.concatenate([p["offsets"] for p in parts])
    ends = np.cumsum(cnt)
    for k, w in enumerate(want):
        got = offsets[ends[k] - cnt[k]:ends[k]]
        assert len(got) == len(w) and (got == w).all(), k
