"""Host logic of the range-sharded count (femto_b200/sharded.py) with world_size 2 on CPU (gloo).

The per-rank step is an oracle-backed stand-in for count_shard_kernel that honours shard residency
(a rank may only evaluate Occ at rows of its own data blocks), so the routing -- state exchange by
owner rank, termination, results returning home -- is exercised exactly as on GPUs.  The CUDA step
itself is covered by tests/test_gpu_sharded.py (-m gpu, needs 2 GPUs)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import corpus
import femto_b200 as fb
from femto_b200 import sharded


def oracle_step_fn(index_path, rank, world, plen, flat, offs):
    """count_shard_kernel's contract (fm_kernels.cuh ShardArgs) restated with the oracle."""
    from oracle.bindings import Oracle
    o = Oracle(index_path)
    info = o.header_info()
    n, bs, nb = info["total_length"], info["block_size"], info["nblocks"]
    blocks = [b for b in range(nb) if sharded.shard_of_block(b, bs, n, world) == rank]
    lo_row = blocks[0] * bs if blocks else 0
    hi_row = min(n, (blocks[-1] + 1) * bs) if blocks else 0

    def owner(row):
        return sharded.shard_of_block(row // bs, bs, n, world)

    def step(states, dest):
        st = states.numpy()
        for k in range(st.shape[0]):
            pid, f, l, i, obA, meta = (int(x) for x in st[k])
            phase, home = meta & 15, meta >> 4
            pat = flat[offs[pid]:offs[pid] + plen[pid]]
            if phase == 3:
                m = len(pat)
                if m == 0:
                    f, l, i = 0, n - 1, 0
                else:
                    c = int(pat[m - 1])
                    f, l, i = o.C(c), o.C(c + 1) - 1, m - 1
                phase = 0
            d = -1
            while True:
                if phase == 2 or (phase == 0 and (f > l or i == 0)):
                    phase, d = 2, home
                    break
                c = int(pat[i - 1])
                if phase == 0 and c >= 261:
                    f, l, i = n, n - 1, i - 1
                    continue
                if phase == 0 and f == 0:
                    obA, phase = o.C(c), 1
                    continue
                row = f - 1 if phase == 0 else l
                if not (lo_row <= row < hi_row):
                    d = owner(row)
                    break
                r = o.occ(c, row)[0]
                if phase == 0:
                    obA, phase = r, 1
                else:
                    f, l, i, phase = obA, r - 1, i - 1, 0
            st[k] = (pid, f, l, i, obA, phase | (home << 4))
            dest[k] = d

    return step


def _worker(rank, world, port, index_path, pats, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    plen, flat, offs = fb.flatten_patterns(pats)
    npat = len(pats)
    lo, hi = npat * rank // world, npat * (rank + 1) // world
    step = oracle_step_fn(index_path, rank, world, plen, flat, offs)
    first, last, rounds = sharded.sharded_count(step, lo, hi, rank, world, "cpu")
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), first=first.numpy(), last=last.numpy(), rounds=rounds)
    dist.destroy_process_group()


def oracle_walk_fn(index_path, rank, world):
    """walk_kernel's shard mode (fm_kernels.cuh WalkArgs::state) restated with the oracle."""
    from oracle.bindings import Oracle
    o = Oracle(index_path)
    info = o.header_info()
    n, bs, nb = info["total_length"], info["block_size"], info["nblocks"]
    blocks = [b for b in range(nb) if sharded.shard_of_block(b, bs, n, world) == rank]
    lo_row = blocks[0] * bs if blocks else 0
    hi_row = min(n, (blocks[-1] + 1) * bs) if blocks else 0

    def walk(states, dest):
        st = states.numpy()
        for k in range(st.shape[0]):
            slot, row, steps, meta = (int(x) for x in st[k])
            home = meta >> 4
            if meta & 15 == 2:
                dest[k] = home
                continue
            while True:
                if not (lo_row <= row < hi_row):
                    st[k] = (slot, row, steps, home << 4)
                    dest[k] = sharded.shard_of_block(row // bs, bs, n, world)
                    break
                ch, nxt, off = o.back_step(row)
                if off >= 0:
                    st[k] = (slot, off + steps, steps, 2 | (home << 4))
                    dest[k] = home
                    break
                assert nxt >= 0
                row, steps = nxt, steps + 1

    return walk


def _locate_worker(rank, world, port, index_path, pats, max_occs, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    plen, flat, offs = fb.flatten_patterns(pats)
    npat = len(pats)
    lo, hi = npat * rank // world, npat * (rank + 1) // world
    step = oracle_step_fn(index_path, rank, world, plen, flat, offs)
    walk = oracle_walk_fn(index_path, rank, world)
    cnt, offsets, r1, r2 = sharded.sharded_locate(step, walk, lo, hi, max_occs, rank, world, "cpu")
    np.savez(os.path.join(out_dir, f"l{rank}.npz"), cnt=cnt.numpy(), offsets=offsets.numpy(), r1=r1, r2=r2)
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("name", ["acgt_64k", "english_100k"])
def test_sharded_count_two_ranks_matches_oracle(name, built_indexes, corpora, tmp_path):
    from oracle.bindings import Oracle
    docs, _ = corpora[name]
    path = built_indexes[name]
    with Oracle(path) as o:
        assert o.header_info()["nblocks"] >= 2          # really sharded
    pats = corpus.sample_patterns(docs, 120, [1, 2, 3, 5, 8, 12], seed=91)
    pats += [np.zeros(0, dtype=np.uint16), np.array([2], dtype=np.uint16), np.array([300, 70], dtype=np.uint16)]
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), path, pats, str(tmp_path)), nprocs=world, join=True)
    with Oracle(path) as o:
        of, ol = o.count([p if (p < 261).all() else p for p in pats[:-1]])
    got_f = np.concatenate([np.load(tmp_path / f"r{r}.npz")["first"] for r in range(world)])
    got_l = np.concatenate([np.load(tmp_path / f"r{r}.npz")["last"] for r in range(world)])
    assert (got_f[:-1] == of).all() and (got_l[:-1] == ol).all()
    # symbol outside the alphabet -> empty range [n, n-1]
    with Oracle(path) as o:
        n = o.header_info()["total_length"]
    assert got_l[-1] - got_f[-1] + 1 <= 0
    rounds = [int(np.load(tmp_path / f"r{r}.npz")["rounds"]) for r in range(world)]
    assert rounds[0] == rounds[1] and rounds[0] >= 1       # states really travelled


def test_exchange_routes_rows_by_destination(tmp_path):
    """Single-process sanity of the bucketing (world 1 group)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(_free_port())
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        st = torch.arange(30, dtype=torch.int64).reshape(5, 6)
        out = sharded.exchange(st, torch.zeros(5, dtype=torch.int32), 1)
        assert torch.equal(out, st)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,max_occs", [("acgt_64k", 7), ("english_100k", 40)])
def test_sharded_locate_two_ranks_matches_oracle(name, max_occs, built_indexes, corpora, tmp_path):
    """count -> clipped row ranges -> sampled-SA walks hopping between the two shards."""
    from oracle.bindings import Oracle
    docs, _ = corpora[name]
    path = built_indexes[name]
    pats = corpus.sample_patterns(docs, 90, [2, 3, 4, 6, 9, 14], seed=97)
    pats += [np.zeros(0, dtype=np.uint16), np.array([300, 70], dtype=np.uint16)]
    world = 2
    mp.spawn(_locate_worker, args=(world, _free_port(), path, pats, max_occs, str(tmp_path)), nprocs=world, join=True)
    with Oracle(path) as o:
        want = o.locate(pats[:-1], max_occs)
    parts = [np.load(tmp_path / f"l{r}.npz") for r in range(world)]
    cnt = np.concatenate([p["cnt"] for p in parts])
    offsets = np.concatenate([p["offsets"] for p in parts])
    ends = np.cumsum(cnt)
    for k, w in enumerate(want):
        got = offsets[ends[k] - cnt[k]:ends[k]]
        assert len(got) == len(w) and (got == w).all(), k
    assert cnt[-1] == 0                                      # symbol outside the alphabet
    assert int(parts[0]["r2"]) == int(parts[1]["r2"]) >= 1   # walks really changed shard


def test_expand_ranges_clip_rule():
    first = torch.tensor([5, 10, 20, 30, 7], dtype=torch.int64)
    last = torch.tensor([4, 12, 24, 35, 7], dtype=torch.int64)   # empty, 3 rows, 5 rows, 6 rows, 1 row
    rows, cnt = sharded.expand_ranges(first, last, 4)
    # last - first > max_occs cuts to max_occs rows: 5 rows (one over) are kept, 6 rows become 4
    assert cnt.tolist() == [0, 3, 5, 4, 1]
    assert rows.tolist() == [10, 11, 12, 20, 21, 22, 23, 24, 30, 31, 32, 33, 7]


# ---- host side of the device-initiated exchange: replicating the batch (the only collective of a mesh batch)
def _gather_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    # equal-length batch: int16 symbols (alpha_t) [n, m]
    mine = torch.from_numpy(rng.integers(1, 261, (5, 7)).astype(np.int16))
    allp = sharded.gather_uniform_batch(mine, world)
    # ragged batch: this rank has rank+3 patterns
    lens = rng.integers(0, 9, rank + 3).astype(np.int32)
    flat = rng.integers(1, 261, int(lens.sum())).astype(np.int16)
    plen_all, flat_all, offs_all, pid_lo = sharded.gather_ragged_batch(torch.from_numpy(lens), torch.from_numpy(flat), world)
    np.savez(os.path.join(out_dir, f"g{rank}.npz"), mine=mine.numpy(), allp=allp.numpy(), lens=lens, flat=flat,
             plen_all=plen_all.numpy(), flat_all=flat_all.numpy(), offs_all=offs_all.numpy(), pid_lo=pid_lo)
    dist.destroy_process_group()


def test_batch_gather_two_ranks(tmp_path):
    world = 2
    mp.spawn(_gather_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"g{r}.npz") for r in range(world)]
    want_uniform = np.concatenate([p["mine"] for p in parts])
    want_lens = np.concatenate([p["lens"] for p in parts])
    want_flat = np.concatenate([p["flat"] for p in parts])
    lo = 0
    for r, p in enumerate(parts):
        assert (p["allp"] == want_uniform).all()
        assert (p["plen_all"] == want_lens).all() and (p["flat_all"] == want_flat).all()
        assert (p["offs_all"] == np.cumsum(want_lens) - want_lens).all()
        assert int(p["pid_lo"]) == lo
        lo += len(p["lens"])
