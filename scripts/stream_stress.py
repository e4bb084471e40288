"""Randomised stress of the streamed host-buffer count (fm_count_flat / fm_count_bytes launched ahead of
their copies, gated by arrival marks) against the device-resident path on the same patterns.

    python scripts/stream_stress.py [iterations] [seed]

Every iteration draws a batch shape -- equal lengths or ragged, 128 Ki .. 640 Ki patterns (the streamed
range), pattern lengths 1 .. 40, alpha_t symbols or raw bytes, counts only or ranges -- runs it through
the host-buffer call and compares every result with fm_count_device on the resident copy of the same batch.
Prints one summary line; exits non-zero on the first difference."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import __graft_entry__ as g

g.build()
import corpus
import femto_b200 as fb

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
docs = [corpus.english_like(300000, 5), corpus.random_bytes(200000, 6)]
path = "/tmp/femto_b200_cache/stress_idx"
if not os.path.exists(os.path.join(path, "_femto_index")):
    os.makedirs("/tmp/femto_b200_cache", exist_ok=True)
    fb.build_index_host(docs, path, block_size=131072, bucket_size=16384, chunk_size=2048)
text = np.concatenate([np.frombuffer(d, dtype=np.uint8) for d in docs])
dev = torch.device("cuda", 0)
ix = fb.Index(path, device=0)
lib = ix.lib
t0 = time.time()
streamed = plain = 0
for it in range(iters):
    n = int(rng.integers(1 << 17, 5 << 17))
    uniform = bool(rng.integers(0, 2))
    as_bytes = bool(rng.integers(0, 2))
    want_last = bool(rng.integers(0, 4))
    if uniform:
        plen = np.full(n, int(rng.integers(1, 41)), dtype=np.int32)
    else:
        plen = rng.integers(1, 41, n).astype(np.int32)
    offs = np.zeros(n, dtype=np.int64)
    offs[1:] = np.cumsum(plen[:-1], dtype=np.int64)
    total = int(plen.sum())
    start = rng.integers(0, len(text) - 41, n)
    idx = np.repeat(start, plen) + (np.arange(total) - np.repeat(offs, plen))
    sym = text[idx]
    if rng.integers(0, 3) == 0:                      # some patterns that die early
        sym = sym.copy()
        sym[rng.integers(0, total, total // 50)] = rng.integers(0, 256, total // 50)
    flat16 = sym.astype(np.uint16) + 5
    first = np.empty(n, dtype=np.int64)
    last = np.empty(n, dtype=np.int64)
    lastp = fb._ptr(last, C.c_int64) if want_last else None
    if as_bytes:
        rc = lib.fm_count_bytes(ix.h, n, fb._ptr(plen, C.c_int32), fb._ptr(sym, C.c_uint8), fb._ptr(offs, C.c_int64),
                                fb._ptr(first, C.c_int64), lastp)
    else:
        rc = lib.fm_count_flat(ix.h, n, fb._ptr(plen, C.c_int32), fb._ptr(flat16, C.c_uint16), fb._ptr(offs, C.c_int64),
                               fb._ptr(first, C.c_int64), lastp)
    assert rc == 0, lib.fm_last_error()
    d_plen = torch.from_numpy(plen).to(dev)
    d_flat = torch.from_numpy(flat16.view(np.int16)).to(dev)
    d_offs = torch.from_numpy(offs).to(dev)
    d_first = torch.empty(n, dtype=torch.int64, device=dev)
    d_last = torch.empty(n, dtype=torch.int64, device=dev)
    ix.count_device(n, d_plen.data_ptr(), d_flat.data_ptr(), d_offs.data_ptr(), d_first.data_ptr(), d_last.data_ptr(), 0)
    torch.cuda.synchronize()
    rf, rl = d_first.cpu().numpy(), d_last.cpu().numpy()
    ok = (first == rf).all() and (last == rl).all() if want_last else (first == rl - rf + 1).all()
    if not ok:
        print(f"MISMATCH at iteration {it}: n={n} uniform={uniform} bytes={as_bytes} want_last={want_last}")
        sys.exit(1)
print(f"stream stress OK: {iters} randomised batches (seed {seed}), {time.time() - t0:.1f}s, "
      f"all results equal to the device-resident path")
