"""Tuning aid: the locate walk kernel over 1 Mi rows of the bench index, at FEMTO_B200_WALK_CTAS of the
environment.  usage: FEMTO_B200_WALK_CTAS=8 python scripts/walk_sweep.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import __graft_entry__ as g; g.build()
import femto_b200 as fb
from femto_b200 import build_gpu, sharded
dev = torch.device("cuda", 0)
path = "/tmp/femto_b200_cache/bytes_4096MiB_seed2_v1"
text = build_gpu.synthetic_bytes(4096 << 20, 2, dev, None)
if not os.path.exists(os.path.join(path, "_femto_index")):
    os.makedirs("/tmp/femto_b200_cache", exist_ok=True)
    build_gpu.build_index_gpu([text], path + ".tmp")
    os.rename(path + ".tmp", path)
n, m = 1 << 20, 32
gen = torch.Generator(device=dev); gen.manual_seed(5)
starts = torch.randint(0, text.numel() - m, (n,), generator=gen, device=dev)
pats = (text[starts[:, None] + torch.arange(m, device=dev)[None, :]].to(torch.int16) + 5).contiguous()
del text
ix = fb.Index(path, device=0)
d_plen = torch.full((n,), m, dtype=torch.int32, device=dev)
d_offs = torch.arange(n, dtype=torch.int64, device=dev) * m
f = torch.empty(n, dtype=torch.int64, device=dev); l = torch.empty_like(f)
ix.count_device(n, d_plen.data_ptr(), pats.data_ptr(), d_offs.data_ptr(), f.data_ptr(), l.data_ptr(), 0)
rows, _ = sharded.expand_ranges(f, l, 2**31 - 1)
rows = rows.contiguous(); out = torch.empty_like(rows)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(6):
    if it == 1: ev0.record()
    assert ix.lib.fm_locate_rows_device(ix.h, rows.numel(), rows.data_ptr(), out.data_ptr(), 0) == 0
ev1.record(); torch.cuda.synchronize()
print(f"WALK_CTAS={os.environ.get('FEMTO_B200_WALK_CTAS', 'default')}: {ev0.elapsed_time(ev1) / 5:.4f} ms per launch over {rows.numel()} rows, checksum {int(out.sum())}")
