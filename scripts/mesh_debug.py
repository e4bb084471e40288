"""Debug aid: the mesh count at scale with ALL shards on one GPU (in-process), against the replica kernel.
usage: python scripts/mesh_debug.py [corpus_mib] [npats] [nshards] [max_ctas] [window]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import __graft_entry__ as g; g.build()
import femto_b200 as fb
from femto_b200 import build_gpu, sharded

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
npats = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 18
nsh = int(sys.argv[3]) if len(sys.argv) > 3 else 2
max_ctas = int(sys.argv[4]) if len(sys.argv) > 4 else 444 // nsh
window = int(sys.argv[5]) if len(sys.argv) > 5 else 0
m = 32
dev = torch.device("cuda", 0)
path = f"/tmp/femto_b200_cache/dbg_bytes_{mib}"
text = build_gpu.synthetic_bytes(mib << 20, 2, dev, None)
if not os.path.exists(os.path.join(path, "_femto_index")):
    os.makedirs("/tmp/femto_b200_cache", exist_ok=True)
    build_gpu.build_index_gpu([text], path, block_size=(mib << 20) // 8, log=print)
gen = torch.Generator(device=dev); gen.manual_seed(5)
starts = torch.randint(0, text.numel() - m, (npats * nsh,), generator=gen, device=dev)
pats = (text[starts[:, None] + torch.arange(m, device=dev)[None, :]].to(torch.int16) + 5).contiguous()
del text
full = fb.Index(path, device=0)
n = npats * nsh
d_plen = torch.full((n,), m, dtype=torch.int32, device=dev)
d_offs = torch.arange(n, dtype=torch.int64, device=dev) * m
rf = torch.empty(n, dtype=torch.int64, device=dev); rl = torch.empty_like(rf)
full.count_device(n, d_plen.data_ptr(), pats.data_ptr(), d_offs.data_ptr(), rf.data_ptr(), rl.data_ptr(), 0)
torch.cuda.synchronize()
ixs = [fb.Index(path, device=0, shard=r, nshards=nsh) for r in range(nsh)]
print("shards", [(int(i.info.first_row), int(i.info.end_row)) for i in ixs], "blocks", int(full.info.num_blocks))
meshes = [sharded.Mesh(ix, r, nsh, window=window, connect=False) for r, ix in enumerate(ixs)]
sharded.Mesh.connect_local(meshes)
for mm in meshes:
    mm.set_limits(max_ctas=max_ctas, timeout_seconds=float(os.environ.get("MESH_TIMEOUT", "3")))
streams = [torch.cuda.Stream(device=dev) for _ in range(nsh)]
fs = [torch.full((npats,), -7, dtype=torch.int64, device=dev) for _ in range(nsh)]
ls = [torch.full((npats,), -7, dtype=torch.int64, device=dev) for _ in range(nsh)]
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for r, mm in enumerate(meshes):
        mm.launch_count(None, pats, None, m, r * npats, npats, fs[r], ls[r], stream=streams[r].cuda_stream)
    try:
        stats = [mm.finish(stream=streams[r].cuda_stream) for r, mm in enumerate(meshes)]
    except Exception as e:
        print("FAILED", e)
        for r, mm in enumerate(meshes):
            try:
                print(r, mm.finish(stream=streams[r].cuda_stream))
            except Exception as e2:
                print(r, e2)
        got = torch.cat(fs)
        print("undelivered", int((got == -7).sum()))
        break
    dt = time.perf_counter() - t0
    ok = bool((torch.cat(fs) == rf).all() and (torch.cat(ls) == rl).all())
    print(f"rep {rep}: {dt*1e3:.2f} ms, {n/dt/1e6:.1f} M patterns/s, exact={ok}", stats)
