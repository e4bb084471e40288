"""Debug aid: mesh count, one process per GPU (CUDA IPC inboxes), no collective in the data path: every
rank generates the whole batch itself.  Launch with torchrun.
usage: torchrun ... scripts/mesh_debug_mp.py [corpus_mib] [npats_per_rank] [window] [reps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import __graft_entry__ as g
import femto_b200 as fb
from femto_b200 import build_gpu, sharded

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
npats = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
window = int(sys.argv[3]) if len(sys.argv) > 3 else 0
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
m = 32
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
path = f"/tmp/femto_b200_cache/dbg_bytes_{mib}"
text = build_gpu.synthetic_bytes(mib << 20, 2, dev, None)
if rank == 0 and not os.path.exists(os.path.join(path, "_femto_index")):
    os.makedirs("/tmp/femto_b200_cache", exist_ok=True)
    build_gpu.build_index_gpu([text], path, block_size=(mib << 20) // 16)
dist.barrier()
gen = torch.Generator(device=dev); gen.manual_seed(5)
starts = torch.randint(0, text.numel() - m, (npats * world,), generator=gen, device=dev)
pats = (text[starts[:, None] + torch.arange(m, device=dev)[None, :]].to(torch.int16) + 5).contiguous()
del text
ix = fb.Index(path, device=local, shard=rank, nshards=world)
mesh = sharded.Mesh(ix, rank, world, window=window)
mesh.set_limits(timeout_seconds=float(os.environ.get("MESH_TIMEOUT", "4")))
first = torch.full((npats,), -7, dtype=torch.int64, device=dev); last = torch.full_like(first, -7)
for rep in range(reps):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    mesh.launch_count(None, pats, None, m, rank * npats, npats, first, last)
    try:
        st = mesh.finish()
        dt = time.perf_counter() - t0
        print(f"[rank {rank}] rep {rep}: {dt*1e3:.2f} ms -> {npats*world/dt/1e6:.1f} M patterns/s (job) {st}", flush=True)
    except Exception as e:
        print(f"[rank {rank}] rep {rep} FAILED: {e}; undelivered {int((first == -7).sum())}", flush=True)
        break
mesh.close(); ix.close()
# parity against the replica kernel on this rank's share
full = fb.Index(path, device=local)
n = npats
d_plen = torch.full((n,), m, dtype=torch.int32, device=dev)
d_offs = torch.arange(n, dtype=torch.int64, device=dev) * m
rf = torch.empty(n, dtype=torch.int64, device=dev); rl = torch.empty_like(rf)
mine = pats[rank * npats:(rank + 1) * npats].contiguous()
full.count_device(n, d_plen.data_ptr(), mine.data_ptr(), d_offs.data_ptr(), rf.data_ptr(), rl.data_ptr(), 0)
torch.cuda.synchronize()
print(f"[rank {rank}] exact={bool((first == rf).all() and (last == rl).all())}", flush=True)
full.close()
dist.destroy_process_group()
