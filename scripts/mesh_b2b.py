"""Debug aid: mesh count launched back to back (as bench.py's sharded leg does), with and without the
all-gather of the patterns in between.  torchrun ... scripts/mesh_b2b.py [corpus_mib] [npats_per_rank]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import femto_b200 as fb
from femto_b200 import build_gpu, sharded
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
npats = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
m = 32
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
path = f"/tmp/femto_b200_cache/dbg_bytes_{mib}"
text = build_gpu.synthetic_bytes(mib << 20, 2, dev, None)
if rank == 0 and not os.path.exists(os.path.join(path, "_femto_index")):
    os.makedirs("/tmp/femto_b200_cache", exist_ok=True)
    build_gpu.build_index_gpu([text], path, block_size=(mib << 20) // 16)
dist.barrier()
gen = torch.Generator(device=dev); gen.manual_seed(5 + rank)
starts = torch.randint(0, text.numel() - m, (npats,), generator=gen, device=dev)
mine = (text[starts[:, None] + torch.arange(m, device=dev)[None, :]].to(torch.int16) + 5).contiguous()
del text
ix = fb.Index(path, device=local, shard=rank, nshards=world)
mesh = sharded.Mesh(ix, rank, world)
first = torch.empty(npats, dtype=torch.int64, device=dev); last = torch.empty_like(first)
allp = sharded.gather_uniform_batch(mine, world)
for mode in ("no gather, finish each", "no gather, back to back", "gather, back to back", "gather, finish each"):
    keep = []
    dist.barrier(); torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    for s in range(8):
        if mode.startswith("gather"):
            keep.append(sharded.gather_uniform_batch(mine, world))
            src = keep[-1]
        else:
            src = allp
        mesh.launch_count(None, src, None, m, rank * npats, npats, first, last)
        if "finish each" in mode:
            mesh.finish()
    ev1.record()
    st = mesh.finish()
    torch.cuda.synchronize()
    print(f"[rank {rank}] {mode}: {ev0.elapsed_time(ev1) / 8:.3f} ms per step (wall {(time.perf_counter() - t0) / 8 * 1e3:.3f}), rounds {st['rounds']}", flush=True)
mesh.close(); ix.close()
dist.destroy_process_group()
