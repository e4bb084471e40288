#!/usr/bin/env python
"""bench.py -- patterns/sec of batched FM-index count() on a synthetic corpus (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU count()

A "step" is one pass of the hot path over one batch of synthetic patterns: 1 Mi text-sampled
length-32 patterns counted against the index of a 4 GiB i.i.d. uniform byte corpus
(BASELINE.json configs[1]).  Per-GPU work is fixed as N grows (each rank holds a replica of the
index and counts its own batch: weak scaling, no data-path collective).

The index is femto's unchanged on-disk format.  It is built once per box into --cache-dir by
femto_b200.build_gpu (GPU suffix sort + the byte-identical host emitter) -- the reference's own
builder runs at ~1 MB/s and would need over an hour; both arms read the same files.

Printed JSON (one line, rank 0):
  value      whole-job patterns/s with the pattern batch already resident in HBM (CUDA events
             around K launches of the count kernel through fm_count_device, max over ranks)
  e2e        the same through the host-buffer C-ABI call fm_count_flat: pinned host patterns in,
             host first/last out, copies inside the timed region (the call streams: kernel launched
             ahead of the copies, gated by an arrival counter); h2d/d2h bytes = what the call copied
             (e2e.copy_only_ms_per_step = the same bytes over the same pinned buffers with no kernel:
             the ceiling a host-buffer call has on this box; e2e_bytes = the call fed raw text bytes,
             fm_count_bytes, half the host->device traffic)
  locate     BASELINE configs[2] through fm_locate_flat, pinned buffers in and out; its own roofline
             from the walk kernel's counters; locate.whole_batch = the same over all 1 Mi patterns
  roofline   algorithmic HBM bytes per launch / kernel time, against MEASURED_PEAKS.json;
             reference_layout = bytes the reference algorithm dereferences on the on-disk layout
  sharded    N > 1: the same batches on the index split in N BWT row ranges, one per GPU, pattern
             states exchanged by the kernels themselves over NVLink peer memory; compared bit for
             bit with the replica leg
  cpu_baseline  the unmodified reference (oracle/_ref) counting a bounded sample of the same
             batch on this box's host cores, 1 server thread as shipped; the sample doubles as
             the in-run parity check (GPU first/last must equal the reference's)

Other workloads: --patterns zipf --kind english --corpus-mib 16384 (BASELINE configs[3], ragged batches);
--kind acgt --corpus-mib 64 --npats 10000 --plen 16 (configs[0]); --parallelism sharded (configs[4]: an
index larger than one GPU, built and loaded by BWT row range, no replica leg).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PLEN_DEFAULT = 32


# ------------------------------------------------------------------------------------------------
def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--corpus-mib", type=int, default=4096, help="synthetic corpus size (MiB); 4096 = headline")
    ap.add_argument("--kind", choices=["bytes", "acgt", "english"], default="bytes",
                    help="english: Zipf words over a fixed 50k-word vocabulary, many documents (BASELINE configs[3])")
    ap.add_argument("--doc-mib", type=int, default=1, help="--kind english: document size (MiB)")
    ap.add_argument("--english-piece-mib", type=int, default=16384,
                    help="--kind english: the corpus is made of generator streams of this size, seeds seed, seed+1, ...")
    ap.add_argument("--chunk-size", type=int, default=2048,
                    help="rows per document chunk of the index (default 2048 as the reference; 0 = build without "
                         "document chunks: they are not on the count / locate path, and with thousands of documents "
                         "they outweigh the rest of the index)")
    ap.add_argument("--block-rows-log2", type=int, default=27,
                    help="log2 of the rows per data block of the index (default 27 = 128 Mi as the reference; smaller "
                         "values give a small corpus several blocks, i.e. something to shard)")
    ap.add_argument("--plen-min", type=int, default=8, help="--patterns zipf: shortest pattern")
    ap.add_argument("--plen-max", type=int, default=256, help="--patterns zipf: longest pattern")
    ap.add_argument("--npats", type=int, default=1 << 20)
    ap.add_argument("--plen", type=int, default=PLEN_DEFAULT)
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--patterns", choices=["text", "random", "zipf"], default="text",
                    help="text: sampled from the corpus (every pattern occurs, all plen-1 steps run; headline); "
                         "random: uniform random symbols of the corpus alphabet (die after a few steps); "
                         "zipf: text-sampled, lengths Zipf(s=1) over [--plen-min, --plen-max] (BASELINE configs[3])")
    ap.add_argument("--lanes", type=int, default=0, help="lanes per pattern group (2, 4 or 8); 0 = engine default")
    ap.add_argument("--sched", choices=["merged", "pair"], default="merged", help="count kernel schedule")
    ap.add_argument("--block-bytes", type=int, default=0, help="rank block size of the HBM image (128/64/32)")
    ap.add_argument("--levels", type=int, default=0,
                    help="wavelet-tree levels per rank block read: 1, 2 (paired) or 4 (quad); default: the library's")
    ap.add_argument("--parallelism", choices=["replica", "sharded"], default="replica",
                    help="N>1: replicate the index and split patterns (default; the sharded leg runs beside it), or "
                         "sharded ONLY (BASELINE configs[4], a corpus larger than one GPU): every rank builds and "
                         "loads just its BWT row range")
    ap.add_argument("--exchange", choices=["mesh", "nccl"], default="mesh",
                    help="--parallelism sharded: device-initiated exchange (default) or round 1's host-driven NCCL loop")
    ap.add_argument("--cache-dir", default=os.environ.get("FEMTO_B200_CACHE", "/tmp/femto_b200_cache"))
    ap.add_argument("--locate-npats", type=int, default=100000, help="patterns of the locate leg (configs[2])")
    ap.add_argument("--cpu-sample-seconds", type=float, default=15.0)
    ap.add_argument("--ref-procs", type=int, default=0, help="--impl reference: worker processes (0 = all cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-big-locate", action="store_true", help="skip the locate leg over the whole batch")
    ap.add_argument("--no-sharded-leg", action="store_true",
                    help="N>1: skip the leg that runs the same batches on the index range-sharded over the N GPUs")
    ap.add_argument("--mesh-window", type=int, default=0, help="sharded leg: own patterns in flight per rank (0 = default)")
    ap.add_argument("--pointer-api", action="store_true",
                    help="also time fm_count (the reference's parallel_count prototype: one pointer per pattern)")
    ap.add_argument("--ref-worker", nargs=3, metavar=("INDEX", "PATS_NPZ", "OUT_NPZ"), help=argparse.SUPPRESS)
    return ap.parse_args()


def log(msg):
    print(f"[bench rank {os.environ.get('RANK', '0')}] {msg}", file=sys.stderr, flush=True)


def index_name(args):
    docs = f"_docs{args.doc_mib}MiB" if args.kind == "english" else ""
    if args.chunk_size != 2048:
        docs += f"_chunk{args.chunk_size}"
    if args.block_rows_log2 != 27:
        docs += f"_block2p{args.block_rows_log2}"
    return f"{args.kind}_{args.corpus_mib}MiB{docs}_seed{args.seed}_v1"


def english_pieces(args, device):
    """The English-like corpus (BASELINE configs[3] / [4]) = streams of --english-piece-mib (16 GiB) of the
    Zipf-word generator, stream k seeded seed + k ("8 x 16 GiB of the config-4 generator,
    different seeds"); up to 16 GiB it is one stream."""
    from femto_b200 import build_gpu
    n, piece = args.corpus_mib << 20, args.english_piece_mib << 20
    for k, s in enumerate(range(0, n, piece)):
        yield build_gpu.synthetic_english(min(piece, n - s), args.seed + k, device)


def corpus_tensor(args, device):
    import torch
    from femto_b200 import build_gpu
    n = args.corpus_mib << 20
    if args.kind == "english":
        pieces = list(english_pieces(args, device))
        return pieces[0] if len(pieces) == 1 else torch.cat(pieces)
    alphabet = b"ACGT" if args.kind == "acgt" else None
    return build_gpu.synthetic_bytes(n, args.seed, device, alphabet)


def corpus_docs(args, text):
    """The corpus as documents: one for the byte / ACGT corpora, --doc-mib pieces for the English-like one."""
    if args.kind != "english":
        return [text]
    step = args.doc_mib << 20
    return [text[i:i + step] for i in range(0, text.numel(), step)]


def ensure_index(args, device, rank, world):
    """Rank 0 builds the index if the cache does not hold it; everyone returns its path."""
    import torch
    path = os.path.join(args.cache_dir, index_name(args))
    done = os.path.join(path, "_femto_index")
    info = {"built": False}
    if rank == 0 and not os.path.exists(done):
        from femto_b200 import build_gpu
        os.makedirs(args.cache_dir, exist_ok=True)
        tmp = path + ".building"
        subprocess.run(["rm", "-rf", tmp, path], check=False)
        log(f"building index {index_name(args)} (one-time, cached in {args.cache_dir})")
        text = corpus_tensor(args, device)
        t = build_gpu.build_index_gpu(corpus_docs(args, text), tmp, chunk_size=args.chunk_size,
                                      block_size=1 << args.block_rows_log2, log=log)
        del text
        torch.cuda.empty_cache()
        os.rename(tmp, path)
        info = {"built": True, **{k: round(v, 2) for k, v in t.items()}}
        log(f"index built: {info}")
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    return path, info


def sample_patterns(args, text, batch_id, rank):
    """Text-sampled patterns (every pattern occurs, so all plen-1 backward steps execute)."""
    import torch
    g = torch.Generator(device=text.device)
    g.manual_seed((args.seed + 1) * 1000003 + batch_id * 9176 + rank * 131)
    n = text.numel()
    if args.patterns == "random":
        sym = torch.randint(0, 256, (args.npats, args.plen), generator=g, device=text.device)
        if args.kind == "acgt":
            sym = torch.tensor(list(b"ACGT"), device=text.device)[sym & 3]
        return (sym.to(torch.int16) + 5).contiguous()
    starts = torch.randint(0, n - args.plen + 1, (args.npats,), generator=g, device=text.device)
    idx = starts[:, None] + torch.arange(args.plen, device=text.device)[None, :]
    return (text[idx].to(torch.int16) + 5).contiguous()      # [npats, plen] alpha_t symbols


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.dev = device_index
        self.proc = None
        self.file = None

    def start(self):
        try:
            self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.dev), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.file, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.file.flush()
        rows = [r.strip().split(", ") for r in open(self.file.name) if r.strip()]
        os.unlink(self.file.name)
        sm, reasons = [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                out["sm_max_mhz"] = float(r[2])
                for nm, v in zip(names, r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        if sm:
            out["sm_mhz"] = float(np.median(sm))
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# reference CPU runs (subprocess workers: the parent may hold a CUDA context)

def ref_worker(index, pats_npz, out_npz):
    from oracle.bindings import Reference
    d = np.load(pats_npz)
    with Reference(index) as r:
        r.count_flat(d["plen"][:64], d["flat"], d["offs"][:64])          # touch the index / warm caches
        t0 = time.perf_counter()
        f, l = r.count_flat(d["plen"], d["flat"], d["offs"])
        dt = time.perf_counter() - t0
    np.savez(out_npz, first=f, last=l, seconds=dt)


def run_reference(index, pats2d, nprocs):
    """Count pats2d [n, plen] with the unmodified reference in nprocs processes; returns
    (first, last, wall_seconds_of_slowest_worker)."""
    n, m = pats2d.shape
    return run_reference_ragged(index, np.full(n, m, dtype=np.int32),
                                np.ascontiguousarray(pats2d.reshape(-1)).view(np.uint16),
                                np.arange(n, dtype=np.int64) * m, nprocs)


def run_reference_ragged(index, plen, flat, offs, nprocs):
    """The same for patterns of any lengths: pattern i = flat[offs[i] .. offs[i] + plen[i])."""
    n = len(plen)
    nprocs = max(1, min(nprocs, n))
    tmp = tempfile.mkdtemp(prefix="femto_ref_")
    procs = []
    bounds = [n * i // nprocs for i in range(nprocs + 1)]
    for i in range(nprocs):
        lo, hi = bounds[i], bounds[i + 1]
        f0, f1 = int(offs[lo]), int(offs[hi - 1] + plen[hi - 1])
        np.savez(os.path.join(tmp, f"in{i}.npz"), plen=np.ascontiguousarray(plen[lo:hi]),
                 flat=np.ascontiguousarray(flat[f0:max(f1, f0 + 1)]),
                 offs=np.ascontiguousarray(offs[lo:hi] - f0))
        procs.append(subprocess.Popen([sys.executable, os.path.abspath(__file__), "--ref-worker", index,
                                       os.path.join(tmp, f"in{i}.npz"), os.path.join(tmp, f"out{i}.npz")]))
    for p in procs:
        if p.wait() != 0:
            raise RuntimeError("reference worker failed")
    outs = [np.load(os.path.join(tmp, f"out{i}.npz")) for i in range(nprocs)]
    first = np.concatenate([o["first"] for o in outs])
    last = np.concatenate([o["last"] for o in outs])
    secs = max(float(o["seconds"]) for o in outs)
    subprocess.run(["rm", "-rf", tmp], check=False)
    return first, last, secs


# ------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.ref_worker:
        ref_worker(*args.ref_worker)
        return

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference" and rank != 0:
        return                                     # rank 0 alone runs the CPU arm

    import torch
    import __graft_entry__ as entry
    cached = os.path.exists(os.path.join(args.cache_dir, index_name(args), "_femto_index"))
    if args.impl == "b200" or not cached:
        # (the reference arm needs this repo's builder only when the index is not in the cache yet: the
        # reference's own builder would take over an hour; its timed path never touches the library)
        entry.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); run it on the B200 box")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1 and args.impl == "b200":
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)

    if args.impl == "b200" and args.parallelism == "sharded":
        run_sharded(args, rank, world, local, device)    # builds its own index, on all GPUs
        return

    index_path, build_info = ensure_index(args, device, rank, world if args.impl == "b200" else 1)
    text = corpus_tensor(args, device)
    workload = (f"count() of {args.npats} {'text-sampled' if args.patterns == 'text' else 'uniform-random'} "
                f"length-{args.plen} patterns on a {args.corpus_mib} MiB "
                f"synthetic {'uniform byte' if args.kind == 'bytes' else 'ACGT'} corpus (1 document), "
                f"default index params (block 128Mi rows, bucket 1Mi, mark_period 20, chunk 2048)")

    if args.impl == "reference":
        run_reference_arm(args, index_path, text, workload)
        return

    import femto_b200 as fb
    from femto_b200 import _lib
    lib = _lib.load()

    if args.patterns == "zipf":
        run_ragged(args, fb, lib, index_path, build_info, text, rank, world, local, device)
        return

    if args.block_bytes:
        assert lib.fm_set_default_block_bytes(args.block_bytes) == 0
    if args.levels:
        assert lib.fm_set_default_levels_per_block(args.levels) == 0
    t0 = time.time()
    ix = fb.Index(index_path, device=local)
    load_s = time.time() - t0
    block_bytes = int(ix.info.rank_block_size)
    levels = int(ix.info.levels_per_block)
    if levels > 1 and args.lanes:
        ix.set_count_schedule(True, args.lanes)
    elif args.lanes or args.sched != "merged":
        ix.set_count_schedule(args.sched == "merged", args.lanes or {128: 4, 64: 2, 32: 1}[block_bytes])
    log(f"index resident: {ix.info.hbm_bytes / 2**30:.2f} GiB HBM, loaded in {load_s:.1f}s, "
        f"max code length {ix.info.max_code_len}")

    # distinct pattern batches (cycled) -- each batch touches ~32 KB x npats of a multi-GB image,
    # far beyond the 126 MB L2, so no L2 flush is needed between steps
    nbatch = min(args.steps + args.warmup, 4)
    batches = [sample_patterns(args, text, b, rank) for b in range(nbatch)]
    del text
    torch.cuda.empty_cache()
    npats, m = args.npats, args.plen
    d_plen = torch.full((npats,), m, dtype=torch.int32, device=device)
    d_offs = torch.arange(npats, dtype=torch.int64, device=device) * m
    d_first = torch.empty(npats, dtype=torch.int64, device=device)
    d_last = torch.empty(npats, dtype=torch.int64, device=device)
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(b):
        ix.count_device(npats, d_plen.data_ptr(), batches[b % nbatch].data_ptr(), d_offs.data_ptr(),
                        d_first.data_ptr(), d_last.data_ptr(), stream)

    # ---- value: inputs resident in HBM ------------------------------------------------------
    for w in range(args.warmup):
        step_resident(w)
    barrier()
    launches0 = ix.kernel_launches()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for s in range(args.steps):
        step_resident(args.warmup + s)
    ev1.record()
    barrier()
    kernel_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    gpu_launches = ix.kernel_launches() - launches0
    results_gpu = (d_first.cpu().numpy().copy(), d_last.cpu().numpy().copy())
    last_batch = batches[(args.warmup + args.steps - 1) % nbatch]

    # ---- e2e: host buffers through the C ABI (pinned in, pinned out) -------------------------
    h_flat = [b.cpu().pin_memory() for b in batches]
    h_plen = torch.full((npats,), m, dtype=torch.int32).pin_memory()
    h_offs = (torch.arange(npats, dtype=torch.int64) * m).pin_memory()
    h_first = torch.empty(npats, dtype=torch.int64).pin_memory()
    h_last = torch.empty(npats, dtype=torch.int64).pin_memory()
    import ctypes as C

    def step_e2e(b):
        rc = lib.fm_count_flat(ix.h, npats, C.cast(h_plen.data_ptr(), C.POINTER(C.c_int32)),
                               C.cast(h_flat[b % nbatch].data_ptr(), C.POINTER(C.c_uint16)),
                               C.cast(h_offs.data_ptr(), C.POINTER(C.c_int64)),
                               C.cast(h_first.data_ptr(), C.POINTER(C.c_int64)),
                               C.cast(h_last.data_ptr(), C.POINTER(C.c_int64)))
        if rc:
            raise RuntimeError(f"fm_count_flat rc={rc}: {lib.fm_last_error()}")

    for w in range(args.warmup):
        step_e2e(w)
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        step_e2e(args.warmup + s)
    barrier()
    e2e_s = time.perf_counter() - t0
    assert (h_first.numpy() == results_gpu[0]).all() and (h_last.numpy() == results_gpu[1]).all(), \
        "host-buffer and device-buffer paths disagree"
    # bytes the call really copied: a batch of equal-length, densely packed patterns travels without
    # its plen / offs arrays (fm_last_transfer)
    _h2d, _d2h = C.c_int64(0), C.c_int64(0)
    lib.fm_last_transfer(ix.h, C.byref(_h2d), C.byref(_d2h))
    h2d, d2h = int(_h2d.value), int(_d2h.value)

    # ---- e2e with the patterns as raw text bytes (fm_count_bytes: half the host->device traffic) ----
    h_text = [(b.cpu() - 5).to(torch.uint8).pin_memory() for b in batches]

    def step_e2e8(b):
        rc = lib.fm_count_bytes(ix.h, npats, C.cast(h_plen.data_ptr(), C.POINTER(C.c_int32)),
                                C.cast(h_text[b % nbatch].data_ptr(), C.POINTER(C.c_uint8)),
                                C.cast(h_offs.data_ptr(), C.POINTER(C.c_int64)),
                                C.cast(h_first.data_ptr(), C.POINTER(C.c_int64)),
                                C.cast(h_last.data_ptr(), C.POINTER(C.c_int64)))
        if rc:
            raise RuntimeError(f"fm_count_bytes rc={rc}: {lib.fm_last_error()}")

    for w in range(args.warmup):
        step_e2e8(w)
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        step_e2e8(args.warmup + s)
    barrier()
    e2e8_s = time.perf_counter() - t0
    assert (h_first.numpy() == results_gpu[0]).all() and (h_last.numpy() == results_gpu[1]).all(), \
        "byte-pattern and device-buffer paths disagree"
    lib.fm_last_transfer(ix.h, C.byref(_h2d), C.byref(_d2h))
    h2d8, d2h8 = int(_h2d.value), int(_d2h.value)

    # ---- what the host side can deliver: the copies of one step alone (same bytes, same pinned buffers, both
    # directions at once on two streams, every rank at the same time) -- the ceiling of any host-buffer call
    def copy_only_ms(nbytes_in, nbytes_out):
        d_in = torch.empty(nbytes_in, dtype=torch.uint8, device=device)
        d_out = torch.empty(nbytes_out, dtype=torch.uint8, device=device)
        src = torch.empty(nbytes_in, dtype=torch.uint8).pin_memory()
        dst = torch.empty(nbytes_out, dtype=torch.uint8).pin_memory()
        s_in, s_out = torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)
        best = None
        for it in range(args.warmup + args.steps):
            if it == args.warmup:
                barrier()
                t0 = time.perf_counter()
            with torch.cuda.stream(s_in):
                d_in.copy_(src, non_blocking=True)
            with torch.cuda.stream(s_out):
                dst.copy_(d_out, non_blocking=True)
            s_in.synchronize()
            s_out.synchronize()
        barrier()
        return (time.perf_counter() - t0) / args.steps * 1e3

    copy_ms = copy_only_ms(h2d, d2h)
    copy8_ms = copy_only_ms(h2d8, d2h8)

    # ---- optional: the pointer-array prototype of parallel_count (gathers on the host first) ----
    pointer_api = None
    if args.pointer_api:
        addr = (h_flat[0].data_ptr() + np.arange(npats, dtype=np.int64) * (2 * m)).astype(np.uint64)
        ptrs = (C.c_void_p * npats).from_buffer(addr)
        plen_c = (C.c_int * npats).from_buffer(h_plen.numpy())
        lib.fm_count.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]

        def step_ptr():
            rc = lib.fm_count(ix.h, npats, plen_c, ptrs, C.c_void_p(h_first.data_ptr()), C.c_void_p(h_last.data_ptr()))
            if rc:
                raise RuntimeError(f"fm_count rc={rc}: {lib.fm_last_error()}")

        step_ptr()
        t0 = time.perf_counter()
        for _ in range(3):
            step_ptr()
        dt = (time.perf_counter() - t0) / 3
        assert (h_first.numpy() == results_gpu[0]).all() if nbatch == 1 else True
        pointer_api = {"api": "fm_count (alpha_t** pats, as parallel_count)", "ms_per_batch": round(dt * 1e3, 3),
                       "value": round(npats / dt, 1), "unit": "patterns/s"}

    # ---- locate (BASELINE configs[2]): text-sampled patterns, count + SA-sample walk, host buffers ----
    from femto_b200 import sharded as _sh

    def locate_leg(nloc, reps):
        """fm_locate_flat on the first nloc patterns of batch 0 (pinned buffers in and out), then the walk
        kernel alone on the same rows (device-resident, CUDA events) with its counters for the roofline."""
        loc_cap = nloc * 8
        h_noccs = torch.zeros(nloc, dtype=torch.int32).pin_memory()
        h_lstart = torch.zeros(nloc, dtype=torch.int64).pin_memory()
        h_lout = torch.zeros(loc_cap, dtype=torch.int64).pin_memory()

        def step_locate():
            rc = lib.fm_locate_flat(ix.h, nloc, C.cast(h_plen.data_ptr(), C.POINTER(C.c_int32)),
                                    C.cast(h_flat[0].data_ptr(), C.POINTER(C.c_uint16)),
                                    C.cast(h_offs.data_ptr(), C.POINTER(C.c_int64)), 2**31 - 1,
                                    C.cast(h_noccs.data_ptr(), C.POINTER(C.c_int32)),
                                    C.cast(h_lstart.data_ptr(), C.POINTER(C.c_int64)),
                                    C.cast(h_lout.data_ptr(), C.POINTER(C.c_int64)), loc_cap)
            if rc:
                raise RuntimeError(f"fm_locate_flat rc={rc}: {lib.fm_last_error()}")

        step_locate()                                                                             # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            step_locate()
        barrier()
        locate_s = (time.perf_counter() - t0) / reps
        # the walk kernel alone: rows of the same patterns, resident in HBM
        ix.count_device(nloc, d_plen.data_ptr(), batches[0].data_ptr(), d_offs.data_ptr(),
                        d_first.data_ptr(), d_last.data_ptr(), stream)
        rows, _cnt = _sh.expand_ranges(d_first[:nloc], d_last[:nloc], 2**31 - 1)
        rows = rows.contiguous()
        d_woff = torch.empty_like(rows)
        wev0, wev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for it in range(reps + 1):
            if it == 1:
                wev0.record()
            rc = lib.fm_locate_rows_device(ix.h, rows.numel(), rows.data_ptr(), d_woff.data_ptr(), stream)
            assert rc == 0
        wev1.record()
        torch.cuda.synchronize()
        walk_ms = wev0.elapsed_time(wev1) / reps
        assert (d_woff.cpu().numpy() == h_lout.numpy()[:rows.numel()]).all(), "walk kernel and fm_locate_flat disagree"
        return {"nloc": nloc, "locate_s": locate_s, "walk_ms": walk_ms, "rows": rows.cpu().numpy(),
                "noccs": h_noccs.numpy(), "lstart": h_lstart.numpy(), "lout": h_lout.numpy()}

    nloc = min(args.locate_npats, npats)
    leg = locate_leg(nloc, 5)
    locate_s, noccs, lstart, lout = leg["locate_s"], leg["noccs"], leg["lstart"], leg["lout"]
    locate_results = int(noccs.sum())
    big = locate_leg(npats, 3) if npats > nloc and not args.no_big_locate else None

    # ---- sharded leg (N > 1): the same index split by BWT row range over the N GPUs, pattern states
    # exchanged by the GPUs themselves (fm_mesh_*), results compared with the replica leg's -----------
    sharded_leg = None
    if world > 1 and not args.no_sharded_leg:
        try:
            sharded_leg = run_mesh_leg(args, fb, index_path, batches, nbatch, step_resident, d_first, d_last,
                                       rank, world, local, device)
        except SystemExit:
            raise                                  # a parity failure must fail the run
        except Exception as e:  # noqa: BLE001    (the replica numbers of this line stand on their own)
            log(f"sharded leg failed: {e}")
            sharded_leg = {"error": str(e)[:300]}

    # ---- max over ranks ---------------------------------------------------------------------
    times = torch.tensor([kernel_ms, e2e_s * 1e3, e2e8_s * 1e3, copy_ms, copy8_ms], dtype=torch.float64, device=device)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    kernel_ms, e2e_ms, e2e8_ms, copy_ms, copy8_ms = (float(x) for x in times)
    ms_per_step = kernel_ms / args.steps
    value = npats * world / (ms_per_step / 1e3)
    e2e_value = npats * world / (e2e_ms / args.steps / 1e3)

    if rank != 0:
        ix.close()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline: algorithmic bytes of one launch (instrumented replay of the last batch) ----
    hb = h_flat[(args.warmup + args.steps - 1) % nbatch].numpy().reshape(-1).view(np.uint16)
    st = ix.count_stats(h_plen.numpy(), hb, h_offs.numpy())
    # per distinct rank block: 128 B payload line + 16 B node record (paired-/quad-level blocks answer
    # 2 / 4 levels each and come with an 8 B record of the next node; of a quad-level block a query
    # uses 64 B of bits and an 8 B header entry); per backward-search step: the symbol's 16 B OccRec,
    # shared by the step's two Occ (+ the 16 B BucketRec in the layouts whose root is not addressed
    # by row); per pattern: 2 B/symbol + 4+8 B length/offset + 16 B result
    # (quad: 64 B of bits + the 8 B header entry; the exit entry of the root block rides in the symbol
    # record for codes of up to 8 bits and is not read)
    per_block = {1: block_bytes + 16, 2: block_bytes + 8, 4: 64 + 8}[levels]
    per_step = 16 if levels == 4 else 32
    alg_bytes = (st["distinct_block_reads"] * per_block + st["steps"] * per_step +
                 npats * (m * 2 + 28))
    peak, peak_src = measured_peaks()
    achieved = alg_bytes / (ms_per_step / 1e3) / 1e9
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": int(alg_bytes), "kernel": "count_kernel",
                "kernel_ms": round(ms_per_step, 4),
                "rank_blocks_requested": st["block_reads"], "rank_blocks_distinct": st["distinct_block_reads"],
                "occ_evals": st["occ_evals"], "backward_steps": st["steps"]}
    layout_key = {1: "plain", 2: "paired", 4: "quad"}[levels] + str(block_bytes)
    roofline["layout"] = layout_key
    traffic_file = os.path.join(ROOT, "profiles", "count_kernel_traffic.json")
    if os.path.exists(traffic_file):  # dram bytes of one launch from the committed ncu --set full capture
        try:
            roofline["traffic"] = json.load(open(traffic_file)).get(layout_key, {}).get("dram_bytes_per_launch")
        except Exception:
            pass
    if roofline["traffic"] and args.corpus_mib == 4096 and args.kind == "bytes" and args.patterns == "text" \
            and npats == 1 << 20 and m == 32:
        # what DRAM really moved (whole 128-byte lines; ncu capture of this workload) over the live kernel
        # time: the figure north_star's "ncu-reported HBM GB/s vs peak" refers to, next to the stricter
        # algorithmic one above
        roofline["dram_gb_s"] = round(roofline["traffic"] / (ms_per_step / 1e3) / 1e9, 1)
        roofline["dram_frac"] = round(roofline["traffic"] / (ms_per_step / 1e3) / 1e9 / peak, 4)
    # The count kernel is a chain of dependent random reads; the rate at which this GPU serves such
    # reads (measured right here by fm_probe_random_reads with the same access size) is the
    # ceiling that binds before HBM bandwidth does.
    try:
        probe = ix.probe_random_reads(block_bytes, steps=300)
        ach = st["distinct_block_reads"] / (ms_per_step / 1e3)
        roofline["random_access"] = {
            "bytes_per_access": block_bytes, "achieved_gaccess_s": round(ach / 1e9, 2),
            "peak_gaccess_s": round(probe["accesses_per_s"] / 1e9, 2),
            "frac": round(ach / probe["accesses_per_s"], 4), "peak_gb_s": round(probe["gb_per_s"], 1),
            "peak_source": "fm_probe_random_reads: dependent uniform random reads over the same image, all SMs"}
    except Exception as e:  # noqa: BLE001
        roofline["random_access"] = {"error": str(e)}

    # ---- locate roofline: the walk kernel's counters (instrumented replay) over its CUDA-event time ----
    def locate_roofline(lg):
        ws = ix.walk_stats(lg["rows"])
        nrows = int(lg["rows"].shape[0])
        # per wavelet-tree block: 64 B of bits + 12 B of header entries + the 8 B exit entry; per mark block
        # the 128 B line + the symbol's 8 B mark record and 16 B occ record; per LF step the 16 B bucket
        # record; 8 B per SA sample; 16 B per row in and out
        alg = (ws["wtree_blocks"] * (64 + 12 + 8) + ws["mark_blocks"] * (128 + 8 + 16) + ws["lf_steps"] * 16 +
               ws["sa_samples"] * 8 + nrows * 16)
        t = lg["walk_ms"] / 1e3
        acc = ws["wtree_blocks"] + ws["mark_blocks"] + ws["sa_samples"]
        traffic = None
        if nrows == 1 << 20 and args.corpus_mib == 4096 and args.kind == "bytes" and args.patterns == "text" and levels == 4:
            try:  # dram bytes of this launch from the committed ncu --set full capture of the same workload
                traffic = json.load(open(traffic_file)).get("walk_quad128_1Mi_rows", {}).get("dram_bytes_per_launch")
            except Exception:
                pass
        r = {"bound": "hbm", "kernel": "walk_kernel (locate)", "kernel_ms": round(lg["walk_ms"], 4), "rows": nrows,
             "achieved": round(alg / t / 1e9, 1), "peak": peak, "unit": "GB/s", "frac": round(alg / t / 1e9 / peak, 4),
             "algorithmic_bytes_per_launch": int(alg), "traffic": traffic, **ws,
             "lf_steps_per_row": round(ws["lf_steps"] / max(nrows, 1), 2)}
        ra = roofline.get("random_access", {})
        if "peak_gaccess_s" in ra:
            r["random_access"] = {"achieved_gaccess_s": round(acc / t / 1e9, 2), "peak_gaccess_s": ra["peak_gaccess_s"],
                                  "frac": round(acc / t / 1e9 / ra["peak_gaccess_s"], 4)}
        return r

    locate_roof = locate_roofline(leg)
    big_locate = None
    if big is not None:
        big_locate = {"patterns": big["nloc"], "occurrences": int(big["noccs"].sum()),
                      "value": round(big["nloc"] / big["locate_s"], 1), "unit": "patterns/s",
                      "ms_per_batch": round(big["locate_s"] * 1e3, 3), "api": "fm_locate_flat",
                      "roofline": locate_roofline(big)}

    # ---- cpu_baseline + in-run parity: the unmodified reference on a bounded sample ----------
    cpu = None
    parity = None
    locate_cpu = None
    if not args.no_cpu_baseline and world == 1:  # reference CPU leg: rank 0 at N=1 only
        from oracle.bindings import have_reference
        pats_host = last_batch.cpu().numpy()
        if have_reference():
            probe = 256
            _, _, ps = run_reference(index_path, pats_host[:probe], 1)
            rate = probe / max(ps, 1e-6)
            sample = int(max(probe, min(npats, rate * args.cpu_sample_seconds)))
            rf, rl, secs = run_reference(index_path, pats_host[:sample], 1)
            ok = bool((rf == results_gpu[0][:sample]).all() and (rl == results_gpu[1][:sample]).all())
            parity = {"checked_patterns": sample, "bit_exact_vs_reference": ok}
            cpu = {"value": round(sample / secs, 1), "unit": "patterns/s", "cores": 1, "kind": "reference",
                   "sample": f"first {sample} patterns of the last timed batch, parallel_count via oracle/_ref "
                             f"(1 server thread as shipped, src/main/server.c:3597), {secs:.1f}s"}
            if not ok:
                raise SystemExit("PARITY FAILURE: GPU first/last differ from the reference on the bench batch")
            # locate: the reference on a bounded sample of the locate batch (in-process, 1 server thread)
            from oracle.bindings import Reference
            lsample = int(max(64, min(nloc, rate * 0.5 * min(args.cpu_sample_seconds, 10.0))))
            lp2d = h_flat[0].numpy()[:lsample].view(np.uint16)
            with Reference(index_path) as r:
                t0 = time.perf_counter()
                rloc = r.locate([lp2d[i] for i in range(lsample)], 2**31 - 1)
                lsecs = time.perf_counter() - t0
            lok = all((lout[lstart[i]:lstart[i] + noccs[i]] == rloc[i]).all() and len(rloc[i]) == noccs[i]
                      for i in range(lsample))
            locate_cpu = {"value": round(lsample / lsecs, 1), "unit": "patterns/s", "cores": 1, "kind": "reference",
                          "sample": f"first {lsample} patterns of the locate batch, parallel_locate via oracle/_ref, "
                                    f"{lsecs:.1f}s", "bit_exact_vs_reference": bool(lok)}
            if not lok:
                raise SystemExit("PARITY FAILURE: GPU locate offsets differ from the reference")
            # SURVEY section 8d's byte definition: what the reference ALGORITHM dereferences on the ON-DISK layout
            # (group searches, varbyte scans, 64-byte segments ...), counted by the instrumented C restatement on a
            # small sample of the same batch.  Reported beside the image's own algorithmic bytes, never mixed with
            # them: it says what the load-time transcoding saves, it is not a roof for this kernel.
            try:
                from oracle.bindings import Oracle
                ns = min(512, npats)
                with Oracle(index_path) as o:
                    o.reset_counters()
                    of_, ol_ = o.count_flat(np.full(ns, m, dtype=np.int32),
                                            np.ascontiguousarray(pats_host[:ns].reshape(-1)).view(np.uint16),
                                            np.arange(ns, dtype=np.int64) * m)
                    cnt = o.counters()
                assert (of_ == results_gpu[0][:ns]).all() and (ol_ == results_gpu[1][:ns]).all()
                per_pat = cnt["bytes"] / ns
                roofline["reference_layout"] = {
                    "bytes_per_pattern": round(per_pat, 1), "occ_per_pattern": round(cnt["occ_calls"] / ns, 2),
                    "levels_per_occ": round(cnt["levels"] / max(cnt["occ_calls"], 1), 2),
                    "equivalent_gb_s": round(per_pat * npats / (ms_per_step / 1e3) / 1e9, 1),
                    "image_bytes_per_pattern": round(alg_bytes / npats, 1),
                    "note": "bytes the reference algorithm reads per pattern on the on-disk layout (oracle/fm_oracle.c "
                            "fmo_counters, SURVEY 8d) against the bytes the transcoded image needs; `equivalent_gb_s` "
                            "= on-disk bytes x patterns / kernel time, above HBM peak because the image needs "
                            "~%.0fx fewer bytes per pattern" % (per_pat / (alg_bytes / npats))}
            except Exception as e:  # noqa: BLE001
                roofline["reference_layout"] = {"error": str(e)[:200]}
        else:
            from oracle.bindings import Oracle
            sample = 2000
            with Oracle(index_path) as o:
                t0 = time.perf_counter()
                of, ol = o.count_flat(np.full(sample, m, dtype=np.int32),
                                      np.ascontiguousarray(pats_host[:sample].reshape(-1)).view(np.uint16),
                                      np.arange(sample, dtype=np.int64) * m)
                secs = time.perf_counter() - t0
            ok = bool((of == results_gpu[0][:sample]).all() and (ol == results_gpu[1][:sample]).all())
            parity = {"checked_patterns": sample, "bit_exact_vs_oracle": ok}
            cpu = {"value": round(sample / secs, 1), "unit": "patterns/s", "cores": 1, "kind": "port",
                   "sample": f"first {sample} patterns of the last timed batch, oracle/fm_oracle.c, {secs:.1f}s"}
            if not ok:
                raise SystemExit("PARITY FAILURE: GPU first/last differ from the oracle on the bench batch")

    out = {
        "metric": "patterns/sec (count)", "value": round(value, 1), "unit": "patterns/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
        "data": "synthetic", "impl": "b200",
        # the first four keys describe the workload and are the same in both arms; `engine` is this arm's own
        "config": {"workload": workload, "patterns_per_gpu_per_step": npats, "pattern_length": m,
                   "index": index_name(args),
                   "engine": {"index_hbm_gib": round(ix.info.hbm_bytes / 2**30, 2),
                              "index_load_s": round(load_s, 1), "index_build": build_info,
                              "parallelism": f"replica x{world} (patterns split, no collective)",
                              "rank_block_bytes": block_bytes, "levels_per_block": levels,
                              "count_schedule": os.environ.get("FEMTO_B200_COUNT_SCHED") or
                                                f"{args.sched}/{args.lanes or 'default'} lanes per pattern group",
                              "l2_policy": "inputs larger than L2: each step reads ~%.1f GB of a %.1f GiB image; "
                                           "%d distinct batches cycled" % (alg_bytes / 1e9, ix.info.hbm_bytes / 2**30, nbatch)}},
        "e2e": {"value": round(e2e_value, 1), "unit": "patterns/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": round(e2e_ms / args.steps, 4),
                "api": "fm_count_flat (pinned host buffers in/out; kernel streamed behind the copies)",
                "copy_only_ms_per_step": round(copy_ms, 4),
                "frac_of_copy_ceiling": round(copy_ms / (e2e_ms / args.steps), 4)},
        "e2e_bytes": {"value": round(npats * world / (e2e8_ms / args.steps / 1e3), 1), "unit": "patterns/s",
                      "h2d_bytes_per_step": h2d8, "d2h_bytes_per_step": d2h8,
                      "ms_per_step": round(e2e8_ms / args.steps, 4),
                      "api": "fm_count_bytes (patterns as raw text bytes, as the reference's tools read them; "
                             "the kernel widens them to alpha_t as it reads)",
                      "copy_only_ms_per_step": round(copy8_ms, 4),
                      "frac_of_copy_ceiling": round(copy8_ms / (e2e8_ms / args.steps), 4)},
        "e2e_pointer_api": pointer_api,
        "gpu_launches": int(gpu_launches), "clocks": clocks, "roofline": roofline,
        "cpu_baseline": cpu, "parity": parity,
        "locate": {"metric": "patterns/sec (locate, count + SA-sample walk, host buffers in/out)",
                   "value": round(nloc / locate_s, 1), "unit": "patterns/s", "patterns": nloc,
                   "occurrences": locate_results, "ms_per_batch": round(locate_s * 1e3, 3),
                   "api": "fm_locate_flat", "cpu_baseline": locate_cpu, "roofline": locate_roof,
                   "whole_batch": big_locate},
    }
    if sharded_leg is not None:
        if "value" in sharded_leg:
            sharded_leg["per_gpu_vs_replica"] = round(sharded_leg["value"] / value, 4)
        out["sharded"] = sharded_leg
    print(json.dumps(out), flush=True)
    ix.close()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def sample_zipf_patterns(args, text, batch_id, rank):
    """Text-sampled patterns with Zipf(s=1) lengths over [plen_min, plen_max] (P(len = plen_min + k) ~ 1/(k+1)),
    each inside one document.  Returns device tensors (plen int32, flat int16, offs int64)."""
    import torch
    dev = text.device
    g = torch.Generator(device=dev)
    g.manual_seed((args.seed + 1) * 1000003 + batch_id * 9176 + rank * 131 + 77)
    lo, hi = args.plen_min, args.plen_max
    w = 1.0 / torch.arange(1, hi - lo + 2, dtype=torch.float64, device=dev)
    plen = (torch.multinomial(w, args.npats, replacement=True, generator=g) + lo).to(torch.int64)
    doc = (args.doc_mib << 20) if args.kind == "english" else text.numel()
    ndocs = max(1, text.numel() // doc)
    d = torch.randint(0, ndocs, (args.npats,), generator=g, device=dev)
    u = torch.rand(args.npats, generator=g, device=dev, dtype=torch.float64)
    start = d * doc + (u * (doc - plen).to(torch.float64)).to(torch.int64)
    offs = torch.cumsum(plen, 0) - plen
    total = int(plen.sum())
    owner = torch.repeat_interleave(torch.arange(args.npats, device=dev), plen)
    pos = start[owner] + (torch.arange(total, device=dev) - offs[owner])
    flat = (text[pos].to(torch.int16) + 5).contiguous()
    return plen.to(torch.int32).contiguous(), flat, offs.contiguous()


def run_ragged(args, fb, lib, index_path, build_info, text, rank, world, local, device):
    """BASELINE configs[3]: count() of patterns of mixed lengths (Zipf over [8, 256]) on the English-like
    many-document corpus.  Same legs and timing rules as the main mode (value: batches resident in HBM, CUDA
    events; e2e: fm_count_flat with pinned host buffers; roofline from the instrumented replay, with the lane
    groups' activity per descent iteration; cpu_baseline = the unmodified reference on a bounded sample of
    the same batch, which is also the in-run parity check)."""
    import ctypes as C
    import torch
    t0 = time.time()
    ix = fb.Index(index_path, device=local)
    load_s = time.time() - t0
    block_bytes, levels = int(ix.info.rank_block_size), int(ix.info.levels_per_block)
    log(f"index resident: {ix.info.hbm_bytes / 2**30:.2f} GiB HBM, loaded in {load_s:.1f}s, "
        f"max code length {ix.info.max_code_len}, {ix.info.num_documents} documents")
    nbatch = min(args.steps + args.warmup, 4)
    batches = [sample_zipf_patterns(args, text, b, rank) for b in range(nbatch)]
    del text
    torch.cuda.empty_cache()
    npats = args.npats
    d_first = torch.empty(npats, dtype=torch.int64, device=device)
    d_last = torch.empty(npats, dtype=torch.int64, device=device)
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(b):
        pl, fl, of = batches[b % nbatch]
        ix.count_device(npats, pl.data_ptr(), fl.data_ptr(), of.data_ptr(), d_first.data_ptr(), d_last.data_ptr(), stream)

    for w in range(args.warmup):
        step_resident(w)
    barrier()
    launches0 = ix.kernel_launches()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for s in range(args.steps):
        step_resident(args.warmup + s)
    ev1.record()
    barrier()
    kernel_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    gpu_launches = ix.kernel_launches() - launches0
    res = (d_first.cpu().numpy().copy(), d_last.cpu().numpy().copy())
    last_b = (args.warmup + args.steps - 1) % nbatch

    host = [tuple(t.cpu().pin_memory() for t in b) for b in batches]
    h_first = torch.empty(npats, dtype=torch.int64).pin_memory()
    h_last = torch.empty(npats, dtype=torch.int64).pin_memory()

    def step_e2e(b):
        pl, fl, of = host[b % nbatch]
        rc = lib.fm_count_flat(ix.h, npats, C.cast(pl.data_ptr(), C.POINTER(C.c_int32)),
                               C.cast(fl.data_ptr(), C.POINTER(C.c_uint16)), C.cast(of.data_ptr(), C.POINTER(C.c_int64)),
                               C.cast(h_first.data_ptr(), C.POINTER(C.c_int64)),
                               C.cast(h_last.data_ptr(), C.POINTER(C.c_int64)))
        if rc:
            raise RuntimeError(f"fm_count_flat rc={rc}: {lib.fm_last_error()}")

    for w in range(args.warmup):
        step_e2e(w)
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        step_e2e(args.warmup + s)
    barrier()
    e2e_s = time.perf_counter() - t0
    assert (h_first.numpy() == res[0]).all() and (h_last.numpy() == res[1]).all(), "host and device paths disagree"
    _h2d, _d2h = C.c_int64(0), C.c_int64(0)
    lib.fm_last_transfer(ix.h, C.byref(_h2d), C.byref(_d2h))

    times = torch.tensor([kernel_ms, e2e_s * 1e3], dtype=torch.float64, device=device)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_per_step = float(times[0]) / args.steps
    e2e_ms = float(times[1]) / args.steps
    if rank != 0:
        ix.close()
        return

    pl, fl, of = (t.numpy() for t in host[last_b])
    flu = fl.view(np.uint16)
    st = ix.count_stats(pl, flu, of)
    symbols = int(pl.sum())
    per_block = {1: block_bytes + 16, 2: block_bytes + 8, 4: 64 + 8}[levels]
    per_step = 16 if levels == 4 else 32
    alg_bytes = st["distinct_block_reads"] * per_block + st["steps"] * per_step + symbols * 2 + npats * 28
    peak, peak_src = measured_peaks()
    achieved = alg_bytes / (ms_per_step / 1e3) / 1e9
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": int(alg_bytes), "kernel": "count_kernel",
                "kernel_ms": round(ms_per_step, 4), "rank_blocks_requested": st["block_reads"],
                "rank_blocks_distinct": st["distinct_block_reads"], "occ_evals": st["occ_evals"],
                "backward_steps": st["steps"],
                # warp divergence: of the lane groups of a warp inside the descent loop, how many still had a
                # block to evaluate (groups wait for the deepest code / the longest pattern of their warp)
                "descent_group_slots": st.get("group_slots"), "descent_groups_active": st.get("groups_active"),
                "active_group_fraction": (round(st["groups_active"] / st["group_slots"], 4)
                                          if st.get("group_slots") else None)}
    try:
        probe = ix.probe_random_reads(block_bytes, steps=300)
        ach = st["distinct_block_reads"] / (ms_per_step / 1e3)
        roofline["random_access"] = {"bytes_per_access": block_bytes, "achieved_gaccess_s": round(ach / 1e9, 2),
                                     "peak_gaccess_s": round(probe["accesses_per_s"] / 1e9, 2),
                                     "frac": round(ach / probe["accesses_per_s"], 4)}
    except Exception as e:  # noqa: BLE001
        roofline["random_access"] = {"error": str(e)}

    cpu = parity = None
    if not args.no_cpu_baseline and world == 1:
        from oracle.bindings import have_reference
        if have_reference():
            probe_n = 256
            _, _, ps = run_reference_ragged(index_path, pl[:probe_n], flu, of[:probe_n], 1)
            rate = probe_n / max(ps, 1e-6)
            sample = int(max(probe_n, min(npats, rate * args.cpu_sample_seconds)))
            rf, rl, secs = run_reference_ragged(index_path, pl[:sample], flu, of[:sample], 1)
            ok = bool((rf == res[0][:sample]).all() and (rl == res[1][:sample]).all())
            parity = {"checked_patterns": sample, "bit_exact_vs_reference": ok}
            cpu = {"value": round(sample / secs, 1), "unit": "patterns/s", "cores": 1, "kind": "reference",
                   "sample": f"first {sample} patterns of the last timed batch, parallel_count via oracle/_ref "
                             f"(1 server thread as shipped, src/main/server.c:3597), {secs:.1f}s"}
            if not ok:
                raise SystemExit("PARITY FAILURE: GPU first/last differ from the reference on the bench batch")
    workload = (f"count() of {npats} text-sampled patterns, lengths Zipf(s=1) over [{args.plen_min}, {args.plen_max}] "
                f"(mean {symbols / npats:.1f}), on a {args.corpus_mib} MiB synthetic English-like corpus "
                f"({ix.info.num_documents} documents of {args.doc_mib} MiB), default index params"
                + ("" if args.chunk_size == 2048 else f" except chunk_size {args.chunk_size} (no document chunks)" if args.chunk_size <= 0
                   else f" except chunk_size {args.chunk_size}"))
    out = {
        "metric": "patterns/sec (count)", "value": round(npats * world / (ms_per_step / 1e3), 1), "unit": "patterns/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "impl": "b200",
        "config": {"workload": workload, "patterns_per_gpu_per_step": npats, "symbols_per_step": symbols,
                   "index": index_name(args), "index_hbm_gib": round(ix.info.hbm_bytes / 2**30, 2),
                   "index_load_s": round(load_s, 1), "index_build": build_info, "max_code_len": int(ix.info.max_code_len),
                   "parallelism": f"replica x{world}", "rank_block_bytes": block_bytes, "levels_per_block": levels,
                   "l2_policy": "inputs larger than L2; %d distinct batches cycled" % nbatch},
        "e2e": {"value": round(npats * world / (e2e_ms / 1e3), 1), "unit": "patterns/s",
                "h2d_bytes_per_step": int(_h2d.value), "d2h_bytes_per_step": int(_d2h.value),
                "ms_per_step": round(e2e_ms, 4), "api": "fm_count_flat (pinned host buffers in/out, streamed)"},
        "gpu_launches": int(gpu_launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
    }
    print(json.dumps(out), flush=True)
    ix.close()


def run_mesh_leg(args, fb, index_path, batches, nbatch, step_resident, d_first, d_last, rank, world, local, device):
    """The batches of the replica leg counted on the index split in `world` BWT row ranges (one per GPU):
    per step one NCCL all-gather replicates the ranks' patterns, then ONE persistent kernel per GPU runs
    the whole batch, states hopping between the GPUs through peer memory.  Timed like the main leg
    (CUDA events, barrier + synchronize on both sides, max over ranks); the last step's results must be
    bit-identical to the replica leg's for the same batch."""
    import torch
    import torch.distributed as dist
    from femto_b200 import sharded
    npats, m = args.npats, args.plen
    t0 = time.time()
    ixs = fb.Index(index_path, device=local, shard=rank, nshards=world)
    load_s = time.time() - t0
    mesh = sharded.Mesh(ixs, rank, world, window=args.mesh_window)
    s_first = torch.empty(npats, dtype=torch.int64, device=device)
    s_last = torch.empty(npats, dtype=torch.int64, device=device)

    # two buffers for the replicated batch, used in turn: the all-gather of step k+2 is ordered behind the kernel
    # of step k on the stream
    gathered = [torch.empty((world * npats, m), dtype=batches[0].dtype, device=device) for _ in range(2)]

    def step(b):
        allp = sharded.gather_uniform_batch(batches[b % nbatch], world, out=gathered[b % 2])
        mesh.launch_count(None, allp, None, m, rank * npats, npats, s_first, s_last)
        return None

    keep = []
    for w in range(args.warmup):
        keep.append(step(w))
    stats = mesh.finish()
    keep.clear()
    dist.barrier()
    torch.cuda.synchronize()
    launches0 = ixs.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for s in range(args.steps):
        keep.append(step(args.warmup + s))
    ev1.record()
    stats = mesh.finish()
    dist.barrier()
    torch.cuda.synchronize()
    keep.clear()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=device)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    launches = ixs.kernel_launches() - launches0
    # parity: the replica kernel on the batch of the last sharded step
    step_resident(args.warmup + args.steps - 1)
    torch.cuda.synchronize()
    ok = torch.tensor([int(bool((s_first == d_first).all() and (s_last == d_last).all()))], device=device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    tot = torch.tensor([stats["sent"], stats["received"], stats["rounds"], stats["occ_pairs"], stats["occ_singles"],
                        stats["empty_polls"]], dtype=torch.float64, device=device)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    info = ixs.info
    leg = {"metric": "patterns/sec (count), index range-sharded by data block over the GPUs",
           "value": round(npats * world / (float(ms[0]) / args.steps / 1e3), 1), "unit": "patterns/s",
           "ms_per_step": round(float(ms[0]) / args.steps, 4), "bit_exact_vs_replica": bool(int(ok[0])),
           "exchange": "device-initiated: persistent kernel per GPU, 32-byte states stored into the owner's inbox over "
                       "NVLink peer memory; one NCCL all-gather of the patterns per step inside the timed region",
           "gpu_launches_per_step": launches / args.steps,
           "states_sent_per_pattern": round(float(tot[0]) / (npats * world), 2),
           "occ_pairs": int(tot[3]), "occ_singles": int(tot[4]), "eval_rounds": int(tot[2]),
           "empty_polls": int(tot[5]),
           "shard_rows_rank0": [int(info.first_row), int(info.end_row)],
           "shard_hbm_gib": round(info.hbm_bytes / 2**30, 2), "shard_load_s": round(load_s, 1)}
    mesh.close()
    ixs.close()
    if not leg["bit_exact_vs_replica"]:
        raise SystemExit("PARITY FAILURE: sharded results differ from the replica leg's")
    return leg


def sharded_text(args, device):
    """The corpus as build_dist.ByteText (one byte per position; 128 GiB of corpus = 137 GB of a B200's HBM)
    and a function (batch, rank) -> text-sampled patterns.  Every rank generates the same text; nothing here
    holds the corpus twice (the English-like text is laid out document by document, one 16 GiB stream at a
    time)."""
    import torch
    from femto_b200 import build_dist, build_gpu
    n, m = args.corpus_mib << 20, args.plen
    if args.kind != "english":
        alphabet = b"ACGT" if args.kind == "acgt" else None
        # counter-based generator: the first n bytes of a longer text are the text of length n
        data = build_gpu.synthetic_bytes(n + 1 + build_gpu.PAD, args.seed, device, alphabet)
        data[n:] = 0
        text = data[:n]
        return (build_dist.ByteText(data, np.array([n + 1], dtype=np.int64)),
                lambda b, r: sample_patterns(args, text, b, r))
    D = args.doc_mib << 20
    ndocs = (n + D - 1) // D
    data = torch.zeros(n + ndocs + build_gpu.PAD, dtype=torch.uint8, device=device)
    ends, pos, carry = [], 0, 0          # carry = bytes of the document being filled that are already in place
    for t in english_pieces(args, device):
        off = 0
        if carry:                         # finish the document the previous stream left open
            k = min(D - carry, t.numel())
            data[pos:pos + k] = t[:k]
            pos, off, carry = pos + k, k, carry + k
            if carry == D:
                pos, carry = pos + 1, 0
                ends.append(pos)
        q = (t.numel() - off) // D        # whole documents: one strided copy
        if q:
            data[pos:pos + q * (D + 1)].view(q, D + 1)[:, :D] = t[off:off + q * D].view(q, D)
            ends += [pos + (i + 1) * (D + 1) for i in range(q)]
            pos, off = pos + q * (D + 1), off + q * D
        k = t.numel() - off
        if k:
            data[pos:pos + k] = t[off:]
            pos, carry = pos + k, k
        del t
    if carry:
        pos += 1
        ends.append(pos)
    assert pos == n + ndocs and len(ends) == ndocs
    B = build_dist.ByteText(data, np.array(ends, dtype=np.int64))
    full = n // D                          # patterns are taken from inside the whole documents

    def sample(batch_id, rank):
        g = torch.Generator(device=device)
        g.manual_seed((args.seed + 1) * 1000003 + batch_id * 9176 + rank * 131)
        doc = torch.randint(0, max(full, 1), (args.npats,), generator=g, device=device)
        room = (D if full else n) - m + 1
        start = doc * (D + 1) + torch.randint(0, room, (args.npats,), generator=g, device=device)
        idx = start[:, None] + torch.arange(m, device=device)[None, :]
        return (data[idx].to(torch.int16) + 5).contiguous()

    return B, sample


def run_sharded(args, rank, world, local, device):
    """BASELINE configs[4]: an index too large for one GPU.  Every rank builds the data blocks it will serve
    (femto_b200/build_dist.py; text replicated, suffixes split by BWT row range, no exchange), opens them as
    its shard, and the batches run through the device-initiated exchange (fm_mesh_count): one NCCL all-gather
    of the patterns per step, then one persistent kernel per GPU.  No replica exists to compare with at this
    size: rank 0 counts a bounded sample of its own patterns with the unmodified reference on the index on
    disk (cpu_baseline + parity).  --exchange nccl runs round 1's host-driven loop instead."""
    import torch
    import torch.distributed as dist
    import femto_b200 as fb
    from femto_b200 import build_dist, sharded
    if world < 2:
        raise SystemExit("--parallelism sharded needs --gpus >= 2 (launch with torchrun)")
    npats, m = args.npats, args.plen
    index_path = os.path.join(args.cache_dir, index_name(args))
    cached = [os.path.exists(os.path.join(index_path, "_femto_index"))]
    dist.broadcast_object_list(cached, src=0)
    B, sample = sharded_text(args, device)
    build_info = {"built": False}
    if not cached[0]:
        tmp = index_path + ".building"
        if rank == 0:
            os.makedirs(args.cache_dir, exist_ok=True)
            subprocess.run(["rm", "-rf", tmp, index_path], check=False)
            log(f"building index {index_name(args)} on {world} GPUs")
        dist.barrier()
        t = build_dist.build_index_distributed(B, B.doc_ends, tmp, rank, world, chunk_size=args.chunk_size, log=log,
                                               block_size=1 << args.block_rows_log2,
                                               nthreads=max(1, (os.cpu_count() or 8) // world))
        if rank == 0:
            os.rename(tmp, index_path)
        dist.barrier()
        build_info = {"built": True, "builder": f"build_dist x{world}", **{k: round(v, 2) for k, v in t.items()}}
        log(f"index built: {build_info}")
    nbatch = 2
    batches = [sample(b, rank) for b in range(nbatch)]
    del B, sample                      # the text must be gone before the shard is loaded (137 GB at configs[4])
    torch.cuda.empty_cache()

    t0 = time.time()
    ix = fb.Index(index_path, device=local, shard=rank, nshards=world)
    load_s = time.time() - t0
    s_first = torch.empty(npats, dtype=torch.int64, device=device)
    s_last = torch.empty(npats, dtype=torch.int64, device=device)
    gathered = [torch.empty((world * npats, m), dtype=batches[0].dtype, device=device) for _ in range(2)]
    mesh = None
    rounds = 0
    if args.exchange == "mesh":
        mesh = sharded.Mesh(ix, rank, world, window=args.mesh_window)

        def step(b):
            allp = sharded.gather_uniform_batch(batches[b % nbatch], world, out=gathered[b % 2])
            mesh.launch_count(None, allp, None, m, rank * npats, npats, s_first, s_last)
    else:
        d_plen = torch.full((world * npats,), m, dtype=torch.int32, device=device)
        d_offs = torch.arange(world * npats, dtype=torch.int64, device=device) * m

        def step(b):
            nonlocal rounds
            allp = sharded.gather_uniform_batch(batches[b % nbatch], world, out=gathered[b % 2])
            fn = sharded.cuda_step_fn(ix, d_plen, allp, d_offs, world)
            f, l, rounds = sharded.sharded_count(fn, rank * npats, (rank + 1) * npats, rank, world, device)
            s_first.copy_(f)
            s_last.copy_(l)

    for w in range(args.warmup):
        step(w)
    if mesh:
        mesh.finish()
    dist.barrier()
    torch.cuda.synchronize()
    launches0 = ix.kernel_launches()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for s in range(args.steps):
        step(args.warmup + s)
    ev1.record()
    stats = mesh.finish() if mesh else {}
    dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=device)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop()
    gpu_launches = ix.kernel_launches() - launches0
    sent = torch.tensor([float(stats.get("sent", 0))], dtype=torch.float64, device=device)
    dist.all_reduce(sent, op=dist.ReduceOp.SUM)
    found = torch.tensor([int(bool(((s_last - s_first + 1) >= 1).all()))], device=device)
    dist.all_reduce(found, op=dist.ReduceOp.MIN)

    # cpu_baseline + parity: the unmodified reference on a bounded sample of rank 0's last batch
    cpu = parity = None
    if rank == 0 and not args.no_cpu_baseline:
        from oracle.bindings import have_reference
        if have_reference():
            last_b = batches[(args.warmup + args.steps - 1) % nbatch].cpu().numpy()
            probe = 256
            _, _, ps = run_reference(index_path, last_b[:probe], 1)
            sample = int(max(probe, min(npats, probe / max(ps, 1e-6) * args.cpu_sample_seconds)))
            rf, rl, secs = run_reference(index_path, last_b[:sample], 1)
            ok = bool((rf == s_first[:sample].cpu().numpy()).all() and (rl == s_last[:sample].cpu().numpy()).all())
            cpu = {"value": round(sample / secs, 1), "unit": "patterns/s", "cores": 1, "kind": "reference",
                   "sample": f"{sample} patterns of rank 0's last batch, unmodified reference, 1 server thread"}
            parity = {"checked_patterns": sample, "bit_exact_vs_reference": ok}
            if not ok:
                raise SystemExit("PARITY FAILURE: sharded results differ from the reference's")
    if rank == 0:
        ms_per_step = float(ms[0]) / args.steps
        total = npats * world
        how = ("device-initiated: persistent kernel per GPU, 32-byte states stored into the owner's inbox over NVLink "
               "peer memory" if mesh else f"host-driven: NCCL all-to-all per round, {rounds} rounds per batch")
        out = {
            "metric": "patterns/sec (count)", "value": round(total / (ms_per_step / 1e3), 1), "unit": "patterns/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
            "data": "synthetic", "impl": "b200",
            "config": {"workload": f"count() of {npats} text-sampled length-{m} patterns per GPU on a "
                                   f"{args.corpus_mib} MiB synthetic {args.kind} corpus, index range-sharded by data "
                                   f"block over {world} GPUs",
                       "patterns_per_gpu_per_step": npats, "pattern_length": m, "index": index_name(args),
                       "engine": {"parallelism": f"index range-sharded x{world} by data block", "exchange": how,
                                  "index_build": build_info, "shard_load_s": round(load_s, 1),
                                  "shard_rows_rank0": [int(ix.info.first_row), int(ix.info.end_row)],
                                  "shard_hbm_gib": round(ix.info.hbm_bytes / 2**30, 2),
                                  "states_sent_per_pattern": round(float(sent[0]) / total, 2),
                                  "l2": "inputs far larger than L2 (no flush)"}},
            "gpu_launches": int(gpu_launches), "clocks": clocks, "all_patterns_found": bool(int(found[0])),
            "cpu_baseline": cpu, "parity": parity,
        }
        print(json.dumps(out), flush=True)
    if mesh:
        mesh.close()
    ix.close()
    dist.barrier()
    dist.destroy_process_group()


def run_reference_arm(args, index_path, text, workload):
    """--impl reference: the reference's own CPU count() on this box's host cores."""
    from oracle.bindings import have_reference
    nprocs = args.ref_procs or (os.cpu_count() or 1)
    m = args.plen
    if not have_reference():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libfemto_ref.so did not travel"}))
        return
    pats = sample_patterns(args, text, 0, 0).cpu().numpy()
    del text
    # size the per-step sample so that warmup+steps finish within a few minutes
    probe = 64 * nprocs
    _, _, ps = run_reference(index_path, pats[:probe], nprocs)
    rate = probe / max(ps, 1e-6)
    budget = 150.0 / max(1, args.steps + args.warmup)
    sample = int(max(probe, min(args.npats, rate * budget)))
    log(f"reference probe: {rate:.0f} patterns/s with {nprocs} processes; {sample} patterns per step")
    secs = []
    for s in range(args.warmup + args.steps):
        lo = (s * sample) % max(1, args.npats - sample + 1)
        _, _, dt = run_reference(index_path, pats[lo:lo + sample], nprocs)
        if s >= args.warmup:
            secs.append(dt)
    per_step = float(np.mean(secs))
    value = sample / per_step
    # the whole batch once, when it fits the time the arm may take
    full = None
    if args.npats / max(value, 1.0) <= 90.0 and sample < args.npats:
        _, _, dt = run_reference(index_path, pats, nprocs)
        full = {"patterns": int(args.npats), "seconds": round(dt, 2), "value": round(args.npats / dt, 1)}
    out = {
        "metric": "patterns/sec (count)", "value": round(value, 1), "unit": "patterns/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(per_step * 1e3, 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
        "data": "synthetic", "impl": "reference",
        # the same workload keys and values as the b200 arm; what this arm timed of it is in `engine`
        "config": {"workload": workload, "patterns_per_gpu_per_step": int(args.npats), "pattern_length": m,
                   "index": index_name(args),
                   "engine": {"patterns_timed_per_step": sample, "whole_batch_once": full,
                              "note": "each step counts a bounded sample of the batch (a rate; the whole batch runs once "
                                      "at the end when it fits); index written by femto_b200's byte-identical emitter "
                                      "(the reference's builder needs >1 h for this corpus), opened by the reference "
                                      "reader"}},
        "cpu_baseline": {"value": round(value, 1), "unit": "patterns/s", "cores": nprocs, "kind": "reference",
                         "sample": f"{sample} patterns per step; {nprocs} processes, each an unmodified "
                                   f"femto server (1 worker thread, as shipped) on a slice of the batch"},
        "e2e": {"value": round(value, 1), "unit": "patterns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


def _stdout_only_for_the_result():
    """Libraries (NCCL's version banner, torchrun notices) write to fd 1; the contract is ONE JSON
    line on stdout.  Point fd 1 at stderr for the whole run and keep the real stdout for print()."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, "w", buffering=1)


if __name__ == "__main__":
    _stdout_only_for_the_result()
    main()
