"""Range-sharded count and locate across the GPUs of one box (second case).

When an index exceeds one GPU's HBM it is partitioned by BWT row range at data-block granularity
(``fm_open_shard``: block b lives on rank b*G/nblocks; the header tables are replicated).  A
backward-search step needs Occ at rows ``first-1`` and ``last``; the LF mapping scatters those rows
over the whole BWT, so instead of fetching index data a pattern's 48-byte STATE travels to the rank
that owns the row it needs next:

    loop until no state is left anywhere:
        every rank: advance its states while their rows are resident      (count_shard_kernel)
        every rank: finished states that are home -> results
        all ranks : all-to-all of the remaining states, keyed by the rank owning their next row

One all-to-all round per dependent remote row: at most 2 per backward-search step.  The exchange
is NCCL ``all_to_all_single`` over NVLink/NVSwitch (gloo on CPU for the host-logic tests), preceded
by one ``all_gather`` of the per-destination counts that doubles as the termination test: one host
synchronisation per round.  State volume is 48 B x patterns per round, far below link bandwidth --
the cost is the ~2(m-1) rounds, which is why batches should be large.

Locate works the same way (``sharded_locate_rows``): the state of a sampled-SA walk is {result slot,
row, LF steps so far, home}; every rank follows LF from its states while their rows are resident
and a mark has not been reached (``walk_kernel`` in shard mode), then the states are exchanged by
the rank owning their next row; a walk takes fewer than ``mark_period`` steps, so at most that
many rounds.  ``sharded_locate`` chains the two: count, expand the ranges into rows with the
reference's clipping rule, walk.

The per-rank step function is pluggable so that the routing logic can be tested on CPU with an
oracle-backed step (tests/test_sharded_routing.py) and run on GPUs with the CUDA kernel.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist

def shard_of_block(b: int, block_size: int, total_length: int, nshards: int) -> int:
    """The rank holding data block b (fm_open_shard; fm_format.hpp shard_of_block): contiguous block ranges
    balanced by rows -- a block goes to the shard its middle row falls into."""
    if nshards <= 1 or total_length <= 0:
        return 0
    return min(nshards - 1, ((2 * b + 1) * block_size * nshards) // (2 * total_length))


STATE_WORDS = 6          # pid, first, last, i, obA, meta = phase | home << 4
PHASE_NEW, PHASE_DONE = 3, 2

# step_fn(state[n,6] int64, dest[n] int32) -> None: advances states in place, fills dest
StepFn = Callable[[torch.Tensor, torch.Tensor], None]


def new_states(pid_lo: int, pid_hi: int, home: int, device) -> torch.Tensor:
    n = pid_hi - pid_lo
    st = torch.zeros((n, STATE_WORDS), dtype=torch.int64, device=device)
    st[:, 0] = torch.arange(pid_lo, pid_hi, dtype=torch.int64, device=device)
    st[:, 5] = PHASE_NEW | (home << 4)
    return st


def exchange(states: torch.Tensor, dest: torch.Tensor, world: int, group=None) -> torch.Tensor:
    """All-to-all of state rows by destination rank (uneven splits)."""
    recv, _ = route(states, dest, world, group)
    return recv


def route(states: torch.Tensor, dest: torch.Tensor, world: int, group=None) -> Tuple[torch.Tensor, int]:
    """One exchange round.  dest[k] in [0, world) = rank that must see state k next; dest[k] == world
    drops the state (it has delivered its result).  ONE small collective carries every rank's send
    counts to everybody -- which gives each rank its receive counts AND the number of states left in
    the whole job -- then one all-to-all moves the states.  One host synchronisation per round.
    Returns (received states, states left anywhere before this exchange)."""
    words = states.shape[1]
    order = torch.argsort(dest)
    counts = torch.bincount(dest, minlength=world + 1)[:world].to(torch.int64)
    matrix = torch.empty((world, world), dtype=torch.int64, device=states.device)
    dist.all_gather_into_tensor(matrix.view(-1), counts, group=group)
    m = matrix.tolist()                                       # the round's only host synchronisation
    left = sum(sum(r) for r in m)
    rank = dist.get_rank(group)
    sc, rc = m[rank], [m[r][rank] for r in range(world)]
    if left == 0:
        return states[:0], 0
    send = states[order[:sum(sc)]].contiguous()
    recv = torch.empty((sum(rc), words), dtype=torch.int64, device=states.device)
    dist.all_to_all_single(recv.view(-1), send.view(-1), [c * words for c in rc], [c * words for c in sc],
                           group=group)
    return recv, left


def sharded_count(step_fn: StepFn, pid_lo: int, pid_hi: int, rank: int, world: int, device,
                  group=None, max_rounds: int = 100000) -> Tuple[torch.Tensor, torch.Tensor, int]:
    """Count patterns [pid_lo, pid_hi) (this rank's share of a batch replicated on every rank).
    Returns (first, last, rounds) for this rank's patterns."""
    n_mine = pid_hi - pid_lo
    # one spare slot at the end takes the writes of states that are not finished-and-home, so that
    # delivering results needs no data-dependent indexing (no host synchronisation)
    first = torch.zeros(n_mine + 1, dtype=torch.int64, device=device)
    last = torch.zeros(n_mine + 1, dtype=torch.int64, device=device)
    states = new_states(pid_lo, pid_hi, rank, device)
    rounds = 0
    while True:
        dest = torch.full((states.shape[0],), rank, dtype=torch.int32, device=device)
        if states.shape[0]:
            step_fn(states, dest)
        # finished states that are already home deliver their result and leave the job
        home_done = ((states[:, 5] & 15) == PHASE_DONE) & (dest == rank)
        slot = torch.where(home_done, states[:, 0] - pid_lo, torch.full_like(states[:, 0], n_mine))
        first[slot] = states[:, 1]
        last[slot] = states[:, 2]
        dest = torch.where(home_done, torch.full_like(dest, world), dest)
        states, left = route(states, dest.long(), world, group)
        if left == 0:
            break
        rounds += 1
        if rounds > max_rounds:
            raise RuntimeError("sharded_count did not terminate")
    return first[:n_mine], last[:n_mine], rounds


def cuda_step_fn(ix, d_plen: torch.Tensor, d_flat: torch.Tensor, d_offs: torch.Tensor, nshards: int) -> StepFn:
    """Step function backed by count_shard_kernel (fm_count_shard_step)."""
    from . import _check

    def step(states: torch.Tensor, dest: torch.Tensor) -> None:
        assert states.is_cuda and states.is_contiguous() and dest.is_contiguous()
        stream = torch.cuda.current_stream().cuda_stream
        _check(ix.lib.fm_count_shard_step(ix.h, states.shape[0], states.data_ptr(), d_plen.data_ptr(),
                                          d_flat.data_ptr(), d_offs.data_ptr(), dest.data_ptr(), nshards, stream),
               "fm_count_shard_step")

    return step


# ---- locate ---------------------------------------------------------------------------------------
WALK_WORDS = 4           # result slot, row (text offset once finished), LF steps so far, phase | home << 4

# walk_fn(state[n,4] int64, dest[n] int32) -> None
WalkFn = Callable[[torch.Tensor, torch.Tensor], None]


def sharded_locate_rows(walk_fn: WalkFn, rows: torch.Tensor, rank: int, world: int, device,
                        group=None, max_rounds: int = 100000) -> Tuple[torch.Tensor, int]:
    """SA[row] for this rank's ``rows`` (global BWT rows, any shard).  Collective: every rank calls it
    with its own rows (possibly none).  Returns (offsets aligned with rows, exchange rounds)."""
    n = int(rows.shape[0])
    out = torch.full((n + 1,), -1, dtype=torch.int64, device=device)   # spare slot: see sharded_count
    states = torch.zeros((n, WALK_WORDS), dtype=torch.int64, device=device)
    states[:, 0] = torch.arange(n, dtype=torch.int64, device=device)
    states[:, 1] = rows.to(device=device, dtype=torch.int64)
    states[:, 3] = rank << 4
    rounds = 0
    while True:
        dest = torch.full((states.shape[0],), rank, dtype=torch.int32, device=device)
        if states.shape[0]:
            walk_fn(states, dest)
        home_done = ((states[:, 3] & 15) == PHASE_DONE) & (dest == rank)
        slot = torch.where(home_done, states[:, 0], torch.full_like(states[:, 0], n))
        out[slot] = states[:, 1]
        dest = torch.where(home_done, torch.full_like(dest, world), dest)
        states, left = route(states, dest.long(), world, group)
        if left == 0:
            break
        rounds += 1
        if rounds > max_rounds:
            raise RuntimeError("sharded_locate_rows did not terminate")
    return out[:n], rounds


def take_walk_status(ix) -> int:
    """Status word of the caller-stream walk launches enqueued so far (fm_take_status): 0, or the code
    of a malformed walk.  sharded_locate checks it after its exchange loop."""
    import ctypes as C
    from . import _check
    st = C.c_int(0)
    _check(ix.lib.fm_take_status(ix.h, torch.cuda.current_stream().cuda_stream, C.byref(st)), "fm_take_status")
    return int(st.value)


def expand_ranges(first: torch.Tensor, last: torch.Tensor, max_occs: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Rows first..last of every non-empty range, clipped the way parallel_locate clips them
    (src/main/server.c:4411-4415: ``last - first > max_occs`` cuts to max_occs rows, so a range
    exactly one over keeps max_occs + 1).  Returns (rows, number of rows per range)."""
    cnt = torch.clamp(last - first + 1, min=0)
    over = (last - first) > max_occs
    cnt = torch.where(over, torch.full_like(cnt, max_occs), cnt)
    starts = torch.cumsum(cnt, 0) - cnt
    total = int(cnt.sum().item())
    owner = torch.repeat_interleave(torch.arange(cnt.shape[0], device=cnt.device), cnt)
    rows = first[owner] + (torch.arange(total, device=cnt.device) - starts[owner])
    return rows, cnt


def sharded_locate(step_fn: StepFn, walk_fn: WalkFn, pid_lo: int, pid_hi: int, max_occs: int, rank: int,
                   world: int, device, group=None):
    """parallel_locate over a range-sharded index for patterns [pid_lo, pid_hi) of a batch replicated
    on every rank.  Returns (noccs[n], offsets concatenated in pattern order, rounds_count, rounds_walk)."""
    first, last, r1 = sharded_count(step_fn, pid_lo, pid_hi, rank, world, device, group)
    rows, cnt = expand_ranges(first, last, max_occs)
    offs, r2 = sharded_locate_rows(walk_fn, rows, rank, world, device, group)
    return cnt, offs, r1, r2


def cuda_walk_fn(ix, nshards: int) -> WalkFn:
    """Walk function backed by walk_kernel in shard mode (fm_locate_shard_step)."""
    from . import _check

    def walk(states: torch.Tensor, dest: torch.Tensor) -> None:
        assert states.is_cuda and states.is_contiguous() and dest.is_contiguous()
        stream = torch.cuda.current_stream().cuda_stream
        _check(ix.lib.fm_locate_shard_step(ix.h, states.shape[0], states.data_ptr(), dest.data_ptr(), nshards, stream),
               "fm_locate_shard_step")

    return walk


# ---- device-initiated exchange ("mesh", femto_b200/csrc/fm_mesh.cuh) ----------------------------------
class Mesh:
    """One rank's end of the device-initiated exchange (``fm_mesh_*``): a persistent kernel per GPU
    stores pattern states straight into the inbox of the GPU owning the row they need next, over
    NVLink peer memory.  Nothing here touches the data path: this class creates the inbox, hands its
    CUDA IPC handle to the other ranks (``torch.distributed``), and launches one kernel per batch.

    Several ranks living in ONE process (tests on a single GPU) are wired with ``Mesh.connect_local``.
    """

    HANDLE_BYTES = 64

    def __init__(self, ix, rank: int, world: int, window: int = 0, cap_log2: int = 0, group=None,
                 connect: bool = True):
        import ctypes as C
        from . import _check
        self.ix, self.rank, self.world, self.group = ix, rank, world, group
        self.lib = ix.lib
        h = C.c_void_p()
        _check(self.lib.fm_mesh_create(ix.h, rank, world, window, cap_log2, C.byref(h)), "fm_mesh_create")
        self.h = h
        if connect and world > 1:
            buf = C.create_string_buffer(self.HANDLE_BYTES)
            _check(self.lib.fm_mesh_export(self.h, buf, self.HANDLE_BYTES), "fm_mesh_export")
            handles = [None] * world
            dist.all_gather_object(handles, bytes(buf.raw), group=group)
            blob = b"".join(handles)
            _check(self.lib.fm_mesh_connect(self.h, blob, self.HANDLE_BYTES), "fm_mesh_connect")

    @staticmethod
    def connect_local(meshes) -> None:
        """Wire the meshes of all ranks of ONE process to each other (no IPC)."""
        import ctypes as C
        from . import _check
        arr = (C.c_void_p * len(meshes))(*[m.h for m in meshes])
        for m in meshes:
            _check(m.lib.fm_mesh_connect_local(m.h, arr), "fm_mesh_connect_local")

    def set_limits(self, max_ctas: int = 0, timeout_seconds: float = 0.0) -> None:
        from . import _check
        _check(self.lib.fm_mesh_set_limits(self.h, max_ctas, timeout_seconds), "fm_mesh_set_limits")

    def launch_count(self, d_plen, d_flat, d_offs, uniform_len: int, pid_lo: int, n_mine: int, d_first, d_last,
                     stream: Optional[int] = None) -> None:
        """Asynchronous: this rank's share of the batch (global pattern ids [pid_lo, pid_lo+n_mine))."""
        from . import _check
        st = torch.cuda.current_stream().cuda_stream if stream is None else stream
        ptr = lambda t: 0 if t is None else t.data_ptr()
        _check(self.lib.fm_mesh_count(self.h, ptr(d_plen), ptr(d_flat), ptr(d_offs), uniform_len, pid_lo, n_mine,
                                      ptr(d_first), ptr(d_last), st), "fm_mesh_count")

    def launch_locate_rows(self, d_rows, d_offsets, stream: Optional[int] = None) -> None:
        from . import _check
        st = torch.cuda.current_stream().cuda_stream if stream is None else stream
        _check(self.lib.fm_mesh_locate_rows(self.h, int(d_rows.shape[0]), d_rows.data_ptr(), d_offsets.data_ptr(), st),
               "fm_mesh_locate_rows")

    def finish(self, stream: Optional[int] = None) -> dict:
        """Wait for the batch; raises if the kernel gave up.  Returns its counters."""
        import ctypes as C
        from . import _check
        st = torch.cuda.current_stream().cuda_stream if stream is None else stream
        status = C.c_int(0)
        stats = (C.c_uint64 * 8)()
        rc = self.lib.fm_mesh_finish(self.h, st, C.byref(status), stats)
        names = ["sent", "received", "rounds", "occ_pairs", "occ_singles", "empty_polls", "injected"]  # rounds = warp rounds
        out = {k: int(stats[i]) for i, k in enumerate(names)}
        if rc:
            from . import FemtoError
            msg = self.lib.fm_last_error()
            raise FemtoError(rc, "fm_mesh_finish", f"{msg.decode(errors='replace') if msg else ''} "
                                                   f"[rank {self.rank}: {out}]")
        return out

    def close(self) -> None:
        if getattr(self, "h", None):
            self.lib.fm_mesh_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def gather_uniform_batch(my_pats: torch.Tensor, world: int, group=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Replicate a batch of equal-length patterns: every rank contributes my_pats [n, m] (same n on
    every rank) and receives [world * n, m]; rank r's patterns are ids [r*n, (r+1)*n).  This
    all-gather is the only collective of a mesh batch (NCCL over NVLink; gloo in the CPU tests)."""
    if world == 1:
        return my_pats
    if out is None:  # (callers in a loop pass a buffer: a fresh device allocation per batch synchronises the device,
        #              and is slow once peer access is enabled)
        out = torch.empty((world * my_pats.shape[0], my_pats.shape[1]), dtype=my_pats.dtype, device=my_pats.device)
    # as bytes: NCCL has no 16-bit integer type
    dist.all_gather_into_tensor(out.view(torch.uint8), my_pats.contiguous().view(torch.uint8), group=group)
    return out


def gather_ragged_batch(plen: torch.Tensor, flat: torch.Tensor, world: int, group=None):
    """Replicate a batch of patterns of any lengths.  Every rank contributes its plen [n_r] and flat
    symbols; returns (plen_all, flat_all, offs_all, pid_lo) with rank r's patterns at ids
    [pid_lo_r, pid_lo_r + n_r).  Two small all-gathers size the exchange, two padded ones move it."""
    dev = plen.device
    if world == 1:
        offs = torch.cumsum(plen.to(torch.int64), 0) - plen.to(torch.int64)
        return plen, flat, offs, 0
    rank = dist.get_rank(group)
    sizes = torch.tensor([plen.shape[0], flat.shape[0]], dtype=torch.int64, device=dev)
    all_sizes = torch.empty((world, 2), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_sizes.view(-1), sizes, group=group)
    all_sizes = all_sizes.cpu()
    max_n, max_f = int(all_sizes[:, 0].max()), int(all_sizes[:, 1].max())
    pl = torch.zeros(max_n, dtype=plen.dtype, device=dev)
    pl[:plen.shape[0]] = plen
    fl = torch.zeros(max_f, dtype=flat.dtype, device=dev)
    fl[:flat.shape[0]] = flat
    pl_all = torch.empty((world, max_n), dtype=plen.dtype, device=dev)
    fl_all = torch.empty((world, max_f), dtype=flat.dtype, device=dev)
    dist.all_gather_into_tensor(pl_all.view(-1).view(torch.uint8), pl.view(torch.uint8), group=group)
    dist.all_gather_into_tensor(fl_all.view(-1).view(torch.uint8), fl.view(torch.uint8), group=group)
    plen_all = torch.cat([pl_all[r, :int(all_sizes[r, 0])] for r in range(world)])
    flat_all = torch.cat([fl_all[r, :int(all_sizes[r, 1])] for r in range(world)])
    offs_all = torch.cumsum(plen_all.to(torch.int64), 0) - plen_all.to(torch.int64)
    pid_lo = int(all_sizes[:rank, 0].sum())
    return plen_all, flat_all, offs_all, pid_lo


def mesh_locate(mesh: "Mesh", first: torch.Tensor, last: torch.Tensor, max_occs: int):
    """parallel_locate's second half over the mesh: the ranges of this rank's patterns (from a mesh
    count) expanded into rows with the reference's clip rule, every row walked to a sampled-SA mark by
    the persistent walk kernels.  Collective.  Returns (rows per pattern, offsets in pattern order)."""
    rows, cnt = expand_ranges(first, last, max_occs)
    out = torch.empty(max(int(rows.shape[0]), 1), dtype=torch.int64, device=first.device)
    mesh.launch_locate_rows(rows.contiguous(), out)
    mesh.finish()
    return cnt, out[:rows.shape[0]]
