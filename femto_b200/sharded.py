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
