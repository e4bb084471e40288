"""Range-sharded count across the GPUs of one box (second case).

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
is NCCL ``all_to_all_single`` over NVLink/NVSwitch (gloo on CPU for the host-logic tests); state
volume is 48 B x patterns per round, far below link bandwidth -- the cost is the ~2(m-1) rounds,
which is why batches should be large.

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
    order = torch.argsort(dest.long(), stable=True)
    send = states[order].contiguous()
    send_counts = torch.bincount(dest.long(), minlength=world).to(torch.int64)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    sc, rc = send_counts.tolist(), recv_counts.tolist()
    recv = torch.empty((sum(rc), STATE_WORDS), dtype=torch.int64, device=states.device)
    dist.all_to_all_single(recv.view(-1), send.view(-1), [c * STATE_WORDS for c in rc], [c * STATE_WORDS for c in sc],
                           group=group)
    return recv


def sharded_count(step_fn: StepFn, pid_lo: int, pid_hi: int, rank: int, world: int, device,
                  group=None, max_rounds: int = 100000) -> Tuple[torch.Tensor, torch.Tensor, int]:
    """Count patterns [pid_lo, pid_hi) (this rank's share of a batch replicated on every rank).
    Returns (first, last, rounds) for this rank's patterns."""
    n_mine = pid_hi - pid_lo
    first = torch.zeros(n_mine, dtype=torch.int64, device=device)
    last = torch.zeros(n_mine, dtype=torch.int64, device=device)
    states = new_states(pid_lo, pid_hi, rank, device)
    rounds = 0
    while True:
        dest = torch.full((states.shape[0],), rank, dtype=torch.int32, device=device)
        if states.shape[0]:
            step_fn(states, dest)
        # finished states that are already home become results
        phase = states[:, 5] & 15
        home_done = (phase == PHASE_DONE) & (dest == rank)
        if bool(home_done.any()):
            done = states[home_done]
            idx = done[:, 0] - pid_lo
            first[idx] = done[:, 1]
            last[idx] = done[:, 2]
            keep = ~home_done
            states, dest = states[keep], dest[keep]
        remaining = torch.tensor([states.shape[0]], dtype=torch.int64, device=device)
        dist.all_reduce(remaining, group=group)
        if int(remaining.item()) == 0:
            break
        states = exchange(states, dest, world, group)
        rounds += 1
        if rounds > max_rounds:
            raise RuntimeError("sharded_count did not terminate")
    return first, last, rounds


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
