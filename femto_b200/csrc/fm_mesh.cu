// fm_mesh.cu -- persistent kernels of the BWT-range-sharded query path (design: fm_mesh.cuh).
//
//   mesh_count_kernel : do_string_query's backward search (src/main/server.c:713-946) over an index
//                       split by data block (src/main/index.h:83-100); a pattern's state
//                       {id, first | C+Occ(c,first-1), last, i, pending symbol, phase, home} moves to the
//                       GPU that owns the row its next Occ needs and comes home with [first, last].
//   mesh_walk_kernel  : the sampled-SA walk of do_back_query / do_context_query
//                       (server.c:2228-2359, 2627-2795); state {slot, row, LF steps, home}.
//
// Both evaluate rank over quad-level blocks exactly as the single-GPU kernels do (fm_rank.cuh), so a
// state computes the same numbers wherever it is.  One round of a warp (16 lane groups, one state each):
//   1. fill the idle groups -- from the inbox slots copied to shared memory during the previous round, else
//      with new patterns of the rank's own batch;
//   2. evaluate, warp-synchronously, every state whose row(s) are resident here;
//   3. store the states that left in the previous round into their owners' inboxes (their slot indices were
//      requested then: the atomics' latency lies behind a whole round);
//   4. route every state: finished and home -> result; next row elsewhere -> leaves (slot index requested
//      now); next row here -> stays for the next round.
// Nothing in a round waits for a memory access other than the evaluation's own rank-block reads.
#include "fm_mesh.cuh"

#include <cstdlib>

#include "fm_rank.cuh"

namespace fmb {
namespace {

// Two 64-bit words with one instruction.  A vector access is a set of scalar accesses: each 64-bit
// word is read / written whole (single-copy atomic), the pair is not -- which is why EVERY word of a
// message carries the tag.
__device__ __forceinline__ void st_volatile_v2(ulonglong2* p, const ulonglong2 v) {
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(v.x), "l"(v.y) : "memory");
}
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// A state in flight: four 64-bit words, each = tag (16 bits) << 48 | 48 bits of payload:
//   word 0: id (32 bits) | meta << 32      meta = phase | home << 2 | what to evaluate next << 10
//   word 1: A    word 2: B     rows or C+Occ values, two's complement in 48 bits (B may be -1)
//   word 3: i (32 bits) | symbol of the pending step << 32
// A word is written and read whole, so a message is complete exactly when all four tags are the
// expected one -- whatever order the words arrive in over NVLink.
struct MeshState {
  int64_t A = 0, B = 0;
  uint32_t id = 0;
  int32_t i = 0;
  // bits 0-1 phase, 2-9 home rank, 10 / 11: the next evaluation covers Occ(c, A-1) / Occ(c, B) (decided by the
  // rank that routed the state: it knows the shard boundaries and the bucket size, the receiver just does it),
  // 16-31 count: the symbol of the pending step, pattern[i-1] (travels with the state: the rank that finishes a
  // step reads the next symbol WHILE it evaluates, so no rank waits for a pattern read)
  uint32_t meta = 0;
  __device__ __forceinline__ int phase() const { return static_cast<int>(meta & 3u); }
  __device__ __forceinline__ int home() const { return static_cast<int>((meta >> 2) & 0xffu); }
  __device__ __forceinline__ bool doA() const { return (meta >> 10) & 1u; }
  __device__ __forceinline__ bool doB() const { return (meta >> 11) & 1u; }
  __device__ __forceinline__ int c() const { return static_cast<int>(meta >> 16); }
  __device__ __forceinline__ void set_phase(int p) { meta = (meta & ~3u) | static_cast<uint32_t>(p); }
  __device__ __forceinline__ void set_c(int c) { meta = (meta & 0xffffu) | (static_cast<uint32_t>(c) << 16); }
  __device__ __forceinline__ void set_do(bool a, bool b) {
    meta = (meta & ~0xc00u) | (a ? 0x400u : 0u) | (b ? 0x800u : 0u);
  }
};
constexpr int kPhaseA = 0;     // count: needs Occ(c, first-1) then Occ(c, last); A = first, B = last.  walk: walking
constexpr int kPhaseB = 1;     // count: A = C[c]+Occ(c,first-1) is known, needs Occ(c, last)
constexpr int kPhaseDone = 2;  // finished: travels home; count: A = first, B = last; walk: A = offset
constexpr int kPhaseNew = 3;   // just injected (never travels): the pattern's symbols are still on their way from memory
constexpr unsigned long long kPayloadMask = (1ull << 48) - 1;
constexpr unsigned long long kTagMask = ~kPayloadMask;
constexpr unsigned long long kNoBlock = ~0ull;

__device__ __forceinline__ int64_t sext48(unsigned long long v) {
  return static_cast<int64_t>(v << 16) >> 16;
}
// the four payload words of a message (the tag is added when the slot index is known)
__device__ __forceinline__ void pack_state(const MeshState& s, ulonglong2& m0, ulonglong2& m1) {
  m0.x = static_cast<unsigned long long>(s.id) | (static_cast<unsigned long long>(s.meta & 0xfffu) << 32);
  m0.y = static_cast<unsigned long long>(s.A) & kPayloadMask;
  m1.x = static_cast<unsigned long long>(s.B) & kPayloadMask;
  m1.y = static_cast<unsigned long long>(static_cast<uint32_t>(s.i)) | (static_cast<unsigned long long>(s.meta >> 16) << 32);
}
__device__ __forceinline__ void unpack_state(MeshState& s, unsigned long long w0, unsigned long long w1,
                                             unsigned long long w2, unsigned long long w3) {
  s.id = static_cast<uint32_t>(w0);
  s.A = sext48(w1);
  s.B = sext48(w2);
  s.i = static_cast<int32_t>(static_cast<uint32_t>(w3));
  s.meta = (static_cast<uint32_t>(w0 >> 32) & 0xfffu) | (static_cast<uint32_t>(w3 >> 32) << 16);
}

// Per CTA.  stage: one 32-byte inbox slot per lane, copied in with cp.async; blk: per warp and ring the cursors
// (next unconsumed index) of its two blocks, entry [kMeshMaxRanks] = the cursors the staged copies were made
// from; peer_ring / start: kernel parameters that are indexed with a run-time value.
// out: the payload words of the state that left each lane group in the previous round (it is stored one round
// after its slot index was requested).
struct MeshShared {
  ulonglong2 stage[kThreads / 32][32][2];
  ulonglong2 peek[kThreads / 32][32];   // first two words of the slot at the cursor of block (ring = lane >> 1, lane & 1)
  ulonglong2 out[kThreads / 32][32][2];
  unsigned long long blk[kThreads / 32][kMeshMaxRanks + 1][2];
  uint4* peer_ring[kMeshMaxRanks];
  int64_t start[kMeshMaxRanks + 1];
  unsigned char owner_tab[kMeshOwnerTab];  // owner of data block b (when block_size is a power of two and nblocks fits)
  unsigned cnt[kThreads / 32][8];  // per warp: sent, received, rounds, counters 3 and 4 of the kernel, -, injected, -
};
__device__ __forceinline__ void mesh_count_up(MeshShared& sh, int wic, int lane, int k, unsigned v) {
  if (lane == 0 && v) sh.cnt[wic][k] += v;
}

struct MeshWarp {
  int lane, sub, gleader, wic;
  bool published = false;
};

// tag of ring index idx: batch field (1..255) << 8 | low 8 bits of the ring's lap, in the top 16 bits.  Never
// 0, so a cleared slot is never valid; a slot is rewritten every lap, and the rings are cleared before the
// batch field repeats (fm_api.cu), so a stale word never carries the expected tag.
__device__ __forceinline__ unsigned long long mesh_tag(const MeshWarp& w, const MeshArgs& a, unsigned long long idx) {
  return a.eptag | (((idx >> a.cap_shift) & 0xffull) << 48);
}
__device__ __forceinline__ const ulonglong2* mesh_slot(const uint4* ring, int src, unsigned long long idx, const MeshWarp& w,
                                                       const MeshArgs& a) {
  return reinterpret_cast<const ulonglong2*>(ring + ((static_cast<size_t>(src) << a.cap_shift) + (idx & a.cap_mask)) * 2);
}

// owner(row) = (row / block_size) * world / nblocks (the reference's block -> file map, partitioned): shard r
// starts at block ceil(r * nblocks / world); start[r] is that block's first row.
__device__ __forceinline__ int mesh_owner(const MeshArgs& a, const MeshShared& sh, int64_t row) {
  if (a.block_shift >= 0) return sh.owner_tab[row >> a.block_shift];
  int o = 0;
  for (int r = 1; r < a.world; r++) o += row >= sh.start[r] ? 1 : 0;
  return o;
}

// ---- consuming the inbox -----------------------------------------------------------------------
// A warp owns TWO blocks of 16 consecutive indices on every ring (the index of the next unconsumed slot of
// each is kept in shared memory).  Blocks are handed out by a ticket counter in this rank's own memory, in
// index order, to whichever warp has just used one up -- a warp that is busy takes tickets more slowly, so
// the load follows the warps' speed, and the unconsumed span of a ring never exceeds the states in flight
// plus the blocks the warps hold.  Every block a warp owns is polled (nothing arrives where nobody looks); a
// ticket is requested in one round and its block joins the polled set in the next, so the atomic's latency is
// never waited for.
//
// Reading is decoupled from consuming: at the END of its fetch a warp copies the remaining slots of the two
// blocks it owns on one ring (one block per half warp; the rings take turns) into shared memory with
// cp.async (L2 only: the slots are written by other GPUs); it looks at them at the START of its next fetch,
// after a whole round of evaluations, so taking states from the inbox costs no memory round trip on the
// round's critical path.  A slot whose four tags are not (yet) the expected ones is simply not there yet.
struct MeshInbox {
  unsigned long long pend;  // lane r < world: the ticket requested for ring r in the previous round
  int pend_which;           // which of the ring's two blocks it replaces, -1: none outstanding
  unsigned todo;            // lane r: blocks of ring r used up and not yet replaced (bit per block)
  int ring;                 // the ring whose blocks are staged (warp-uniform)
  bool staged;
};

__device__ __forceinline__ unsigned long long mesh_ticket(const MeshArgs& a, int ring) {
  return atomicAdd(&a.ctl->head_block[ring], 1ull) * kMeshBlock;
}

__device__ __forceinline__ int mesh_next_ring(const MeshArgs& a, int ring) {  // this rank's own ring is never written
  ring++;
  if (ring == a.rank) ring++;
  if (ring >= a.world) ring = a.rank == 0 ? 1 : 0;
  return ring;
}

__device__ __forceinline__ void mesh_stage_issue(const MeshWarp& w, const MeshArgs& a, MeshInbox& in, MeshShared& sh) {
  const int h = w.lane >> 4, j = w.lane & 15;
  const unsigned long long cur = sh.blk[w.wic][in.ring][h];
  if (cur != kNoBlock && j < kMeshBlock - static_cast<int>(cur & (kMeshBlock - 1))) {
    const ulonglong2* src = mesh_slot(a.ring, in.ring, cur + j, w, a);
    const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(&sh.stage[w.wic][w.lane][0]));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16), "l"(src + 1) : "memory");
  }
  // and a look at the slot under the cursor of EVERY block the warp owns (one lane per block): a block receives
  // its 16 messages in a burst when the ring's fill frontier passes it, so this tells the next fetch where to go
  if (a.world > 2 && w.lane < 2 * a.world) {
    const unsigned long long pc = sh.blk[w.wic][w.lane >> 1][w.lane & 1];
    if (pc != kNoBlock) {
      const ulonglong2* src = mesh_slot(a.ring, w.lane >> 1, pc, w, a);
      const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(&sh.peek[w.wic][w.lane]));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  if (j == 0) sh.blk[w.wic][kMeshMaxRanks][h] = cur;  // what the staged slots belong to
  in.staged = true;
}

// Fill idle groups from the slots staged by the previous fetch, then stage the next ones.  needers: ballot
// of the leader lanes of the groups without a state; updated.  Warp-collective; called every round.
__device__ __forceinline__ void mesh_take_inbox(MeshWarp& w, const MeshArgs& a, MeshInbox& in, MeshShared& sh,
                                                unsigned& needers, MeshState& s, bool& have) {
  if (a.world == 1) return;
  // tickets requested in the previous round have arrived: their blocks join the polled set; blocks still
  // waiting for a replacement request theirs
  if (w.lane < a.world && w.lane != a.rank) {
    if (in.pend_which >= 0) {
      sh.blk[w.wic][w.lane][in.pend_which] = in.pend;
      in.pend_which = -1;
    }
    if (in.todo) {
      const int wh = (in.todo & 1u) ? 0 : 1;
      in.pend = mesh_ticket(a, w.lane);
      in.pend_which = wh;
      in.todo &= ~(1u << wh);
    }
  }
  __syncwarp();
  const int want = __popc(needers);
  if (in.staged && want > 0) {
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    const int h = w.lane >> 4, j = w.lane & 15;
    const unsigned long long cur = sh.blk[w.wic][in.ring][h];
    const bool same = cur == sh.blk[w.wic][kMeshMaxRanks][h];  // (a block that joined the set since was not staged)
    bool valid = false;
    ulonglong2 m0 = make_ulonglong2(0, 0), m1 = m0;
    if (same && cur != kNoBlock && j < kMeshBlock - static_cast<int>(cur & (kMeshBlock - 1))) {
      m0 = sh.stage[w.wic][w.lane][0];
      m1 = sh.stage[w.wic][w.lane][1];
      const unsigned long long tag = mesh_tag(w, a, cur + j);
      valid = (m0.x & kTagMask) == tag && (m0.y & kTagMask) == tag && (m1.x & kTagMask) == tag &&
              (m1.y & kTagMask) == tag;  // all four words of THIS message have landed
    }
    const unsigned vm = __ballot_sync(kFull, valid);
    // messages in order from each block's cursor
    const int v0 = __ffs(~(vm & 0xffffu)) - 1, v1 = __ffs(~(vm >> 16)) - 1;
    const int take0 = min(v0, want), take1 = min(v1, want - take0);
    const int take = take0 + take1;
    if (take > 0) {
      // the k-th idle group takes the k-th message: first block 0's, then block 1's
      const int k = __popc(needers & ((1u << w.gleader) - 1u));
      const int src = k < take0 ? k : min(16 + (k - take0), 31);
      const unsigned long long w0 = __shfl_sync(kFull, m0.x, src), w1 = __shfl_sync(kFull, m0.y, src),
                               w2 = __shfl_sync(kFull, m1.x, src), w3 = __shfl_sync(kFull, m1.y, src);
      if (!have && k < take) {
        unpack_state(s, w0, w1, w2, w3);
        have = true;
      }
      needers = __ballot_sync(kFull, !have && w.sub == 0);
      mesh_count_up(sh, w.wic, w.lane, 1, take);
      __syncwarp();
      // advance the cursors; a block that is used up leaves the polled set and is replaced by a new ticket
      if (w.lane == in.ring) {
#pragma unroll
        for (int hh = 0; hh < 2; hh++) {
          const int tk = hh ? take1 : take0;
          if (tk > 0) {
            unsigned long long nx = sh.blk[w.wic][in.ring][hh] + tk;
            if ((nx & (kMeshBlock - 1)) == 0) {
              nx = kNoBlock;
              in.todo |= 1u << hh;
            }
            sh.blk[w.wic][in.ring][hh] = nx;
          }
        }
      }
      __syncwarp();
    }
    if (take0 == v0 && take1 == v1) {  // drained as far as filled: another ring
      int next = mesh_next_ring(a, in.ring);
      if (a.world > 2) {  // the first one after this whose blocks were seen to hold something, else simply the next
        bool has = false;
        if (w.lane < 2 * a.world && (w.lane >> 1) != a.rank && (w.lane >> 1) != in.ring) {
          const unsigned long long pc = sh.blk[w.wic][w.lane >> 1][w.lane & 1];
          has = pc != kNoBlock && (sh.peek[w.wic][w.lane].x & kTagMask) == mesh_tag(w, a, pc);
        }
        const unsigned lanes = __ballot_sync(kFull, has);
        if (lanes) {
          unsigned rings = 0;  // bit r: ring r has a block with data
          for (int r = 0; r < a.world; r++) rings |= ((lanes >> (2 * r)) & 3u) ? 1u << r : 0u;
          const unsigned after = rings & ~((2u << in.ring) - 1u);
          next = __ffs(after ? after : rings) - 1;
        }
      }
      in.ring = next;
    }
    in.staged = false;
  }
  if (!in.staged) mesh_stage_issue(w, a, in, sh);
}

// ---- leaving states ------------------------------------------------------------------------------
// Step 1: claim their ring indices (one atomic per destination and warp, on counters in this rank's own
// memory).  Returns this lane's index (valid when it sends).  The stores follow one round later.
__device__ __forceinline__ unsigned long long mesh_send_claim(MeshWarp& w, const MeshArgs& a, MeshShared& sh, bool& send,
                                                             int dest) {
  if (send && (dest < 0 || dest >= a.world)) {  // cannot happen with a well-formed index; never store out of bounds
    if (w.sub == 0) atomicExch(&a.ctl->status, 2);
    send = false;
  }
  const bool mine = send && w.sub == 0;
  mesh_count_up(sh, w.wic, w.lane, 0, __popc(__ballot_sync(kFull, mine)));
  // one atomic per destination and warp: the lanes that send to the same rank find each other with match.any,
  // the lowest of them claims the indices for all (same-address atomics from every lane would serialise in L2:
  // measured 3x slower at two ranks, where every state goes to the one other rank)
  const unsigned peers = __match_any_sync(kFull, mine ? dest : -1);
  const int leader = __ffs(peers) - 1;
  unsigned long long base = 0;
  if (mine && w.lane == leader) base = atomicAdd(&a.ctl->out_tail[dest], static_cast<unsigned long long>(__popc(peers)));
  base = __shfl_sync(kFull, base, leader);
  return base + __popc(peers & lanemask_lt());
}
__device__ __forceinline__ void mesh_send_store(const MeshWarp& w, const MeshArgs& a, const MeshShared& sh, int dest,
                                                unsigned long long idx) {
  if (dest >= 0 && w.sub == 0) {
    const unsigned long long tag = mesh_tag(w, a, idx);
    ulonglong2 m0 = sh.out[w.wic][w.lane][0], m1 = sh.out[w.wic][w.lane][1];
    m0.x |= tag; m0.y |= tag; m1.x |= tag; m1.y |= tag;
    ulonglong2* slot = const_cast<ulonglong2*>(mesh_slot(sh.peer_ring[dest], a.rank, idx, w, a));
    st_volatile_v2(slot, m0);
    st_volatile_v2(slot + 1, m1);
  }
}

// Results were delivered: count them (a reduction: nobody waits for the counter).  The rank's "all my
// results are home" flag is raised by the first warps that find themselves idle afterwards (mesh_idle_exit).
__device__ __forceinline__ void mesh_delivered(const MeshWarp& w, const MeshArgs& a, unsigned delivered_mask) {
  if (delivered_mask && w.lane == 0) {
    const unsigned long long n = __popc(delivered_mask);
    atomicAdd(&a.ctl->done_count, n);
    atomicAdd(&a.ctl->inflight, 0ull - n);
  }
}

// Idle warp: true when the batch is over everywhere (or has failed).  idle counts this warp's consecutive idle
// polls of about a microsecond each.
__device__ __forceinline__ bool mesh_idle_exit(MeshWarp& w, const MeshArgs& a, unsigned& idle) {
  const bool ok = w.lane >= a.world || ld_volatile_u64(&a.ctl->rank_done[w.lane]) >= a.epoch;
  if (__all_sync(kFull, ok)) return true;
  if (!w.published) {  // all results of this rank's own patterns are home: tell every rank (any idle warp may, once)
    const bool all_home = ld_volatile_u64(&a.ctl->done_count) >= static_cast<unsigned long long>(a.n_mine);
    if (all_home) {
      if (w.lane < a.world) st_volatile_u64(&a.peer_ctl[w.lane]->rank_done[a.rank], a.epoch);
      w.published = true;
    }
  }
  int st = 0;
  if (w.lane == 0) st = *reinterpret_cast<volatile int*>(&a.ctl->status);
  if (__shfl_sync(kFull, st, 0)) return true;
  if (++idle > a.timeout_polls) {
    if (w.lane == 0) atomicExch(&a.ctl->status, 1);
    return true;
  }
  __nanosleep(idle < 8 ? 100u << (idle >> 1) : 1000u);
  return false;
}

__device__ __forceinline__ void mesh_flush_stats(const MeshWarp& w, const MeshArgs& a, const MeshShared& sh) {
  __syncwarp();
  if (w.lane < 8 && sh.cnt[w.wic][w.lane]) atomicAdd(&a.ctl->stats[w.lane], static_cast<unsigned long long>(sh.cnt[w.wic][w.lane]));
}

__device__ __forceinline__ MeshWarp mesh_warp_init(const MeshArgs& a, MeshInbox& in, MeshShared& sh) {
  MeshWarp w;
  w.lane = threadIdx.x & 31;
  w.sub = w.lane & 1;
  w.gleader = w.lane & ~1;
  w.wic = threadIdx.x >> 5;
  if (threadIdx.x < (kThreads / 32) * 8) (&sh.cnt[0][0])[threadIdx.x] = 0;
  if (threadIdx.x < kMeshMaxRanks) sh.peer_ring[threadIdx.x] = a.peer_ring[threadIdx.x];
  if (threadIdx.x <= kMeshMaxRanks) sh.start[threadIdx.x] = a.shard_start[threadIdx.x];
  if (a.block_shift >= 0)
    for (int b = threadIdx.x; b < kMeshOwnerTab; b += kThreads) {
      const int64_t row = static_cast<int64_t>(b) << a.block_shift;
      int o = 0;
      for (int r = 1; r < a.world; r++) o += row >= a.shard_start[r] ? 1 : 0;
      sh.owner_tab[b] = static_cast<unsigned char>(o);
    }
  in.pend = 0;
  in.pend_which = -1;
  in.todo = 0;
  if (w.lane < kMeshMaxRanks) {
    const bool ring = w.lane < a.world && w.lane != a.rank;
    sh.blk[w.wic][w.lane][0] = ring ? mesh_ticket(a, w.lane) : kNoBlock;
    sh.blk[w.wic][w.lane][1] = ring ? mesh_ticket(a, w.lane) : kNoBlock;
  }
  // different warps start their rotation at different rings
  in.ring = a.world > 1 ? static_cast<int>((blockIdx.x * (kThreads / 32) + w.wic) % (a.world - 1)) : 0;
  if (in.ring >= a.rank) in.ring++;
  in.staged = false;
  __syncthreads();
  return w;
}

// ---- new patterns of the own batch --------------------------------------------------------------
// A warp claims batch-local ids in chunks and hands them to its idle groups over the following rounds.  The
// claim is a three-stage pipeline, one stage per round, so that no round waits for it: read the rank's
// counters -> if fewer than `window` own patterns are unfinished, add the chunk to `injected` -> use the ids.
constexpr unsigned kFeedChunk = 32;
struct MeshFeed {
  unsigned pool_next = 0, pool_end = 0;  // batch-local ids this warp may hand out (warp-uniform; ids fit 32 bits)
  unsigned long long p0 = 0;             // lane 0: what the previous stage requested (in-flight count, then the claim)
  int stage = 0;
  bool exhausted = false;
};

__device__ __forceinline__ void mesh_feed_advance(const MeshWarp& w, const MeshArgs& a, MeshFeed& f) {
  if (f.exhausted || f.pool_next < f.pool_end) return;
  const unsigned long long n = static_cast<unsigned long long>(a.n_mine);
  if (f.stage == 0) {
    if (w.lane == 0) f.p0 = ld_volatile_u64(&a.ctl->inflight);
    f.stage = 1;
  } else if (f.stage == 1) {
    int dec = 0;  // 0 window full: look again, 1 chunk requested
    if (w.lane == 0 && static_cast<long long>(f.p0) < static_cast<long long>(a.window)) {
      dec = 1;
      f.p0 = atomicAdd(&a.ctl->injected, static_cast<unsigned long long>(kFeedChunk));
      atomicAdd(&a.ctl->inflight, static_cast<unsigned long long>(kFeedChunk));
    }
    f.stage = __shfl_sync(kFull, dec, 0) ? 2 : 0;
  } else {
    const unsigned long long base = __shfl_sync(kFull, f.p0, 0);
    const unsigned long long got = base >= n ? 0ull : min(static_cast<unsigned long long>(kFeedChunk), n - base);
    if (w.lane == 0 && got < kFeedChunk)  // ids past the end of the batch are not in flight
      atomicAdd(&a.ctl->inflight, 0ull - (kFeedChunk - got));
    if (got == 0) f.exhausted = true;
    else { f.pool_next = static_cast<unsigned>(base); f.pool_end = static_cast<unsigned>(base + got); }
    f.stage = 0;
  }
}

// the batch-local id for this group, or -1.  Warp-collective.
__device__ __forceinline__ int64_t mesh_feed_take(MeshWarp& w, MeshShared& sh, MeshFeed& f, unsigned needers, bool have) {
  const unsigned avail = f.pool_end - f.pool_next;
  if (!needers || !avail) return -1;
  const unsigned take = min(static_cast<unsigned>(__popc(needers)), avail);
  const unsigned k = __popc(needers & ((1u << w.gleader) - 1u));
  const unsigned idx = f.pool_next + k;
  f.pool_next += take;
  mesh_count_up(sh, w.wic, w.lane, 6, take);
  return (!have && k < take) ? static_cast<int64_t>(idx) : -1;
}

// ---------------------------------------------------------------------------------------------
// Where a state goes next, and what is evaluated there: 0.. = that rank (possibly this one), -1 = its result
// is home (deliver).  Sets the state's doA / doB bits for the rank that evaluates.
__device__ __forceinline__ int mesh_count_route(const DevImage& im, const MeshArgs& a, const MeshShared& sh,
                                                const int64_t* s_C, MeshState& s) {
  if (s.phase() == kPhaseNew) {  // [C[c], C[c+1]-1] for the pattern's last symbol (server.c:781-801)
    const int m = s.i, c0 = static_cast<int>(s.A);
    s.set_c(static_cast<int>(s.B));
    if (m <= 0) {  // empty pattern: every row (server.c:782-808)
      s.A = 0; s.B = im.total_length - 1; s.i = 0;
    } else {
      if (c0 >= kAlphaDev) { s.A = im.total_length; s.B = s.A - 1; }  // get_C(ch>=ALPHA_SIZE), index.c:1545
      else { s.A = s_C[c0]; s.B = s_C[c0 + 1] - 1; }
      s.i = m - 1;
    }
    s.set_phase(kPhaseA);
  }
  if (s.phase() == kPhaseA && s.A <= s.B && s.i > 0 && s.c() >= kAlphaDev) {  // symbol outside the alphabet: empty range
    s.A = im.total_length; s.B = s.A - 1; s.i--;
  }
  if (s.phase() == kPhaseA && (s.A > s.B || s.i == 0)) s.set_phase(kPhaseDone);  // ends the reference's loop (server.c:832-841)
  s.set_do(false, false);
  if (s.phase() == kPhaseDone) return s.home() == a.rank ? -1 : s.home();
  if (s.phase() == kPhaseA && s.A == 0) {  // Occ(c,-1) = 0 without touching the index (server.c:847-851)
    s.A = s_C[s.c()];
    s.set_phase(kPhaseB);
  }
  const int oB = mesh_owner(a, sh, s.B);
  if (s.phase() == kPhaseB) {
    s.set_do(false, true);
    return oB;
  }
  const int64_t rowA = s.A - 1;
  const int oA = mesh_owner(a, sh, rowA);
  // both rows in one evaluation when they lie on one rank and in one bucket, else Occ(c, first-1) now and
  // Occ(c, last) in a round of its own
  const bool same = im.bucket_shift >= 0 ? (rowA >> im.bucket_shift) == (s.B >> im.bucket_shift)
                                         : rowA / im.bucket_size == s.B / im.bucket_size;
  s.set_do(true, oA == oB && same);
  return oA;
}

__device__ __forceinline__ void mesh_count_deliver(const MeshArgs& a, const MeshState& s) {
  const int64_t slot = static_cast<int64_t>(s.id) - a.pid_lo;
  if (a.last) { a.first[slot] = s.A; a.last[slot] = s.B; }
  else a.first[slot] = s.B - s.A + 1;  // parallel_count with last==NULL (femto.c:313-318)
}

// HINT: rank blocks are read with the L2 evict-first policy, so that they do not push the inbox rings and the
// in-flight patterns' symbols (both re-read within microseconds) out of L2.
template <int MINB, bool HINT>
__global__ void __launch_bounds__(kThreads, MINB) mesh_count_kernel(const DevImage im, const MeshArgs a) {
  const uint64_t pol = HINT ? l2_evict_first_policy() : 0;
  __shared__ __align__(16) MeshShared sh;
  __shared__ int64_t s_C[kAlphaDev + 1];  // C[] of the whole index (replicated header table)
  for (int t = threadIdx.x; t <= kAlphaDev; t += kThreads) s_C[t] = im.C[t];
  MeshInbox inbox;
  MeshWarp w = mesh_warp_init(a, inbox, sh);
  MeshFeed feed;
  feed.exhausted = a.n_mine == 0;
  MeshState s;
  bool have = false;
  int out_dest = -1;                 // >= 0: a state left this group in the previous round (payload in sh.out)
  unsigned long long out_idx = 0;    // its slot index (requested then)
  unsigned idle = 0;

  for (;;) {
    // ---- 1. states for the idle groups: inbox first, then new patterns of the own batch.  A new pattern's
    // symbols are requested here and looked at after the evaluations below (fresh).
    bool fresh = false;
    unsigned needers = __ballot_sync(kFull, !have && w.sub == 0);
    mesh_take_inbox(w, a, inbox, sh, needers, s, have);
    mesh_feed_advance(w, a, feed);
    {
      const int64_t k = mesh_feed_take(w, sh, feed, needers, have);
      if (k >= 0) {
        s.id = static_cast<uint32_t>(a.pid_lo + k);
        const int m = a.uniform_len > 0 ? a.uniform_len : a.plen[s.id];
        const uint16_t* pat = a.flat + (a.uniform_len > 0 ? static_cast<int64_t>(s.id) * m : a.offs[s.id]);
        s.i = m;
        s.A = m > 0 ? pat[m - 1] : 0;  // raw symbols until the routing step turns them into a range
        s.B = m > 1 ? pat[m - 2] : 0;
        s.meta = static_cast<uint32_t>(kPhaseNew) | (static_cast<uint32_t>(a.rank) << 2);
        have = true;
        fresh = true;
      }
    }
    if (!__any_sync(kFull, have || out_dest >= 0)) {
      if (mesh_idle_exit(w, a, idle)) break;
      continue;
    }
    idle = 0;
    mesh_count_up(sh, w.wic, w.lane, 2, 1);

    // ---- 2. the Occ evaluations the states ask for (their rows are resident: that is why they are here)
    {
      const bool doA = have && !fresh && s.doA(), doB = have && !fresh && s.doB();
      const bool any = doA || doB;
      if (__any_sync(kFull, any)) {
        uint32_t idxA = 0, idxB = 0, base = 0, node = 0, leaf = 0, rexit = 0;
        int L = 0;
        int64_t ob = 0;
        bool actA = false, actB = false;
        int cn = 0;  // the symbol of the step after this one, read while this one is evaluated
        if (doB && s.i >= 2) {
          const int m = a.uniform_len > 0 ? a.uniform_len : 0;
          const uint16_t* pat = a.flat + (m ? static_cast<int64_t>(s.id) * m : a.offs[s.id]);
          cn = pat[s.i - 2];
        }
        if (any) {
          int64_t g = 0, g2;
          uint32_t ra = 0, rb = 0;
          if (doA) split_row(im, s.A - 1, g, ra);
          if (doB) split_row(im, s.B, doA ? g2 : g, rb);
          const int4 rv = __ldg(reinterpret_cast<const int4*>(im.occ + g * kAlphaStride + s.c()));
          ob = rec_occ_base(rv);
          leaf = static_cast<uint32_t>(rv.z);
          rexit = static_cast<uint32_t>(rv.w);
          base = static_cast<uint32_t>(g * im.root_stride);
          node = rexit >> 4;
          idxA = ra + 1;
          idxB = rb + 1;
          if (leaf) {  // else: symbol absent from the bucket, Occ is the bucket base (index.c:2080-2089)
            L = 31 - __clz(leaf);
            actA = doA;
            actB = doB;
          }
        }
        mesh_count_up(sh, w.wic, w.lane, 3, __popc(__ballot_sync(kFull, doA && doB && w.sub == 0)));
        mesh_count_up(sh, w.wic, w.lane, 4, __popc(__ballot_sync(kFull, any && !(doA && doB) && w.sub == 0)));
        quad_descend_pair<HINT>(im, actA, actB, idxA, idxB, base, node, leaf, L, rexit, w.sub, pol);
        if (any) {
          const int64_t resA = ob + (leaf ? idxA : 0u), resB = ob + (leaf ? idxB : 0u);
          if (doA && doB) { s.A = resA; s.B = resB - 1; s.i--; s.set_phase(kPhaseA); s.set_c(cn); }
          else if (doA) { s.A = resA; s.set_phase(kPhaseB); }
          else { s.B = resB - 1; s.i--; s.set_phase(kPhaseA); s.set_c(cn); }  // A already holds C[c]+Occ(c,first-1)
        }
      }
    }

    // ---- 3. the states that left in the previous round: their slot indices have long arrived
    mesh_send_store(w, a, sh, out_dest, out_idx);
    out_dest = -1;

    // ---- 4. route every state: result, leaves, or stays
    {
      int dest = -2;  // -2: no state
      if (have) dest = mesh_count_route(im, a, sh, s_C, s);
      const bool deliver = dest == -1;
      if (deliver) {
        if (w.sub == 0) mesh_count_deliver(a, s);
        have = false;
      }
      mesh_delivered(w, a, __ballot_sync(kFull, deliver && w.sub == 0));
      bool send = dest >= 0 && dest != a.rank;
      out_idx = mesh_send_claim(w, a, sh, send, dest);
      if (send) {
        if (w.sub == 0) pack_state(s, sh.out[w.wic][w.lane][0], sh.out[w.wic][w.lane][1]);
        out_dest = dest;
        have = false;
      }
    }
  }
  mesh_flush_stats(w, a, sh);
}

// ---------------------------------------------------------------------------------------------
// Sampled-SA walks over the range-sharded index.  State: id = result slot at the home rank, A = BWT
// row (the text offset once finished, -1 for a malformed walk), i = LF steps taken so far.
// Route: the rank that must see the state next, or -1 = deliver here.
__device__ __forceinline__ int mesh_walk_route(const DevImage& im, const MeshArgs& a, const MeshShared& sh,
                                               const MeshWarp& w, MeshState& s) {
  if (s.phase() != kPhaseDone && (s.A < 0 || s.A >= im.total_length)) {  // not a row of this index
    if (w.sub == 0) atomicExch(&a.ctl->status, 2);
    s.A = -1;
    s.set_phase(kPhaseDone);
  }
  if (s.phase() == kPhaseDone) return s.home() == a.rank ? -1 : s.home();
  return mesh_owner(a, sh, s.A);
}

template <int MINB>
__global__ void __launch_bounds__(kThreads, MINB) mesh_walk_kernel(const DevImage im, const MeshArgs a) {
  __shared__ __align__(16) MeshShared sh;
  MeshInbox inbox;
  MeshWarp w = mesh_warp_init(a, inbox, sh);
  MeshFeed feed;
  feed.exhausted = a.n_mine == 0;
  MeshState s;
  bool have = false;
  int out_dest = -1;
  unsigned long long out_idx = 0;
  unsigned idle = 0;
  unsigned long long n_quad = 0, n_mark = 0, n_sample = 0;  // (rank_* helpers count in 64 bits)

  for (;;) {
    bool fresh = false;
    unsigned needers = __ballot_sync(kFull, !have && w.sub == 0);
    mesh_take_inbox(w, a, inbox, sh, needers, s, have);
    mesh_feed_advance(w, a, feed);
    {
      const int64_t k = mesh_feed_take(w, sh, feed, needers, have);
      if (k >= 0) {
        s.id = static_cast<uint32_t>(k);
        s.A = a.rows[k];  // looked at after this round's steps
        s.B = 0;
        s.i = 0;
        s.meta = static_cast<uint32_t>(kPhaseA) | (static_cast<uint32_t>(a.rank) << 2);
        have = true;
        fresh = true;
      }
    }
    if (!__any_sync(kFull, have || out_dest >= 0)) {
      if (mesh_idle_exit(w, a, idle)) break;
      continue;
    }
    idle = 0;
    mesh_count_up(sh, w.wic, w.lane, 2, 1);

    // one LF step with mark test for every walking state whose row is resident (do_back_query, server.c:2228-2359)
    {
      const bool step = have && !fresh && s.phase() == kPhaseA && s.A >= im.first_row && s.A < im.end_row;
      if (__any_sync(kFull, step)) {
        int64_t g = 0;
        uint32_t rb = 0, ch = 0, count = 0;
        uint64_t markval_base = 0;
        if (step) split_row(im, s.A, g, rb);
        quad_wtree_rank(im, step, g, rb, w.sub, ch, count, markval_base, n_quad);
        const bool ok = step && ch < static_cast<uint32_t>(kAlphaDev) && count > 0;
        int64_t occ_base = 0, offset = -1;
        mark_lookup<2, kQuadBlockWords>(im, ok, g, ch, count, markval_base, w.sub, offset, occ_base, n_mark, n_sample);
        if (step) {
          if (!ok) {
            if (w.sub == 0) atomicExch(&a.ctl->status, 2);
            s.A = -1;
            s.set_phase(kPhaseDone);
          } else if (offset >= 0) {
            s.A = offset + s.i;
            s.set_phase(kPhaseDone);
          } else if (ch <= static_cast<uint32_t>(kEscSeofDev) || s.i > (1 << 30)) {  // unmarked document start
            if (w.sub == 0) atomicExch(&a.ctl->status, 2);
            s.A = -1;
            s.set_phase(kPhaseDone);
          } else {
            s.A = occ_base + count - 1;  // LF
            s.i++;
          }
        }
      }
    }

    mesh_send_store(w, a, sh, out_dest, out_idx);
    out_dest = -1;

    {
      int dest = -2;
      if (have) dest = mesh_walk_route(im, a, sh, w, s);
      const bool deliver = dest == -1;
      if (deliver) {
        if (w.sub == 0) a.out_offset[s.id] = s.A;
        have = false;
      }
      mesh_delivered(w, a, __ballot_sync(kFull, deliver && w.sub == 0));
      bool send = dest >= 0 && dest != a.rank;
      out_idx = mesh_send_claim(w, a, sh, send, dest);
      if (send) {
        if (w.sub == 0) pack_state(s, sh.out[w.wic][w.lane][0], sh.out[w.wic][w.lane][1]);
        out_dest = dest;
        have = false;
      }
    }
  }
  mesh_count_up(sh, w.wic, w.lane, 3, static_cast<unsigned>(n_quad));
  mesh_count_up(sh, w.wic, w.lane, 4, static_cast<unsigned>(n_mark + n_sample));
  mesh_flush_stats(w, a, sh);
}

cudaError_t launch_mesh(const void* kernel, const DevImage& im, const MeshArgs& a, int sm_count, int max_ctas,
                        cudaStream_t stream, int64_t* launch_counter) {
  if (im.levels != 4) return cudaErrorInvalidValue;  // quad-level image only
  if (a.world < 1 || a.world > kMeshMaxRanks || a.cap_shift < 4 || a.cap_shift > 30) return cudaErrorInvalidValue;
  int bps = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kernel, kThreads, 0);
  if (e != cudaSuccess) return e;
  if (bps < 1) bps = 1;
  int grid = sm_count * bps;
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  // cooperative launch: every CTA of the grid is resident at once -- each warp owns blocks of the inbox, so a
  // CTA that is not running would leave its blocks unread
  DevImage im_copy = im;
  MeshArgs a_copy = a;
  void* args[2] = {&im_copy, &a_copy};
  static const bool plain = [] { const char* v = std::getenv("FEMTO_B200_MESH_COOP"); return v && v[0] == '0'; }();
  if (plain) e = cudaLaunchKernel(kernel, dim3(grid), dim3(kThreads), args, 0, stream);  // debugging aid
  else e = cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(kThreads), args, 0, stream);
  if (e != cudaSuccess) return e;
  if (launch_counter) ++*launch_counter;
  return cudaGetLastError();
}

// resident CTAs per SM the kernels are compiled for: 4 (64 registers; default) or 3 (80)
int mesh_ctas_per_sm() {
  static const int minb = [] { const char* v = std::getenv("FEMTO_B200_MESH_CTAS"); return v && v[0] == '3' ? 3 : 4; }();
  return minb;
}

}  // namespace

cudaError_t launch_mesh_count(const DevImage& im, const MeshArgs& a, int sm_count, int max_ctas, cudaStream_t stream,
                              int64_t* launch_counter) {
  static const bool hint = [] { const char* v = std::getenv("FEMTO_B200_MESH_EF"); return !(v && v[0] == '0'); }();
  const void* k3 = hint ? reinterpret_cast<const void*>(&mesh_count_kernel<3, true>)
                        : reinterpret_cast<const void*>(&mesh_count_kernel<3, false>);
  const void* k4 = hint ? reinterpret_cast<const void*>(&mesh_count_kernel<4, true>)
                        : reinterpret_cast<const void*>(&mesh_count_kernel<4, false>);
  return launch_mesh(mesh_ctas_per_sm() == 3 ? k3 : k4, im, a, sm_count, max_ctas, stream, launch_counter);
}

cudaError_t launch_mesh_walk(const DevImage& im, const MeshArgs& a, int sm_count, int max_ctas, cudaStream_t stream,
                             int64_t* launch_counter) {
  return launch_mesh(mesh_ctas_per_sm() == 3 ? reinterpret_cast<const void*>(&mesh_walk_kernel<3>)
                                             : reinterpret_cast<const void*>(&mesh_walk_kernel<4>),
                     im, a, sm_count, max_ctas, stream, launch_counter);
}

}  // namespace fmb
