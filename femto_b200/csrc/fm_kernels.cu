// fm_kernels.cu -- hand-written sm_100a kernels of the FM-index query engine.
//
// What runs here, per reference function (paths relative to the reference tree):
//   count_*_kernel : do_string_query's backward search loop (src/main/server.c:713-946) with both
//                    Occ evaluations of a step -- header_occs_request(HDR_BACK) + block_request(OCCS)
//                    (src/main/index.c:1698-1765, 1973-2100) -> wtree_occs (src/main/wtree.c:1081-1115)
//                    -> bseq_rank (wtree.c:635-763) -- as one rank-block read per wavelet-tree level.
//   walk_kernel    : do_back_query (server.c:2228-2359) = wtree_rank (wtree.c:1117-1148) + mark test +
//                    sampled-SA read (index.c:2037-2140) + LF; iterated for locate
//                    (do_context_query, server.c:2627-2795, backward half) and document extract
//                    (do_extract_document_query, server.c:6364-6437).
//   occ_kernel     : a batch of single C[ch]+Occ(ch,row) evaluations (leaf interface cross-check).
//
// Execution model.  A rank block is read by a group of LPQ adjacent lanes through the read-only path
// and popcounted under position masks, lane partials combined with __shfl_xor_sync.  Three block
// layouts (fm_image.hpp): plain (word 0 = ones before the block, then (BW-1)*32 payload bits: one
// wavelet-tree level per read), paired (two levels per read) and quad (four levels per 128-byte
// read, the default: a lane holds one 32-byte sector, fetched with ONE 256-bit load, and builds
// both words' masks from one 64-bit shift).  Warps are persistent: a lane group that finishes its
// pattern pulls the next one from a global atomic queue (one atomic per warp for all the groups that
// ask together), so dead patterns and mixed lengths do not idle lanes for the rest of the batch; a
// queue position may be gated by an arrival counter (streamed host-buffer batches).  No tensor cores:
// the work is bound by the RATE of dependent random HBM reads (profiles/r01_quad_schedules.md).
//
// Schedules of the count kernel:
//   sync  : a pattern owns LPQ lanes that advance BOTH ranks together, warp-synchronously per
//           backward-search step; when both positions fall into the same rank block (93% of the
//           iterations on the 4 GiB byte corpus) the block is read once and evaluated for both.
//           Default for every layout.
//   split : quad blocks only; one lane per Occ, the two lanes of a pattern meet once per step.
//   pair  : plain blocks only; a pattern owns 2*LPQ lanes, one sub-group per Occ (the first kernel).
// After the count of a locate call: clip_kernel / cub scan / expand_rows_kernel turn the ranges into
// rows on the device.
#include "fm_kernels.cuh"
#include "fm_rank.cuh"

#include <algorithm>
#include <climits>
#include <cstdlib>

#include <cub/device/device_scan.cuh>

namespace fmb {
namespace {

// the shard holding `row`: shard_of_block (fm_format.hpp) of its data block
__device__ __forceinline__ int shard_of_row(int64_t row, int64_t block_size, int64_t total_length, int nshards) {
  const int64_t s = ((2 * (row / block_size) + 1) * block_size * nshards) / (2 * total_length);
  return static_cast<int>(s < nshards - 1 ? s : nshards - 1);
}

// Shared by both count schedules: retire a finished pattern, pull the next one from the queue.
// "first > last || i == 0" ends the reference's while loop (server.c:832-841).
struct PatternState {
  int64_t f = 0, l = -1, pid = -1;
  int i = 0;
  const uint16_t* pat = nullptr;
  bool have = false, exhausted = false;
};

// Symbol k of a pattern.  SYM8: the batch holds raw text bytes (CountArgs::sym8), symbol = byte +
// CHARACTER_OFFSET (strtoalpha, src/main/index_types.h:85-97); s.pat is then a byte address.
template <bool SYM8>
__device__ __forceinline__ int pat_sym(const uint16_t* pat, int k) {
  if constexpr (SYM8) return static_cast<int>(reinterpret_cast<const uint8_t*>(pat)[k]) + 5;
  else return pat[k];
}

template <bool SYM8 = false>
__device__ __forceinline__ void retire_and_fetch(PatternState& s, bool can_retire, const DevImage& im,
                                                 const CountArgs& a, unsigned long long* work, int lane,
                                                 int gleader) {
  if (s.have && can_retire && (s.f > s.l || s.i == 0)) {
    if (lane == gleader) {
      if (a.last) { a.first[s.pid] = s.f; a.last[s.pid] = s.l; }
      else a.first[s.pid] = s.l - s.f + 1;  // parallel_count with last==NULL (femto.c:313-318)
    }
    s.have = false;
  }
  const bool need = !s.have && !s.exhausted;
  // ONE atomic per warp for all the groups that need a pattern (equal-length batches retire a warp's
  // patterns together, and at launch every group of the grid asks at once): the groups take
  // consecutive queue positions in lane order
  const unsigned askers = __ballot_sync(kFull, need && lane == gleader);
  unsigned long long idx = 0;
  if (askers) {
    const int first_asker = __ffs(askers) - 1;
    if (lane == first_asker) idx = atomicAdd(work, static_cast<unsigned long long>(__popc(askers)));
    idx = __shfl_sync(kFull, idx, first_asker) + __popc(askers & ((1u << gleader) - 1u));
  }
  if (need) {
    bool arrived = static_cast<int64_t>(idx) < a.npats;
    if (arrived && a.avail) {  // streamed batch: wait until the copy stream has delivered this pattern
      // ld.acquire.sys: the loads of plen / offs / symbols below are ordered AFTER the load that
      // observes the arrival mark (a plain or volatile load would allow them to be satisfied first;
      // their addresses are known before the spin), and the mark is observed at system scope, where
      // the copy stream's write (cuStreamWriteValue64 behind the chunk's copies) lands
      for (unsigned spins = 0; ld_acquire_sys(a.avail) <= idx; spins++) {
        __nanosleep(400);
        // Give up after ~0.1 s without touching the pattern (its bytes may not be there): the host
        // then repeats the batch with the copies ahead of the kernel.  This is what happens under a
        // profiler that serialises the streams, and it keeps a broken copy stream from hanging the GPU.
        if (spins > (1u << 18) || *reinterpret_cast<volatile int32_t*>(a.stalled)) {
          atomicExch(a.stalled, 1);
          arrived = false;
          break;
        }
      }
    }
    if (arrived) {
      s.pid = static_cast<int64_t>(idx);
      const int m = a.uniform_len > 0 ? a.uniform_len : a.plen[s.pid];
      const int64_t so = a.uniform_len > 0 ? s.pid * m : a.offs[s.pid];
      s.pat = SYM8 ? reinterpret_cast<const uint16_t*>(reinterpret_cast<const uint8_t*>(a.flat) + so) : a.flat + so;
      if (m <= 0) {  // empty pattern: every row (server.c:782-808)
        s.f = 0; s.l = im.total_length - 1; s.i = 0;
      } else {
        const int c = pat_sym<SYM8>(s.pat, m - 1);
        if (c >= kAlphaDev) { s.f = im.total_length; s.l = s.f - 1; }  // get_C(ch>=ALPHA_SIZE), index.c:1545
        else { s.f = __ldg(im.C + c); s.l = __ldg(im.C + c + 1) - 1; }
        s.i = m - 1;
      }
      s.have = true;
    } else {
      s.exhausted = true;
    }
  }
}

__device__ __forceinline__ void flush_stats(unsigned long long* stats, int lane, unsigned long long a,
                                            unsigned long long b, unsigned long long c, unsigned long long d) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(kFull, a, o);
    b += __shfl_xor_sync(kFull, b, o);
    c += __shfl_xor_sync(kFull, c, o);
    d += __shfl_xor_sync(kFull, d, o);
  }
  if (lane == 0) {
    atomicAdd(stats + 0, a);
    atomicAdd(stats + 1, b);
    atomicAdd(stats + 2, c);
    atomicAdd(stats + 3, d);
  }
}

// ---------------------------------------------------------------------------------------------
// count, "pair" schedule
template <int LPQ, int BW, bool STATS>
__global__ void __launch_bounds__(kThreads) count_pair_kernel(const DevImage im, const CountArgs a,
                                                               unsigned long long* __restrict__ work,
                                                               unsigned long long* __restrict__ stats) {
  unsigned long long n_reads = 0, n_distinct = 0, n_occ = 0, n_steps = 0;
  constexpr int GL = 2 * LPQ;  // lanes per pattern
  const int lane = threadIdx.x & 31;
  const int sub = lane & (LPQ - 1);
  const int which = (lane / LPQ) & 1;  // 0: Occ(c, first-1)   1: Occ(c, last)
  const int gleader = lane & ~(GL - 1);
  PatternState s;

  for (;;) {
    retire_and_fetch(s, true, im, a, work, lane, gleader);
    if (!__any_sync(kFull, s.have)) break;

    const bool stepping = s.have && s.f <= s.l && s.i > 0;
    int c = 0;
    int64_t row = 0;
    bool q = false, badc = false;
    if (stepping) {
      c = s.pat[s.i - 1];
      badc = c >= kAlphaDev;
      row = which ? s.l : s.f - 1;
      q = !badc && row >= 0;  // first == 0: Occ(c,-1) = 0 without touching the index (server.c:847-851)
    }
    int64_t r = occ_descend<LPQ, BW, STATS>(im, q, c, row, sub, &n_reads, &n_distinct);
    if (STATS && sub == 0) {
      n_occ += q ? 1 : 0;
      n_steps += (stepping && which == 0) ? 1 : 0;
    }
    if (stepping && !q && !badc) r = __ldg(im.C + c);
    const int64_t other = __shfl_xor_sync(kFull, r, LPQ);
    if (stepping) {
      if (badc) {
        s.f = im.total_length; s.l = s.f - 1;
      } else {
        s.f = which ? other : r;
        s.l = (which ? r : other) - 1;
      }
      s.i--;
    }
  }
  if (STATS) flush_stats(stats, lane, n_reads, n_distinct, n_occ, n_steps);
}

// ---------------------------------------------------------------------------------------------
// The 64 region bytes of one quad-level block held by ONE lane (split schedule below).
struct QuadLine {
  uint32_t r[4][4];  // r[l][w]: word w of the block's level-l region
  __device__ __forceinline__ void load(const uint4* __restrict__ blocks, uint32_t blk) {
    const uint4* b = blocks + static_cast<size_t>(blk) * (kQuadBlockWords / 4) + 4;
    const uint4 x0 = __ldg(b), x1 = __ldg(b + 1), x2 = __ldg(b + 2), x3 = __ldg(b + 3);
    // word w of region l is block word 16 + 8 (w >> 1) + 2 l + (w & 1)   (fm_image.hpp)
    r[0][0] = x0.x; r[0][1] = x0.y; r[1][0] = x0.z; r[1][1] = x0.w;
    r[2][0] = x1.x; r[2][1] = x1.y; r[3][0] = x1.z; r[3][1] = x1.w;
    r[0][2] = x2.x; r[0][3] = x2.y; r[1][2] = x2.z; r[1][3] = x2.w;
    r[2][2] = x3.x; r[2][3] = x3.y; r[3][2] = x3.z; r[3][3] = x3.w;
  }
  // ones of region L at positions >= x (inv == 0) or < x (inv == ~0)
  template <int L>
  __device__ __forceinline__ uint32_t one_sided(int x, uint32_t inv) const {
    uint32_t c = 0;
#pragma unroll
    for (int w = 0; w < 4; w++) c += __popc(r[L][w] & (shr_clamp(kFull, max(x - 32 * w, 0)) ^ inv));
    return c;
  }
  // ones of region L within [a, e), 0 <= a <= e <= kQuadPos
  template <int L>
  __device__ __forceinline__ uint32_t range(int a, int e) const {
    uint32_t c = 0;
#pragma unroll
    for (int w = 0; w < 4; w++)
      c += __popc(r[L][w] & shr_clamp(kFull, max(a - 32 * w, 0)) & ~shr_clamp(kFull, max(e - 32 * w, 0)));
    return c;
  }
  // All four levels for one position: j = positions of the block's stretch up to and including
  // ours, nib = the path, h = H[nib & ~1], H[nib | 1].  Returns the 1-based index at exit nib.
  __device__ __forceinline__ uint32_t eval(const uint2 h, uint32_t nib, int j) const {
    uint32_t c = one_sided<0>(j, kFull);
    uint32_t b = (nib >> 3) & 1u;
    j = b ? c : j - c;
    c = one_sided<1>(b ? kQuadPos - j : j, b ? 0u : kFull);  // child 1 is stored backward from the end
    b = (nib >> 2) & 1u;
    j = b ? c : j - c;
    int a = static_cast<int>(h.x >> 24) - (b ? j : 0);
    c = range<2>(a, a + j);
    b = (nib >> 1) & 1u;
    j = b ? c : j - c;
    a = static_cast<int>(h.y >> 24) - (b ? j : 0);
    c = range<3>(a, a + j);
    b = nib & 1u;
    j = b ? c : j - c;
    return ((b ? h.y : h.x) & 0xffffffu) + static_cast<uint32_t>(j);
  }
};

// ---------------------------------------------------------------------------------------------
// count over quad-level blocks, one lane per Occ ("split" schedule).
//
// A pattern owns two adjacent lanes: lane 0 evaluates C[c] + Occ(c, first-1), lane 1 evaluates
// C[c] + Occ(c, last).  Each lane runs a complete, independent Occ: its own row, bucket, record
// and blocks, all four levels of a block from the 16 registers it loaded itself.  Consequences:
//   * no shuffle and no packed arithmetic inside the descent -- the lanes meet once per step, to
//     exchange the two results;
//   * one block evaluation per lane and iteration instead of two per lane pair;
//   * rows in different buckets or different blocks need no special case;
//   * when both rows lie in the same block (the common case once the range is narrow) the two
//     lanes load the same addresses and the hardware merges them into one line request.
// The root block of a step is addressed from the row alone (root area, fm_image.hpp) and requested
// together with the (bucket, symbol) record; the next symbol is fetched one step ahead.
//
// Measured (profiles/r01_quad_schedules.md): 1.06 G warp instructions per 1 Mi-pattern launch
// against 1.72 G of the two-lane schedule, but 237 M patterns/s against 360 M: the header entry
// and the exit entry can only be requested once the record has arrived, so a step still pays
// record -> header in series, and four 128-bit loads per lane and block keep the load/store unit's
// queue full (MIO / LG throttle 17 % of the stall samples).  Kept selectable
// (fm_set_count_schedule(ix, 1, 1)) as the simpler formulation; not the default.
template <int MINB, bool STATS, int THREADS>
__global__ void __launch_bounds__(THREADS, MINB) count_quad_split_kernel(const DevImage im, const CountArgs a,
                                                                          unsigned long long* __restrict__ work,
                                                                          unsigned long long* __restrict__ stats) {
  unsigned long long n_ranks = 0, n_blocks = 0, n_occ = 0, n_steps = 0;
  const int lane = threadIdx.x & 31;
  const int sub = lane & 1;
  const int gleader = lane & ~1;
  const uint2* __restrict__ exits = reinterpret_cast<const uint2*>(im.quads);
  PatternState s;
  int c = 0, cn = 0;

  for (;;) {
    {
      const int64_t pid0 = s.pid;
      retire_and_fetch(s, true, im, a, work, lane, gleader);
      if (s.have && s.pid != pid0 && s.i > 0) cn = s.pat[s.i - 1];  // a new pattern: its first symbol to extend by
    }
    if (!__any_sync(kFull, s.have)) break;

    // ---- set-up of one backward-search step: this lane's row, record and root block
    bool stepping = s.have && s.f <= s.l && s.i > 0;
    bool act = false;
    uint32_t idx = 0, base = 0, node = 0, leaf = 0, rexit = 0;
    int L = 0;
    int64_t ob = 0;
    QuadLine line;
    if (stepping) {
      c = cn;
      if (s.i >= 2) cn = s.pat[s.i - 2];  // the next step's symbol, one step ahead
      if (STATS && sub == 0) n_steps++;
      if (c >= kAlphaDev) {  // symbol outside the alphabet: empty range
        s.f = im.total_length; s.l = s.f - 1; s.i--;
        stepping = false;
      } else {
        const int64_t row = sub ? s.l : s.f - 1;
        if (row < 0) {  // Occ(c,-1) = 0 without touching the index (server.c:847-851)
          ob = __ldg(im.C + c);
        } else {
          int64_t g;
          uint32_t rb;
          split_row(im, row, g, rb);
          base = static_cast<uint32_t>(g * im.root_stride);
          line.load(im.blocks, base + (rb >> 7));  // independent of the record read below
          const int4 rv = __ldg(reinterpret_cast<const int4*>(im.occ + g * kAlphaStride + c));
          ob = rec_occ_base(rv);
          leaf = static_cast<uint32_t>(rv.z);
          node = static_cast<uint32_t>(rv.w) >> 4;
          rexit = static_cast<uint32_t>(rv.w);
          if (STATS) n_occ++;
          if (leaf) {  // else: symbol absent from the bucket, Occ is the bucket base (index.c:2080-2089)
            L = 31 - __clz(leaf);
            idx = rb + 1;
            act = true;
          }
        }
      }
    }

    // ---- descend: four wavelet-tree levels per iteration
    int lvl = 0;
    while (__any_sync(kFull, act)) {
      const uint32_t p = act ? idx - 1 : 0u;
      const uint32_t blk = base + (p >> 7);
      if (STATS) {
        const uint32_t oblk = __shfl_xor_sync(kFull, blk, 1);
        const bool oact = __shfl_xor_sync(kFull, act ? 1 : 0, 1) != 0;
        if (act) {
          n_ranks++;
          n_blocks += (sub == 1 && oact && oblk == blk) ? 0 : 1;
        }
      }
      if (act) {
        const uint32_t nib = quad_path(leaf, L, lvl);
        uint2 ex = make_uint2(0, 0);
        if (lvl + 4 < L) {
          if (lvl == 0 && (rexit & kRootExitDirect)) ex = make_uint2(rexit & ~kRootExitDirect, 0u);
          else ex = ldg_pinned(exits + (static_cast<size_t>(node) * 16 + nib));
        }
        const uint2 h = ldg_pinned(reinterpret_cast<const uint2*>(im.blocks + static_cast<size_t>(blk) * (kQuadBlockWords / 4)) + (nib >> 1));
        if (lvl > 0) line.load(im.blocks, blk);  // the root block was requested during set-up
        idx = line.eval(h, nib, static_cast<int>(p & 127u) + 1);
        act = idx != 0 && lvl + 4 < L;
        base = ex.x;
        node = ex.y;
      }
      lvl += 4;
    }
    // lane 0 now holds the new first, lane 1 the new last + 1
    ob += idx;
    const int64_t other = __shfl_xor_sync(kFull, ob, 1);
    if (stepping) {
      s.f = sub ? other : ob;
      s.l = (sub ? ob : other) - 1;
      s.i--;
    }
  }
  if (STATS) flush_stats(stats, lane, n_ranks, n_blocks, n_occ, n_steps);
}

// ---------------------------------------------------------------------------------------------
// count, "sync" schedule: one group of LPQ lanes per pattern advances both ranks of a step.
// EXP (measurement variants of the quad branch, profiles/r01_quad_schedules.md; results unchanged):
//   1 = position B always re-reads its line, also when it is A's (an L1 hit; the first version),
//   2 = 150 extra dependent ALU instructions per iteration (issue / ALU sensitivity).
template <int LPQ, int BW, int MINB, bool STATS, int LV = 1, int EXP = 0, bool SYM8 = false>
__global__ void __launch_bounds__(kThreads, MINB) count_sync_kernel(const DevImage im, const CountArgs a,
                                                                     unsigned long long* __restrict__ work,
                                                                     unsigned long long* __restrict__ stats) {
  constexpr uint32_t BITS = (BW - 1) * 32;
  unsigned long long n_ranks = 0, n_blocks = 0, n_occ = 0, n_steps = 0;
  unsigned long long n_slots = 0, n_active = 0;  // lane groups inside descent iterations: all of them / those with a block to evaluate
  const int lane = threadIdx.x & 31;
  const int sub = lane & (LPQ - 1);
  const int gleader = lane & ~(LPQ - 1);
  PatternState s;
  bool cross_pending = false;  // row `last` lies in another bucket than `first-1`: its descent is the next round
  int c = 0;
  int64_t obA = 0, obB = 0;    // Occ bases; become C[c]+Occ(c,first-1) and C[c]+Occ(c,last)

  for (;;) {
    retire_and_fetch<SYM8>(s, !cross_pending, im, a, work, lane, gleader);
    if (!__any_sync(kFull, s.have)) break;

    // ---- set-up of one round (normally a whole backward-search step), all groups together
    bool stepping = s.have && (cross_pending || (s.f <= s.l && s.i > 0));
    bool actA = false, actB = false;
    uint32_t idxA = 0, idxB = 0, base = 0, node = 0, leaf = 0, rexit = 0;
    int L = 0;
    if (stepping) {
      int64_t g;
      uint32_t rb;
      split_row(im, s.l, g, rb);
      bool hasA = false, cross = false;
      uint32_t rbA = 0;
      if (!cross_pending) {
        c = pat_sym<SYM8>(s.pat, s.i - 1);
        if (STATS && sub == 0) n_steps++;
        if (c >= kAlphaDev) {  // symbol outside the alphabet: empty range
          s.f = im.total_length; s.l = s.f - 1; s.i--;
          stepping = false;
        } else if (s.f != 0) {
          int64_t gA;
          split_row(im, s.f - 1, gA, rbA);
          hasA = true;
          cross = gA != g;  // first-1 sits in another bucket: this round does A, the next one B
          if (cross) g = gA;
        }
      }
      if (stepping) {
        // the symbol's record; the bucket's root comes with it (quad: root blocks are addressed by
        // row, the root QuadRec is named by the record) or from the bucket record (other layouts)
        const int4 rv = __ldg(reinterpret_cast<const int4*>(im.occ + g * kAlphaStride + c));
        uint4 br = make_uint4(static_cast<uint32_t>(g * im.root_stride), static_cast<uint32_t>(rv.w) >> 4, 0, 0);
        if (LV != 4) br = __ldg(reinterpret_cast<const uint4*>(im.buckets + g));
        if (LV == 4) rexit = static_cast<uint32_t>(rv.w);
        const int64_t ob = rec_occ_base(rv);
        leaf = static_cast<uint32_t>(rv.z);
        if (cross_pending) {  // second round of a cross-bucket step: row `last`
          obB = ob;
          idxB = rb + 1;
          actB = true;
          cross_pending = false;
          if (STATS && sub == 0) n_occ++;
        } else if (!hasA) {   // Occ(c,-1) = 0 without touching the index (server.c:847-851)
          obA = __ldg(im.C + c);
          obB = ob;
          idxB = rb + 1;
          actB = true;
          if (STATS && sub == 0) n_occ++;
        } else {
          obA = ob;
          idxA = rbA + 1;
          actA = true;
          if (STATS && sub == 0) n_occ += cross ? 1 : 2;
          if (cross) {
            cross_pending = true;
          } else {
            obB = ob;
            idxB = rb + 1;
            actB = true;
          }
        }
        if (leaf == 0) {  // symbol absent from the bucket: Occ is the bucket base (index.c:2080-2089)
          actA = actB = false;
        } else {
          base = br.x;
          node = br.y;
          L = 31 - __clz(leaf);
        }
      }
    }
    const bool jobA = actA, jobB = actB;
    const bool finish = stepping && !cross_pending;

    // ---- descend: one level (two with paired-level blocks) per iteration for every group
    int lvl = 0;
    if constexpr (LV == 4) {
      static_assert(LV != 4 || (LPQ == 2 && BW == kQuadBlockWords), "quad-level blocks: 2 lanes, 128 bytes");
      while (__any_sync(kFull, actA || actB)) {
        const bool any = actA || actB;
        const uint32_t pA = actA ? idxA - 1 : 0u, pB = actB ? idxB - 1 : 0u;
        const uint32_t kA = pA >> 7, kB = pB >> 7;
        const bool two = actA && actB && kA != kB;
        const uint32_t blkA = base + (actA ? kA : kB), blkB = base + (actB ? kB : kA);
        const uint32_t nib = any ? quad_path(leaf, L, lvl) : 0u;
        // p serves position A, q position B; q is read only when B lies in another block -- the
        // load / L1 path, not issue slots, is what this kernel saturates first
        QuadWords p, q;
        p.clear();
        q.clear();
        uint2 hp = make_uint2(0, 0), hq = hp, ex = hp;
        if (any) {
          p.load(im.blocks, blkA, sub);
          hp = quad_header(im.blocks, blkA, nib >> 1);
          if (EXP == 1 || two) {
            q.load(im.blocks, blkB, sub);
            hq = quad_header(im.blocks, blkB, nib >> 1);
          }
          if (lvl + 4 < L) {
            // codes of 5..8 bits carry the first block of their second (and last) quad node in the record
            if (lvl == 0 && (rexit & kRootExitDirect)) ex = make_uint2(rexit & ~kRootExitDirect, 0u);
            else ex = __ldg(reinterpret_cast<const uint2*>(im.quads[node].exit[nib]));
          }
        }
        if (EXP == 2) {
          uint32_t x = p.d[0] ^ nib;
#pragma unroll
          for (int t = 0; t < 50; t++) x = (x ^ (x >> 7)) + 0x9e3779b9u;
          if (x == 0x12345u && lvl > 1000) actA = false;  // never true; keeps the chain alive
        }
        const int lb = 64 * sub;
        int jA = static_cast<int>(pA & 127u) + 1, jB = static_cast<int>(pB & 127u) + 1;
        uint32_t b;
        if (EXP == 1 || __any_sync(kFull, two)) {  // some pattern of the warp needs two blocks (rare once ranges are narrow)
          if (EXP != 1 && !two) {
            hq = hp;
#pragma unroll
            for (int t = 0; t < 8; t++) q.d[t] = p.d[t];
          }
          b = quad_eval_pair(p, q, hp, hq, nib, lb, jA, jB);
        } else {                                   // every pattern evaluates both positions from one line
          hq = hp;
          b = quad_eval_pair(p, p, hp, hp, nib, lb, jA, jB);
        }
        if (STATS && any && sub == 0) {
          n_blocks += two ? 2 : 1;
          n_ranks += (actA ? 1 : 0) + (actB ? 1 : 0);
        }
        if (STATS && sub == 0) {
          n_slots++;
          n_active += any ? 1 : 0;
        }
        lvl += 4;
        if (any) {
          if (actA) {
            idxA = ((b ? hp.y : hp.x) & 0xffffffu) + static_cast<uint32_t>(jA);
            actA = idxA != 0 && lvl < L;
          }
          if (actB) {
            idxB = ((b ? hq.y : hq.x) & 0xffffffu) + static_cast<uint32_t>(jB);
            actB = idxB != 0 && lvl < L;
          }
          base = ex.x;
          node = ex.y;
        }
      }
    } else if constexpr (LV == 2) {
      constexpr uint32_t B = kPairedSlicePos * (BW / kPairedSliceWords);
      while (__any_sync(kFull, actA || actB)) {
        const bool any = actA || actB;
        uint32_t kA, kB, offA, offB;
        paired_split<BW>(actA ? idxA - 1 : 0u, kA, offA);
        paired_split<BW>(actB ? idxB - 1 : 0u, kB, offB);
        const bool two = actA && actB && kA != kB;
        const bool has2 = lvl + 2 <= L;
        const uint32_t b1 = any ? (leaf >> (L - lvl - 1)) & 1u : 0u;
        const uint32_t b2 = (any && has2) ? (leaf >> (L - lvl - 2)) & 1u : 0u;
        BlockWords<LPQ, BW> p, q;
        p.clear();
        q.clear();
        if (any) {
          p.load(im.blocks, base + (actA ? kA : kB), sub);
          q.load(im.blocks, base + (actB ? kB : kA), sub);
        }
        uint2 gc = make_uint2(0, 0);
        if (any && lvl + 2 < L) gc = __ldg(reinterpret_cast<const uint2*>(im.supers[node].gc[2 * b1 + b2]));
        // level one: ranks in the super node
        const uint32_t px = group_sum<LPQ>(paired_count<LPQ, BW, kPairedX>(p, static_cast<int>(offA) + 1, sub) |
                                           (paired_count<LPQ, BW, kPairedX>(q, static_cast<int>(offB) + 1, sub) << 16));
        const uint32_t cxA = px & 0xffffu, cxB = px >> 16;
        const uint32_t onesA = p.w[0] + cxA, onesB = q.w[0] + cxB;
        const uint32_t i1A = b1 ? onesA : idxA - onesA;  // wtree.c:1109-1110
        const uint32_t i1B = b1 ? onesB : idxB - onesB;
        // level two: ranks in child b1 over the bits stored in the same block
        const uint32_t jA = b1 ? cxA : offA + 1 - cxA, jB = b1 ? cxB : offB + 1 - cxB;
        const uint32_t hiA = b1 ? B - jA : jA, hiB = b1 ? B - jB : jB;
        const uint32_t pr = group_sum<LPQ>(paired_count<LPQ, BW, kPairedR>(p, static_cast<int>(hiA), sub) |
                                           (paired_count<LPQ, BW, kPairedR>(q, static_cast<int>(hiB), sub) << 16));
        const uint32_t h1A = paired_h1<LPQ, BW>(p, b1), h1B = paired_h1<LPQ, BW>(q, b1);
        const uint32_t o2A = b1 ? h1A - (pr & 0xffffu) : h1A + (pr & 0xffffu);
        const uint32_t o2B = b1 ? h1B - (pr >> 16) : h1B + (pr >> 16);
        if (STATS && any && sub == 0) {
          n_blocks += two ? 2 : 1;
          n_ranks += (actA ? 1 : 0) + (actB ? 1 : 0);
        }
        lvl += 2;
        if (any) {
          if (actA) {
            idxA = (i1A == 0) ? 0u : (!has2 ? i1A : (b2 ? o2A : i1A - o2A));
            actA = idxA != 0 && lvl < L;
          }
          if (actB) {
            idxB = (i1B == 0) ? 0u : (!has2 ? i1B : (b2 ? o2B : i1B - o2B));
            actB = idxB != 0 && lvl < L;
          }
          base = gc.x;
          node = gc.y;
        }
      }
    } else
    while (__any_sync(kFull, actA || actB)) {
      const bool any = actA || actB;
      const uint32_t pA = actA ? idxA - 1 : 0u;
      const uint32_t pB = actB ? idxB - 1 : 0u;
      const uint32_t kA = pA / BITS, kB = pB / BITS;
      const uint32_t offA = pA - kA * BITS, offB = pB - kB * BITS;
      const bool two = actA && actB && kA != kB;  // the two positions need different blocks
      // p serves position A, q position B.  When both positions share a block (the common case) or
      // only one is active, q re-reads p's line: an L1 hit, no second HBM access.
      BlockWords<LPQ, BW> p, q;
      p.clear();
      q.clear();
      if (any) {
        p.load(im.blocks, base + (actA ? kA : kB), sub);
        q.load(im.blocks, base + (actB ? kB : kA), sub);
      }
      uint4 nr = make_uint4(0, 0, 0, 0);
      if (any && lvl + 1 < L) nr = __ldg(reinterpret_cast<const uint4*>(im.nodes + node));
      const uint32_t packed = group_sum<LPQ>(p.count_upto(offA, sub) | (q.count_upto(offB, sub) << 16));
      const uint32_t onesA = group_lane0<LPQ>(p.w[0]) + (packed & 0xffffu);
      const uint32_t onesB = group_lane0<LPQ>(q.w[0]) + (packed >> 16);
      if (STATS && any && sub == 0) {
        n_blocks += two ? 2 : 1;
        n_ranks += (actA ? 1 : 0) + (actB ? 1 : 0);
      }
      lvl++;
      if (any) {
        const uint32_t b = (leaf >> (L - lvl)) & 1u;
        if (actA) { idxA = b ? onesA : idxA - onesA; actA = idxA != 0 && lvl < L; }  // wtree.c:1109-1110
        if (actB) { idxB = b ? onesB : idxB - onesB; actB = idxB != 0 && lvl < L; }
        base = b ? nr.y : nr.x;
        node = b ? nr.w : nr.z;
      }
    }
    if (jobA) obA += idxA;
    if (jobB) obB += idxB;
    if (finish) { s.f = obA; s.l = obB - 1; s.i--; }
  }
  if (STATS) {
    flush_stats(stats, lane, n_ranks, n_blocks, n_occ, n_steps);
    flush_stats(stats + 4, lane, n_slots, n_active, 0, 0);
  }
}

// ---------------------------------------------------------------------------------------------
// MINB: resident CTAs per SM the kernel is compiled for (0 = the compiler's choice, 64 registers = 4 CTAs).  The
// walk is a chain of dependent reads with little arithmetic between them: it is bound by how many chains an SM
// holds (ncu: 82 % of the warp cycles wait for a load, issue slots 24 % busy at 4 CTAs), so the locate kernel
// trades registers for warps.
template <int LPQ, int BW, int MODE, int LV = 1, int MINB = 0>
__global__ void __launch_bounds__(kThreads, MINB) walk_kernel(const DevImage im, const WalkArgs a,
                                                         unsigned long long* __restrict__ work) {
  constexpr uint32_t BITS = (BW - 1) * 32;
  const int lane = threadIdx.x & 31;
  const int sub = lane & (LPQ - 1);
  const int gleader = lane & ~(LPQ - 1);

  int64_t row = 0, rid = -1, steps = 0, nsteps = 0, sym_off = 0, meta = 0;
  bool have = false, exhausted = false;
  // counters of the instrumented launches (WalkArgs::stats): LF steps, wavelet-tree blocks, mark blocks, SA samples
  unsigned long long n_steps = 0, n_quad = 0, n_mark = 0, n_sample = 0;
  // kWalkShard: park the state for the rank that must see it next
  auto park = [&](int64_t value, int64_t phase, int dest) {
    if (lane == gleader) {
      int64_t* st = a.state + rid * kWalkStateWords;
      st[1] = value;
      st[2] = steps;
      st[3] = (meta & ~int64_t(15)) | phase;
      a.dest[rid] = dest;
    }
  };

  for (;;) {
    const bool need = !have && !exhausted;
    // ONE atomic per warp for all the groups that need a row: consecutive queue positions in lane order
    const unsigned askers = __ballot_sync(kFull, need && lane == gleader);
    unsigned long long idx = 0;
    if (askers) {
      const int first_asker = __ffs(askers) - 1;
      if (lane == first_asker) idx = atomicAdd(work, static_cast<unsigned long long>(__popc(askers)));
      idx = __shfl_sync(kFull, idx, first_asker) + __popc(askers & ((1u << gleader) - 1u));
    }
    if (need) {
      if (static_cast<int64_t>(idx) < a.nrows) {
        rid = static_cast<int64_t>(idx);
        have = true;
        if (MODE == kWalkShard) {
          const int64_t* st = a.state + rid * kWalkStateWords;
          row = st[1];
          steps = st[2];
          meta = st[3];
          if ((meta & 15) == 2) {  // already finished: on its way home
            if (lane == gleader) a.dest[rid] = static_cast<int>(meta >> 4);
            have = false;
          }
        } else {
          row = a.rows[rid];
          steps = 0;
        }
        if (MODE == kWalkExtract) {
          nsteps = a.nsteps[rid];
          sym_off = a.sym_off[rid];
          if (nsteps <= 0) have = false;
        }
      } else {
        exhausted = true;
      }
    }
    if (!__any_sync(kFull, have || !exhausted)) break;

    bool act = have;
    if (act && (row < im.first_row || row >= im.end_row)) {
      if (MODE == kWalkShard && row >= 0 && row < im.total_length) {
        // the row lives on another rank: the state travels there
        park(row, 0, shard_of_row(row, a.block_size, im.total_length, a.nshards));
      } else if (MODE == kWalkShard) {
        if (lane == gleader) atomicExch(a.status, 1);
        park(-1, 2, static_cast<int>(meta >> 4));
      } else if (lane == gleader) {
        atomicExch(a.status, 1);
        if (MODE != kWalkExtract) a.out_offset[rid] = -1;
      }
      have = false;
      act = false;
    }

    // wtree_rank: descend by the bit found at each node (wtree.c:1117-1148)
    int64_t g = 0;
    uint32_t rb = 0, base = 0, node = 0, idx1 = 0, ch = 0;
    uint64_t markval_base = 0;
    if (act) {
      split_row(im, row, g, rb);
      if (LV != 4) {  // (quad: the root block is addressed by the row; quad_wtree_rank reads the record beside it)
        const uint4 br = __ldg(reinterpret_cast<const uint4*>(im.buckets + g));
        base = br.x;
        node = br.y;
        markval_base = static_cast<uint64_t>(br.z) | (static_cast<uint64_t>(br.w) << 32);
      }
      idx1 = rb + 1;
    }
    bool desc = act;
    (void)desc; (void)base; (void)node;
    if (act && lane == gleader) n_steps++;
    if constexpr (LV == 4) {
      static_assert(LV != 4 || (LPQ == 2 && BW == kQuadBlockWords), "quad-level blocks: 2 lanes, 128 bytes");
      quad_wtree_rank(im, act, g, rb, sub, ch, idx1, markval_base, n_quad);
    } else if constexpr (LV == 2) {
      constexpr uint32_t B = kPairedSlicePos * (BW / kPairedSliceWords);
      while (__any_sync(kFull, desc)) {
        uint32_t k, off;
        paired_split<BW>(desc ? idx1 - 1 : 0u, k, off);
        BlockWords<LPQ, BW> w;
        w.clear();
        uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0;
        uint2 ci = make_uint2(0, 0);
        if (desc) {
          w.load(im.blocks, base + k, sub);
          const uint4* rec = reinterpret_cast<const uint4*>(im.supers + node);
          r0 = __ldg(rec);                                          // grandchildren 0, 1
          r1 = __ldg(rec + 1);                                      // grandchildren 2, 3
          ci = __ldg(reinterpret_cast<const uint2*>(rec + 2));      // child_info
        }
        const uint32_t cx = group_sum<LPQ>(paired_count<LPQ, BW, kPairedX>(w, static_cast<int>(off) + 1, sub));
        const uint32_t bit1 = paired_bit<LPQ, BW, kPairedX>(w, off);
        const uint32_t ones1 = w.w[0] + cx;
        const uint32_t i1 = bit1 ? ones1 : idx1 - ones1;
        const uint32_t j = max(bit1 ? cx : off + 1 - cx, 1u);  // >= 1: the position itself is one of them
        const uint32_t rp = bit1 ? B - j : j - 1;              // region bit of the child at index i1-1
        const uint32_t cr = group_sum<LPQ>(paired_count<LPQ, BW, kPairedR>(w, static_cast<int>(bit1 ? B - j : j), sub));
        const uint32_t bit2 = paired_bit<LPQ, BW, kPairedR>(w, rp);
        const uint32_t h1 = paired_h1<LPQ, BW>(w, bit1);
        const uint32_t ones2 = bit1 ? h1 - cr : h1 + cr;
        if (desc) {
          const uint32_t info1 = bit1 ? ci.y : ci.x;
          if (info1 & kChildLeaf) {
            idx1 = i1;
            ch = info1 & 0xffffu;
            desc = false;
          } else {
            idx1 = bit2 ? ones2 : i1 - ones2;
            const uint4 rr = bit1 ? r1 : r0;
            const uint32_t gbase = bit2 ? rr.z : rr.x, ginfo = bit2 ? rr.w : rr.y;
            if (ginfo & kChildLeaf) {
              ch = ginfo & 0xffffu;
              desc = false;
            } else {
              base = gbase;
              node = ginfo;
            }
          }
        }
      }
    } else
    while (__any_sync(kFull, desc)) {
      const uint32_t p = desc ? idx1 - 1 : 0u;
      const uint32_t k = p / BITS;
      const uint32_t off = p - k * BITS;
      uint4 nr = make_uint4(0, 0, 0, 0);
      if (desc) nr = __ldg(reinterpret_cast<const uint4*>(im.nodes + node));
      uint32_t ones, bit;
      block_rank<LPQ, BW, true>(im.blocks, base + k, off, desc, sub, ones, bit);
      if (desc) {
        idx1 = bit ? ones : (idx1 - ones);
        const uint32_t info = bit ? nr.w : nr.z;
        if (info & kChildLeaf) {
          ch = info & 0xffffu;
          desc = false;
        } else {
          base = bit ? nr.y : nr.x;
          node = info;
        }
      }
    }
    const uint32_t count = idx1;  // this row holds the count-th occurrence of ch in the bucket

    // mark test: rank over the symbol's mark bit-vector at its occurrence number (index.c:2102-2140)
    const bool ok = act && ch < static_cast<uint32_t>(kAlphaDev) && count > 0;
    int64_t occ_base = 0, offset = -1;
    if constexpr (MODE == kWalkExtract) {  // extraction follows LF only: the mark bit-vectors are not read
      if (ok) occ_base = rec_occ_base(__ldg(reinterpret_cast<const int4*>(im.occ + static_cast<size_t>(g) * kAlphaStride + ch)));
    } else {
      mark_lookup<LPQ, BW>(im, ok, g, ch, count, markval_base, sub, offset, occ_base, n_mark, n_sample);
    }
    // LF: row' = C[ch] + occs before the bucket + count - 1; stop at a document boundary
    // (ch <= ESCAPE_CODE_SEOF, server.c:2341-2346)
    const int64_t next = (ch <= static_cast<uint32_t>(kEscSeofDev)) ? -1 : occ_base + count - 1;

    if (act && !ok) {
      if (lane == gleader) atomicExch(a.status, 2);
      if (MODE == kWalkShard) park(-1, 2, static_cast<int>(meta >> 4));
      else if (lane == gleader && MODE != kWalkExtract) a.out_offset[rid] = -1;
      have = false;
    } else if (act) {
      if (MODE == kWalkShard) {
        if (offset >= 0) {
          park(offset + steps, 2, static_cast<int>(meta >> 4));
          have = false;
        } else if (next < 0 || steps > im.total_length) {
          if (lane == gleader) atomicExch(a.status, 3);
          park(-1, 2, static_cast<int>(meta >> 4));
          have = false;
        } else {
          row = next;  // resident or not is decided at the top of the next iteration
          steps++;
        }
      } else if (MODE == kWalkLocate) {
        if (offset >= 0) {
          if (lane == gleader) a.out_offset[rid] = offset + steps;
          have = false;
        } else if (next < 0 || steps > im.total_length) {  // unmarked document start: index violates should_mark()
          if (lane == gleader) { atomicExch(a.status, 3); a.out_offset[rid] = -1; }
          have = false;
        } else {
          row = next;
          steps++;
        }
      } else if (MODE == kWalkStep) {
        if (lane == gleader) {
          a.out_ch[rid] = static_cast<int32_t>(ch);
          a.out_next[rid] = next;
          a.out_offset[rid] = offset;
        }
        have = false;
      } else {  // extract
        if (lane == gleader) a.out_sym[sym_off + (nsteps - 1 - steps)] = static_cast<uint16_t>(ch);
        steps++;
        if (steps == nsteps) {
          have = false;
        } else if (next < 0) {
          if (lane == gleader) atomicExch(a.status, 4);
          have = false;
        } else {
          row = next;
        }
      }
    }
  }
  if (a.stats) flush_stats(a.stats, lane, n_steps, n_quad, n_mark, n_sample);
}

// ---------------------------------------------------------------------------------------------
template <int LPQ, int BW, int LV = 1>
__global__ void __launch_bounds__(kThreads) occ_kernel(const DevImage im, const OccArgs a) {
  constexpr int QPW = 32 / LPQ;
  const int lane = threadIdx.x & 31;
  const int sub = lane & (LPQ - 1);
  const int qi = lane / LPQ;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t base = warp * QPW; base < a.n; base += nwarps * QPW) {
    const int64_t item = base + qi;
    int c = 0;
    int64_t row = 0;
    bool q = false;
    if (item < a.n) {
      c = a.ch[item];
      row = a.rows[item];
      q = c < kAlphaDev && row >= im.first_row && row < im.end_row;
    }
    const int64_t r = occ_any<LPQ, BW, LV>(im, q, c, row, sub);
    if (item < a.n && sub == 0) a.out[item] = q ? r : -1;
  }
}

// ---------------------------------------------------------------------------------------------
// Range-sharded count: advance pattern states while the BWT rows they need are resident here.
template <int LPQ, int BW, int LV = 1>
__global__ void __launch_bounds__(kThreads) count_shard_kernel(const DevImage im, const ShardArgs a) {
  constexpr int QPW = 32 / LPQ;
  const int lane = threadIdx.x & 31;
  const int sub = lane & (LPQ - 1);
  const int qi = lane / LPQ;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t base = warp * QPW; base < a.n; base += nwarps * QPW) {
    const int64_t item = base + qi;
    bool running = item < a.n;
    int64_t pid = 0, f = 0, l = -1, obA = 0;
    int i = 0, phase = 2, home = 0, dest = -1;
    const uint16_t* pat = nullptr;
    if (running) {
      const int64_t* s = a.state + item * kShardStateWords;
      pid = s[0]; f = s[1]; l = s[2]; i = static_cast<int>(s[3]); obA = s[4];
      phase = static_cast<int>(s[5] & 15);
      home = static_cast<int>(s[5] >> 4);
      pat = a.flat + a.offs[pid];
      if (phase == 3) {  // new pattern: [C[c], C[c+1]-1] for its last symbol (server.c:781-801)
        const int m = a.plen[pid];
        if (m <= 0) { f = 0; l = im.total_length - 1; i = 0; }
        else {
          const int c = pat[m - 1];
          if (c >= kAlphaDev) { f = im.total_length; l = f - 1; }
          else { f = __ldg(im.C + c); l = __ldg(im.C + c + 1) - 1; }
          i = m - 1;
        }
        phase = 0;
      }
    }
    for (;;) {
      bool q = false;
      int c = 0;
      int64_t row = 0;
      if (running) {
        if (phase == 2) {
          dest = home; running = false;
        } else if (phase == 0 && (f > l || i == 0)) {
          phase = 2; dest = home; running = false;
        } else {
          c = pat[i - 1];
          if (phase == 0 && c >= kAlphaDev) {
            f = im.total_length; l = f - 1; i--;
          } else if (phase == 0 && f == 0) {
            obA = __ldg(im.C + c); phase = 1;  // Occ(c,-1) = 0 (server.c:847-851)
          } else {
            row = phase == 0 ? f - 1 : l;
            if (row >= im.first_row && row < im.end_row) {
              q = true;
            } else {
              dest = shard_of_row(row, a.block_size, im.total_length, a.nshards);
              running = false;
            }
          }
        }
      }
      if (!__any_sync(kFull, running)) break;
      const int64_t r = occ_any<LPQ, BW, LV>(im, q, c, row, sub);
      if (q) {
        if (phase == 0) { obA = r; phase = 1; }
        else { f = obA; l = r - 1; i--; phase = 0; }
      }
    }
    if (item < a.n && sub == 0) {
      int64_t* s = a.state + item * kShardStateWords;
      s[1] = f; s[2] = l; s[3] = i; s[4] = obA;
      s[5] = static_cast<int64_t>(phase) | (static_cast<int64_t>(home) << 4);
      a.dest[item] = dest;
    }
  }
}

template <typename K>
int blocks_per_sm(K kernel) {
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kThreads, 0) != cudaSuccess || n < 1) n = 1;
  return n;
}

inline int grid_for(int64_t groups_needed, int groups_per_block, int sm_count, int bps) {
  int64_t blocks = (groups_needed + groups_per_block - 1) / groups_per_block;
  const int64_t cap = static_cast<int64_t>(sm_count) * bps;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace

// sched: pair schedule = lanes per Occ query (4 or 8); sync schedule = 1000 + 10*lanes + MINB.
cudaError_t launch_count(const DevImage& im, const CountArgs& a, unsigned long long* d_work, int sched, int sm_count,
                         cudaStream_t stream, int64_t* launch_counter, unsigned long long* d_stats) {
  if (a.npats <= 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(d_work, 0, sizeof(unsigned long long), stream);
  if (e != cudaSuccess) return e;
  const int code = (im.levels - 1) * 1000000 + im.block_words * 10000 + sched;
#define FM_LAUNCH(KERNEL, LANES_PER_PATTERN)                                                             \
  do {                                                                                                   \
    static const int bps = blocks_per_sm(KERNEL);                                                        \
    KERNEL<<<grid_for(a.npats, kThreads / (LANES_PER_PATTERN), sm_count, bps), kThreads, 0, stream>>>(    \
        im, a, d_work, d_stats);                                                                         \
  } while (0)
#define FM_SYNC(BW, LANES, MINB)                                                                         \
  case (BW) * 10000 + 1000 + 10 * (LANES) + (MINB):                                                      \
    if (d_stats) FM_LAUNCH((count_sync_kernel<LANES, BW, MINB, true>), LANES);                           \
    else FM_LAUNCH((count_sync_kernel<LANES, BW, MINB, false>), LANES);                                  \
    break;
#define FM_SYNC2(BW, LANES, MINB)                                                                        \
  case 1000000 + (BW) * 10000 + 1000 + 10 * (LANES) + (MINB):                                            \
    if (d_stats) FM_LAUNCH((count_sync_kernel<LANES, BW, MINB, true, 2>), LANES);                        \
    else FM_LAUNCH((count_sync_kernel<LANES, BW, MINB, false, 2>), LANES);                               \
    break;
#define FM_SYNC4(MINB)                                                                                   \
  case 3000000 + 32 * 10000 + 1000 + 10 * 2 + (MINB):                                                    \
    if (d_stats) FM_LAUNCH((count_sync_kernel<2, 32, MINB, true, 4>), 2);                                \
    else FM_LAUNCH((count_sync_kernel<2, 32, MINB, false, 4>), 2);                                       \
    break;
  // raw text bytes instead of alpha_t symbols: the default schedule of the quad image only
  if (a.sym8) {
    if (code != 3000000 + 32 * 10000 + 1000 + 10 * 2 + 4 || d_stats) return cudaErrorNotSupported;
    FM_LAUNCH((count_sync_kernel<2, 32, 4, false, 4, 0, true>), 2);
    if (launch_counter) ++*launch_counter;
    return cudaGetLastError();
  }
#define FM_SYNC4EXP(MINB, EXP)                                                                           \
  case 3000000 + 32 * 10000 + 1000 + 70 + 10 * (EXP) + (MINB):                                           \
    if (d_stats) FM_LAUNCH((count_sync_kernel<2, 32, MINB, true, 4, EXP>), 2);                           \
    else FM_LAUNCH((count_sync_kernel<2, 32, MINB, false, 4, EXP>), 2);                                  \
    break;
#define FM_QUAD_SPLIT(CODE, THREADS, MINB)                                                               \
  case 3000000 + 32 * 10000 + 1000 + (CODE):                                                             \
    do {                                                                                                 \
      auto kern = d_stats ? count_quad_split_kernel<MINB, true, THREADS>                                 \
                          : count_quad_split_kernel<MINB, false, THREADS>;                               \
      int bps = 0;                                                                                       \
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, THREADS, 0) != cudaSuccess || bps < 1) bps = 1; \
      kern<<<grid_for(a.npats, (THREADS) / 2, sm_count, bps), THREADS, 0, stream>>>(im, a, d_work, d_stats); \
    } while (0);                                                                                         \
    break;
#define FM_PAIR(BW, LANES)                                                                               \
  case (BW) * 10000 + (LANES):                                                                           \
    if (d_stats) FM_LAUNCH((count_pair_kernel<LANES, BW, true>), 2 * (LANES));                           \
    else FM_LAUNCH((count_pair_kernel<LANES, BW, false>), 2 * (LANES));                                  \
    break;
  switch (code) {
    // the schedules fm_set_count_schedule / default_count_sched can select (fm_api.cu)
    FM_PAIR(32, 4) FM_PAIR(32, 8) FM_PAIR(16, 4) FM_PAIR(16, 2) FM_PAIR(8, 2)
    FM_SYNC(32, 8, 5) FM_SYNC(32, 4, 4) FM_SYNC(32, 2, 3)
    FM_SYNC(16, 4, 6) FM_SYNC(16, 2, 4) FM_SYNC(16, 1, 4)
    FM_SYNC(8, 2, 6) FM_SYNC(8, 1, 4)
    FM_SYNC2(32, 4, 6) FM_SYNC2(32, 2, 4) FM_SYNC2(32, 1, 3) FM_SYNC2(16, 2, 4) FM_SYNC2(16, 1, 4)
    FM_SYNC4(4)
    FM_QUAD_SPLIT(67, 128, 9)  /* split schedule: 60 + k */
#ifdef FM_TUNING_VARIANTS
    // register-budget / occupancy sweeps and the perturbation variants behind
    // profiles/r01_count_schedule_sweep.md, r01_paired_level_sweep.md, r01_quad_schedules.md
    // (selected with FEMTO_B200_COUNT_SCHED); not part of the product build
    FM_SYNC(32, 4, 5) FM_SYNC(32, 2, 4) FM_SYNC(16, 4, 5) FM_SYNC(16, 2, 5) FM_SYNC(16, 1, 3)
    FM_SYNC(8, 2, 5) FM_SYNC(8, 1, 5) FM_SYNC(8, 1, 6)
    FM_SYNC2(32, 4, 4) FM_SYNC2(32, 4, 5) FM_SYNC2(32, 2, 3) FM_SYNC2(32, 2, 5) FM_SYNC2(32, 1, 2)
    FM_SYNC2(16, 2, 5) FM_SYNC2(16, 2, 6) FM_SYNC2(16, 1, 3) FM_SYNC2(16, 1, 5)
    FM_SYNC4(3) FM_SYNC4(5) FM_SYNC4(6) FM_SYNC4(8)
    FM_QUAD_SPLIT(65, 256, 5)
    FM_SYNC4EXP(4, 1) FM_SYNC4EXP(5, 1) FM_SYNC4EXP(4, 2)  /* codes 1084, 1085, 1094 */
#endif
    default: return cudaErrorInvalidValue;
  }
#undef FM_SYNC
#undef FM_SYNC2
#undef FM_SYNC4
#undef FM_QUAD_SPLIT
#undef FM_SYNC4EXP
#undef FM_PAIR
#undef FM_LAUNCH
  if (launch_counter) ++*launch_counter;
  return cudaGetLastError();
}

template <int LPQ, int BW, int LV = 1>
static cudaError_t launch_walk_cfg(const DevImage& im, const WalkArgs& a, WalkMode mode, unsigned long long* d_work,
                                   int sm_count, cudaStream_t stream) {
  const int gpb = kThreads / LPQ;
#ifdef FM_TUNING_VARIANTS
  if (mode == kWalkLocate && LV == 4) {
    // CTAs per SM: FEMTO_B200_WALK_CTAS = 4 (64 registers), 6 (40) or 8 (32); measured 1.69 / 1.85 / 2.01 ms per
    // 1 Mi rows (profiles/r02_walk_kernel.md): more warps with spilled registers are slower, 4 stays
    static const int want = [] { const char* v = std::getenv("FEMTO_B200_WALK_CTAS"); return v ? std::atoi(v) : kWalkLocateCtas; }();
    if (want >= 8) {
      static const int bps = blocks_per_sm(walk_kernel<LPQ, BW, kWalkLocate, LV, 8>);
      walk_kernel<LPQ, BW, kWalkLocate, LV, 8><<<grid_for(a.nrows, gpb, sm_count, bps), kThreads, 0, stream>>>(im, a, d_work);
    } else if (want >= 6) {
      static const int bps = blocks_per_sm(walk_kernel<LPQ, BW, kWalkLocate, LV, 6>);
      walk_kernel<LPQ, BW, kWalkLocate, LV, 6><<<grid_for(a.nrows, gpb, sm_count, bps), kThreads, 0, stream>>>(im, a, d_work);
    } else {
      static const int bps = blocks_per_sm(walk_kernel<LPQ, BW, kWalkLocate, LV>);
      walk_kernel<LPQ, BW, kWalkLocate, LV><<<grid_for(a.nrows, gpb, sm_count, bps), kThreads, 0, stream>>>(im, a, d_work);
    }
  } else
#endif
  if (mode == kWalkLocate) {
    static const int bps = blocks_per_sm(walk_kernel<LPQ, BW, kWalkLocate, LV>);
    walk_kernel<LPQ, BW, kWalkLocate, LV><<<grid_for(a.nrows, gpb, sm_count, bps), kThreads, 0, stream>>>(im, a, d_work);
  } else if (mode == kWalkStep) {
    static const int bps = blocks_per_sm(walk_kernel<LPQ, BW, kWalkStep, LV>);
    walk_kernel<LPQ, BW, kWalkStep, LV><<<grid_for(a.nrows, gpb, sm_count, bps), kThreads, 0, stream>>>(im, a, d_work);
  } else if (mode == kWalkShard) {
    static const int bps = blocks_per_sm(walk_kernel<LPQ, BW, kWalkShard, LV>);
    walk_kernel<LPQ, BW, kWalkShard, LV><<<grid_for(a.nrows, gpb, sm_count, bps), kThreads, 0, stream>>>(im, a, d_work);
  } else {
    static const int bps = blocks_per_sm(walk_kernel<LPQ, BW, kWalkExtract, LV>);
    walk_kernel<LPQ, BW, kWalkExtract, LV><<<grid_for(a.nrows, gpb, sm_count, bps), kThreads, 0, stream>>>(im, a, d_work);
  }
  return cudaGetLastError();
}

// lanes per query for the walk / occ kernels: 4 or 8 at 128-byte blocks, 2 or 4 at 64, 1 or 2 at 32
static int walk_lanes(const DevImage& im, int lpq) {
  // at least one 128-bit load per lane; paired-level blocks: whole 32-byte slices per lane
  const int max_lanes = im.levels == 4 ? 2 : im.levels == 2 ? im.block_words / 8 : im.block_words / 4;
  int lanes = lpq;
  while (lanes > max_lanes) lanes >>= 1;
  if (lanes < max_lanes / 2) lanes = max_lanes / 2;
  return lanes < 1 ? 1 : lanes;
}

cudaError_t launch_walk(const DevImage& im, const WalkArgs& a, WalkMode mode, unsigned long long* d_work, int lpq,
                        int sm_count, cudaStream_t stream, int64_t* launch_counter) {
  if (a.nrows <= 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(d_work, 0, sizeof(unsigned long long), stream);
  if (e != cudaSuccess) return e;
  switch ((im.levels - 1) * 10000 + im.block_words * 100 + walk_lanes(im, lpq)) {
    case 33202: e = launch_walk_cfg<2, 32, 4>(im, a, mode, d_work, sm_count, stream); break;
    case 13204: e = launch_walk_cfg<4, 32, 2>(im, a, mode, d_work, sm_count, stream); break;
    case 13202: e = launch_walk_cfg<2, 32, 2>(im, a, mode, d_work, sm_count, stream); break;
    case 11602: e = launch_walk_cfg<2, 16, 2>(im, a, mode, d_work, sm_count, stream); break;
    case 11601: e = launch_walk_cfg<1, 16, 2>(im, a, mode, d_work, sm_count, stream); break;
    case 3208: e = launch_walk_cfg<8, 32>(im, a, mode, d_work, sm_count, stream); break;
    case 3204: e = launch_walk_cfg<4, 32>(im, a, mode, d_work, sm_count, stream); break;
    case 1604: e = launch_walk_cfg<4, 16>(im, a, mode, d_work, sm_count, stream); break;
    case 1602: e = launch_walk_cfg<2, 16>(im, a, mode, d_work, sm_count, stream); break;
    case 802: e = launch_walk_cfg<2, 8>(im, a, mode, d_work, sm_count, stream); break;
    case 801: e = launch_walk_cfg<1, 8>(im, a, mode, d_work, sm_count, stream); break;
    default: return cudaErrorInvalidValue;
  }
  if (launch_counter) ++*launch_counter;
  return e;
}

cudaError_t launch_count_shard(const DevImage& im, const ShardArgs& a, int lpq, int sm_count, cudaStream_t stream,
                               int64_t* launch_counter) {
  if (a.n <= 0) return cudaSuccess;
#define FM_SHARD(LANES, BW, ...)                                                                          \
  do {                                                                                                    \
    static const int bps = blocks_per_sm(count_shard_kernel<LANES, BW, ##__VA_ARGS__>);                   \
    count_shard_kernel<LANES, BW, ##__VA_ARGS__>                                                          \
        <<<grid_for(a.n, kThreads / (LANES), sm_count, bps), kThreads, 0, stream>>>(im, a);               \
  } while (0)
  switch ((im.levels - 1) * 10000 + im.block_words * 100 + walk_lanes(im, lpq)) {
    case 33202: FM_SHARD(2, 32, 4); break;
    case 13204: FM_SHARD(4, 32, 2); break;
    case 13202: FM_SHARD(2, 32, 2); break;
    case 11602: FM_SHARD(2, 16, 2); break;
    case 11601: FM_SHARD(1, 16, 2); break;
    case 3208: FM_SHARD(8, 32); break;
    case 3204: FM_SHARD(4, 32); break;
    case 1604: FM_SHARD(4, 16); break;
    case 1602: FM_SHARD(2, 16); break;
    case 802: FM_SHARD(2, 8); break;
    case 801: FM_SHARD(1, 8); break;
    default: return cudaErrorInvalidValue;
  }
#undef FM_SHARD
  if (launch_counter) ++*launch_counter;
  return cudaGetLastError();
}

namespace {

// noccs as 64-bit counts for the scan (a clipped range can still hold 2^31 rows)
__global__ void __launch_bounds__(kThreads) clip_kernel(int64_t n, const int64_t* __restrict__ first,
                                                         const int64_t* __restrict__ last, int64_t max_occs,
                                                         int32_t* __restrict__ noccs, int64_t* __restrict__ cnt) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t f = first[i];
  int64_t l = last[i];
  int64_t c = 0;
  if (f <= l) {
    if (l - f > max_occs) l = f + max_occs - 1;  // do_locate_query's clip (server.c:4411-4415)
    c = l - f + 1;
    if (c < 0) c = 0;  // max_occs < 0: the clip leaves last < first, i.e. no rows (as the reference)
  }
  noccs[i] = static_cast<int32_t>(c);
  cnt[i] = c;
}

__global__ void __launch_bounds__(kThreads) total_kernel(int64_t n, const int64_t* __restrict__ cnt,
                                                          const int64_t* __restrict__ start, int64_t* __restrict__ total) {
  if (blockIdx.x == 0 && threadIdx.x == 0) *total = n ? start[n - 1] + cnt[n - 1] : 0;
}

// one thread per output row: the pattern it belongs to is the last one starting at or before it
__global__ void __launch_bounds__(kThreads) expand_rows_kernel(int64_t npats, int64_t total,
                                                                const int64_t* __restrict__ first,
                                                                const int64_t* __restrict__ start,
                                                                int64_t* __restrict__ rows) {
  const int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= total) return;
  int64_t lo = 0, hi = npats;  // upper_bound(start, k) - 1; patterns without rows share their successor's start
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(start + mid) <= k) lo = mid; else hi = mid;
  }
  rows[k] = __ldg(first + lo) + (k - __ldg(start + lo));
}

}  // namespace

size_t expand_scratch_bytes(int64_t npats) {
  size_t temp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, temp, static_cast<const int64_t*>(nullptr), static_cast<int64_t*>(nullptr),
                                static_cast<int>(std::min<int64_t>(npats, INT32_MAX)));
  // layout used by launch_clip_and_scan: [counts, 256-byte aligned][cub temporary storage]
  const size_t cnt = (size_t(std::max<int64_t>(npats, 1)) * sizeof(int64_t) + 255) & ~size_t(255);
  return cnt + ((temp + 255) & ~size_t(255)) + 256;
}

cudaError_t launch_clip_and_scan(int64_t npats, const int64_t* d_first, const int64_t* d_last, int max_occs,
                                 int32_t* d_noccs, int64_t* d_out_start, int64_t* d_total, void* d_scratch,
                                 size_t scratch_bytes, cudaStream_t stream, int64_t* launch_counter) {
  if (npats <= 0 || npats > INT32_MAX) return cudaErrorInvalidValue;
  const size_t cnt_bytes = size_t(npats) * sizeof(int64_t);
  if (scratch_bytes < cnt_bytes) return cudaErrorInvalidValue;
  int64_t* d_cnt = static_cast<int64_t*>(d_scratch);
  void* d_temp = static_cast<char*>(d_scratch) + ((cnt_bytes + 255) & ~size_t(255));
  size_t temp = scratch_bytes - ((cnt_bytes + 255) & ~size_t(255));
  const int grid = static_cast<int>((npats + kThreads - 1) / kThreads);
  clip_kernel<<<grid, kThreads, 0, stream>>>(npats, d_first, d_last, max_occs, d_noccs, d_cnt);
  cudaError_t e = cub::DeviceScan::ExclusiveSum(d_temp, temp, d_cnt, d_out_start, static_cast<int>(npats), stream);
  if (e != cudaSuccess) return e;
  total_kernel<<<1, 32, 0, stream>>>(npats, d_cnt, d_out_start, d_total);
  if (launch_counter) *launch_counter += 3;
  return cudaGetLastError();
}

cudaError_t launch_expand_rows(int64_t npats, int64_t total, const int64_t* d_first, const int64_t* d_out_start,
                               int64_t* d_rows, cudaStream_t stream, int64_t* launch_counter) {
  if (total <= 0) return cudaSuccess;
  const int64_t grid = (total + kThreads - 1) / kThreads;
  if (grid > INT32_MAX) return cudaErrorInvalidValue;
  expand_rows_kernel<<<static_cast<int>(grid), kThreads, 0, stream>>>(npats, total, d_first, d_out_start, d_rows);
  if (launch_counter) ++*launch_counter;
  return cudaGetLastError();
}

namespace {

// Dependent random reads with the access shape of the query kernels: LPA lanes x 16 bytes per
// access, the next address of a chain derived from the data just read.
template <int LPA>
__global__ void __launch_bounds__(kThreads) probe_kernel(const uint4* __restrict__ base, uint64_t n_units, int steps,
                                                         unsigned long long* __restrict__ sink) {
  const int lane = threadIdx.x & 31;
  const int sub = lane & (LPA - 1);
  const uint64_t group = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) / LPA;
  uint32_t xa = static_cast<uint32_t>(group * 2654435761u + 12345u);
  uint32_t xb = static_cast<uint32_t>(group * 2246822519u + 977u);
  for (int s = 0; s < steps; s++) {
    const uint64_t ua = (static_cast<uint64_t>(xa) * n_units) >> 32;  // uniform in [0, n_units)
    const uint64_t ub = (static_cast<uint64_t>(xb) * n_units) >> 32;
    const uint4 va = __ldg(base + ua * LPA + sub);
    const uint4 vb = __ldg(base + ub * LPA + sub);
    uint32_t fa = va.x ^ va.y ^ va.z ^ va.w, fb = vb.x ^ vb.y ^ vb.z ^ vb.w;
#pragma unroll
    for (int o = LPA / 2; o > 0; o >>= 1) {
      fa ^= __shfl_xor_sync(kFull, fa, o);
      fb ^= __shfl_xor_sync(kFull, fb, o);
    }
    xa = (xa ^ fa) * 2654435761u + 0x9e3779b9u + static_cast<uint32_t>(s);
    xb = (xb ^ fb) * 2246822519u + 0x85ebca6bu + static_cast<uint32_t>(s);
  }
  if ((xa ^ xb) == 0x12345678u && sub == 0) atomicAdd(sink, 1ull);  // keeps the chains alive
}

}  // namespace

cudaError_t launch_probe(const uint4* base, uint64_t n_units, int bytes_per_access, int steps, int sm_count,
                         cudaStream_t stream, unsigned long long* d_sink, int64_t* accesses) {
  if (n_units == 0 || steps <= 0) return cudaErrorInvalidValue;
  int bps = 1;
  int64_t groups = 0;
#define FM_PROBE(LPA)                                                                    \
  do {                                                                                   \
    bps = blocks_per_sm(probe_kernel<LPA>);                                              \
    groups = static_cast<int64_t>(sm_count) * bps * kThreads / (LPA);                    \
    probe_kernel<LPA><<<sm_count * bps, kThreads, 0, stream>>>(base, n_units, steps, d_sink); \
  } while (0)
  switch (bytes_per_access) {
    case 128: FM_PROBE(8); break;
    case 64: FM_PROBE(4); break;
    case 32: FM_PROBE(2); break;
    default: return cudaErrorInvalidValue;
  }
#undef FM_PROBE
  if (accesses) *accesses = groups * 2 * steps;
  return cudaGetLastError();
}

cudaError_t launch_occ(const DevImage& im, const OccArgs& a, unsigned long long* /*d_work*/, int lpq, int sm_count,
                       cudaStream_t stream, int64_t* launch_counter) {
  if (a.n <= 0) return cudaSuccess;
#define FM_OCC(LANES, BW, ...)                                                                            \
  do {                                                                                                    \
    static const int bps = blocks_per_sm(occ_kernel<LANES, BW, ##__VA_ARGS__>);                           \
    occ_kernel<LANES, BW, ##__VA_ARGS__>                                                                  \
        <<<grid_for(a.n, kThreads / (LANES), sm_count, bps), kThreads, 0, stream>>>(im, a);               \
  } while (0)
  switch ((im.levels - 1) * 10000 + im.block_words * 100 + walk_lanes(im, lpq)) {
    case 33202: FM_OCC(2, 32, 4); break;
    case 13204: FM_OCC(4, 32, 2); break;
    case 13202: FM_OCC(2, 32, 2); break;
    case 11602: FM_OCC(2, 16, 2); break;
    case 11601: FM_OCC(1, 16, 2); break;
    case 3208: FM_OCC(8, 32); break;
    case 3204: FM_OCC(4, 32); break;
    case 1604: FM_OCC(4, 16); break;
    case 1602: FM_OCC(2, 16); break;
    case 802: FM_OCC(2, 8); break;
    case 801: FM_OCC(1, 8); break;
    default: return cudaErrorInvalidValue;
  }
#undef FM_OCC
  if (launch_counter) ++*launch_counter;
  return cudaGetLastError();
}

}  // namespace fmb
