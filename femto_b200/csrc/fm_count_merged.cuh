// fm_count_merged.cuh -- the count kernel, "merged pair" schedule (included by fm_kernels.cu).
//
// A backward-search step needs Occ(c, first-1) and Occ(c, last) (reference src/main/server.c:842-936).
// After the first few steps the two rows are neighbours, so at every wavelet-tree level they fall
// into the SAME 128-byte rank block 93% of the time (measured on the 4 GiB byte corpus).  This
// kernel gives a pattern ONE group of LPQ lanes that advances both ranks together:
//
//   micro-op = one rank-block read by the group (+ the node record when the level completes)
//     * both positions in the same block  -> one read, two masked popcounts, one packed reduction
//     * positions in different blocks     -> two micro-ops for that level (A, then B)
//     * the two rows in different buckets -> two descents in sequence (rare: wide ranges only)
//
// Every loop iteration each group of the warp performs exactly one micro-op (or a step set-up that
// reads the 16-byte OccRec / BucketRec), so a warp keeps 32/LPQ independent HBM line reads in
// flight and never reads a line twice for the same step.  Groups pull new patterns from a global
// atomic queue when theirs is finished.
#pragma once

namespace fmb {
namespace {

__device__ __forceinline__ uint32_t top_mask(int n) {
  // the n (>=0) most significant bits set; n >= 32 gives all ones (funnel shift clamps at 32)
  return __funnelshift_rc(0u, 0xFFFFFFFFu, static_cast<uint32_t>(n));
}

// Ranks at two offsets of ONE block (offB may equal offA).  Warp-collective.
template <int LPQ>
__device__ __forceinline__ void block_rank2(const uint4* __restrict__ blocks, uint32_t blk, uint32_t offA,
                                            uint32_t offB, bool active, int sub, uint32_t& onesA,
                                            uint32_t& onesB) {
  constexpr int WPL = kBlockWords / LPQ;
  constexpr int VPL = WPL / 4;
  uint32_t w[WPL];
#pragma unroll
  for (int t = 0; t < WPL; t++) w[t] = 0;
  if (active) {
    const uint4* p = blocks + static_cast<size_t>(blk) * (kBlockWords / 4) + sub * VPL;
#pragma unroll
    for (int v = 0; v < VPL; v++) {
      const uint4 x = __ldg(p + v);
      w[4 * v + 0] = x.x; w[4 * v + 1] = x.y; w[4 * v + 2] = x.z; w[4 * v + 3] = x.w;
    }
  }
  const uint32_t hdr_word = w[0];
  if (sub == 0) w[0] = 0;  // word 0 of the block is the header, not payload
  // bits of this lane's words (counted from its first bit) that lie at or before each position
  const int lane_bit0 = 32 * WPL * sub;
  const int nbA = static_cast<int>(offA) + 33 - lane_bit0;
  const int nbB = static_cast<int>(offB) + 33 - lane_bit0;
  uint32_t cA = 0, cB = 0;
#pragma unroll
  for (int t = 0; t < WPL; t++) {
    cA += __popc(w[t] & top_mask(max(nbA - 32 * t, 0)));
    cB += __popc(w[t] & top_mask(max(nbB - 32 * t, 0)));
  }
  uint32_t packed = cA | (cB << 16);  // each count <= 992
#pragma unroll
  for (int o = LPQ / 2; o > 0; o >>= 1) packed += __shfl_xor_sync(kFull, packed, o);
  const uint32_t hdr = __shfl_sync(kFull, hdr_word, 0, LPQ);
  onesA = hdr + (packed & 0xffffu);
  onesB = hdr + (packed >> 16);
}

template <int LPQ, int MINB, bool STATS>
__global__ void __launch_bounds__(kThreads, MINB) count_merged_kernel(const DevImage im, const CountArgs a,
                                                                 unsigned long long* __restrict__ work,
                                                                 unsigned long long* __restrict__ stats) {
  unsigned long long n_ranks = 0, n_blocks = 0, n_occ = 0, n_steps = 0;
  const int lane = threadIdx.x & 31;
  const int sub = lane & (LPQ - 1);
  const int gleader = lane & ~(LPQ - 1);

  // pattern state
  int64_t f = 0, l = -1, pid = -1;
  int i = 0;
  const uint16_t* pat = nullptr;
  bool have = false, exhausted = false;
  // step / job state
  bool descending = false;     // a wavelet-tree descent is in progress
  bool cross_pending = false;  // row B lies in another bucket: its descent follows A's
  bool jobA = false, jobB = false, actA = false, actB = false, half = false;
  uint32_t base = 0, node = 0, leaf = 0, idxA = 0, idxB = 0, savedA = 0;
  int L = 0, lvl = 0, c = 0;
  int64_t obA = 0, obB = 0;  // Occ bases, turned into the step's results C[c]+Occ when a descent ends

  for (;;) {
    // ---- retire and fetch ("first > last || i == 0" ends the reference's loop, server.c:832-841)
    if (have && !descending && !cross_pending && (f > l || i == 0)) {
      if (lane == gleader) {
        if (a.last) { a.first[pid] = f; a.last[pid] = l; }
        else a.first[pid] = l - f + 1;  // parallel_count with last==NULL (femto.c:313-318)
      }
      have = false;
    }
    const bool need = !have && !exhausted;
    unsigned long long idx = 0;
    if (need && lane == gleader) idx = atomicAdd(work, 1ull);
    idx = __shfl_sync(kFull, idx, gleader);
    if (need) {
      if (static_cast<int64_t>(idx) < a.npats) {
        pid = static_cast<int64_t>(idx);
        const int m = a.plen[pid];
        pat = a.flat + a.offs[pid];
        if (m <= 0) {  // empty pattern: every row (server.c:782-808)
          f = 0; l = im.total_length - 1; i = 0;
        } else {
          const int c0 = pat[m - 1];
          if (c0 >= kAlphaDev) { f = im.total_length; l = f - 1; }  // get_C(ch>=ALPHA_SIZE), index.c:1545
          else { f = __ldg(im.C + c0); l = __ldg(im.C + c0 + 1) - 1; }
          i = m - 1;
        }
        have = true;
      } else {
        exhausted = true;
      }
    }
    if (!__any_sync(kFull, have)) break;

    const bool run_op = descending;  // groups that were descending at the top of this iteration

    // ---- step set-up (no rank block is read in this iteration by this group)
    if (have && !descending && (cross_pending || (f <= l && i > 0))) {
      bool finish = false;
      int64_t g = 0;
      uint32_t rb = 0;
      if (!cross_pending) {
        c = pat[i - 1];
        if (STATS && sub == 0) n_steps++;
        if (c >= kAlphaDev) {  // symbol outside the alphabet: empty range
          f = im.total_length; l = f - 1; i--;
        } else {
          split_row(im, l, g, rb);
          const int4 rv = __ldg(reinterpret_cast<const int4*>(im.occ + g * kAlphaStride + c));
          obB = static_cast<int64_t>(static_cast<uint32_t>(rv.x)) | (static_cast<int64_t>(rv.y) << 32);
          leaf = static_cast<uint32_t>(rv.z);
          if (STATS && sub == 0) n_occ++;
          if (f == 0) {  // Occ(c,-1) = 0 without touching the index (server.c:847-851)
            obA = __ldg(im.C + c);
            jobA = false; jobB = true;
          } else {
            int64_t gA;
            uint32_t rbA;
            split_row(im, f - 1, gA, rbA);
            if (STATS && sub == 0) n_occ++;
            if (gA == g) {
              obA = obB;
              jobA = jobB = true;
              idxA = rbA + 1;
            } else {  // first-1 in another bucket: descend there first, row B afterwards
              const int4 ra = __ldg(reinterpret_cast<const int4*>(im.occ + gA * kAlphaStride + c));
              obA = static_cast<int64_t>(static_cast<uint32_t>(ra.x)) | (static_cast<int64_t>(ra.y) << 32);
              leaf = static_cast<uint32_t>(ra.z);
              cross_pending = true;
              jobA = true; jobB = false;
              g = gA;
              idxA = rbA + 1;
            }
          }
          idxB = rb + 1;
          if (leaf == 0) {  // symbol absent from this bucket: Occ is the bucket base (index.c:2080-2089)
            finish = !cross_pending;  // obA / obB already hold the results
          } else {
            descending = true;
          }
        }
      } else {  // second descent of a cross-bucket step: row B
        split_row(im, l, g, rb);
        const int4 rv = __ldg(reinterpret_cast<const int4*>(im.occ + g * kAlphaStride + c));
        obB = static_cast<int64_t>(static_cast<uint32_t>(rv.x)) | (static_cast<int64_t>(rv.y) << 32);
        leaf = static_cast<uint32_t>(rv.z);
        cross_pending = false;
        jobA = false; jobB = true;
        idxB = rb + 1;
        if (leaf == 0) finish = true;
        else descending = true;
      }
      if (descending) {
        const uint4 br = __ldg(reinterpret_cast<const uint4*>(im.buckets + g));
        base = br.x;
        node = br.y;
        L = 31 - __clz(leaf);
        lvl = 0;
        half = false;
        actA = jobA;
        actB = jobB;
      }
      if (finish) { f = obA; l = obB - 1; i--; }
    }

    // ---- one micro-op for every group that was descending
    const uint32_t pA = (run_op && actA) ? idxA - 1 : 0u;
    const uint32_t pB = (run_op && actB) ? idxB - 1 : 0u;
    const uint32_t kA = pA / kBitsPerBlock, kB = pB / kBitsPerBlock;
    const uint32_t offA = pA - kA * kBitsPerBlock, offB = pB - kB * kBitsPerBlock;
    const bool both = actA && actB;
    const bool same = both && kA == kB;
    const bool tgtA = actA && !half;
    const bool split_first = both && !same && !half;  // handles A only, level not complete yet
    const bool completes = run_op && !split_first;
    const uint32_t blk = base + (tgtA ? kA : kB);
    const uint32_t o1 = tgtA ? offA : offB;
    const uint32_t o2 = same ? offB : o1;
    uint4 nr = make_uint4(0, 0, 0, 0);
    if (completes && lvl + 1 < L) nr = __ldg(reinterpret_cast<const uint4*>(im.nodes + node));
    uint32_t r1, r2;
    block_rank2<LPQ>(im.blocks, blk, o1, o2, run_op, sub, r1, r2);
    if (STATS && run_op && sub == 0) {
      n_blocks++;
      n_ranks += same ? 2 : 1;
    }
    if (run_op) {
      if (split_first) {
        savedA = r1;
        half = true;
      } else {
        uint32_t onesA, onesB;
        if (half) { onesA = savedA; onesB = r1; }
        else if (same) { onesA = r1; onesB = r2; }
        else { onesA = r1; onesB = r1; }  // only one of the two is active
        half = false;
        lvl++;
        const uint32_t b = (leaf >> (L - lvl)) & 1u;
        if (actA) { idxA = b ? onesA : idxA - onesA; actA = idxA != 0; }  // index -= occs[!bit] (wtree.c:1109)
        if (actB) { idxB = b ? onesB : idxB - onesB; actB = idxB != 0; }
        if (lvl == L || !(actA || actB)) {  // leaf reached, or both counts are already zero
          if (jobA) obA += idxA;
          if (jobB) obB += idxB;
          descending = false;
          if (!cross_pending) { f = obA; l = obB - 1; i--; }
        } else {
          base = b ? nr.y : nr.x;
          node = b ? nr.w : nr.z;
        }
      }
    }
  }
  if (STATS) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      n_ranks += __shfl_xor_sync(kFull, n_ranks, o);
      n_blocks += __shfl_xor_sync(kFull, n_blocks, o);
      n_occ += __shfl_xor_sync(kFull, n_occ, o);
      n_steps += __shfl_xor_sync(kFull, n_steps, o);
    }
    if (lane == 0) {
      atomicAdd(stats + 0, n_ranks);
      atomicAdd(stats + 1, n_blocks);
      atomicAdd(stats + 2, n_occ);
      atomicAdd(stats + 3, n_steps);
    }
  }
}

}  // namespace
}  // namespace fmb
