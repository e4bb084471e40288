// fm_rank.cuh -- device functions shared by the query kernels (fm_kernels.cu) and the range-sharded
// mesh kernels (fm_mesh.cu): reading and evaluating rank blocks of the three image layouts
// (fm_image.hpp) and the full C[c] + Occ(c,row) descents built from them.
//
//   block_rank / occ_descend        plain blocks: one wavelet-tree level per read
//   paired_* / occ_descend_paired   paired-level blocks: two levels per read
//   QuadWords / quad_* / occ_descend_quad
//                                   quad-level blocks: four levels per 128-byte read; a lane of a
//                                   2-lane group holds one 32-byte sector, fetched with ONE 256-bit load
//   quad_eval_pair                  both positions of a backward-search step from one or two quad blocks
// Reference functions restated: wtree_occs (src/main/wtree.c:1081-1115) over bseq_rank
// (wtree.c:635-763), with header_occs_request / get_bucket_occs folded into OccRec::occ_base
// (src/main/index.c:1698-1765, 1828-1843).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

#include "fm_image.hpp"

namespace fmb {
namespace {
constexpr unsigned kFull = 0xffffffffu;
constexpr int kThreads = 256;
constexpr int kWalkLocateCtas = 4;  // resident CTAs per SM of the locate walk over the quad image (fm_kernels.cu walk_kernel)
constexpr int kEscSeofDev = 2;  // ESCAPE_CODE_SEOF, src/main/index_types.h:42-48
constexpr int kAlphaDev = 261;

__device__ __forceinline__ uint32_t popc_top(uint32_t w, int keep) {
  // ones among the `keep` most significant bits of w: keep >= 32 counts the whole word, keep <= 0
  // nothing.  shr.b32 clamps shift amounts above 31 to 32 (result 0), so only the lower bound of
  // the shift needs an explicit max.
  uint32_t shifted;
  const uint32_t sh = static_cast<uint32_t>(max(32 - keep, 0));
  asm("shr.b32 %0, %1, %2;" : "=r"(shifted) : "r"(w), "r"(sh));
  return __popc(shifted);
}

// The WPL words of rank block `blk` that belong to lane `sub` of its group.
template <int LPQ, int BW>
struct BlockWords {
  static constexpr int WPL = BW / LPQ;
  uint32_t w[WPL];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int t = 0; t < WPL; t++) w[t] = 0;
  }
  __device__ __forceinline__ void load(const uint4* __restrict__ blocks, uint32_t blk, int sub) {
    const uint32_t* base = reinterpret_cast<const uint32_t*>(blocks) + static_cast<size_t>(blk) * BW + sub * WPL;
    if (WPL >= 4) {
#pragma unroll
      for (int v = 0; v < WPL / 4; v++) {
        const uint4 x = __ldg(reinterpret_cast<const uint4*>(base) + v);
        w[4 * v + 0] = x.x; w[4 * v + 1] = x.y; w[4 * v + 2] = x.z; w[4 * v + 3] = x.w;
      }
    } else {
      const uint2 x = __ldg(reinterpret_cast<const uint2*>(base));
      w[0] = x.x; w[1] = x.y;
    }
  }
  // ones among this lane's payload bits at or before payload offset `off` (header word excluded)
  __device__ __forceinline__ uint32_t count_upto(uint32_t off, int sub) const {
    const int nb = static_cast<int>(off) + 33 - 32 * WPL * sub;
    uint32_t c = 0;
#pragma unroll
    for (int t = 0; t < WPL; t++) {
      const uint32_t word = (t == 0 && sub == 0) ? 0u : w[t];
      c += popc_top(word, nb - 32 * t);
    }
    return c;
  }
};

template <int LPQ>
__device__ __forceinline__ uint32_t group_sum(uint32_t v) {
#pragma unroll
  for (int o = LPQ / 2; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

template <int LPQ>
__device__ __forceinline__ uint32_t group_lane0(uint32_t v) {
  return LPQ == 1 ? v : __shfl_sync(kFull, v, 0, LPQ);
}

// Rank at one offset of one block, cooperatively by the LPQ lanes of a group (warp-collective).
//   blk : rank block index, off : 0-based bit offset inside the block payload
// Returns ones in the node's sequence up to and including the addressed bit, optionally the bit.
template <int LPQ, int BW, bool WANT_BIT>
__device__ __forceinline__ void block_rank(const uint4* __restrict__ blocks, uint32_t blk, uint32_t off,
                                           bool active, int sub, uint32_t& ones_incl, uint32_t& bit) {
  constexpr int WPL = BW / LPQ;
  BlockWords<LPQ, BW> b;
  b.clear();
  if (active) b.load(blocks, blk, sub);
  const uint32_t cnt = group_sum<LPQ>(b.count_upto(off, sub));
  ones_incl = group_lane0<LPQ>(b.w[0]) + cnt;
  if (WANT_BIT) {
    const int wq = (static_cast<int>(off) + 32) >> 5;  // block word holding the bit
    const int t_sel = wq % WPL;
    uint32_t mine = 0;
#pragma unroll
    for (int t = 0; t < WPL; t++) mine = (t == t_sel) ? b.w[t] : mine;
    const uint32_t word = LPQ == 1 ? mine : __shfl_sync(kFull, mine, wq / WPL, LPQ);
    bit = (word >> (31 - (off & 31))) & 1u;
  } else {
    bit = 0;
  }
}

__device__ __forceinline__ void split_row(const DevImage& im, int64_t row, int64_t& g, uint32_t& rb) {
  if (im.bucket_shift >= 0) {
    g = row >> im.bucket_shift;
    rb = static_cast<uint32_t>(row) & static_cast<uint32_t>(im.bucket_size - 1);
  } else {
    g = row / im.bucket_size;
    rb = static_cast<uint32_t>(row - g * im.bucket_size);
  }
  g -= im.first_bucket;
}

__device__ __forceinline__ int64_t rec_occ_base(const int4& rv) {
  return static_cast<int64_t>(static_cast<uint32_t>(rv.x)) | (static_cast<int64_t>(rv.y) << 32);
}

// C[c] + Occ(c,row) for the group's query (uniform across its LPQ lanes).  Warp-collective:
// every lane of the warp must call it; inactive groups pass active=false and get 0.
// STATS (instrumented launches only): n_reads counts rank blocks requested, n_distinct counts them
// once when the partner sub-group (the other Occ of the same backward-search step) asks for the
// same block at the same level.
template <int LPQ, int BW, bool STATS = false>
__device__ __forceinline__ int64_t occ_descend(const DevImage& im, bool active, int c, int64_t row, int sub,
                                               unsigned long long* n_reads = nullptr,
                                               unsigned long long* n_distinct = nullptr) {
  constexpr uint32_t BITS = (BW - 1) * 32;
  int64_t occ_base = 0;
  uint32_t leaf = 0, base = 0, node = 0, idx1 = 0;
  int L = 0;
  if (active) {
    int64_t g;
    uint32_t rb;
    split_row(im, row, g, rb);
    const int4 rv = __ldg(reinterpret_cast<const int4*>(im.occ + g * kAlphaStride + c));
    occ_base = rec_occ_base(rv);
    leaf = static_cast<uint32_t>(rv.z);
    if (leaf) {
      const uint4 br = __ldg(reinterpret_cast<const uint4*>(im.buckets + g));
      base = br.x;
      node = br.y;
      L = 31 - __clz(leaf);
      idx1 = rb + 1;  // 1-based index into the bucket, as wtree_occs (wtree.c:1097)
    }
  }
  bool desc = active && leaf != 0;
  int lvl = 0;
  while (__any_sync(kFull, desc)) {
    const uint32_t p = desc ? idx1 - 1 : 0u;
    const uint32_t k = p / BITS;
    const uint32_t off = p - k * BITS;
    uint4 nr = make_uint4(0, 0, 0, 0);
    if (desc && lvl + 1 < L) nr = __ldg(reinterpret_cast<const uint4*>(im.nodes + node));
    uint32_t ones, bit;
    block_rank<LPQ, BW, false>(im.blocks, base + k, off, desc, sub, ones, bit);
    if (STATS) {
      const uint32_t mine = desc ? base + k : 0xffffffffu;
      const uint32_t partner = __shfl_xor_sync(kFull, mine, LPQ);
      if (desc && sub == 0) {
        ++*n_reads;
        const bool second_of_pair = ((threadIdx.x & 31) / LPQ) & 1;
        if (!(second_of_pair && partner == mine)) ++*n_distinct;
      }
    }
    if (desc) {
      lvl++;
      const uint32_t b = (leaf >> (L - lvl)) & 1u;
      idx1 = b ? ones : (idx1 - ones);  // index -= occs[!bit]  (wtree.c:1109)
      if (idx1 == 0 || lvl == L) {
        desc = false;
      } else {
        base = b ? nr.y : nr.x;
        node = b ? nr.w : nr.z;
      }
    }
  }
  return occ_base + static_cast<int64_t>(leaf ? idx1 : 0u);
}

// ---- paired-level blocks (fm_image.hpp): a lane owns SPL = BW/8/LPQ whole 32-byte slices ----------
constexpr int kPairedX = 2;  // first X word of a slice
constexpr int kPairedR = 5;  // first children-region word of a slice

template <int BW>
__device__ __forceinline__ void paired_split(uint32_t p, uint32_t& k, uint32_t& off) {
  constexpr uint32_t B = kPairedSlicePos * (BW / kPairedSliceWords);
  k = p / B;
  off = p - k * B;
}

// ones among the first n bits of the block's X stretch (FIRST = kPairedX) or children region
// (FIRST = kPairedR) that this lane holds
template <int LPQ, int BW, int FIRST>
__device__ __forceinline__ uint32_t paired_count(const BlockWords<LPQ, BW>& b, int n, int sub) {
  constexpr int SPL = BW / kPairedSliceWords / LPQ;
  int keep = n - kPairedSlicePos * SPL * sub;
  uint32_t c = 0;
#pragma unroll
  for (int sl = 0; sl < SPL; sl++) {
#pragma unroll
    for (int t = 0; t < 3; t++) c += popc_top(b.w[sl * kPairedSliceWords + FIRST + t], keep - 32 * t);
    keep -= kPairedSlicePos;
  }
  return c;
}

// bit q of the X stretch / children region, fetched from the lane that holds it
template <int LPQ, int BW, int FIRST>
__device__ __forceinline__ uint32_t paired_bit(const BlockWords<LPQ, BW>& b, uint32_t q) {
  constexpr int SPL = BW / kPairedSliceWords / LPQ;
  const uint32_t gs = q / kPairedSlicePos, r = q - gs * kPairedSlicePos;
  const int t_sel = static_cast<int>(gs % SPL) * kPairedSliceWords + FIRST + static_cast<int>(r >> 5);
  uint32_t mine = 0;
#pragma unroll
  for (int sl = 0; sl < SPL; sl++)
#pragma unroll
    for (int t = 0; t < 3; t++) {
      const int wi = sl * kPairedSliceWords + FIRST + t;
      mine = (wi == t_sel) ? b.w[wi] : mine;
    }
  const uint32_t word = LPQ == 1 ? mine : __shfl_sync(kFull, mine, gs / SPL, LPQ);
  return (word >> (31 - (r & 31))) & 1u;
}

// header word h1 of child b1: slice b1's second word
template <int LPQ, int BW>
__device__ __forceinline__ uint32_t paired_h1(const BlockWords<LPQ, BW>& b, uint32_t b1) {
  constexpr int SPL = BW / kPairedSliceWords / LPQ;
  if (SPL >= 2) return group_lane0<LPQ>(b1 ? b.w[kPairedSliceWords + 1] : b.w[1]);
  return __shfl_sync(kFull, b.w[1], b1, LPQ);
}

// occ_descend over paired-level blocks: two wavelet-tree levels per block read.
template <int LPQ, int BW>
__device__ __forceinline__ int64_t occ_descend_paired(const DevImage& im, bool active, int c, int64_t row, int sub) {
  constexpr uint32_t B = kPairedSlicePos * (BW / kPairedSliceWords);
  int64_t occ_base = 0;
  uint32_t leaf = 0, base = 0, node = 0, idx1 = 0;
  int L = 0;
  if (active) {
    int64_t g;
    uint32_t rb;
    split_row(im, row, g, rb);
    const int4 rv = __ldg(reinterpret_cast<const int4*>(im.occ + g * kAlphaStride + c));
    occ_base = rec_occ_base(rv);
    leaf = static_cast<uint32_t>(rv.z);
    if (leaf) {
      const uint4 br = __ldg(reinterpret_cast<const uint4*>(im.buckets + g));
      base = br.x;
      node = br.y;
      L = 31 - __clz(leaf);
      idx1 = rb + 1;
    }
  }
  bool desc = active && leaf != 0;
  int lvl = 0;
  while (__any_sync(kFull, desc)) {
    uint32_t k, off;
    paired_split<BW>(desc ? idx1 - 1 : 0u, k, off);
    const bool has2 = lvl + 2 <= L;
    const uint32_t b1 = desc ? (leaf >> (L - lvl - 1)) & 1u : 0u;
    const uint32_t b2 = (desc && has2) ? (leaf >> (L - lvl - 2)) & 1u : 0u;
    BlockWords<LPQ, BW> w;
    w.clear();
    if (desc) w.load(im.blocks, base + k, sub);
    uint2 gc = make_uint2(0, 0);
    if (desc && lvl + 2 < L) gc = __ldg(reinterpret_cast<const uint2*>(im.supers[node].gc[2 * b1 + b2]));
    const uint32_t cx = group_sum<LPQ>(paired_count<LPQ, BW, kPairedX>(w, static_cast<int>(off) + 1, sub));
    const uint32_t ones1 = w.w[0] + cx;
    const uint32_t i1 = b1 ? ones1 : idx1 - ones1;        // index -= occs[!bit]  (wtree.c:1109)
    const uint32_t j = b1 ? cx : off + 1 - cx;            // of the child's first i1 bits, those stored here
    const uint32_t hi = b1 ? B - j : j;
    const uint32_t cr = group_sum<LPQ>(paired_count<LPQ, BW, kPairedR>(w, static_cast<int>(hi), sub));
    const uint32_t h1 = paired_h1<LPQ, BW>(w, b1);
    const uint32_t ones2 = b1 ? h1 - cr : h1 + cr;
    const uint32_t i2 = b2 ? ones2 : i1 - ones2;
    lvl += 2;
    if (desc) {
      idx1 = has2 ? i2 : i1;
      if (i1 == 0) idx1 = 0;
      desc = idx1 != 0 && lvl < L;
      base = gc.x;
      node = gc.y;
    }
  }
  return occ_base + static_cast<int64_t>(leaf ? idx1 : 0u);
}

// ---- quad-level blocks (fm_image.hpp): 2 lanes per group; lane `sub` holds words 2 sub and
// 2 sub + 1 (bits [64 sub, 64 sub + 64)) of each of the four regions --------------------------------
__device__ __forceinline__ uint32_t shr_clamp(uint32_t v, int s) {  // v >> s, 0 when s >= 32 (s >= 0)
  uint32_t r;
  asm("shr.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(s));
  return r;
}

// Loads the scheduler must not sink towards their first use: they are issued early on purpose, so
// that their latency overlaps the block read (ptxas otherwise moves them behind the evaluation).
__device__ __forceinline__ uint2 ldg_pinned(const uint2* ptr) {
  uint2 v;
  asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(ptr) : "memory");
  return v;
}
struct QuadWords {
  uint32_t d[8];  // d[2 l], d[2 l + 1]: this lane's two words of region l
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int t = 0; t < 8; t++) d[t] = 0;
  }
  __device__ __forceinline__ void load(const uint4* __restrict__ blocks, uint32_t blk, int sub) {
    // this lane's 32 bytes = one sector, fetched with ONE 256-bit load (sm_100: LDG.E.256): half the
    // load instructions and L1 wavefronts of two 128-bit loads
    const uint4* b = blocks + static_cast<size_t>(blk) * (kQuadBlockWords / 4) + 4 + 2 * sub;
    asm("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7])
        : "l"(b));
  }
  // the same load with an L2 eviction hint (pol from l2_evict_first_policy()): a block read once and not
  // again soon should not push out the lines other parts of a kernel keep coming back to
  __device__ __forceinline__ void load_hint(const uint4* __restrict__ blocks, uint32_t blk, int sub, uint64_t pol) {
    const uint4* b = blocks + static_cast<size_t>(blk) * (kQuadBlockWords / 4) + 4 + 2 * sub;
    asm("ld.global.nc.L2::cache_hint.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8], %9;"
        : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7])
        : "l"(b), "l"(pol));
  }
  // Masks of this lane's 64 bits at positions >= x (x relative to the lane's first bit, any
  // integer): .x for the first word, .y for the second.  ONE 64-bit shift of all-ones serves both
  // words and needs no upper clamp (shr.b64 yields 0 for amounts >= 64); the lower clamp is the
  // only other instruction.
  static __device__ __forceinline__ uint2 mask_ge(int x) {
    const uint32_t ux = static_cast<uint32_t>(max(x, 0));
    unsigned long long m;
    asm("shr.b64 %0, %1, %2;" : "=l"(m) : "l"(0xffffffffffffffffull), "r"(ux));
    return make_uint2(static_cast<uint32_t>(m >> 32), static_cast<uint32_t>(m));
  }
  // ones of region L within [a, b), a and b given relative to this lane's first bit
  template <int L>
  __device__ __forceinline__ uint32_t range(int a, int b) const {
    const uint2 ma = mask_ge(a), mb = mask_ge(b);
    return __popc(d[2 * L] & ma.x & ~mb.x) + __popc(d[2 * L + 1] & ma.y & ~mb.y);
  }
  // ones of region L at positions >= x (inv == 0) or < x (inv == ~0)
  template <int L>
  __device__ __forceinline__ uint32_t one_sided(int x, uint32_t inv) const {
    const uint2 m = mask_ge(x);
    return __popc(d[2 * L] & (m.x ^ inv)) + __popc(d[2 * L + 1] & (m.y ^ inv));
  }
  // bit `pos` of region L, fetched from the lane that holds it
  template <int L>
  __device__ __forceinline__ uint32_t bit(uint32_t pos) const {
    const uint32_t wi = pos >> 5;
    const uint32_t mine = (wi & 1u) ? d[2 * L + 1] : d[2 * L];
    const uint32_t word = __shfl_sync(kFull, mine, wi >> 1, 2);
    return (word >> (31 - (pos & 31))) & 1u;
  }
};

__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint2 quad_header_hint(const uint4* __restrict__ blocks, uint32_t blk, uint32_t path3,
                                                  uint64_t pol) {
  uint2 v;
  const uint2* p = reinterpret_cast<const uint2*>(blocks + static_cast<size_t>(blk) * (kQuadBlockWords / 4)) + path3;
  asm("ld.global.nc.L2::cache_hint.v2.u32 {%0, %1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pol));
  return v;
}

// H[2 path3], H[2 path3 + 1]: both exits below the level-3 node and, in their top bytes, the
// anchors of the level-2 and level-3 nodes on the path
__device__ __forceinline__ uint2 quad_header(const uint4* __restrict__ blocks, uint32_t blk, uint32_t path3) {
  return __ldg(reinterpret_cast<const uint2*>(blocks + static_cast<size_t>(blk) * (kQuadBlockWords / 4)) + path3);
}

// the four path bits of one block for a code of L bits of which lvl are consumed; a code that
// ends inside the block is extended with 0 bits
__device__ __forceinline__ uint32_t quad_path(uint32_t leaf, int L, int lvl) {
  const int rem = L - lvl;
  return rem >= 4 ? (leaf >> (rem - 4)) & 15u : (leaf << (4 - rem)) & 15u;
}

// One block for one position: j = positions of the block's stretch up to and including ours.
// Returns the 1-based index in the node at exit `nib`.
__device__ __forceinline__ uint32_t quad_levels(const QuadWords& w, const uint2 h, uint32_t nib, int j, int sub) {
  const int lb = 64 * sub;
  uint32_t c = group_sum<2>(popc_top(w.d[0], j - lb) + popc_top(w.d[1], j - lb - 32));
  uint32_t b = (nib >> 3) & 1u;
  j = b ? c : j - c;
  int a = b ? kQuadPos - j : 0;
  c = group_sum<2>(w.range<1>(a - lb, a + j - lb));
  b = (nib >> 2) & 1u;
  j = b ? c : j - c;
  a = static_cast<int>(h.x >> 24) - (b ? j : 0);
  c = group_sum<2>(w.range<2>(a - lb, a + j - lb));
  b = (nib >> 1) & 1u;
  j = b ? c : j - c;
  a = static_cast<int>(h.y >> 24) - (b ? j : 0);
  c = group_sum<2>(w.range<3>(a - lb, a + j - lb));
  b = nib & 1u;
  j = b ? c : j - c;
  return ((b ? h.y : h.x) & 0xffffffu) + static_cast<uint32_t>(j);
}

// occ_descend over quad-level blocks: four wavelet-tree levels per block read.
__device__ __forceinline__ int64_t occ_descend_quad(const DevImage& im, bool active, int c, int64_t row, int sub) {
  int64_t occ_base = 0;
  uint32_t leaf = 0, base = 0, node = 0, idx1 = 0;
  int L = 0;
  if (active) {
    int64_t g;
    uint32_t rb;
    split_row(im, row, g, rb);
    const int4 rv = __ldg(reinterpret_cast<const int4*>(im.occ + g * kAlphaStride + c));
    occ_base = rec_occ_base(rv);
    leaf = static_cast<uint32_t>(rv.z);
    if (leaf) {
      const uint4 br = __ldg(reinterpret_cast<const uint4*>(im.buckets + g));
      base = br.x;
      node = br.y;
      L = 31 - __clz(leaf);
      idx1 = rb + 1;
    }
  }
  bool desc = active && leaf != 0;
  int lvl = 0;
  while (__any_sync(kFull, desc)) {
    const uint32_t p = desc ? idx1 - 1 : 0u;
    const uint32_t blk = base + (p >> 7);
    const uint32_t nib = desc ? quad_path(leaf, L, lvl) : 0u;
    QuadWords w;
    w.clear();
    uint2 h = make_uint2(0, 0), ex = make_uint2(0, 0);
    if (desc) {
      w.load(im.blocks, blk, sub);
      h = quad_header(im.blocks, blk, nib >> 1);
      if (lvl + 4 < L) ex = __ldg(reinterpret_cast<const uint2*>(im.quads[node].exit[nib]));
    }
    const uint32_t r = quad_levels(w, h, nib, static_cast<int>(p & 127u) + 1, sub);
    lvl += 4;
    if (desc) {
      idx1 = r;
      desc = idx1 != 0 && lvl < L;
      base = ex.x;
      node = ex.y;
    }
  }
  return occ_base + static_cast<int64_t>(leaf ? idx1 : 0u);
}

// LV = wavelet-tree levels per block of the image: 1, 2 (paired) or 4 (quad)
template <int LPQ, int BW, int LV>
__device__ __forceinline__ int64_t occ_any(const DevImage& im, bool active, int c, int64_t row, int sub) {
  if constexpr (LV == 4) return occ_descend_quad(im, active, c, row, sub);
  else if constexpr (LV == 2) return occ_descend_paired<LPQ, BW>(im, active, c, row, sub);
  else return occ_descend<LPQ, BW, false>(im, active, c, row, sub);
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}


// All four levels of one quad block for the two positions of a backward-search step: p / hp serve
// position A, q / hq position B (the same objects when both positions lie in one block).  jA, jB:
// positions of the block's stretch up to and including ours, replaced by the count at the exit;
// returns the last path bit.  Warp-collective (one shuffle per level carries both counts).
__device__ __forceinline__ uint32_t quad_eval_pair(const QuadWords& p, const QuadWords& q, const uint2 hp,
                                                   const uint2 hq, uint32_t nib, int lb, int& jA, int& jB) {
  // level 0: the node's own stretch, prefix [0, j)
  uint32_t c = group_sum<2>(p.one_sided<0>(jA - lb, kFull) | (q.one_sided<0>(jB - lb, kFull) << 16));
  uint32_t b = (nib >> 3) & 1u;
  jA = b ? (c & 0xffffu) : jA - (c & 0xffffu);
  jB = b ? (c >> 16) : jB - (c >> 16);
  // level 1: child 0 forward from 0, child 1 backward from the end of the region: one-sided too
  uint32_t inv = b ? 0u : kFull;
  c = group_sum<2>(p.one_sided<1>((b ? kQuadPos - jA : jA) - lb, inv) |
                   (q.one_sided<1>((b ? kQuadPos - jB : jB) - lb, inv) << 16));
  b = (nib >> 2) & 1u;
  jA = b ? (c & 0xffffu) : jA - (c & 0xffffu);
  jB = b ? (c >> 16) : jB - (c >> 16);
  // level 2: anchored at the header's level-2 anchor
  int aA = static_cast<int>(hp.x >> 24) - (b ? jA : 0) - lb;
  int aB = static_cast<int>(hq.x >> 24) - (b ? jB : 0) - lb;
  c = group_sum<2>(p.range<2>(aA, aA + jA) | (q.range<2>(aB, aB + jB) << 16));
  b = (nib >> 1) & 1u;
  jA = b ? (c & 0xffffu) : jA - (c & 0xffffu);
  jB = b ? (c >> 16) : jB - (c >> 16);
  // level 3
  aA = static_cast<int>(hp.y >> 24) - (b ? jA : 0) - lb;
  aB = static_cast<int>(hq.y >> 24) - (b ? jB : 0) - lb;
  c = group_sum<2>(p.range<3>(aA, aA + jA) | (q.range<3>(aB, aB + jB) << 16));
  b = nib & 1u;
  jA = b ? (c & 0xffffu) : jA - (c & 0xffffu);
  jB = b ? (c >> 16) : jB - (c >> 16);
  return b;
}


// Both Occ of one backward-search round over quad-level blocks, for every group of the warp
// (warp-collective; the loop of count_sync_kernel's quad branch as a function, used by the mesh
// kernels).  On entry idxA / idxB are the 1-based positions in the bucket's root sequence of the
// rows asked for (actA / actB say which of them), base / node the root quad node, leaf the symbol's
// code (1 << L | code), rexit OccRec::root_exit.  On return idxA / idxB are the occurrences of the
// symbol among the bucket's rows up to and including the row (wtree_occs, wtree.c:1081-1115).
// HINT: the block reads carry the L2 evict-first policy `pol`.
template <bool HINT = false>
__device__ __forceinline__ void quad_descend_pair(const DevImage& im, bool actA, bool actB, uint32_t& idxA,
                                                  uint32_t& idxB, uint32_t base, uint32_t node, uint32_t leaf, int L,
                                                  uint32_t rexit, int sub, uint64_t pol = 0) {
  int lvl = 0;
  while (__any_sync(kFull, actA || actB)) {
    const bool any = actA || actB;
    const uint32_t pA = actA ? idxA - 1 : 0u, pB = actB ? idxB - 1 : 0u;
    const uint32_t kA = pA >> 7, kB = pB >> 7;
    const bool two = actA && actB && kA != kB;
    const uint32_t blkA = base + (actA ? kA : kB), blkB = base + (actB ? kB : kA);
    const uint32_t nib = any ? quad_path(leaf, L, lvl) : 0u;
    QuadWords p, q;
    p.clear();
    q.clear();
    uint2 hp = make_uint2(0, 0), hq = hp, ex = hp;
    if (any) {
      if (HINT) {
        p.load_hint(im.blocks, blkA, sub, pol);
        hp = quad_header_hint(im.blocks, blkA, nib >> 1, pol);
        if (two) {
          q.load_hint(im.blocks, blkB, sub, pol);
          hq = quad_header_hint(im.blocks, blkB, nib >> 1, pol);
        }
      } else {
        p.load(im.blocks, blkA, sub);
        hp = quad_header(im.blocks, blkA, nib >> 1);
        if (two) {
          q.load(im.blocks, blkB, sub);
          hq = quad_header(im.blocks, blkB, nib >> 1);
        }
      }
      if (lvl + 4 < L) {
        if (lvl == 0 && (rexit & kRootExitDirect)) ex = make_uint2(rexit & ~kRootExitDirect, 0u);
        else ex = __ldg(reinterpret_cast<const uint2*>(im.quads[node].exit[nib]));
      }
    }
    const int lb = 64 * sub;
    int jA = static_cast<int>(pA & 127u) + 1, jB = static_cast<int>(pB & 127u) + 1;
    uint32_t b;
    if (__any_sync(kFull, two)) {
      if (!two) {
        hq = hp;
#pragma unroll
        for (int t = 0; t < 8; t++) q.d[t] = p.d[t];
      }
      b = quad_eval_pair(p, q, hp, hq, nib, lb, jA, jB);
    } else {
      hq = hp;
      b = quad_eval_pair(p, p, hp, hp, nib, lb, jA, jB);
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
}

// wtree_rank over quad-level blocks (wtree.c:1117-1148): follows the bits found at row rb of local
// bucket g down to a leaf.  Returns the symbol there and `count` = this row holds the count-th
// occurrence of it in the bucket.  Warp-collective; 2 lanes per group.
//   * the root block is addressed by the row alone (root area, fm_image.hpp) and requested FIRST;
//     the bucket record (root QuadRec, base of the SA samples) is read beside it, not before it;
//   * both header sectors of a block are requested together with its bit regions, so that the
//     anchors and exit counts, whose position depends on the bits found, are L1 hits by the time
//     they are needed instead of a second trip to memory.
__device__ __forceinline__ void quad_wtree_rank(const DevImage& im, bool act, int64_t g, uint32_t rb, int sub,
                                                uint32_t& ch, uint32_t& count, uint64_t& markval_base,
                                                unsigned long long& n_blocks) {
  const uint32_t* words = reinterpret_cast<const uint32_t*>(im.blocks);
  uint32_t base = 0, node = 0, idx1 = 0;
  bool desc = act;
  bool root = true;
  ch = 0;
  markval_base = 0;
  if (act) {
    base = static_cast<uint32_t>(g * im.root_stride);
    idx1 = rb + 1;
  }
  while (__any_sync(kFull, desc)) {
    const uint32_t p = desc ? idx1 - 1 : 0u;
    const uint32_t blk = base + (p >> 7);
    const uint32_t* hw = words + static_cast<size_t>(blk) * kQuadBlockWords;
    QuadWords w;
    w.clear();
    if (desc) {
      w.load(im.blocks, blk, sub);
      // touch both header sectors now (lane `sub` takes sector `sub`)
      uint32_t t0, t1, t2, t3;
      asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(t0), "=r"(t1), "=r"(t2), "=r"(t3) : "l"(hw + 8 * sub));
      n_blocks += (sub == 0);
    }
    if (root) {  // first block of the walk step: the bucket record travels beside it
      if (desc) {
        const uint4 br = __ldg(reinterpret_cast<const uint4*>(im.buckets + g));
        node = br.y;
        markval_base = static_cast<uint64_t>(br.z) | (static_cast<uint64_t>(br.w) << 32);
      }
      root = false;
    }
    const int lb = 64 * sub;
    int j = static_cast<int>(p & 127u) + 1;
    // level 0: follow the bit at our position; j stays >= 1 all the way down
    uint32_t c = group_sum<2>(popc_top(w.d[0], j - lb) + popc_top(w.d[1], j - lb - 32));
    uint32_t b = w.bit<0>(static_cast<uint32_t>(j - 1));
    uint32_t path = b;
    j = b ? c : j - c;
    // level 1
    int a = b ? kQuadPos - j : 0;
    c = group_sum<2>(w.range<1>(a - lb, a + j - lb));
    uint32_t nb = w.bit<1>(static_cast<uint32_t>(b ? a : a + j - 1));
    j = nb ? c : j - c;
    b = nb;
    path = (path << 1) | b;
    // level 2: anchor in the top byte of H[4 * (b1 b2)]
    a = static_cast<int>((desc ? __ldg(hw + (path << 2)) : 0u) >> 24) - (b ? j : 0);
    c = group_sum<2>(w.range<2>(a - lb, a + j - lb));
    nb = w.bit<2>(static_cast<uint32_t>(b ? a : a + j - 1));
    j = nb ? c : j - c;
    b = nb;
    path = (path << 1) | b;
    // level 3: anchor in the top byte of H[2 * (b1 b2 b3) + 1]; the exits' counts below it
    const uint2 h = desc ? __ldg(reinterpret_cast<const uint2*>(hw) + path) : make_uint2(0, 0);
    a = static_cast<int>(h.y >> 24) - (b ? j : 0);
    c = group_sum<2>(w.range<3>(a - lb, a + j - lb));
    nb = w.bit<3>(static_cast<uint32_t>(b ? a : a + j - 1));
    j = nb ? c : j - c;
    path = (path << 1) | nb;
    if (desc) {
      const uint2 ex = __ldg(reinterpret_cast<const uint2*>(im.quads[node].exit[path]));
      idx1 = ((nb ? h.y : h.x) & 0xffffffu) + static_cast<uint32_t>(j);
      if (ex.y & kChildLeaf) {
        ch = ex.y & 0xffffu;
        desc = false;
      } else {
        base = ex.x;
        node = ex.y;
      }
    }
  }
  count = idx1;
}

// The mark test and SA sample of a row that holds the count-th occurrence of ch in local bucket g
// (block_request LOCATION branch, src/main/index.c:2102-2140), and what LF needs: returns
// offset = SA[row] or -1 when the row is not marked, occ_base = C[ch] + occurrences of ch before the
// bucket.  Mark bit-vectors are plain one-level blocks in every image.  Warp-collective.
template <int LPQ, int BW>
__device__ __forceinline__ void mark_lookup(const DevImage& im, bool ok, int64_t g, uint32_t ch, uint32_t count,
                                            uint64_t markval_base, int sub, int64_t& offset, int64_t& occ_base,
                                            unsigned long long& n_mark_blocks, unsigned long long& n_samples) {
  constexpr uint32_t BITS = (BW - 1) * 32;
  uint32_t mark_base = 0, markval_off = 0;
  occ_base = 0;
  if (ok) {
    const size_t rec = static_cast<size_t>(g) * kAlphaStride + ch;
    const uint2 mr = __ldg(reinterpret_cast<const uint2*>(im.mark + rec));
    mark_base = mr.x;
    markval_off = mr.y;
    occ_base = rec_occ_base(__ldg(reinterpret_cast<const int4*>(im.occ + rec)));
    n_mark_blocks += (sub == 0);
  }
  const uint32_t mp = ok ? count - 1 : 0u;
  const uint32_t mk = mp / BITS;
  const uint32_t moff = mp - mk * BITS;
  uint32_t mones, mbit;
  block_rank<LPQ, BW, true>(im.blocks, mark_base + mk, moff, ok, sub, mones, mbit);
  offset = -1;
  if (ok && mbit) {
    offset = __ldg(im.markvals + markval_base + markval_off + (mones - 1));
    n_samples += (sub == 0);
  }
}

}  // namespace
}  // namespace fmb
