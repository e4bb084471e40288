// fm_kernels.cu -- hand-written sm_100a kernels of the FM-index query engine.
//
// What runs here, per reference function (paths relative to the reference tree):
//   count_kernel  : do_string_query's backward search loop (src/main/server.c:713-946) with both
//                   Occ evaluations of a step -- header_occs_request(HDR_BACK) + block_request(OCCS)
//                   (src/main/index.c:1698-1765, 1973-2100) -> wtree_occs (src/main/wtree.c:1081-1115)
//                   -> bseq_rank (wtree.c:635-763) -- as one rank-block read per wavelet-tree level.
//   walk_kernel   : do_back_query (server.c:2228-2359) = wtree_rank (wtree.c:1117-1148) + mark test +
//                   sampled-SA read (index.c:2037-2140) + LF; iterated for locate
//                   (do_context_query, server.c:2627-2795, backward half) and document extract
//                   (do_extract_document_query, server.c:6364-6437).
//   occ_kernel    : a batch of single C[ch]+Occ(ch,row) evaluations (leaf interface cross-check).
//
// Execution model: a rank query is served by LPQ (4 or 8) adjacent lanes that together read ONE
// 128-byte rank block with 128-bit loads (ld.global.nc.v4), popcount their words under a position
// mask and combine with __shfl_xor_sync.  A pattern owns 2*LPQ lanes: one sub-group evaluates
// Occ(c, first-1), the other Occ(c, last), concurrently -- so a warp carries 32/(2*LPQ) patterns and
// 32/LPQ independent 128-byte HBM reads per step.  Warps are persistent: finished pattern groups pull
// the next pattern from a global atomic queue, so early-dying patterns and mixed lengths do not idle
// lanes for the rest of the batch.  No tensor cores: this is HBM-latency/bandwidth-bound integer work.
#include "fm_kernels.cuh"

namespace fmb {
namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kThreads = 256;
constexpr int kEscSeofDev = 2;       // ESCAPE_CODE_SEOF, src/main/index_types.h:42-48
constexpr int kAlphaDev = 261;

__device__ __forceinline__ uint32_t popc_top(uint32_t w, int n) {
  // ones among the n (0..32) most significant bits of w
  const uint32_t mask = static_cast<uint32_t>(0xFFFFFFFF00000000ull >> n);
  return __popc(w & mask);
}

// Rank inside one 128-byte block, cooperatively by the LPQ lanes of a sub-group.
//   blk : rank block index, off : 0-based bit offset inside the block payload (0..991)
// Returns ones in the node's sequence up to and including the addressed bit (header + in-block),
// and optionally the bit itself.  Inactive sub-groups issue no loads but take part in the shuffles.
template <int LPQ, bool WANT_BIT>
__device__ __forceinline__ void block_rank(const uint4* __restrict__ blocks, uint32_t blk, uint32_t off,
                                           bool active, int sub, uint32_t& ones_incl, uint32_t& bit) {
  constexpr int WPL = 32 / LPQ;  // 32-bit words per lane
  constexpr int VPL = WPL / 4;   // 128-bit loads per lane
  uint32_t w[WPL];
#pragma unroll
  for (int t = 0; t < WPL; t++) w[t] = 0;
  if (active) {
    const uint4* p = blocks + static_cast<size_t>(blk) * 8 + sub * VPL;
#pragma unroll
    for (int v = 0; v < VPL; v++) {
      const uint4 x = __ldg(p + v);
      w[4 * v + 0] = x.x; w[4 * v + 1] = x.y; w[4 * v + 2] = x.z; w[4 * v + 3] = x.w;
    }
  }
  // block bit space: word 0 is the header, payload bit `off` sits at position 32+off
  const int upto = static_cast<int>(off) + 33;  // number of block bit positions up to and incl. the bit
  uint32_t cnt = 0;
#pragma unroll
  for (int t = 0; t < WPL; t++) {
    const int wi = sub * WPL + t;
    int n = upto - 32 * wi;
    n = max(0, min(32, n));
    const uint32_t word = (t == 0 && sub == 0) ? 0u : w[t];
    cnt += popc_top(word, n);
  }
#pragma unroll
  for (int o = LPQ / 2; o > 0; o >>= 1) cnt += __shfl_xor_sync(kFull, cnt, o);
  const uint32_t hdr = __shfl_sync(kFull, w[0], 0, LPQ);
  ones_incl = hdr + cnt;
  if (WANT_BIT) {
    const int wq = (static_cast<int>(off) + 32) >> 5;
    const int t_sel = wq % WPL;
    uint32_t mine = 0;
#pragma unroll
    for (int t = 0; t < WPL; t++) mine = (t == t_sel) ? w[t] : mine;
    const uint32_t word = __shfl_sync(kFull, mine, wq / WPL, LPQ);
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

// C[c] + Occ(c,row) for the sub-group's query (uniform across its LPQ lanes).  Warp-collective:
// every lane of the warp must call it; inactive sub-groups pass active=false and get 0.
// STATS (instrumented launches only): n_reads counts rank blocks requested, n_distinct counts them
// once when the partner sub-group (the other Occ of the same backward-search step) asks for the
// same block at the same level -- the bytes the step needs by design.
template <int LPQ, bool STATS = false>
__device__ __forceinline__ int64_t occ_descend(const DevImage& im, bool active, int c, int64_t row, int sub,
                                               unsigned long long* n_reads = nullptr,
                                               unsigned long long* n_distinct = nullptr) {
  int64_t occ_base = 0;
  uint32_t leaf = 0, base = 0, node = 0, idx1 = 0;
  int L = 0;
  if (active) {
    int64_t g;
    uint32_t rb;
    split_row(im, row, g, rb);
    const int4 rv = __ldg(reinterpret_cast<const int4*>(im.occ + g * kAlphaStride + c));
    occ_base = static_cast<int64_t>(static_cast<uint32_t>(rv.x)) | (static_cast<int64_t>(rv.y) << 32);
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
    const uint32_t k = p / kBitsPerBlock;
    const uint32_t off = p - k * kBitsPerBlock;
    uint4 nr = make_uint4(0, 0, 0, 0);
    if (desc && lvl + 1 < L) nr = __ldg(reinterpret_cast<const uint4*>(im.nodes + node));
    uint32_t ones, bit;
    block_rank<LPQ, false>(im.blocks, base + k, off, desc, sub, ones, bit);
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

// ---------------------------------------------------------------------------------------------
template <int LPQ, bool STATS>
__global__ void __launch_bounds__(kThreads) count_kernel(const DevImage im, const CountArgs a,
                                                          unsigned long long* __restrict__ work,
                                                          unsigned long long* __restrict__ stats) {
  unsigned long long n_reads = 0, n_distinct = 0, n_occ = 0, n_steps = 0;
  constexpr int GL = 2 * LPQ;  // lanes per pattern
  const int lane = threadIdx.x & 31;
  const int sub = lane & (LPQ - 1);
  const int which = (lane / LPQ) & 1;  // 0: Occ(c, first-1)   1: Occ(c, last)
  const int gleader = lane & ~(GL - 1);

  int64_t f = 0, l = -1, pid = -1;
  int i = 0;
  const uint16_t* pat = nullptr;
  bool have = false, exhausted = false;

  for (;;) {
    // retire: "first > last || i == 0" ends the reference's while loop (server.c:832-841)
    if (have && (f > l || i == 0)) {
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
          const int c = pat[m - 1];
          if (c >= kAlphaDev) { f = im.total_length; l = f - 1; }  // get_C(ch>=ALPHA_SIZE), index.c:1545
          else { f = __ldg(im.C + c); l = __ldg(im.C + c + 1) - 1; }
          i = m - 1;
        }
        have = true;
      } else {
        exhausted = true;
      }
    }
    if (!__any_sync(kFull, have)) break;

    const bool stepping = have && f <= l && i > 0;
    int c = 0;
    int64_t row = 0;
    bool q = false, badc = false;
    if (stepping) {
      c = pat[i - 1];
      badc = c >= kAlphaDev;
      row = which ? l : f - 1;
      q = !badc && row >= 0;  // first == 0: Occ(c,-1) = 0 without touching the index (server.c:847-851)
    }
    int64_t r = occ_descend<LPQ, STATS>(im, q, c, row, sub, &n_reads, &n_distinct);
    if (STATS && sub == 0) {
      n_occ += q ? 1 : 0;
      n_steps += (stepping && which == 0) ? 1 : 0;
    }
    if (stepping && !q && !badc) r = __ldg(im.C + c);
    const int64_t other = __shfl_xor_sync(kFull, r, LPQ);
    if (stepping) {
      if (badc) {
        f = im.total_length; l = f - 1;
      } else {
        f = which ? other : r;
        l = (which ? r : other) - 1;
      }
      i--;
    }
  }
  if (STATS) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      n_reads += __shfl_xor_sync(kFull, n_reads, o);
      n_distinct += __shfl_xor_sync(kFull, n_distinct, o);
      n_occ += __shfl_xor_sync(kFull, n_occ, o);
      n_steps += __shfl_xor_sync(kFull, n_steps, o);
    }
    if (lane == 0) {
      atomicAdd(stats + 0, n_reads);
      atomicAdd(stats + 1, n_distinct);
      atomicAdd(stats + 2, n_occ);
      atomicAdd(stats + 3, n_steps);
    }
  }
}

// ---------------------------------------------------------------------------------------------
template <int LPQ, int MODE>
__global__ void __launch_bounds__(kThreads) walk_kernel(const DevImage im, const WalkArgs a,
                                                         unsigned long long* __restrict__ work) {
  const int lane = threadIdx.x & 31;
  const int sub = lane & (LPQ - 1);
  const int gleader = lane & ~(LPQ - 1);

  int64_t row = 0, rid = -1, steps = 0, nsteps = 0, sym_off = 0;
  bool have = false, exhausted = false;

  for (;;) {
    const bool need = !have && !exhausted;
    unsigned long long idx = 0;
    if (need && lane == gleader) idx = atomicAdd(work, 1ull);
    idx = __shfl_sync(kFull, idx, gleader);
    if (need) {
      if (static_cast<int64_t>(idx) < a.nrows) {
        rid = static_cast<int64_t>(idx);
        row = a.rows[rid];
        steps = 0;
        have = true;
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
      if (lane == gleader) {
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
      const uint4 br = __ldg(reinterpret_cast<const uint4*>(im.buckets + g));
      base = br.x;
      node = br.y;
      markval_base = static_cast<uint64_t>(br.z) | (static_cast<uint64_t>(br.w) << 32);
      idx1 = rb + 1;
    }
    bool desc = act;
    while (__any_sync(kFull, desc)) {
      const uint32_t p = desc ? idx1 - 1 : 0u;
      const uint32_t k = p / kBitsPerBlock;
      const uint32_t off = p - k * kBitsPerBlock;
      uint4 nr = make_uint4(0, 0, 0, 0);
      if (desc) nr = __ldg(reinterpret_cast<const uint4*>(im.nodes + node));
      uint32_t ones, bit;
      block_rank<LPQ, true>(im.blocks, base + k, off, desc, sub, ones, bit);
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
    uint32_t mark_base = 0, markval_off = 0;
    int64_t occ_base = 0;
    if (ok) {
      const size_t rec = static_cast<size_t>(g) * kAlphaStride + ch;
      const uint2 mr = __ldg(reinterpret_cast<const uint2*>(im.mark + rec));
      mark_base = mr.x;
      markval_off = mr.y;
      const int4 rv = __ldg(reinterpret_cast<const int4*>(im.occ + rec));
      occ_base = static_cast<int64_t>(static_cast<uint32_t>(rv.x)) | (static_cast<int64_t>(rv.y) << 32);
    }
    const uint32_t mp = ok ? count - 1 : 0u;
    const uint32_t mk = mp / kBitsPerBlock;
    const uint32_t moff = mp - mk * kBitsPerBlock;
    uint32_t mones, mbit;
    block_rank<LPQ, true>(im.blocks, mark_base + mk, moff, ok, sub, mones, mbit);
    int64_t offset = -1;
    if (ok && mbit) offset = __ldg(im.markvals + markval_base + markval_off + (mones - 1));
    // LF: row' = C[ch] + occs before the bucket + count - 1; stop at a document boundary
    // (ch <= ESCAPE_CODE_SEOF, server.c:2341-2346)
    const int64_t next = (ch <= static_cast<uint32_t>(kEscSeofDev)) ? -1 : occ_base + count - 1;

    if (act && !ok) {
      if (lane == gleader) {
        atomicExch(a.status, 2);
        if (MODE != kWalkExtract) a.out_offset[rid] = -1;
      }
      have = false;
    } else if (act) {
      if (MODE == kWalkLocate) {
        if (offset >= 0) {
          if (lane == gleader) a.out_offset[rid] = offset + steps;
          have = false;
        } else if (next < 0) {  // unmarked document start: the index violates should_mark()
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
}

// ---------------------------------------------------------------------------------------------
template <int LPQ>
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
    const int64_t r = occ_descend<LPQ>(im, q, c, row, sub);
    if (item < a.n && sub == 0) a.out[item] = q ? r : -1;
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
}  // namespace fmb

#include "fm_count_merged.cuh"

namespace fmb {

cudaError_t launch_count(const DevImage& im, const CountArgs& a, unsigned long long* d_work, int lpq, int sm_count,
                         cudaStream_t stream, int64_t* launch_counter, unsigned long long* d_stats) {
  if (a.npats <= 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(d_work, 0, sizeof(unsigned long long), stream);
  if (e != cudaSuccess) return e;
  // lpq encodes the schedule: 4/8 = "pair" (two sub-groups per pattern), 100+{2,4,8} = "merged pair"
  // (fm_count_merged.cuh).  d_stats selects the instrumented twin (never the timed path).
#define FM_LAUNCH_COUNT(KERNEL, LANES_PER_PATTERN)                                                      \
  do {                                                                                                  \
    static const int bps = blocks_per_sm(KERNEL);                                                       \
    KERNEL<<<grid_for(a.npats, kThreads / (LANES_PER_PATTERN), sm_count, bps), kThreads, 0, stream>>>(   \
        im, a, d_work, d_stats);                                                                        \
  } while (0)
  // merged schedule codes: 1000 + 10*lanes + min resident blocks per SM (register budget)
#define FM_MERGED_CASE(LANES, MINB)                                                                      \
    case 1000 + 10 * (LANES) + (MINB):                                                                   \
      if (d_stats) FM_LAUNCH_COUNT((count_merged_kernel<LANES, MINB, true>), LANES);                     \
      else FM_LAUNCH_COUNT((count_merged_kernel<LANES, MINB, false>), LANES);                            \
      break;
  switch (lpq) {
    FM_MERGED_CASE(2, 3) FM_MERGED_CASE(2, 4)
    FM_MERGED_CASE(4, 4) FM_MERGED_CASE(4, 5) FM_MERGED_CASE(4, 6)
    FM_MERGED_CASE(8, 5) FM_MERGED_CASE(8, 6)
    case 8: if (d_stats) FM_LAUNCH_COUNT((count_kernel<8, true>), 16); else FM_LAUNCH_COUNT((count_kernel<8, false>), 16); break;
    case 4: if (d_stats) FM_LAUNCH_COUNT((count_kernel<4, true>), 8); else FM_LAUNCH_COUNT((count_kernel<4, false>), 8); break;
    default: return cudaErrorInvalidValue;
  }
#undef FM_MERGED_CASE
#undef FM_LAUNCH_COUNT
  if (launch_counter) ++*launch_counter;
  return cudaGetLastError();
}

template <int LPQ>
static cudaError_t launch_walk_lpq(const DevImage& im, const WalkArgs& a, WalkMode mode, unsigned long long* d_work,
                                   int sm_count, cudaStream_t stream) {
  const int gpb = kThreads / LPQ;
  if (mode == kWalkLocate) {
    static const int bps = blocks_per_sm(walk_kernel<LPQ, kWalkLocate>);
    walk_kernel<LPQ, kWalkLocate><<<grid_for(a.nrows, gpb, sm_count, bps), kThreads, 0, stream>>>(im, a, d_work);
  } else if (mode == kWalkStep) {
    static const int bps = blocks_per_sm(walk_kernel<LPQ, kWalkStep>);
    walk_kernel<LPQ, kWalkStep><<<grid_for(a.nrows, gpb, sm_count, bps), kThreads, 0, stream>>>(im, a, d_work);
  } else {
    static const int bps = blocks_per_sm(walk_kernel<LPQ, kWalkExtract>);
    walk_kernel<LPQ, kWalkExtract><<<grid_for(a.nrows, gpb, sm_count, bps), kThreads, 0, stream>>>(im, a, d_work);
  }
  return cudaGetLastError();
}

cudaError_t launch_walk(const DevImage& im, const WalkArgs& a, WalkMode mode, unsigned long long* d_work, int lpq,
                        int sm_count, cudaStream_t stream, int64_t* launch_counter) {
  if (a.nrows <= 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(d_work, 0, sizeof(unsigned long long), stream);
  if (e != cudaSuccess) return e;
  e = (lpq == 8) ? launch_walk_lpq<8>(im, a, mode, d_work, sm_count, stream)
                 : launch_walk_lpq<4>(im, a, mode, d_work, sm_count, stream);
  if (launch_counter) ++*launch_counter;
  return e;
}

cudaError_t launch_occ(const DevImage& im, const OccArgs& a, unsigned long long* /*d_work*/, int lpq, int sm_count,
                       cudaStream_t stream, int64_t* launch_counter) {
  if (a.n <= 0) return cudaSuccess;
  if (lpq == 8) {
    static const int bps = blocks_per_sm(occ_kernel<8>);
    occ_kernel<8><<<grid_for(a.n, kThreads / 8, sm_count, bps), kThreads, 0, stream>>>(im, a);
  } else {
    static const int bps = blocks_per_sm(occ_kernel<4>);
    occ_kernel<4><<<grid_for(a.n, kThreads / 4, sm_count, bps), kThreads, 0, stream>>>(im, a);
  }
  if (launch_counter) ++*launch_counter;
  return cudaGetLastError();
}

}  // namespace fmb
