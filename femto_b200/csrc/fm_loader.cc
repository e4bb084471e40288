// fm_loader.cc -- decode a femto index into the host copy of the rank image.
//
// Replaces, as a one-time load step, what the reference does lazily per query:
// open_data_block / b_fault (src/main/index.c:1419-1468, 1222-1342) and the per-rank
// decoding inside bseq_rank (src/main/wtree.c:635-763).
#include "fm_loader.hpp"

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <mutex>
#include <thread>
#include <unordered_map>

namespace fmb {

HostImage::~HostImage() { std::free(rank_words); }

namespace {

struct BucketPlan {
  BucketTables tab;
  std::vector<uint32_t> node_ids;   // internal node ids (heap numbering), sorted
  std::vector<uint32_t> node_offs;  // bseq offset of each from the wtree start (0 = no data)
  std::vector<int64_t> node_bits;
  std::vector<uint32_t> mark_offs;  // per in-use seq: offset of the mark-table bseq from off_marktab
  std::vector<int64_t> mark_bits;
  std::vector<uint32_t> markarr_offs;
  int64_t n_blocks = 0, n_wtree_blocks = 0;
  int64_t n_root_blocks = 0;  // quad layout: blocks of the root node, placed outside [block_base, +n_blocks)
  uint32_t chunk_dir = 0;     // absolute offset of the chunk directory in the block (0 = no chunks)
  int32_t n_chunks = 0;
  int64_t chunk_len = 0, chunk_base = 0;  // bytes of the chunk section; where they go in HostImage::chunk_bytes
  // paired-level layout: internal nodes at even depth own the blocks ("super nodes")
  std::vector<int32_t> super_of;    // per internal node: its index among the bucket's super nodes, or -1
  int64_t n_records = 0;            // NodeRec (plain) or SuperRec (paired) entries of this bucket
  // assigned bases
  int64_t node_base = 0, block_base = 0, markval_base = 0;
  int64_t n_markvals = 0;  // total ones over the bucket's mark tables
};

inline int node_depth(uint32_t heap_id) { return 31 - __builtin_clz(heap_id); }

// rank block geometry chosen at load time: bw 32-bit words per block, (bw-1)*32 payload bits
thread_local int t_block_words = kDefaultBlockWords;
inline int64_t bits_per_block() { return int64_t(t_block_words - 1) * 32; }
inline int64_t blocks_for_bits(int64_t nbits) {
  return std::max<int64_t>(1, (nbits + bits_per_block() - 1) / bits_per_block());
}

// paired-level layout: a block of a super node X holds 96 positions of X per 32-byte slice and the
// same positions' bits of X's two children (see fm_image.hpp)
inline int64_t paired_positions(int bw) { return int64_t(kPairedSlicePos) * (bw / kPairedSliceWords); }
inline int64_t paired_blocks_for_bits(int64_t nbits, int bw) {
  return std::max<int64_t>(1, (nbits + paired_positions(bw) - 1) / paired_positions(bw));
}

void plan_bucket(const Blob& blk, const BlockHeader& bh, int bpb, int bucket, BucketPlan* p, int bw, int levels) {
  t_block_words = bw;
  parse_bucket_tables(blk, bh, bpb, bucket, &p->tab);
  const BucketTables& t = p->tab;
  {  // document chunks: bucket header word 5 = number of chunks, directory right after the header
    const int64_t nch = int64_t(be32(blk.at(size_t(t.off_bucket) + 20, 4)));
    if (nch < 0 || nch > (int64_t(1) << 24)) throw Error(FM_ERR_FORMAT, "bad chunk count");
    p->n_chunks = int32_t(nch);
    if (nch > 0) {
      p->chunk_dir = t.off_bucket + 24;
      const uint8_t* dir = blk.at(p->chunk_dir, 4 * size_t(nch + 1));
      const uint32_t lo = be32(dir), hi = be32(dir + 4 * size_t(nch));
      if (lo != 24 + 4 * uint32_t(nch + 1) || hi < lo) throw Error(FM_ERR_FORMAT, "bad chunk directory");
      blk.at(size_t(t.off_bucket) + hi - 1, 1);
      p->chunk_len = int64_t(hi) - 24;  // directory + chunks
    }
  }
  // wavelet tree directory
  const uint8_t* wt = blk.at(t.off_wtree, 4);
  const size_t wt_avail = size_t(t.off_marktab) - size_t(t.off_wtree);
  const uint32_t n_int = be32(wt);
  if (4 + 8 * size_t(n_int) > wt_avail) throw Error(FM_ERR_FORMAT, "wavelet tree directory truncated");
  if (n_int == 0 || n_int > uint32_t(kAlpha + 1)) throw Error(FM_ERR_FORMAT, "bad wavelet tree node count");
  p->node_ids.resize(n_int);
  p->node_offs.resize(n_int);
  p->node_bits.resize(n_int);
  p->n_blocks = 0;
  for (uint32_t k = 0; k < n_int; k++) {
    p->node_ids[k] = be32(wt + 4 + 8 * size_t(k));
    p->node_offs[k] = be32(wt + 8 + 8 * size_t(k));
    if (k && p->node_ids[k] <= p->node_ids[k - 1]) throw Error(FM_ERR_FORMAT, "wavelet tree directory not sorted");
    int64_t nbits = 0;
    if (p->node_offs[k] != 0) {
      if (p->node_offs[k] >= wt_avail || (p->node_offs[k] & 7u)) throw Error(FM_ERR_FORMAT, "bad bseq offset");
      nbits = bseq_length(open_bseq(wt + p->node_offs[k], wt_avail - p->node_offs[k]));
    }
    p->node_bits[k] = nbits;
    if (levels == 1) p->n_blocks += blocks_for_bits(nbits);
  }
  if (p->node_ids[0] != 1) throw Error(FM_ERR_FORMAT, "wavelet tree has no root");
  p->n_records = int64_t(n_int);
  if (levels > 1) {  // nodes at depth 0, levels, 2*levels, ... own the blocks
    p->super_of.assign(n_int, -1);
    int32_t ns = 0;
    for (uint32_t k = 0; k < n_int; k++) {
      if (node_depth(p->node_ids[k]) % levels == 0) {
        p->super_of[k] = ns++;
        const int64_t nblk = levels == 2 ? paired_blocks_for_bits(p->node_bits[k], bw)
                                         : std::max<int64_t>(1, (p->node_bits[k] + kQuadPos - 1) / kQuadPos);
        // quad layout: the root's blocks live in the directly addressed root area (fm_image.hpp)
        if (levels == 4 && k == 0) p->n_root_blocks = nblk;
        else p->n_blocks += nblk;
      }
    }
    p->n_records = ns;
  }
  p->n_wtree_blocks = p->n_blocks;
  // mark tables / arrays: u32 offsets per in-use symbol, relative to the section start
  const int n = t.n_in_use;
  const size_t mt_avail = size_t(t.off_markarr) - size_t(t.off_marktab);
  const uint8_t* mt = blk.at(t.off_marktab, mt_avail);
  const size_t ma_end = t.off_end > t.off_markarr ? t.off_end : blk.size();
  const size_t ma_avail = ma_end - size_t(t.off_markarr);
  const uint8_t* ma = blk.at(t.off_markarr, ma_avail);
  if (4 * size_t(n) > mt_avail || 4 * size_t(n) > ma_avail) throw Error(FM_ERR_FORMAT, "mark offsets truncated");
  p->mark_offs.resize(size_t(n));
  p->mark_bits.resize(size_t(n));
  p->markarr_offs.resize(size_t(n));
  int64_t mark_ones_total = 0;
  for (int s = 0; s < n; s++) {
    p->mark_offs[size_t(s)] = be32(mt + 4 * size_t(s));
    p->markarr_offs[size_t(s)] = be32(ma + 4 * size_t(s));
    if (p->mark_offs[size_t(s)] >= mt_avail || (p->mark_offs[size_t(s)] & 7u))
      throw Error(FM_ERR_FORMAT, "bad mark table offset");
    if (p->markarr_offs[size_t(s)] > ma_avail) throw Error(FM_ERR_FORMAT, "bad mark array offset");
    int64_t ones = 0;
    const int64_t nbits = bseq_length(open_bseq(mt + p->mark_offs[size_t(s)], mt_avail - p->mark_offs[size_t(s)]), &ones);
    p->mark_bits[size_t(s)] = nbits;
    p->n_blocks += blocks_for_bits(nbits);
    mark_ones_total += ones;
  }
  p->n_markvals = mark_ones_total;
}

// Fill rank blocks for one expanded bit sequence. bits: MSB-first words (zero padded).
// Returns the number of ones.
int64_t fill_blocks(const uint32_t* bits, int64_t nbits, uint32_t* dst_blocks) {
  const int64_t nb = blocks_for_bits(nbits);
  const int64_t nwords = (nbits + 31) / 32;
  const int bw = t_block_words, pw = bw - 1;
  uint32_t ones = 0;
  for (int64_t k = 0; k < nb; k++) {
    uint32_t* w = dst_blocks + k * bw;
    w[0] = ones;
    const int64_t w0 = k * pw;
    for (int j = 0; j < pw; j++) {
      const uint32_t v = (w0 + j < nwords) ? bits[w0 + j] : 0u;
      w[1 + j] = v;
      ones += uint32_t(__builtin_popcount(v));
    }
  }
  return ones;
}

struct Scratch {
  std::vector<uint32_t> bits;
  uint32_t* get(int64_t nbits) {
    const size_t words = size_t((nbits + 31) / 32) + 1;
    if (bits.size() < words) bits.resize(words);
    std::fill(bits.begin(), bits.begin() + words, 0u);
    return bits.data();
  }
};

// ---- paired-level layout ------------------------------------------------------------------
// copy `len` bits from src (MSB-first words) starting at bit src_off to dst starting at dst_off;
// dst bits must be zero beforehand
void copy_bits(uint32_t* dst, int64_t dst_off, const uint32_t* src, int64_t src_off, int64_t len) {
  while (len > 0) {
    const int so = int(src_off & 31), dof = int(dst_off & 31);
    const int take = int(std::min<int64_t>(len, std::min(32 - so, 32 - dof)));
    const uint32_t chunk = (src[src_off >> 5] << so) >> (32 - take);  // top-aligned -> low `take` bits
    dst[dst_off >> 5] |= chunk << (32 - dof - take);
    src_off += take;
    dst_off += take;
    len -= take;
  }
}

inline uint32_t bit_reverse32(uint32_t v) {
  v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
  v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
  v = ((v >> 4) & 0x0f0f0f0fu) | ((v & 0x0f0f0f0fu) << 4);
  return __builtin_bswap32(v);
}

int64_t count_ones(const uint32_t* bits, int64_t from, int64_t to) {  // ones in [from, to)
  int64_t c = 0;
  while (from < to) {
    const int o = int(from & 31);
    const int take = int(std::min<int64_t>(to - from, 32 - o));
    const uint32_t chunk = (bits[from >> 5] << o) >> (32 - take);
    c += __builtin_popcount(chunk);
    from += take;
  }
  return c;
}

// Builds the blocks and SuperRecs of one bucket's wavelet tree; returns the block cursor after them.
int64_t fill_wavelet_paired(const Blob& blk, const BucketPlan& p, HostImage* im, int64_t local_bucket,
                            const std::unordered_map<uint32_t, uint32_t>& leaf_sym) {
  const BucketTables& t = p.tab;
  const int bw = im->block_words;
  const int64_t B = paired_positions(bw);
  const int slices = bw / kPairedSliceWords;
  const int stretch_words = 3 * slices;  // words of X payload per block; the children region has as many
  uint32_t xs[12], rs[12];
  const uint8_t* wt = blk.at(t.off_wtree, 4);
  const size_t wt_avail = size_t(t.off_marktab) - size_t(t.off_wtree);
  const size_t n_int = p.node_ids.size();
  auto index_of = [&](uint32_t id) -> int {
    auto it = std::lower_bound(p.node_ids.begin(), p.node_ids.end(), id);
    return (it != p.node_ids.end() && *it == id) ? int(it - p.node_ids.begin()) : -1;
  };
  auto expand = [&](int k, std::vector<uint32_t>& out) {
    const int64_t nbits = p.node_bits[size_t(k)];
    out.assign(size_t((nbits + 31) / 32) + 2, 0u);
    if (nbits > 0) {
      const int64_t got = bseq_expand(open_bseq(wt + p.node_offs[size_t(k)], wt_avail - p.node_offs[size_t(k)]), out.data(), nbits);
      if (got != nbits) throw Error(FM_ERR_FORMAT, "bseq expansion length mismatch");
    }
  };
  // block base of every super node
  std::vector<int64_t> base(n_int, 0);
  int64_t cursor = p.block_base;
  for (size_t k = 0; k < n_int; k++) {
    if (p.super_of[k] < 0) continue;
    base[k] = cursor;
    cursor += paired_blocks_for_bits(p.node_bits[k], bw);
  }
  std::vector<uint32_t> X, C0, C1;
  for (size_t k = 0; k < n_int; k++) {
    if (p.super_of[k] < 0) continue;
    const uint32_t id = p.node_ids[k];
    const int c0 = index_of(2 * id), c1 = index_of(2 * id + 1);
    const int64_t n = p.node_bits[k];
    expand(int(k), X);
    const int64_t onesX = count_ones(X.data(), 0, n);
    if (c0 >= 0) { expand(c0, C0); if (p.node_bits[size_t(c0)] != n - onesX) throw Error(FM_ERR_FORMAT, "wavelet child 0 length mismatch"); }
    if (c1 >= 0) { expand(c1, C1); if (p.node_bits[size_t(c1)] != onesX) throw Error(FM_ERR_FORMAT, "wavelet child 1 length mismatch"); }
    // blocks
    const int64_t nb = paired_blocks_for_bits(n, bw);
    int64_t ones_before = 0, c0_ones_before = 0, c1_ones_before = 0;
    for (int64_t j = 0; j < nb; j++) {
      uint32_t* w = im->rank_words + (base[k] + j) * bw;
      const int64_t p0 = j * B, p1 = std::min<int64_t>(n, p0 + B), len = p1 - p0;
      const int64_t ones = count_ones(X.data(), p0, p1), z = len - ones;
      const int64_t zeros_before = p0 - ones_before;
      std::fill(xs, xs + stretch_words, 0u);
      std::fill(rs, rs + stretch_words, 0u);
      copy_bits(xs, 0, X.data(), p0, len);
      int64_t c0_ones = 0, c1_ones = 0;
      if (c0 >= 0) {  // child 0: the stretch's zeros, in order, from the front of the region
        copy_bits(rs, 0, C0.data(), zeros_before, z);
        c0_ones = count_ones(C0.data(), zeros_before, zeros_before + z);
      }
      if (c1 >= 0) {  // child 1: the stretch's ones, reversed, from the back
        for (int64_t i = 0; i < ones; i++) {
          const int64_t src = ones_before + i;
          if ((C1[size_t(src >> 5)] >> (31 - (src & 31))) & 1u) {
            const int64_t dst = B - 1 - i;
            rs[dst >> 5] |= 1u << (31 - (dst & 31));
            c1_ones++;
          }
        }
      }
      for (int s = 0; s < slices; s++) {
        uint32_t* sw = w + s * kPairedSliceWords;
        sw[0] = uint32_t(ones_before);
        sw[1] = (s & 1) ? uint32_t(c1_ones_before + c0_ones + c1_ones) : uint32_t(c0_ones_before);
        for (int t = 0; t < 3; t++) {
          sw[2 + t] = xs[3 * s + t];
          sw[5 + t] = rs[3 * s + t];
        }
      }
      c0_ones_before += c0_ones;
      c1_ones_before += c1_ones;
      ones_before += ones;
    }
    // record: children and grandchildren
    SuperRec& sr = im->supers[size_t(p.node_base) + size_t(p.super_of[k])];
    for (uint32_t b1 = 0; b1 < 2; b1++) {
      const uint32_t child = 2 * id + b1;
      const int ci = b1 ? c1 : c0;
      if (ci < 0) {
        auto ls = leaf_sym.find(child);
        sr.child_info[b1] = kChildLeaf | (ls == leaf_sym.end() ? kEndOfBucketSym : ls->second);
      } else {
        sr.child_info[b1] = 0;
      }
      for (uint32_t b2 = 0; b2 < 2; b2++) {
        const uint32_t gc = 2 * child + b2;
        const uint32_t slot = b1 * 2 + b2;
        sr.gc[slot][0] = 0;
        sr.gc[slot][1] = kChildLeaf | kEndOfBucketSym;
        if (ci < 0) continue;  // the child is a leaf: no grandchildren
        const int gi = index_of(gc);
        if (gi >= 0) {
          sr.gc[slot][0] = uint32_t(base[size_t(gi)]);
          sr.gc[slot][1] = uint32_t(p.node_base + p.super_of[size_t(gi)]);
        } else {
          auto ls = leaf_sym.find(gc);
          sr.gc[slot][1] = kChildLeaf | (ls == leaf_sym.end() ? kEndOfBucketSym : ls->second);
        }
      }
    }
    sr.pad[0] = sr.pad[1] = 0;
  }
  BucketRec& br = im->buckets[size_t(local_bucket)];
  br.root_base = uint32_t(base[0]);
  br.root_node = uint32_t(p.node_base + p.super_of[0]);
  return cursor;
}

// ---- quad-level layout (fm_image.hpp) ---------------------------------------------------------
// Builds the blocks and QuadRecs of one bucket's wavelet tree; returns the block cursor after them.
int64_t fill_wavelet_quad(const Blob& blk, const BucketPlan& p, HostImage* im, int64_t local_bucket,
                          const std::unordered_map<uint32_t, uint32_t>& leaf_sym) {
  const BucketTables& t = p.tab;
  constexpr int P = kQuadPos, BW = kQuadBlockWords;
  const uint8_t* wt = blk.at(t.off_wtree, 4);
  const size_t wt_avail = size_t(t.off_marktab) - size_t(t.off_wtree);
  const size_t n_int = p.node_ids.size();
  auto index_of = [&](uint64_t id) -> int {
    if (id > 0xffffffffull) return -1;
    auto it = std::lower_bound(p.node_ids.begin(), p.node_ids.end(), uint32_t(id));
    return (it != p.node_ids.end() && *it == uint32_t(id)) ? int(it - p.node_ids.begin()) : -1;
  };
  auto expand = [&](int k, std::vector<uint32_t>& out) {
    const int64_t nbits = p.node_bits[size_t(k)];
    out.assign(size_t((nbits + 31) / 32) + 2, 0u);
    if (nbits > 0) {
      const int64_t got = bseq_expand(open_bseq(wt + p.node_offs[size_t(k)], wt_avail - p.node_offs[size_t(k)]), out.data(), nbits);
      if (got != nbits) throw Error(FM_ERR_FORMAT, "bseq expansion length mismatch");
    }
  };
  std::vector<int64_t> base(n_int, 0);
  int64_t cursor = p.block_base;
  for (size_t k = 0; k < n_int; k++) {
    if (p.super_of[k] < 0) continue;
    if (k == 0) {  // root: block (row in bucket) / kQuadPos of the bucket's slot in the root area
      base[k] = local_bucket * im->root_stride;
      continue;
    }
    base[k] = cursor;
    cursor += std::max<int64_t>(1, (p.node_bits[k] + P - 1) / P);
  }
  std::vector<uint32_t> seq[16];
  for (size_t k = 0; k < n_int; k++) {
    if (p.super_of[k] < 0) continue;
    const uint64_t id = p.node_ids[k];
    const int64_t n = p.node_bits[k];
    // block-local members: v = 2^l + r  <->  tree node id * 2^l + r
    bool internal[16] = {false};
    int64_t len[16] = {0};
    for (int v = 1; v < 16; v++) {
      const int l = 31 - __builtin_clz(unsigned(v));
      const int mi = (v == 1) ? int(k) : (internal[v >> 1] ? index_of((id << l) + uint64_t(v - (1 << l))) : -1);
      internal[v] = mi >= 0;
      if (internal[v]) {
        expand(mi, seq[v]);
        len[v] = p.node_bits[size_t(mi)];
      }
    }
    int64_t pos_before[32] = {0};  // positions routed to member / exit v by earlier blocks
    const int64_t nb = std::max<int64_t>(1, (n + P - 1) / P);
    for (int64_t j = 0; j < nb; j++) {
      uint32_t* w = im->rank_words + (base[k] + j) * BW;
      uint32_t region[4][4] = {{0}};
      int cnt[32] = {0}, real[32] = {0}, lo[32] = {0}, hi[32] = {0};
      cnt[1] = P;
      real[1] = int(std::max<int64_t>(0, std::min<int64_t>(P, n - j * P)));
      lo[1] = 0;
      hi[1] = P;
      for (int v = 1; v < 16; v++) {
        const int l = 31 - __builtin_clz(unsigned(v));
        const bool forward = v == 1 || (v & 1) == 0;
        int ones = 0;
        if (internal[v] && real[v] > 0) {
          if (pos_before[v] + real[v] > len[v]) throw Error(FM_ERR_FORMAT, "wavelet tree child shorter than its parent says");
          uint32_t y[5] = {0, 0, 0, 0, 0};  // the node's bits for this block's positions, in order
          copy_bits(y, 0, seq[v].data(), pos_before[v], real[v]);
          for (int t = 0; t < 4; t++) ones += __builtin_popcount(y[t]);
          if (forward) {
            copy_bits(region[l], lo[v], y, 0, real[v]);
          } else {  // entry i goes to bit hi-1-i: place the reversed string so that it ends at hi
            uint32_t r[5] = {0, 0, 0, 0, 0};
            for (int t = 0; t < 4; t++) r[t] = bit_reverse32(y[3 - t]);
            copy_bits(region[l], hi[v] - real[v], r, kQuadPos - real[v], real[v]);
          }
        }
        const int c0 = 2 * v, c1 = 2 * v + 1;
        cnt[c1] = real[c1] = ones;
        cnt[c0] = cnt[v] - ones;   // positions past the end of the sequence travel down the 0 side
        real[c0] = real[v] - ones;
        lo[c0] = lo[v];
        hi[c0] = lo[v] + cnt[c0];
        hi[c1] = hi[v];
        lo[c1] = hi[v] - cnt[c1];
      }
      for (int q = 0; q < 16; q++) {
        const int64_t e = pos_before[16 + q];
        if (e >= (int64_t(1) << 24)) throw Error(FM_ERR_FULL, "bucket too large for the quad-level layout");
        const int path3 = q >> 1, v2 = 4 + (path3 >> 1), v3 = 8 + path3;
        const int anchor = (q & 1) ? ((v3 & 1) ? hi[v3] : lo[v3]) : ((v2 & 1) ? hi[v2] : lo[v2]);
        w[q] = uint32_t(e) | (uint32_t(anchor) << 24);
      }
      for (int l = 0; l < 4; l++)
        for (int wi = 0; wi < 4; wi++) w[16 + 8 * (wi >> 1) + 2 * l + (wi & 1)] = region[l][wi];
      for (int v = 1; v < 32; v++) pos_before[v] += real[v];
    }
    for (int v = 1; v < 16; v++)
      if (internal[v] && pos_before[v] != len[v]) throw Error(FM_ERR_FORMAT, "wavelet tree child length mismatch");
    // record: the 16 exits
    QuadRec& qr = im->quads[size_t(p.node_base) + size_t(p.super_of[k])];
    for (int q = 0; q < 16; q++) {
      uint64_t cur = id;
      qr.exit[q][0] = 0;
      qr.exit[q][1] = kChildLeaf | kEndOfBucketSym;
      bool done = false;
      for (int d = 0; d < 4 && !done; d++) {
        const uint64_t child = 2 * cur + uint64_t((q >> (3 - d)) & 1);
        if (index_of(child) >= 0) {
          cur = child;
          continue;
        }
        done = true;  // a leaf: its exit is its code extended with 0 bits
        auto ls = child <= 0xffffffffull ? leaf_sym.find(uint32_t(child)) : leaf_sym.end();
        const bool zero_tail = (q & ((1 << (3 - d)) - 1)) == 0;
        if (ls != leaf_sym.end() && zero_tail) qr.exit[q][1] = kChildLeaf | ls->second;
      }
      if (!done) {
        const int gi = index_of(cur);
        qr.exit[q][0] = uint32_t(base[size_t(gi)]);
        qr.exit[q][1] = uint32_t(p.node_base + p.super_of[size_t(gi)]);
      }
    }
  }
  BucketRec& br = im->buckets[size_t(local_bucket)];
  br.root_base = uint32_t(base[0]);
  br.root_node = uint32_t(p.node_base + p.super_of[0]);
  return cursor;
}

void fill_bucket(const IndexFiles& files, const Blob& blk, const BlockHeader& bh, int64_t blk_num, int bucket,
                 const BucketPlan& p, HostImage* im, int64_t local_bucket, Scratch* scratch,
                 std::atomic<int64_t>* markval_used) {
  const int kBlockWords = im->block_words;
  t_block_words = kBlockWords;
  const BucketTables& t = p.tab;
  const int bpb = files.buckets_per_block();
  const uint8_t* wt = blk.at(t.off_wtree, 4);
  const size_t wt_avail = size_t(t.off_marktab) - size_t(t.off_wtree);
  const size_t n_int = p.node_ids.size();

  if (p.chunk_len > 0)  // chunk directory + chunks, as stored
    std::memcpy(im->chunk_bytes.data() + p.chunk_base, blk.at(p.chunk_dir, size_t(p.chunk_len)), size_t(p.chunk_len));

  // leaf id -> symbol
  std::unordered_map<uint32_t, uint32_t> leaf_sym;
  leaf_sym.reserve(size_t(t.n_in_use) * 2 + 2);
  for (int s = 0; s <= t.n_in_use; s++) leaf_sym[t.leaf[s]] = t.seq_to_ch[s];

  // rank blocks of every internal node, in directory order
  std::vector<int64_t> node_block(n_int);
  int64_t cursor = p.block_base;
  if (im->levels == 2) cursor = fill_wavelet_paired(blk, p, im, local_bucket, leaf_sym);
  if (im->levels == 4) cursor = fill_wavelet_quad(blk, p, im, local_bucket, leaf_sym);
  for (size_t k = 0; k < n_int && im->levels == 1; k++) {
    node_block[k] = cursor;
    const int64_t nbits = p.node_bits[k];
    if (nbits > 0) {
      uint32_t* bits = scratch->get(nbits);
      const int64_t got = bseq_expand(open_bseq(wt + p.node_offs[k], wt_avail - p.node_offs[k]), bits, nbits);
      if (got != nbits) throw Error(FM_ERR_FORMAT, "bseq expansion length mismatch");
      fill_blocks(bits, nbits, im->rank_words + cursor * kBlockWords);
    }
    cursor += blocks_for_bits(nbits);
  }

  // node records
  for (size_t k = 0; k < n_int && im->levels == 1; k++) {
    NodeRec& nr = im->nodes[size_t(p.node_base) + k];
    for (uint32_t b = 0; b < 2; b++) {
      const uint32_t child = p.node_ids[k] * 2 + b;
      auto it = std::lower_bound(p.node_ids.begin(), p.node_ids.end(), child);
      if (it != p.node_ids.end() && *it == child) {
        const size_t ci = size_t(it - p.node_ids.begin());
        nr.child_base[b] = uint32_t(node_block[ci]);
        nr.child_info[b] = uint32_t(p.node_base + int64_t(ci));
      } else {
        auto ls = leaf_sym.find(child);
        nr.child_base[b] = 0;
        nr.child_info[b] = kChildLeaf | (ls == leaf_sym.end() ? kEndOfBucketSym : ls->second);
      }
    }
  }

  BucketRec& br = im->buckets[size_t(local_bucket)];
  if (im->levels == 1) {
    br.root_base = uint32_t(node_block[0]);
    br.root_node = uint32_t(p.node_base);
  }
  br.markval_base = uint64_t(p.markval_base);

  // per-symbol records
  const size_t rec0 = size_t(local_bucket) * kAlphaStride;
  for (int ch = 0; ch < kAlpha; ch++) {
    OccRec& o = im->occ[rec0 + size_t(ch)];
    o.occ_base = files.C(ch) + files.block_occs(ch, blk_num) + int64_t(bucket_occs(blk, bh, bpb, ch, bucket));
    o.leaf = t.in_use[ch] ? t.leaf[t.ch_to_seq[ch]] : 0u;
    o.root_exit = 0;
    if (im->levels == 4 && o.leaf) {  // index of the root QuadRec's exit entry on this symbol's path
      const int len = 31 - __builtin_clz(o.leaf);
      const uint32_t nib = len >= 4 ? (o.leaf >> (len - 4)) & 15u : (o.leaf << (4 - len)) & 15u;
      const uint32_t root = uint32_t(p.node_base + p.super_of[0]);
      o.root_exit = root * 16u + nib;
      // codes of 5..8 bits need exactly one more block, and of its exit entry only the first block:
      // store that directly (kRootExitDirect) and the step saves the read of the entry
      const uint32_t base2 = im->quads[root].exit[nib][0];
      if (len > 4 && len <= 8 && base2 < kRootExitDirect) o.root_exit = kRootExitDirect | base2;
    }
    MarkRec& m = im->mark[rec0 + size_t(ch)];
    m.mark_base = 0;
    m.markval_off = 0;
  }

  // mark bit-vectors and sampled SA values
  const size_t mt_avail = size_t(t.off_markarr) - size_t(t.off_marktab);
  const uint8_t* mt = blk.at(t.off_marktab, mt_avail);
  const size_t ma_end = t.off_end > t.off_markarr ? t.off_end : blk.size();
  const size_t ma_avail = ma_end - size_t(t.off_markarr);
  const uint8_t* ma = blk.at(t.off_markarr, ma_avail);
  const int vbits = num_bits64(uint64_t(bh.total_length));  // text_size_bits, index.c:1445
  int64_t val_cursor = 0;
  for (int s = 0; s < t.n_in_use; s++) {
    const int ch = t.seq_to_ch[s];
    const int64_t nbits = p.mark_bits[size_t(s)];
    uint32_t* bits = scratch->get(nbits);
    const uint32_t off = p.mark_offs[size_t(s)];
    const int64_t got = bseq_expand(open_bseq(mt + off, mt_avail - off), bits, nbits);
    if (got != nbits) throw Error(FM_ERR_FORMAT, "mark table expansion length mismatch");
    const int64_t ones = fill_blocks(bits, nbits, im->rank_words + cursor * kBlockWords);
    MarkRec& m = im->mark[rec0 + size_t(ch)];
    m.mark_base = uint32_t(cursor);
    m.markval_off = uint32_t(val_cursor);
    cursor += blocks_for_bits(nbits);
    // values: `ones` fields of vbits bits, MSB-first (bsW64, buffer_funcs.h:118-127)
    const uint32_t aoff = p.markarr_offs[size_t(s)];
    const size_t need_bytes = size_t((ones * vbits + 7) / 8);
    if (size_t(aoff) + need_bytes > ma_avail) throw Error(FM_ERR_FORMAT, "mark array truncated");
    const uint8_t* src = ma + aoff;
    int64_t* dst = im->markvals.data() + p.markval_base + val_cursor;
    size_t bitpos = 0;
    for (int64_t j = 0; j < ones; j++) {
      uint64_t v = 0;
      int left = vbits;
      while (left > 0) {
        const int o = int(bitpos & 7);
        const int take = std::min(8 - o, left);
        const uint32_t byte = src[bitpos >> 3];
        v = (v << take) | ((byte >> (8 - o - take)) & ((1u << take) - 1u));
        bitpos += size_t(take);
        left -= take;
      }
      dst[j] = int64_t(v);
    }
    val_cursor += ones;
  }
  if (cursor != p.block_base + p.n_blocks) throw Error(FM_ERR_INVALID, "rank block accounting error");
  if (val_cursor != p.n_markvals) throw Error(FM_ERR_FORMAT, "mark table ones do not match their S sums");
  markval_used->fetch_add(val_cursor);
}

template <typename F>
void parallel_for(int64_t n, int nthreads, F&& fn) {
  std::atomic<int64_t> next{0};
  std::mutex err_mu;
  std::unique_ptr<Error> first_err;
  auto worker = [&](int tid) {
    try {
      for (;;) {
        const int64_t i = next.fetch_add(1);
        if (i >= n) break;
        fn(i, tid);
      }
    } catch (const Error& e) {
      std::lock_guard<std::mutex> g(err_mu);
      if (!first_err) first_err.reset(new Error(e));
      next.store(n);
    } catch (const std::exception& e) {
      std::lock_guard<std::mutex> g(err_mu);
      if (!first_err) first_err.reset(new Error(FM_ERR_UNKNOWN, e.what()));
      next.store(n);
    }
  };
  nthreads = int(std::max<int64_t>(1, std::min<int64_t>(nthreads, n)));
  std::vector<std::thread> th;
  for (int t = 1; t < nthreads; t++) th.emplace_back(worker, t);
  worker(0);
  for (auto& t : th) t.join();
  if (first_err) throw *first_err;
}

}  // namespace

namespace {
std::atomic<int> g_default_block_words{0};
}

namespace {
std::atomic<int> g_default_paired{-1};  // levels per block chosen by set_default_levels_per_block, -1 = none
// quad-level blocks are the fastest layout measured on B200 (profiles/r01_levels_per_block.md)
constexpr int kDefaultLevelsPerBlock = 4;
}

// Block size when none was chosen: 64 bytes for the paired layout, 128 bytes for one level per
// block (profiles/r01_paired_level_sweep.md).
int default_block_words(int levels) {
  if (levels == 4) return kQuadBlockWords;  // the quad layout is defined for 128-byte blocks only
  const int v = g_default_block_words.load();
  if (v != 0) return v;
  if (const char* e = std::getenv("FEMTO_B200_BLOCK_BYTES")) {
    const int b = std::atoi(e);
    if (b == 32 || b == 64 || b == 128) return b / 4;
  }
  return levels == 2 ? 16 : kDefaultBlockWords;
}

int default_levels_per_block() {
  int v = g_default_paired.load();
  if (v <= 0) {
    v = kDefaultLevelsPerBlock;
    if (const char* e = std::getenv("FEMTO_B200_LEVELS_PER_BLOCK")) v = std::atoi(e);
  }
  return (v == 1 || v == 2 || v == 4) ? v : kDefaultLevelsPerBlock;
}

bool set_default_levels_per_block(int levels) {
  if (levels != 0 && levels != 1 && levels != 2 && levels != 4) return false;
  g_default_paired.store(levels == 0 ? -1 : levels);
  return true;
}

bool set_default_block_words(int words) {
  if (words != 0 && words != 8 && words != 16 && words != 32) return false;
  g_default_block_words.store(words);
  return true;
}

HostRank host_rank(const uint32_t* rank_words, int block_words, uint32_t base_block, uint32_t index1) {
  const uint32_t p = index1 - 1;
  const uint32_t bits = uint32_t(block_words - 1) * 32;
  const uint32_t k = p / bits, off = p % bits;
  const uint32_t* w = rank_words + (size_t(base_block) + k) * size_t(block_words);
  uint32_t ones = w[0];
  const uint32_t full = off / 32, rem = off % 32;
  for (uint32_t j = 0; j < full; j++) ones += uint32_t(__builtin_popcount(w[1 + j]));
  const uint32_t last = w[1 + full];
  ones += uint32_t(__builtin_popcount(last >> (31 - rem)));
  return HostRank{ones, (last >> (31 - rem)) & 1u};
}

HostPairedRank host_paired_rank(const uint32_t* rank_words, int block_words, uint32_t base_block, uint32_t index1,
                                int follow) {
  const uint32_t B = uint32_t(paired_positions(block_words));
  const uint32_t p = index1 - 1, k = p / B, off = p % B;
  const uint32_t* w = rank_words + (size_t(base_block) + k) * size_t(block_words);
  auto xbit = [&](uint32_t q) { return (w[8 * (q / 96) + 2 + (q % 96) / 32] >> (31 - (q & 31))) & 1u; };
  auto rbit = [&](uint32_t q) { return (w[8 * (q / 96) + 5 + (q % 96) / 32] >> (31 - (q & 31))) & 1u; };
  HostPairedRank r{};
  uint32_t cnt = 0;
  for (uint32_t q = 0; q <= off; q++) cnt += xbit(q);
  r.bit1 = xbit(off);
  const uint32_t b1 = follow < 0 ? r.bit1 : uint32_t(follow);
  const uint32_t ones1 = w[0] + cnt;
  r.index1 = b1 ? ones1 : index1 - ones1;
  const uint32_t j = b1 ? cnt : off + 1 - cnt;  // of the child's first index1 bits, those stored in this block
  const uint32_t hi = b1 ? B - j : j;
  uint32_t rc = 0;
  for (uint32_t q = 0; q < hi; q++) rc += rbit(q);
  const uint32_t h1 = w[8 * b1 + 1];
  r.ones2 = b1 ? h1 - rc : h1 + rc;
  r.bit2 = j == 0 ? 0u : rbit(b1 ? B - j : j - 1);
  return r;
}

HostQuadRank host_quad_rank(const uint32_t* rank_words, uint32_t base_block, uint32_t index1, int path) {
  const uint32_t p = index1 - 1, k = p / kQuadPos;
  uint32_t j = p % kQuadPos + 1;
  const uint32_t* w = rank_words + (size_t(base_block) + k) * size_t(kQuadBlockWords);
  auto rbit = [&](int l, uint32_t pos) {
    const uint32_t wi = pos >> 5;
    return (w[16 + 8 * (wi >> 1) + 2 * uint32_t(l) + (wi & 1)] >> (31 - (pos & 31))) & 1u;
  };
  uint32_t q = 0;
  for (int l = 0; l < 4; l++) {
    const bool back = l > 0 && (q & 1u);
    uint32_t anchor = 0;
    if (l == 1) anchor = back ? kQuadPos : 0;
    if (l == 2) anchor = w[q << 2] >> 24;
    if (l == 3) anchor = w[(q << 1) | 1u] >> 24;
    const uint32_t a = back ? anchor - j : anchor;
    uint32_t cnt = 0;
    for (uint32_t x = a; x < a + j; x++) cnt += rbit(l, x);
    const uint32_t bit = path >= 0 ? (uint32_t(path) >> (3 - l)) & 1u : (j ? rbit(l, back ? anchor - j : anchor + j - 1) : 0u);
    j = bit ? cnt : j - cnt;
    q = (q << 1) | bit;
  }
  return HostQuadRank{q, (w[q] & 0xffffffu) + j};
}

std::unique_ptr<HostImage> build_host_image(const std::string& path, int shard, int nshards, int nthreads,
                                            int block_words, int levels) {
  if (nshards < 1 || shard < 0 || shard >= nshards) throw Error(FM_ERR_PARAM, "bad shard");
  if (levels == 0) levels = default_levels_per_block();
  if (levels != 1 && levels != 2 && levels != 4) throw Error(FM_ERR_PARAM, "levels per block must be 1, 2 or 4");
  if (block_words == 0) block_words = default_block_words(levels);
  if (block_words != 8 && block_words != 16 && block_words != 32) throw Error(FM_ERR_PARAM, "rank block must be 32, 64 or 128 bytes");
  if (levels == 2 && block_words < 16) throw Error(FM_ERR_PARAM, "the paired-level layout needs 64- or 128-byte blocks");
  if (levels == 4 && block_words != kQuadBlockWords) throw Error(FM_ERR_PARAM, "the quad-level layout needs 128-byte blocks");
  if (nthreads <= 0) nthreads = int(std::max(1u, std::thread::hardware_concurrency()));
  auto files = IndexFiles::open(path);
  const BlockHeader& h = files->header();
  const int bpb = files->buckets_per_block();
  std::unique_ptr<HostImage> im(new HostImage());
  im->hdr = h;
  im->block_words = block_words;
  im->levels = levels;
  const int kBlockWords = block_words;

  // data blocks of this shard (shard_of_block, fm_format.hpp: contiguous, balanced by rows)
  int64_t b0 = h.nblocks, b1 = 0;
  for (int64_t b = 0; b < h.nblocks; b++) {
    if (shard_of_block(b, h.block_size, h.total_length, nshards) == shard) { b0 = std::min(b0, b); b1 = std::max(b1, b + 1); }
  }
  if (b0 >= b1) { b0 = b1 = 0; }
  im->first_block = b0;
  im->end_block = b1;
  im->first_row = b0 * int64_t(h.block_size);
  im->end_row = std::min<int64_t>(h.total_length, b1 * int64_t(h.block_size));
  if (b0 == b1) im->end_row = im->first_row;
  im->first_bucket = b0 * bpb;

  im->C.resize(262);
  for (int c = 0; c < 262; c++) im->C[size_t(c)] = files->C(c);
  im->doc_ends.resize(size_t(h.ndocs));
  im->doc_eof_rows.resize(size_t(h.ndocs));
  im->doc_info_off.assign(size_t(h.ndocs) + 1, 0);
  for (int64_t d = 0; d < h.ndocs; d++) {
    im->doc_ends[size_t(d)] = files->doc_end(d);
    im->doc_eof_rows[size_t(d)] = files->doc_eof_row(d);
    const auto info = files->doc_info(d);
    if (info.second) im->doc_info_bytes.insert(im->doc_info_bytes.end(), info.first, info.first + info.second);
    im->doc_info_off[size_t(d) + 1] = int64_t(im->doc_info_bytes.size());
  }

  // map blocks, validate headers
  std::vector<Blob> blobs;
  std::vector<BlockHeader> bhs;
  std::vector<int64_t> bucket0;  // local index of each block's first bucket
  int64_t nb = 0;
  for (int64_t b = b0; b < b1; b++) {
    blobs.push_back(files->map_block(b));
    BlockHeader bh = parse_block_header(blobs.back(), kMagicDataBlock);
    if (bh.block_number != b || bh.block_size != h.block_size || bh.bucket_size != h.bucket_size ||
        bh.total_length != h.total_length)
      throw Error(FM_ERR_FORMAT, "data block header does not match the header block");
    const int64_t expect_rows = std::min<int64_t>(h.block_size, h.total_length - b * int64_t(h.block_size));
    if (bh.size != expect_rows) throw Error(FM_ERR_FORMAT, "data block has an unexpected number of rows");
    if (bh.num_buckets != (bh.size + h.bucket_size - 1) / h.bucket_size)
      throw Error(FM_ERR_FORMAT, "data block has an unexpected number of buckets");
    if (b + 1 < h.nblocks && bh.num_buckets != bpb) throw Error(FM_ERR_FORMAT, "short block in the middle of the index");
    blobs.back().at(0, size_t(kBlockHeaderBytes) + 4 * (size_t(bpb) + 1) + 4 * size_t(kAlpha) * size_t(bh.num_buckets));
    bhs.push_back(bh);
    bucket0.push_back(nb);
    nb += bh.num_buckets;
  }
  im->nbuckets = nb;

  // pass 1: plan every bucket
  std::vector<BucketPlan> plans{size_t(nb)};
  std::vector<std::pair<int32_t, int32_t>> where{size_t(nb)};  // (local block, bucket in block)
  for (size_t lb = 0; lb < blobs.size(); lb++)
    for (int k = 0; k < bhs[lb].num_buckets; k++) where[size_t(bucket0[lb] + k)] = {int32_t(lb), int32_t(k)};
  parallel_for(nb, nthreads, [&](int64_t g, int) {
    const auto [lb, k] = where[size_t(g)];
    plan_bucket(blobs[size_t(lb)], bhs[size_t(lb)], bpb, k, &plans[size_t(g)], block_words, levels);
  });

  int64_t nodes = 0, blocks = 0, vals = 0, wt_blocks = 0;
  int max_len = 0;
  if (levels == 4) {  // root area: bucket g's root blocks start at g * root_stride
    im->root_stride = (int64_t(h.bucket_size) + kQuadPos - 1) / kQuadPos;
    blocks = nb * im->root_stride;
    wt_blocks = blocks;
  }
  for (auto& p : plans) {
    if (p.n_root_blocks > im->root_stride) throw Error(FM_ERR_FORMAT, "bucket longer than bucket_size");
    p.node_base = nodes;
    p.block_base = blocks;
    p.markval_base = vals;
    nodes += p.n_records;
    blocks += p.n_blocks;
    wt_blocks += p.n_wtree_blocks;
    vals += p.n_markvals;
    max_len = std::max(max_len, p.tab.max_len);
  }
  {  // document chunks: one byte range per bucket
    im->chunk_off.assign(size_t(nb) + 1, 0);
    im->chunk_count.assign(size_t(nb), 0);
    im->chunk_dir_rel.assign(size_t(nb), 24);
    int64_t at = 0;
    for (int64_t g = 0; g < nb; g++) {
      plans[size_t(g)].chunk_base = at;
      im->chunk_off[size_t(g)] = at;
      im->chunk_count[size_t(g)] = plans[size_t(g)].n_chunks;
      at += plans[size_t(g)].chunk_len;
    }
    im->chunk_off[size_t(nb)] = at;
    im->chunk_bytes.resize(size_t(at));
  }
  if (blocks >= (int64_t(1) << 32) || nodes >= (int64_t(1) << (levels == 4 ? 27 : 31)))
    throw Error(FM_ERR_FULL, "shard too large for 32-bit rank block indices; use more shards");
  im->max_code_len = max_len;
  im->n_rank_blocks = blocks;
  im->n_wtree_blocks = wt_blocks;
  im->rank_words = static_cast<uint32_t*>(std::calloc(size_t(std::max<int64_t>(blocks, 1)) * kBlockWords, 4));
  if (!im->rank_words) throw Error(FM_ERR_MEM, "out of host memory for the rank image");
  if (levels == 4) im->quads.resize(size_t(nodes));
  else if (levels == 2) im->supers.resize(size_t(nodes));
  else im->nodes.resize(size_t(nodes));
  im->occ.resize(size_t(nb) * kAlphaStride);
  im->mark.resize(size_t(nb) * kAlphaStride);
  im->buckets.resize(size_t(nb));
  im->markvals.resize(size_t(vals));

  // pass 2: decode
  std::atomic<int64_t> used{0};
  std::vector<Scratch> scratch{size_t(std::max(1, nthreads))};
  parallel_for(nb, nthreads, [&](int64_t g, int tid) {
    const auto [lb, k] = where[size_t(g)];
    fill_bucket(*files, blobs[size_t(lb)], bhs[size_t(lb)], b0 + lb, k, plans[size_t(g)], im.get(), g,
                &scratch[size_t(tid)], &used);
  });
  return im;
}

void chunk_documents(const BlockHeader& hdr, int64_t first_row, int64_t end_row, int64_t first_bucket,
                     const std::vector<uint8_t>& chunk_bytes, const std::vector<int64_t>& chunk_off,
                     const std::vector<int32_t>& chunk_count, const std::vector<uint32_t>& chunk_dir_rel,
                     int64_t row, int64_t* first, int64_t* last, std::vector<int64_t>* docs) {
  docs->clear();
  if (hdr.chunk_size <= 0) throw Error(FM_ERR_MISSING, "index was built without document chunks");
  if (row < first_row || row >= end_row) throw Error(FM_ERR_PARAM, "row not resident");
  // chunk number inside the data block, then bucket and chunk in bucket (index.c:2156-2160, 2205-2212)
  const int64_t blk = row / hdr.block_size, row0 = blk * int64_t(hdr.block_size);
  const int64_t blk_rows = std::min<int64_t>(hdr.block_size, hdr.total_length - row0);
  const int64_t cn = (row - row0) / hdr.chunk_size;
  const int64_t per_bucket = hdr.bucket_size / hdr.chunk_size;
  const int64_t bpb = (int64_t(hdr.block_size) + hdr.bucket_size - 1) / hdr.bucket_size;
  const int64_t g = blk * bpb + cn / per_bucket - first_bucket;
  const int64_t c = cn % per_bucket;
  *first = row0 + cn * hdr.chunk_size;
  *last = std::min(*first + hdr.chunk_size - 1, row0 + blk_rows - 1);
  if (g < 0 || g >= int64_t(chunk_count.size()) || c >= chunk_count[size_t(g)])
    throw Error(FM_ERR_FORMAT, "chunk missing from its bucket");
  const uint8_t* sec = chunk_bytes.data() + chunk_off[size_t(g)];
  const int64_t sec_len = chunk_off[size_t(g) + 1] - chunk_off[size_t(g)];
  const uint32_t rel = chunk_dir_rel[size_t(g)];  // directory entries are relative to the bucket start
  const int64_t lo = int64_t(be32(sec + 4 * c)) - rel, hi = int64_t(be32(sec + 4 * (c + 1))) - rel;
  if (lo < 0 || hi < lo || hi > sec_len) throw Error(FM_ERR_FORMAT, "bad chunk bounds");
  // number of documents: chunk_num_docs_bits = num_bits64(chunk_size) bits, MSB-first, then flush
  const int nbits = num_bits64(uint64_t(hdr.chunk_size));
  size_t bit = size_t(lo) * 8;
  const size_t end_bit = size_t(hi) * 8;
  auto get = [&](int n) -> uint64_t {
    uint64_t v = 0;
    for (int i = 0; i < n; i++, bit++) {
      if (bit >= end_bit) throw Error(FM_ERR_FORMAT, "chunk truncated");
      v = (v << 1) | ((sec[bit >> 3] >> (7 - (bit & 7))) & 1u);
    }
    return v;
  };
  const uint64_t ndocs = get(nbits);
  bit = (bit + 7) & ~size_t(7);
  // gamma-coded deltas of (document + 1) (results.c:133-152, 356-371)
  int64_t lastdoc = 0;
  docs->reserve(size_t(ndocs));
  for (uint64_t i = 0; i < ndocs; i++) {
    int zeros = 0;
    while (get(1) == 0) {
      if (++zeros > 62) throw Error(FM_ERR_FORMAT, "bad gamma code in chunk");
    }
    const uint64_t v = (uint64_t(1) << zeros) | get(zeros);
    lastdoc += int64_t(v);
    docs->push_back(lastdoc - 1);
  }
}

}  // namespace fmb
