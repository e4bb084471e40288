// fm_builder.cc -- writes femto's on-disk index format from BWT rows ("next" row f-1 of the
// scope table: the index builder that feeds the query engine).
//
// The output is byte-identical to what the reference's construction path writes for the same
// (L, SA) rows, so that the reference reader opens it and the two engines can be compared on the
// same index.  Reference behaviour restated here (paths relative to the reference tree):
//   bucket layout + map/Huffman stream + marks   compress_bucket            src/main/index.c:309-738
//   Huffman code lengths (bzip2's algorithm)     BZ2_hbMakeCodeLengths      src/main/huffman.c:61-147
//   canonical codes                              BZ2_hbAssignCodes          src/main/huffman.c:152-167
//   wavelet tree over Huffman leaves             wtree_construct            src/main/wtree.c:907-1078
//   RLE-gamma / raw 512-bit segments, S/A0/A1/AP bseq_construct, save_run,  src/main/wtree.c:84-603
//                                                save_segment
//   document chunks (gamma-coded doc ids)        results_create_sort        src/main/results.c:133-197
//                                                bucket_compress_job        src/dcx_cc/dcx.hh:4993-5040
//   data block / header block                    begin/update/finish_*      src/main/index.c:817-1197
//                                                constructor_construct_header src/main/construct.c:293-566
//   marking rule                                 should_mark                src/main/index_types.h:134-144
//   flattened container                          flatten_index              src/main/index.c:2260-2365
#include <sys/stat.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/femto_b200.h"
#include "fm_format.hpp"

namespace fmb {
namespace {

using Bytes = std::vector<uint8_t>;

inline void put32(Bytes& b, uint32_t v) {
  const size_t n = b.size();
  b.resize(n + 4);
  put_be32(b.data() + n, v);
}
inline void set32(Bytes& b, size_t at, uint32_t v) { put_be32(b.data() + at, v); }
inline void put64(Bytes& b, uint64_t v) {
  const size_t n = b.size();
  b.resize(n + 8);
  put_be64(b.data() + n, v);
}
inline void align8(Bytes& b) {
  while (b.size() & 7u) b.push_back(0);
}

// MSB-first bit stream appended to a byte vector; flush() pads the last byte with zeros
// (bsW24 / bsFinishWrite, src/utils/buffer_funcs.h:63-127).
struct BitSink {
  Bytes& out;
  uint64_t acc = 0;
  int n = 0;
  explicit BitSink(Bytes& o) : out(o) {}
  void put(uint64_t v, int bits) {  // bits <= 56 per call
    while (bits > 0) {
      const int take = std::min(bits, 56 - n);
      const uint64_t part = (bits == take) ? (v & ((take == 64) ? ~0ull : ((1ull << take) - 1)))
                                           : ((v >> (bits - take)) & ((1ull << take) - 1));
      acc = (acc << take) | part;
      n += take;
      bits -= take;
      while (n >= 8) {
        out.push_back(uint8_t(acc >> (n - 8)));
        n -= 8;
      }
    }
  }
  void put_wide(uint64_t v, int bits) {  // up to 64 bits
    if (bits > 32) { put(v >> 32, bits - 32); put(v & 0xffffffffull, 32); }
    else put(v, bits);
  }
  void gamma(uint64_t v) {  // Elias gamma: floor(log2 v) zeros, then v in log+1 bits
    const int lg = 63 - __builtin_clzll(v);
    put_wide(0, lg);
    put_wide(v, lg + 1);
  }
  void flush() {
    if (n > 0) {
      out.push_back(uint8_t(acc << (8 - n)));
      n = 0;
    }
    acc = 0;
  }
};

// ------------------------------------------------------------------ Huffman code lengths
// bzip2's heap construction, including its tie behaviour (weights carry the subtree depth in the
// low 8 bits) and the halve-and-retry loop when a code would exceed max_len.
void huffman_code_lengths(const int32_t* freq, int n, int max_len, uint8_t* len_out) {
  std::vector<int32_t> heap(size_t(n) + 2), weight(size_t(n) * 2 + 2), parent(size_t(n) * 2 + 2);
  for (int i = 0; i < n; i++) weight[size_t(i) + 1] = (freq[i] == 0 ? 1 : freq[i]) << 8;
  for (;;) {
    int n_nodes = n, n_heap = 0;
    heap[0] = 0;
    weight[0] = 0;
    parent[0] = -2;
    auto sift_up = [&](int pos) {
      const int32_t item = heap[size_t(pos)];
      while (weight[size_t(item)] < weight[size_t(heap[size_t(pos >> 1)])]) {
        heap[size_t(pos)] = heap[size_t(pos >> 1)];
        pos >>= 1;
      }
      heap[size_t(pos)] = item;
    };
    auto sift_down = [&](int pos) {
      const int32_t item = heap[size_t(pos)];
      for (;;) {
        int child = pos << 1;
        if (child > n_heap) break;
        if (child < n_heap && weight[size_t(heap[size_t(child) + 1])] < weight[size_t(heap[size_t(child)])]) child++;
        if (weight[size_t(item)] < weight[size_t(heap[size_t(child)])]) break;
        heap[size_t(pos)] = heap[size_t(child)];
        pos = child;
      }
      heap[size_t(pos)] = item;
    };
    for (int i = 1; i <= n; i++) {
      parent[size_t(i)] = -1;
      heap[size_t(++n_heap)] = i;
      sift_up(n_heap);
    }
    while (n_heap > 1) {
      const int a = heap[1];
      heap[1] = heap[size_t(n_heap--)];
      sift_down(1);
      const int b = heap[1];
      heap[1] = heap[size_t(n_heap--)];
      sift_down(1);
      n_nodes++;
      parent[size_t(a)] = parent[size_t(b)] = n_nodes;
      const int32_t wa = weight[size_t(a)], wb = weight[size_t(b)];
      weight[size_t(n_nodes)] = int32_t((uint32_t(wa) & 0xffffff00u) + (uint32_t(wb) & 0xffffff00u)) |
                                (1 + std::max(wa & 0xff, wb & 0xff));
      parent[size_t(n_nodes)] = -1;
      heap[size_t(++n_heap)] = n_nodes;
      sift_up(n_heap);
    }
    bool too_long = false;
    for (int i = 1; i <= n; i++) {
      int depth = 0;
      for (int k = i; parent[size_t(k)] >= 0; k = parent[size_t(k)]) depth++;
      len_out[i - 1] = uint8_t(depth);
      if (depth > max_len) too_long = true;
    }
    if (!too_long) return;
    for (int i = 1; i <= n; i++) weight[size_t(i)] = (1 + (weight[size_t(i)] >> 8) / 2) << 8;
  }
}

// ------------------------------------------------------------------ bseq encoder
struct Segment {
  uint64_t w[kSegWords];
  int used = 0;
  int appends = 0;
  void start(bool rle, int firstbit) {
    std::memset(w, 0, sizeof(w));
    used = 0;
    push(rle ? 1 : 0, 1);
    if (rle) push(uint64_t(firstbit), 1);
    appends = 0;
  }
  bool room(int64_t need) const { return need <= int64_t(kSegBits - used); }
  void push(uint64_t v, int n) {  // n in 1..64, v < 2^n
    const int wi = used >> 6, off = used & 63, space = 64 - off;
    if (n <= space) {
      w[wi] |= v << (space - n);
    } else {
      w[wi] |= v >> (n - space);
      w[wi + 1] |= v << (64 - (n - space));
    }
    used += n;
    appends++;
  }
  void push_run(int bit, int n) {  // n copies of bit
    if (n <= 0) return;
    appends += n;
    if (!bit) { used += n; return; }
    int left = n;
    while (left > 0) {
      const int wi = used >> 6, off = used & 63;
      const int take = std::min(64 - off, left);
      const uint64_t ones = (take == 64) ? ~0ull : (((1ull << take) - 1) << (64 - off - take));
      w[wi] |= ones;
      used += take;
      left -= take;
    }
  }
};

inline int gamma_bits(uint64_t run) { return 1 + 2 * (63 - __builtin_clzll(run)); }

class BseqEncoder {
 public:
  explicit BseqEncoder(int force_type = 0) : initial_mode_(force_type) {}

  // bits: MSB-first packed 64-bit words, nbits > 0
  void encode(const uint64_t* bits, int64_t nbits, Bytes* out) {
    reset();
    int last = int(bits[0] >> 63);
    lastbit_ = last;
    mode_ = initial_mode_;
    rle_.start(true, last);
    unc_.start(false, last);
    int64_t run = 0;
    // run extraction, a word at a time
    int64_t pos = 0;
    while (pos < nbits) {
      const int64_t wi = pos >> 6;
      const int off = int(pos & 63);
      uint64_t x = bits[wi] << off;           // remaining bits of this word, MSB first
      int avail = int(std::min<int64_t>(64 - off, nbits - pos));
      if (last) x = ~x;                        // count how long the current bit value continues
      int same = x ? __builtin_clzll(x) : 64;
      if (same >= avail) {
        run += avail;
        pos += avail;
      } else {
        run += same;
        pos += same;
        save_run(last, run);
        last ^= 1;
        lastbit_ = last;
        run = 0;
      }
    }
    if (run) save_run(last, run);
    Segment& act = (mode_ >= 0) ? rle_ : unc_;
    if (act.appends > 0) save_segment();
    assemble(out);
  }

 private:
  void reset() {
    S_.clear(); A0_.clear(); A1_.clear(); AP_.clear(); D_.clear();
    occ_rle_[0] = occ_rle_[1] = occ_unc_[0] = occ_unc_[1] = 0;
    tot_[0] = tot_[1] = 0;
    seg_in_group_ = 0; segnum_ = 0; last_words_ = 0; seg_ap_ = 0;
  }

  static void varbyte(Bytes& s, uint32_t v) {  // 7 bits per byte, LSB first, last byte flagged 0x80
    for (;;) {
      uint8_t b = uint8_t(v & 0x7f);
      v >>= 7;
      if (!v) { s.push_back(b | 0x80); return; }
      s.push_back(b);
    }
  }

  void save_segment() {
    const bool use_rle = mode_ >= 0;
    Segment& act = use_rle ? rle_ : unc_;
    const int64_t* occ = use_rle ? occ_rle_ : occ_unc_;
    last_words_ = (act.used + 63) / 64;
    for (int i = 0; i < kSegWords; i++) D_.push_back(act.w[i]);
    segnum_++;
    varbyte(S_, uint32_t(occ[0]));
    varbyte(S_, uint32_t(occ[1]));
    if (seg_in_group_ == 0) {
      A0_.push_back(uint32_t(tot_[0]));
      A1_.push_back(uint32_t(tot_[1]));
      AP_.push_back(uint32_t(seg_ap_));
    }
    if (++seg_in_group_ >= kSegsPerGroup) seg_in_group_ = 0;
    tot_[0] += occ[0];
    tot_[1] += occ[1];
    occ_rle_[0] = occ_rle_[1] = occ_unc_[0] = occ_unc_[1] = 0;
    mode_ = initial_mode_;
    rle_.start(true, lastbit_);
    unc_.start(false, lastbit_);
    seg_ap_ = int64_t(S_.size());
  }

  void save_run(int bit, int64_t run) {
    int glen = gamma_bits(uint64_t(run));
    for (;;) {
      const bool rle_ok = rle_.room(glen);
      const bool unc_ok = unc_.room(run);
      if (!rle_ok && mode_ == 0) {
        // the RLE form is full: keep it only if it already represents at least a raw segment's worth
        mode_ = (occ_rle_[0] + occ_rle_[1] < kSegBits - 1) ? -1 : 1;
      }
      if (!unc_ok && mode_ == 0) mode_ = 1;
      const bool flush = (!rle_ok && !unc_ok) || (mode_ == 1 && !rle_ok) || (mode_ == -1 && !unc_ok);
      if (!flush) break;
      if (mode_ == -1 && !unc_ok) {  // top the raw segment up with the head of this run
        const int fit = int(std::min<int64_t>(kSegBits - unc_.used, run));
        unc_.push_run(bit, fit);
        occ_unc_[bit] += fit;
        run -= fit;
        if (run) glen = gamma_bits(uint64_t(run));
      }
      save_segment();
      if (run == 0) return;
    }
    if (mode_ >= 0) {
      rle_.push(uint64_t(run), glen);
      occ_rle_[bit] += run;
    }
    if (mode_ <= 0) {
      unc_.push_run(bit, int(run));
      occ_unc_[bit] += run;
    }
  }

  void assemble(Bytes* out) {
    const size_t ng = A0_.size();
    out->clear();
    put32(*out, 0);
    put32(*out, uint32_t(ng));
    const int64_t seg_words = int64_t(kSegWords) * (segnum_ - 1) + last_words_;
    put32(*out, uint32_t(seg_words));
    put32(*out, 0);  // D offset, patched below
    for (uint32_t v : A0_) put32(*out, v);
    for (uint32_t v : A1_) put32(*out, v);
    for (uint32_t v : AP_) put32(*out, v);
    out->insert(out->end(), S_.begin(), S_.end());
    align8(*out);
    set32(*out, 12, uint32_t(out->size()));
    for (int64_t i = 0; i < seg_words; i++) put64(*out, D_[size_t(i)]);
  }

  int initial_mode_;
  int mode_ = 0;
  int lastbit_ = 0;
  Segment rle_, unc_;
  int64_t occ_rle_[2], occ_unc_[2], tot_[2];
  int seg_in_group_ = 0;
  int64_t segnum_ = 0, last_words_ = 0, seg_ap_ = 0;
  Bytes S_;
  std::vector<uint32_t> A0_, A1_, AP_;
  std::vector<uint64_t> D_;
};

// Growable MSB-first bit vector.
struct BitVec {
  std::vector<uint64_t> w;
  uint64_t acc = 0;
  int n = 0;
  int64_t nbits = 0;
  void push(unsigned bit) {
    acc = (acc << 1) | bit;
    nbits++;
    if (++n == 64) { w.push_back(acc); acc = 0; n = 0; }
  }
  void finish() {
    if (n) { w.push_back(acc << (64 - n)); acc = 0; n = 0; }
    if (w.empty()) w.push_back(0);
  }
};

// ------------------------------------------------------------------ one bucket
struct BucketInput {
  int len = 0;
  const uint16_t* L = nullptr;
  const int64_t* offsets = nullptr;  // SA value if marked, else -1
  const int64_t* docs = nullptr;     // document of each row (only when chunk_size > 0)
};

void wtree_build(int alpha_size, const uint32_t* leaf, int len, const uint16_t* seq, BseqEncoder& enc, Bytes* out) {
  // internal nodes = every proper prefix of every leaf id
  std::vector<uint32_t> nodes;
  for (int s = 0; s < alpha_size; s++)
    for (uint32_t v = leaf[s] >> 1; v >= 1; v >>= 1) nodes.push_back(v);
  std::sort(nodes.begin(), nodes.end());
  nodes.erase(std::unique(nodes.begin(), nodes.end()), nodes.end());
  const size_t n_int = nodes.size();
  // path of each symbol: (slot, bit) from the leaf's parent up to the root
  struct Hop { uint32_t slot; uint32_t bit; };
  std::vector<std::vector<Hop>> path{size_t(alpha_size)};
  for (int s = 0; s < alpha_size; s++) {
    for (uint32_t v = leaf[s]; v > 1; v >>= 1) {
      const uint32_t parent = v >> 1;
      const size_t slot = size_t(std::lower_bound(nodes.begin(), nodes.end(), parent) - nodes.begin());
      path[size_t(s)].push_back(Hop{uint32_t(slot), v & 1u});
    }
  }
  std::vector<BitVec> bits(n_int);
  for (int i = 0; i < len; i++)
    for (const Hop& h : path[seq[i]]) bits[h.slot].push(h.bit);

  out->clear();
  put32(*out, uint32_t(n_int));
  const size_t dir_at = out->size();
  out->resize(dir_at + 8 * n_int, 0);
  for (size_t k = 0; k < n_int; k++) set32(*out, dir_at + 8 * k, nodes[k]);
  align8(*out);
  Bytes z;
  for (size_t k = 0; k < n_int; k++) {
    if (bits[k].nbits == 0) continue;  // offset stays 0: "no data"
    const int64_t nb = bits[k].nbits;
    bits[k].finish();
    enc.encode(bits[k].w.data(), nb, &z);
    set32(*out, dir_at + 8 * k + 4, uint32_t(out->size()));
    out->insert(out->end(), z.begin(), z.end());
    align8(*out);
    std::vector<uint64_t>().swap(bits[k].w);
  }
}

// Appends the compressed bucket to dst (dst.size() is the bucket's base) and fills occs[261].
void compress_one_bucket(const BucketInput& in, int64_t total_length, int chunk_size, Bytes& dst, int32_t* occs) {
  const int offset_bits = num_bits64(uint64_t(total_length));
  const int chunk_docs_bits = chunk_size > 0 ? num_bits64(uint64_t(chunk_size)) : 0;
  std::fill(occs, occs + kAlpha, 0);
  for (int i = 0; i < in.len; i++) {
    if (in.L[i] >= kAlpha) throw Error(FM_ERR_PARAM, "BWT symbol out of range");
    occs[in.L[i]]++;
  }
  // symbols in use, Huffman lengths (the end-of-bucket symbol gets frequency 1)
  uint16_t ch_to_seq[kAlpha];
  int32_t freq[kAlpha + 1];
  int n_in_use = 0;
  for (int c = 0; c < kAlpha; c++) {
    if (occs[c] > 0) { ch_to_seq[c] = uint16_t(n_in_use); freq[n_in_use++] = occs[c]; }
    else ch_to_seq[c] = 0xffff;
  }
  const int alpha_size = n_in_use + 1;
  freq[n_in_use] = 1;
  uint8_t code_len[kAlpha + 1];
  huffman_code_lengths(freq, alpha_size, kMaxCodeLen, code_len);
  int minl = 32, maxl = 0;
  for (int s = 0; s < alpha_size; s++) { minl = std::min<int>(minl, code_len[s]); maxl = std::max<int>(maxl, code_len[s]); }
  uint32_t leaf[kAlpha + 1];
  {
    uint32_t vec = 0;
    for (int l = minl; l <= maxl; l++) {
      for (int s = 0; s < alpha_size; s++)
        if (code_len[s] == l) leaf[s] = (vec++) | (1u << l);
      vec <<= 1;
    }
  }

  // per-symbol mark bits and mark records, in row order
  std::vector<BitVec> mark_bits{size_t(n_in_use)};
  std::vector<Bytes> mark_vals{size_t(n_in_use)};
  std::vector<std::unique_ptr<BitSink>> mark_sink;
  std::vector<int64_t> mark_count(size_t(n_in_use), 0);
  for (int s = 0; s < n_in_use; s++) mark_sink.emplace_back(new BitSink(mark_vals[size_t(s)]));
  std::vector<uint16_t> seq(static_cast<size_t>(in.len));
  for (int i = 0; i < in.len; i++) {
    const int s = ch_to_seq[in.L[i]];
    seq[size_t(i)] = uint16_t(s);
    const int64_t off = in.offsets[i];
    if (off != -1) {
      mark_bits[size_t(s)].push(1);
      mark_sink[size_t(s)]->put_wide(uint64_t(off), offset_bits);
      mark_count[size_t(s)]++;
    } else {
      mark_bits[size_t(s)].push(0);
    }
  }
  for (int s = 0; s < n_in_use; s++) mark_sink[size_t(s)]->flush();

  const size_t base = dst.size();
  for (int i = 0; i < 6; i++) put32(dst, 0);
  set32(dst, base + 0, kMagicBucket);
  const int num_chunks = chunk_size > 0 ? (in.len + chunk_size - 1) / chunk_size : 0;
  set32(dst, base + 20, uint32_t(num_chunks));

  // document chunks: directory of num_chunks+1 offsets, then per chunk the number of distinct
  // documents and the gamma-coded ascending (doc+1) deltas
  if (chunk_size > 0) {
    const size_t dir = dst.size();
    for (int i = 0; i < num_chunks + 1; i++) put32(dst, 0);
    std::vector<int64_t> ids;
    for (int c = 0, start = 0; start < in.len; start += chunk_size, c++) {
      const int n = std::min(chunk_size, in.len - start);
      ids.assign(in.docs + start, in.docs + start + n);
      std::sort(ids.begin(), ids.end());
      ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
      set32(dst, dir + 4 * size_t(c), uint32_t(dst.size() - base));
      {
        BitSink bs(dst);
        bs.put_wide(uint64_t(ids.size()), chunk_docs_bits);
        bs.flush();
      }
      {
        BitSink bs(dst);
        int64_t last = 0;
        for (int64_t d : ids) { bs.gamma(uint64_t(d + 1 - last)); last = d + 1; }
        bs.flush();
      }
      set32(dst, dir + 4 * size_t(c + 1), uint32_t(dst.size() - base));
    }
  }

  // symbol map + Huffman lengths
  set32(dst, base + 4, uint32_t(dst.size() - base));
  {
    BitSink bs(dst);
    constexpr int kGroups = (kAlpha + 15) / 16;
    bool used16[kGroups];
    for (int g = 0; g < kGroups; g++) {
      used16[g] = false;
      for (int j = 0; j < 16; j++)
        if (g * 16 + j < kAlpha && occs[g * 16 + j] > 0) used16[g] = true;
    }
    for (int g = 0; g < kGroups; g++) bs.put(used16[g] ? 1 : 0, 1);
    for (int g = 0; g < kGroups; g++)
      if (used16[g])
        for (int j = 0; j < 16; j++) bs.put((g * 16 + j < kAlpha && occs[g * 16 + j] > 0) ? 1 : 0, 1);
    int curr = code_len[0];
    bs.put(uint64_t(curr), 5);
    for (int s = 0; s < alpha_size; s++) {
      while (curr < code_len[s]) { bs.put(2, 2); curr++; }
      while (curr > code_len[s]) { bs.put(3, 2); curr--; }
      bs.put(0, 1);
    }
    bs.flush();
  }
  align8(dst);

  // wavelet tree
  set32(dst, base + 8, uint32_t(dst.size() - base));
  BseqEncoder enc;
  {
    Bytes wt;
    wtree_build(alpha_size, leaf, in.len, seq.data(), enc, &wt);
    dst.insert(dst.end(), wt.begin(), wt.end());
  }
  align8(dst);

  // mark tables
  set32(dst, base + 12, uint32_t(dst.size() - base));
  {
    const size_t sec = dst.size();
    dst.resize(sec + 4 * size_t(n_in_use), 0);
    align8(dst);
    Bytes z;
    for (int s = 0; s < n_in_use; s++) {
      set32(dst, sec + 4 * size_t(s), uint32_t(dst.size() - sec));
      const int64_t nb = mark_bits[size_t(s)].nbits;
      mark_bits[size_t(s)].finish();
      enc.encode(mark_bits[size_t(s)].w.data(), nb, &z);
      dst.insert(dst.end(), z.begin(), z.end());
      align8(dst);
    }
    align8(dst);
  }

  // mark arrays
  set32(dst, base + 16, uint32_t(dst.size() - base));
  {
    const size_t sec = dst.size();
    dst.resize(sec + 4 * size_t(n_in_use), 0);
    align8(dst);
    for (int s = 0; s < n_in_use; s++) {
      set32(dst, sec + 4 * size_t(s), uint32_t(dst.size() - sec));
      const size_t size = size_t((mark_count[size_t(s)] * offset_bits + 7) / 8);
      dst.insert(dst.end(), mark_vals[size_t(s)].begin(), mark_vals[size_t(s)].begin() + size);
    }
  }
  align8(dst);
}

void write_block_header_bytes(Bytes& b, uint32_t magic, int64_t block_number, int64_t nblocks, int64_t total_length,
                              int64_t ndocs, int32_t num_buckets, int32_t size, int32_t block_size,
                              int32_t bucket_size, int32_t mark_period, int32_t chunk_size) {
  put32(b, magic);
  put32(b, kFormatVersion);
  put64(b, uint64_t(block_number));
  put64(b, uint64_t(nblocks));
  put64(b, uint64_t(total_length));
  put64(b, uint64_t(ndocs));
  put32(b, uint32_t(num_buckets));
  put32(b, uint32_t(size));
  put32(b, 0);  // variable_block_size
  put32(b, uint32_t(block_size));
  put32(b, uint32_t(bucket_size));
  put32(b, uint32_t(mark_period));
  put32(b, 1);  // mark_type (set_default_param, index.c:136)
  put32(b, 0);  // variable_chunk_size
  put32(b, uint32_t(chunk_size));
  put32(b, uint32_t(kWtreeSettings));
  put32(b, uint32_t(kAlpha));
  put32(b, kMagicEndOfHeader);
}

void write_file(const std::string& path, const uint8_t* data, size_t n) {
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) throw Error(FM_ERR_IO, "cannot create " + path);
  const size_t w = n ? std::fwrite(data, 1, n, f) : 0;
  const int rc = std::fclose(f);
  if (w != n || rc) throw Error(FM_ERR_IO, "short write to " + path);
}

std::string block_path(const std::string& dir, int64_t id) {
  char buf[64];
  std::snprintf(buf, sizeof(buf), "/%02llx", static_cast<unsigned long long>(id));
  return dir + buf;
}

}  // namespace
}  // namespace fmb

using namespace fmb;

struct fm_builder {
  std::string dir;
  int64_t total_length = 0, ndocs = 0;
  std::vector<int64_t> doc_ends;
  int32_t block_size = 0, bucket_size = 0, chunk_size = 0, mark_period = 0;
  int nthreads = 1;
  int64_t nblocks = 0;
  int bpb = 0;
  // state
  int64_t rows_done = 0;      // rows fully flushed into blocks
  int64_t cur_block = 0;
  std::vector<uint16_t> L;    // rows of the block being accumulated
  std::vector<int64_t> off, docs;
  std::vector<int64_t> occs_total;              // [261] before the current block (within this builder's rows)
  std::vector<std::vector<int64_t>> block_counts; // [block][261]: symbols inside the block
  // range builders (fm_builder_create_range): data blocks [first_block, first_block + range_blocks) only
  bool range = false;
  int64_t first_block = 0, end_row = 0;
  std::vector<int64_t> eof_rows;
  std::vector<std::string> doc_info;
  std::vector<bool> doc_info_set;
};

namespace {

thread_local std::string g_builder_error;
int bfail(int code, const std::string& m) {
  g_builder_error = m;
  std::fprintf(stderr, "femto_b200 builder: %s\n", m.c_str());
  return code;
}

int64_t find_doc(const std::vector<int64_t>& ends, int64_t offset) {
  return int64_t(std::upper_bound(ends.begin(), ends.end(), offset) - ends.begin());
}

void flush_block(fm_builder* b) {
  const int64_t rows = int64_t(b->L.size());
  const int nbuckets = int((rows + b->bucket_size - 1) / b->bucket_size);
  std::vector<Bytes> zb{size_t(nbuckets)};
  std::vector<std::vector<int32_t>> occs(static_cast<size_t>(nbuckets), std::vector<int32_t>(kAlpha));
  std::atomic<int> next{0};
  std::mutex mu;
  std::unique_ptr<Error> err;
  auto worker = [&]() {
    for (;;) {
      const int k = next.fetch_add(1);
      if (k >= nbuckets) return;
      try {
        BucketInput in;
        const int64_t r0 = int64_t(k) * b->bucket_size;
        in.len = int(std::min<int64_t>(b->bucket_size, rows - r0));
        in.L = b->L.data() + r0;
        in.offsets = b->off.data() + r0;
        in.docs = b->chunk_size > 0 ? b->docs.data() + r0 : nullptr;
        compress_one_bucket(in, b->total_length, b->chunk_size, zb[size_t(k)], occs[size_t(k)].data());
      } catch (const Error& e) {
        std::lock_guard<std::mutex> g(mu);
        if (!err) err.reset(new Error(e));
        next.store(nbuckets);
      }
    }
  };
  const int nt = std::max(1, std::min(b->nthreads, nbuckets));
  std::vector<std::thread> th;
  for (int t = 1; t < nt; t++) th.emplace_back(worker);
  worker();
  for (auto& t : th) t.join();
  if (err) throw *err;

  // assemble the data block
  Bytes blk;
  size_t total = 0;
  for (auto& z : zb) total += z.size() + 8;
  blk.reserve(size_t(kBlockHeaderBytes) + 4 * (size_t(b->bpb) + 1) + 4 * size_t(kAlpha) * size_t(nbuckets) + 16 + total);
  write_block_header_bytes(blk, kMagicDataBlock, b->cur_block, b->nblocks, b->total_length, b->ndocs, nbuckets,
                           int32_t(rows), b->block_size, b->bucket_size, b->mark_period,
                           b->chunk_size > 0 ? b->chunk_size : -1);
  const size_t dir_at = blk.size();
  blk.resize(dir_at + 4 * (size_t(b->bpb) + 1), 0);
  const size_t occs_at = blk.size();
  blk.resize(occs_at + 4 * size_t(kAlpha) * size_t(nbuckets), 0);
  align8(blk);
  std::vector<int64_t> since(kAlpha, 0);
  for (int k = 0; k < nbuckets; k++) {
    for (int c = 0; c < kAlpha; c++)
      set32(blk, occs_at + 4 * (size_t(c) * size_t(nbuckets) + size_t(k)), uint32_t(since[size_t(c)]));
    align8(blk);
    if (blk.size() + zb[size_t(k)].size() + 8 >= (size_t(1) << 32)) throw Error(FM_ERR_FULL, "data block exceeds 4 GiB");
    set32(blk, dir_at + 4 * size_t(k), uint32_t(blk.size()));
    blk.insert(blk.end(), zb[size_t(k)].begin(), zb[size_t(k)].end());
    Bytes().swap(zb[size_t(k)]);
    align8(blk);
    set32(blk, dir_at + 4 * size_t(k + 1), uint32_t(blk.size()));
    for (int c = 0; c < kAlpha; c++) since[size_t(c)] += occs[size_t(k)][size_t(c)];
  }
  write_file(block_path(b->dir, b->cur_block + 1), blk.data(), blk.size());

  b->block_counts.push_back(since);
  for (int c = 0; c < kAlpha; c++) b->occs_total[size_t(c)] += since[size_t(c)];
  b->rows_done += rows;
  b->cur_block++;
  b->L.clear();
  b->off.clear();
  b->docs.clear();
}

// Header block (constructor_construct_header, src/main/construct.c:407-460) from the per-block symbol
// counts: C[], occurrences before every block, document ends, rows of the document ends, info strings.
void write_header_block(const std::string& dir, int64_t total_length, int64_t ndocs, const int64_t* doc_ends,
                        int32_t block_size, int32_t bucket_size, int32_t chunk_size, int32_t mark_period,
                        int64_t nblocks, const int64_t* block_counts, const int64_t* eof_rows,
                        const std::vector<std::string>& doc_info, const std::vector<bool>& doc_info_set) {
  Bytes h;
  write_block_header_bytes(h, kMagicHeaderBlock, -1, nblocks, total_length, ndocs, 0, 0, block_size, bucket_size,
                           mark_period, chunk_size > 0 ? chunk_size : -1);
  std::vector<int64_t> total(kAlpha, 0);
  for (int64_t k = 0; k < nblocks; k++)
    for (int c = 0; c < kAlpha; c++) total[size_t(c)] += block_counts[size_t(k) * kAlpha + size_t(c)];
  int64_t sum = 0;
  for (int c = 0; c < kAlpha; c++) { put64(h, uint64_t(sum)); sum += total[size_t(c)]; }
  if (sum != total_length) throw Error(FM_ERR_FORMAT, "block symbol counts do not add up to total_length");
  for (int c = 0; c < kAlpha; c++) {
    int64_t before = 0;
    for (int64_t k = 0; k < nblocks; k++) {
      put64(h, uint64_t(before));
      before += block_counts[size_t(k) * kAlpha + size_t(c)];
    }
  }
  for (int64_t d = 0; d < ndocs; d++) put64(h, uint64_t(doc_ends[size_t(d)]));
  for (int64_t d = 0; d < ndocs; d++) {
    if (eof_rows[size_t(d)] < 0 || eof_rows[size_t(d)] >= ndocs) throw Error(FM_ERR_FORMAT, "document end row missing");
    put64(h, uint64_t(eof_rows[size_t(d)]));
  }
  const size_t info_dir = h.size();
  h.resize(info_dir + 8 * (size_t(ndocs) + 1), 0);
  for (int64_t d = 0; d < ndocs; d++) {
    std::string info = size_t(d) < doc_info_set.size() && doc_info_set[size_t(d)] ? doc_info[size_t(d)] : ("doc" + std::to_string(d));
    put_be64(h.data() + info_dir + 8 * size_t(d), uint64_t(h.size()));
    h.insert(h.end(), info.begin(), info.end());
    put_be64(h.data() + info_dir + 8 * size_t(d + 1), uint64_t(h.size()));
  }
  write_file(block_path(dir, 0), h.data(), h.size());
  const std::string tag = "This is a FEMTO index constructed by femto_b200\n";
  write_file(dir + "/_femto_index", reinterpret_cast<const uint8_t*>(tag.data()), tag.size());
}

int check_build_params(const char* who, const char* out_dir, int64_t total_length, int64_t ndocs, const int64_t* doc_ends,
                       int32_t block_size, int32_t bucket_size, int32_t chunk_size, int32_t mark_period) {
  const std::string w = who;
  if (!out_dir || !doc_ends || total_length <= 0 || ndocs <= 0) return bfail(FM_ERR_PARAM, w + ": bad argument");
  if (block_size <= 0 || bucket_size <= 0 || block_size % bucket_size != 0)
    return bfail(FM_ERR_PARAM, w + ": block_size must be a positive multiple of bucket_size");
  if (chunk_size > 0 && bucket_size % chunk_size != 0)
    return bfail(FM_ERR_PARAM, w + ": bucket_size must be a multiple of chunk_size");
  if (mark_period < 0) return bfail(FM_ERR_PARAM, w + ": negative mark_period");
  for (int64_t d = 0; d < ndocs; d++)
    if (doc_ends[d] <= (d ? doc_ends[d - 1] : 0)) return bfail(FM_ERR_PARAM, w + ": empty or unordered document");
  if (doc_ends[ndocs - 1] != total_length) return bfail(FM_ERR_PARAM, w + ": document ends do not sum to total_length");
  if (mkdir(out_dir, 0777) && errno != EEXIST) return bfail(FM_ERR_IO, std::string("cannot create ") + out_dir);
  return FM_OK;
}

}  // namespace

extern "C" {

int fm_builder_create_range(const char* out_dir, int64_t total_length, int64_t ndocs, const int64_t* doc_ends,
                            int32_t block_size, int32_t bucket_size, int32_t chunk_size, int32_t mark_period,
                            int nthreads, int64_t first_block, int64_t range_blocks, fm_builder_t** out) {
  if (!out) return bfail(FM_ERR_PARAM, "fm_builder_create: bad argument");
  const int rc = check_build_params("fm_builder_create", out_dir, total_length, ndocs, doc_ends, block_size, bucket_size,
                                    chunk_size, mark_period);
  if (rc) return rc;
  const int64_t nblocks = (total_length + block_size - 1) / block_size;
  const bool range = first_block >= 0;
  if (range && (range_blocks < 0 || first_block + range_blocks > nblocks))
    return bfail(FM_ERR_PARAM, "fm_builder_create_range: block range outside the index");
  fm_builder* b = new fm_builder();
  b->dir = out_dir;
  b->total_length = total_length;
  b->ndocs = ndocs;
  b->doc_ends.assign(doc_ends, doc_ends + ndocs);
  b->block_size = block_size;
  b->bucket_size = bucket_size;
  b->chunk_size = chunk_size > 0 ? chunk_size : 0;
  b->mark_period = mark_period;
  b->nthreads = nthreads > 0 ? nthreads : int(std::max(1u, std::thread::hardware_concurrency()));
  b->nblocks = nblocks;
  b->bpb = block_size / bucket_size;
  b->occs_total.assign(kAlpha, 0);
  b->eof_rows.assign(size_t(ndocs), -1);
  b->doc_info.resize(size_t(ndocs));
  b->doc_info_set.assign(size_t(ndocs), false);
  b->range = range;
  b->end_row = total_length;
  if (range) {
    b->first_block = first_block;
    b->cur_block = first_block;
    b->rows_done = std::min<int64_t>(total_length, first_block * int64_t(block_size));
    b->end_row = std::min<int64_t>(total_length, (first_block + range_blocks) * int64_t(block_size));
  }
  *out = b;
  return FM_OK;
}

int fm_builder_create(const char* out_dir, int64_t total_length, int64_t ndocs, const int64_t* doc_ends,
                      int32_t block_size, int32_t bucket_size, int32_t chunk_size, int32_t mark_period,
                      int nthreads, fm_builder_t** out) {
  return fm_builder_create_range(out_dir, total_length, ndocs, doc_ends, block_size, bucket_size, chunk_size, mark_period,
                                 nthreads, -1, 0, out);
}

int fm_builder_set_doc_info(fm_builder_t* b, int64_t doc, const void* info, int64_t len) {
  if (!b || doc < 0 || doc >= b->ndocs || len < 0 || (len && !info)) return bfail(FM_ERR_PARAM, "fm_builder_set_doc_info: bad argument");
  b->doc_info[size_t(doc)].assign(static_cast<const char*>(info), size_t(len));
  b->doc_info_set[size_t(doc)] = true;
  return FM_OK;
}

int fm_builder_append(fm_builder_t* b, int64_t nrows, const uint16_t* L, const int64_t* sa) {
  if (!b || nrows < 0 || (nrows && (!L || !sa))) return bfail(FM_ERR_PARAM, "fm_builder_append: bad argument");
  try {
    int64_t i = 0;
    while (i < nrows) {
      const int64_t row0 = b->rows_done + int64_t(b->L.size());
      if (row0 >= b->end_row) return bfail(FM_ERR_PARAM, "fm_builder_append: more rows than the builder's range holds");
      const int64_t block_rows = std::min<int64_t>(b->block_size, b->total_length - b->rows_done);
      const int64_t take = std::min<int64_t>(nrows - i, block_rows - int64_t(b->L.size()));
      const size_t at = b->L.size();
      b->L.resize(at + size_t(take));
      b->off.resize(at + size_t(take));
      if (b->chunk_size > 0) b->docs.resize(at + size_t(take));
      std::atomic<int> bad{0};
      auto mark_rows = [&](int64_t j0, int64_t j1) {
        const int64_t period = b->mark_period;
        for (int64_t j = j0; j < j1; j++) {
          const int64_t s = sa[i + j];
          if (s < 0 || s >= b->total_length) { bad.store(1); return; }
          const int64_t doc = b->ndocs == 1 ? 0 : find_doc(b->doc_ends, s);
          const int64_t start = doc ? b->doc_ends[size_t(doc - 1)] : 0;
          const int64_t dlen = b->doc_ends[size_t(doc)] - start;
          const int64_t doff = s - start;
          bool mark = false;  // should_mark (index_types.h:134-144)
          if (period != 0) mark = doff == 0 || doff == dlen - 1 || doff % period == 0;
          b->L[at + size_t(j)] = L[i + j];
          b->off[at + size_t(j)] = mark ? s : -1;
          if (b->chunk_size > 0) b->docs[at + size_t(j)] = doc;
          // the first ndocs rows are the suffixes starting with each document's SEOF
          // (constructor_construct_header, construct.c:407-460)
          const int64_t row = row0 + j;
          if (row < b->ndocs) {
            if (!mark) { bad.store(2); return; }
            b->eof_rows[size_t(doc)] = row;
          }
        }
      };
      const int nt = int(std::min<int64_t>(b->nthreads, take >> 18));
      if (nt <= 1) {
        mark_rows(0, take);
      } else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++) th.emplace_back(mark_rows, take * t / nt, take * (t + 1) / nt);
        for (auto& t : th) t.join();
      }
      if (bad.load() == 1) return bfail(FM_ERR_PARAM, "fm_builder_append: suffix array value out of range");
      if (bad.load() == 2) return bfail(FM_ERR_PARAM, "Final EOF character of each document must be marked");
      i += take;
      if (int64_t(b->L.size()) == block_rows) flush_block(b);
    }
  } catch (const Error& e) {
    return bfail(e.code, std::string("fm_builder_append: ") + e.what());
  } catch (const std::bad_alloc&) {
    return bfail(FM_ERR_MEM, "fm_builder_append: out of memory");
  }
  return FM_OK;
}

void fm_builder_abort(fm_builder_t* b) { delete b; }

int fm_builder_finish(fm_builder_t* b) {
  if (!b) return bfail(FM_ERR_PARAM, "fm_builder_finish: null builder");
  std::unique_ptr<fm_builder> guard(b);
  if (b->range) return bfail(FM_ERR_PARAM, "fm_builder_finish: a range builder ends with fm_builder_finish_range");
  if (b->rows_done != b->total_length || !b->L.empty())
    return bfail(FM_ERR_PARAM, "fm_builder_finish: fewer rows appended than total_length");
  try {
    std::vector<int64_t> counts(size_t(b->nblocks) * kAlpha);
    for (int64_t k = 0; k < b->nblocks; k++)
      std::copy(b->block_counts[size_t(k)].begin(), b->block_counts[size_t(k)].end(), counts.begin() + k * kAlpha);
    write_header_block(b->dir, b->total_length, b->ndocs, b->doc_ends.data(), b->block_size, b->bucket_size, b->chunk_size,
                       b->mark_period, b->nblocks, counts.data(), b->eof_rows.data(), b->doc_info, b->doc_info_set);
  } catch (const Error& e) {
    return bfail(e.code, std::string("fm_builder_finish: ") + e.what());
  }
  return FM_OK;
}

int fm_builder_finish_range(fm_builder_t* b, int64_t* block_counts, int64_t* eof_rows) {
  if (!b) return bfail(FM_ERR_PARAM, "fm_builder_finish_range: null builder");
  std::unique_ptr<fm_builder> guard(b);
  if (!b->range || !block_counts || !eof_rows) return bfail(FM_ERR_PARAM, "fm_builder_finish_range: bad argument");
  if (b->rows_done != b->end_row || !b->L.empty())
    return bfail(FM_ERR_PARAM, "fm_builder_finish_range: fewer rows appended than the range holds");
  for (size_t k = 0; k < b->block_counts.size(); k++)
    std::copy(b->block_counts[k].begin(), b->block_counts[k].end(), block_counts + k * kAlpha);
  std::copy(b->eof_rows.begin(), b->eof_rows.end(), eof_rows);
  return FM_OK;
}

int fm_builder_write_header(const char* out_dir, int64_t total_length, int64_t ndocs, const int64_t* doc_ends,
                            int32_t block_size, int32_t bucket_size, int32_t chunk_size, int32_t mark_period,
                            const int64_t* block_counts, const int64_t* eof_rows, const void* const* doc_info,
                            const int64_t* doc_info_len) {
  if (!block_counts || !eof_rows) return bfail(FM_ERR_PARAM, "fm_builder_write_header: bad argument");
  const int rc = check_build_params("fm_builder_write_header", out_dir, total_length, ndocs, doc_ends, block_size,
                                    bucket_size, chunk_size, mark_period);
  if (rc) return rc;
  try {
    std::vector<std::string> info;
    std::vector<bool> info_set;
    if (doc_info && doc_info_len) {
      info.resize(size_t(ndocs));
      info_set.assign(size_t(ndocs), false);
      for (int64_t d = 0; d < ndocs; d++)
        if (doc_info[d] && doc_info_len[d] >= 0) {
          info[size_t(d)].assign(static_cast<const char*>(doc_info[d]), size_t(doc_info_len[d]));
          info_set[size_t(d)] = true;
        }
    }
    const int64_t nblocks = (total_length + block_size - 1) / block_size;
    write_header_block(out_dir, total_length, ndocs, doc_ends, block_size, bucket_size, chunk_size > 0 ? chunk_size : 0,
                       mark_period, nblocks, block_counts, eof_rows, info, info_set);
  } catch (const Error& e) {
    return bfail(e.code, std::string("fm_builder_write_header: ") + e.what());
  }
  return FM_OK;
}

int fm_flatten(const char* index_dir, const char* out_file) {
  if (!index_dir || !out_file) return bfail(FM_ERR_PARAM, "fm_flatten: null argument");
  try {
    auto files = IndexFiles::open(index_dir);
    if (files->flattened()) return bfail(FM_ERR_PARAM, "fm_flatten: input is already flattened");
    const int64_t nb = files->nblocks() + 1;  // header block + data blocks
    const int64_t page = 4096;
    FILE* f = std::fopen(out_file, "wb");
    if (!f) return bfail(FM_ERR_IO, std::string("cannot create ") + out_file);
    Bytes head;
    put32(head, kMagicFlattened);
    put32(head, kFormatVersion);
    put64(head, uint64_t(nb));
    std::vector<int64_t> offs(size_t(nb) + 1, 0);
    head.resize(head.size() + 8 * (size_t(nb) + 1), 0);
    int64_t pos = int64_t(head.size());
    auto pad = [&](int64_t& p) { p = (p + page - 1) / page * page; };
    pad(pos);
    // first pass: offsets
    std::vector<int64_t> sizes(static_cast<size_t>(nb));
    for (int64_t i = 0; i < nb; i++) {
      sizes[size_t(i)] = i == 0 ? int64_t(files->header_blob().size()) : int64_t(files->map_block(i - 1).size());
      offs[size_t(i)] = pos;
      pos += sizes[size_t(i)];
      pad(pos);
    }
    offs[size_t(nb)] = pos;
    for (int64_t i = 0; i <= nb; i++) put_be64(head.data() + 16 + 8 * size_t(i), uint64_t(offs[size_t(i)]));
    auto put = [&](const uint8_t* p, size_t n) { if (n && std::fwrite(p, 1, n, f) != n) throw Error(FM_ERR_IO, "short write"); };
    auto fill_to = [&](int64_t target, int64_t& cur) { static const uint8_t z[4096] = {0}; while (cur < target) { const size_t n = size_t(std::min<int64_t>(4096, target - cur)); put(z, n); cur += int64_t(n); } };
    int64_t cur = 0;
    put(head.data(), head.size());
    cur = int64_t(head.size());
    for (int64_t i = 0; i < nb; i++) {
      fill_to(offs[size_t(i)], cur);
      if (i == 0) put(files->header_blob().data(), files->header_blob().size());
      else { Blob b = files->map_block(i - 1); put(b.data(), b.size()); }
      cur += sizes[size_t(i)];
    }
    fill_to(offs[size_t(nb)], cur);
    if (std::fclose(f)) return bfail(FM_ERR_IO, "fm_flatten: close failed");
  } catch (const Error& e) {
    return bfail(e.code, std::string("fm_flatten: ") + e.what());
  }
  return FM_OK;
}

// test hook: encode one bit sequence exactly as the reference's bseq_construct_forcetype
// (bits MSB-first packed in bytes); returns a malloc()ed buffer.
int fm_debug_bseq_encode(const uint8_t* bits_msb, int64_t nbits, int force_type, uint8_t** out, int64_t* out_len) {
  if (!bits_msb || nbits <= 0 || !out || !out_len) return FM_ERR_PARAM;
  std::vector<uint64_t> w(size_t((nbits + 63) / 64), 0);
  for (int64_t i = 0; i < (nbits + 7) / 8; i++) w[size_t(i >> 3)] |= uint64_t(bits_msb[i]) << (56 - 8 * (i & 7));
  Bytes z;
  BseqEncoder enc(force_type);
  enc.encode(w.data(), nbits, &z);
  *out = static_cast<uint8_t*>(std::malloc(z.size()));
  if (!*out) return FM_ERR_MEM;
  std::memcpy(*out, z.data(), z.size());
  *out_len = int64_t(z.size());
  return FM_OK;
}

// test hook: expand an encoded bit sequence with the loader's decoder (MSB-first bytes out)
int fm_debug_bseq_expand(const uint8_t* z, int64_t zlen, uint8_t* bits_out, int64_t cap_bits, int64_t* nbits) {
  try {
    BseqView v = open_bseq(z, size_t(zlen));
    const int64_t n = bseq_length(v);
    *nbits = n;
    if (n > cap_bits) return FM_ERR_FULL;
    std::vector<uint32_t> w(size_t((n + 31) / 32) + 1, 0);
    bseq_expand(v, w.data(), n);
    for (int64_t i = 0; i < (n + 7) / 8; i++) bits_out[i] = uint8_t(w[size_t(i >> 2)] >> (24 - 8 * (i & 3)));
  } catch (const Error& e) {
    return e.code;
  }
  return FM_OK;
}

void fm_debug_free(void* p) { std::free(p); }

}  // extern "C"
