// fm_debug.cc -- host-side inspection hooks for the rank image (tests of the LOADER only).
//
// These walk the host copy of the image (fm_loader.hpp) with plain loops so that the decoding of
// the on-disk format into rank blocks / node records can be verified on a machine without a GPU.
// They are not part of the query API: nothing in fm_api.cu calls them, they are not declared in
// include/femto_b200.h, and the fm_* query entry points fail when no CUDA device is present.
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "fm_loader.hpp"
#include "fm_stream_plan.hpp"

using namespace fmb;

namespace {
struct DebugImage {
  std::unique_ptr<HostImage> im;
};

bool split(const HostImage& im, int64_t row, int64_t* g, uint32_t* rb) {
  if (row < im.first_row || row >= im.end_row) return false;
  *g = row / im.hdr.bucket_size - im.first_bucket;
  *rb = uint32_t(row % im.hdr.bucket_size);
  return true;
}
}  // namespace

extern "C" {

void* fm_debug_image_open(const char* path, int shard, int nshards, int nthreads, int* err) {
  try {
    auto* d = new DebugImage();
    d->im = build_host_image(path, shard, nshards, nthreads);
    if (err) *err = 0;
    return d;
  } catch (const Error& e) {
    if (err) *err = e.code;
    return nullptr;
  }
}

void fm_debug_image_close(void* h) { delete static_cast<DebugImage*>(h); }

// out[0..7] = n_rank_blocks, n_wtree_blocks, nodes, buckets, markvals, first_row, end_row, max_code_len
void fm_debug_image_stats(void* h, int64_t* out) {
  const HostImage& im = *static_cast<DebugImage*>(h)->im;
  out[0] = im.n_rank_blocks;
  out[1] = im.n_wtree_blocks;
  out[2] = int64_t(im.levels == 4 ? im.quads.size() : im.levels == 2 ? im.supers.size() : im.nodes.size());
  out[3] = im.nbuckets;
  out[4] = int64_t(im.markvals.size());
  out[5] = im.first_row;
  out[6] = im.end_row;
  out[7] = im.max_code_len;
}

// C[ch] + Occ(ch,row) through the image tables; -1 if out of range
int64_t fm_debug_image_occ(void* h, int ch, int64_t row) {
  const HostImage& im = *static_cast<DebugImage*>(h)->im;
  int64_t g;
  uint32_t rb;
  if (ch < 0 || ch >= kAlpha || !split(im, row, &g, &rb)) return -1;
  const OccRec& o = im.occ[size_t(g) * kAlphaStride + size_t(ch)];
  if (!o.leaf) return o.occ_base;
  const BucketRec& br = im.buckets[size_t(g)];
  uint32_t base = br.root_base, node = br.root_node, idx1 = rb + 1;
  const int L = 31 - __builtin_clz(o.leaf);
  if (im.levels == 4) {
    // the root area and OccRec::root_exit must name the same block / exit entry as the bucket record
    if (int64_t(base) != g * im.root_stride) return -3;
    for (int lvl = 0; lvl < L; lvl += 4) {
      const int rem = L - lvl;  // code bits left; a code ending inside the block is extended with 0 bits
      const uint32_t path = rem >= 4 ? (o.leaf >> (rem - 4)) & 15u : (o.leaf << (4 - rem)) & 15u;
      const HostQuadRank r = host_quad_rank(im.rank_words, base, idx1, int(path));
      idx1 = r.index1;
      const uint32_t* ex = im.quads[node].exit[path];
      if (lvl == 0) {  // the record's shortcut to this entry
        if (o.root_exit & kRootExitDirect) {
          if (L <= 4 || L > 8 || (o.root_exit & ~kRootExitDirect) != ex[0]) return -3;
        } else if (ex != &im.quads[0].exit[0][0] + 2 * size_t(o.root_exit)) {
          return -3;
        }
      }
      if (rem <= 4) {
        if (!(ex[1] & kChildLeaf) || (ex[1] & 0xffffu) != uint32_t(ch)) return -2;
        break;
      }
      if (ex[1] & kChildLeaf) return -2;
      if (idx1 == 0) break;
      base = ex[0];
      node = ex[1];
    }
    return o.occ_base + idx1;
  }
  if (im.levels == 2) {
    for (int lvl = 0; lvl < L; lvl += 2) {
      const uint32_t b1 = (o.leaf >> (L - lvl - 1)) & 1u;
      const HostPairedRank r = host_paired_rank(im.rank_words, im.block_words, base, idx1, int(b1));
      const SuperRec& sr = im.supers[node];
      idx1 = r.index1;
      if (lvl + 1 == L) {
        if (!(sr.child_info[b1] & kChildLeaf)) return -2;
        break;
      }
      if (sr.child_info[b1] & kChildLeaf) return -2;
      if (idx1 == 0) break;
      const uint32_t b2 = (o.leaf >> (L - lvl - 2)) & 1u;
      idx1 = b2 ? r.ones2 : idx1 - r.ones2;
      if (idx1 == 0) break;
      const uint32_t* gc = sr.gc[2 * b1 + b2];
      if (lvl + 2 == L) {
        if (!(gc[1] & kChildLeaf)) return -2;
        break;
      }
      if (gc[1] & kChildLeaf) return -2;
      base = gc[0];
      node = gc[1];
    }
    return o.occ_base + idx1;
  }
  for (int lvl = 1; lvl <= L; lvl++) {
    const HostRank r = host_rank(im.rank_words, im.block_words, base, idx1);
    const uint32_t b = (o.leaf >> (L - lvl)) & 1u;
    idx1 = b ? r.ones : idx1 - r.ones;
    if (idx1 == 0) break;
    if (lvl < L) {
      const NodeRec& nr = im.nodes[node];
      if (nr.child_info[b] & kChildLeaf) return -2;  // image inconsistent with the leaf code
      base = nr.child_base[b];
      node = nr.child_info[b];
    }
  }
  return o.occ_base + idx1;
}

// one LF step with mark test; returns 0 or a negative error
int fm_debug_image_back_step(void* h, int64_t row, int32_t* ch_out, int64_t* next, int64_t* offset) {
  const HostImage& im = *static_cast<DebugImage*>(h)->im;
  int64_t g;
  uint32_t rb;
  if (!split(im, row, &g, &rb)) return -1;
  const BucketRec& br = im.buckets[size_t(g)];
  uint32_t base = br.root_base, node = br.root_node, idx1 = rb + 1, ch = 0;
  for (int guard = 0; guard < 64 && im.levels == 4; guard++) {
    const HostQuadRank r = host_quad_rank(im.rank_words, base, idx1, -1);
    const uint32_t* ex = im.quads[node].exit[r.exit];
    idx1 = r.index1;
    if (ex[1] & kChildLeaf) { ch = ex[1] & 0xffffu; break; }
    base = ex[0];
    node = ex[1];
  }
  for (int guard = 0; guard < 64 && im.levels == 2; guard++) {
    const HostPairedRank r = host_paired_rank(im.rank_words, im.block_words, base, idx1, -1);
    const SuperRec& sr = im.supers[node];
    idx1 = r.index1;
    if (sr.child_info[r.bit1] & kChildLeaf) { ch = sr.child_info[r.bit1] & 0xffffu; break; }
    idx1 = r.bit2 ? r.ones2 : idx1 - r.ones2;
    const uint32_t* gc = sr.gc[2 * r.bit1 + r.bit2];
    if (gc[1] & kChildLeaf) { ch = gc[1] & 0xffffu; break; }
    base = gc[0];
    node = gc[1];
  }
  for (int guard = 0; guard < 64 && im.levels == 1; guard++) {
    const HostRank r = host_rank(im.rank_words, im.block_words, base, idx1);
    idx1 = r.bit ? r.ones : idx1 - r.ones;
    const NodeRec& nr = im.nodes[node];
    const uint32_t info = nr.child_info[r.bit];
    if (info & kChildLeaf) { ch = info & 0xffffu; break; }
    base = nr.child_base[r.bit];
    node = info;
  }
  if (ch >= uint32_t(kAlpha) || idx1 == 0) return -2;
  const size_t rec = size_t(g) * kAlphaStride + ch;
  const HostRank m = host_rank(im.rank_words, im.block_words, im.mark[rec].mark_base, idx1);
  *ch_out = int32_t(ch);
  *offset = m.bit ? im.markvals[size_t(br.markval_base) + im.mark[rec].markval_off + m.ones - 1] : -1;
  *next = ch <= uint32_t(kEscSeof) ? -1 : im.occ[rec].occ_base + idx1 - 1;
  return 0;
}

}  // extern "C"

// Documents of the chunk holding `row` (HostImage tables; fmb::chunk_documents).  Returns the number
// of documents (docs filled up to cap), or -(error code).
extern "C" int64_t fm_debug_image_chunk(void* h, int64_t row, int64_t* first, int64_t* last, int64_t* docs,
                                        int64_t cap) {
  const fmb::HostImage& im = *static_cast<DebugImage*>(h)->im;
  try {
    std::vector<int64_t> d;
    fmb::chunk_documents(im.hdr, im.first_row, im.end_row, im.first_bucket, im.chunk_bytes, im.chunk_off,
                         im.chunk_count, im.chunk_dir_rel, row, first, last, &d);
    for (size_t i = 0; i < d.size() && int64_t(i) < cap; i++) docs[i] = d[i];
    return int64_t(d.size());
  } catch (const fmb::Error& e) {
    return -int64_t(e.code);
  }
}

// The chunk plan of a streamed count (fm_stream_plan.hpp) for an in-order batch, as count_host walks
// it: out receives 5 int64 per chunk {kernel (0/1), first pattern, end pattern, first symbol, end
// symbol}; returns the number of chunks, -1 when the batch is too small to be streamed, -2 when
// out_cap (in chunks) is too small.
extern "C" int64_t fm_debug_stream_plan2(int64_t npats, const int32_t* plen, const int64_t* offs, int64_t flat_len,
                                         int sym_bytes, int64_t* out, int64_t out_cap);
extern "C" int64_t fm_debug_stream_plan(int64_t npats, const int32_t* plen, const int64_t* offs, int64_t flat_len,
                                        int64_t* out, int64_t out_cap) {
  return fm_debug_stream_plan2(npats, plen, offs, flat_len, 2, out, out_cap);
}
// sym_bytes: 2 = alpha_t symbols (fm_count_flat), 1 = raw text bytes (fm_count_bytes)
extern "C" int64_t fm_debug_stream_plan2(int64_t npats, const int32_t* plen, const int64_t* offs, int64_t flat_len,
                                         int sym_bytes, int64_t* out, int64_t out_cap) {
  using namespace fmb;
  if (npats < kStreamMinBatch) return -1;
  const int64_t mid = stream_split(npats);
  const int64_t half_lo[2] = {0, mid}, half_hi[2] = {mid, npats};
  int64_t k = 0, fdone = 0;
  for (int h = 0; h < 2; h++) {
    for (int64_t lo = half_lo[h]; lo < half_hi[h]; lo += stream_chunk_at(lo), k++) {
      const int64_t hi = std::min(half_hi[h], lo + stream_chunk_at(lo));
      const int64_t fend = std::max(fdone, stream_symbol_cut(plen, offs, hi, npats, flat_len, sym_bytes));
      if (k >= out_cap) return -2;
      int64_t* o = out + 5 * k;
      o[0] = h; o[1] = lo; o[2] = hi; o[3] = fdone; o[4] = fend;
      fdone = fend;
    }
  }
  return k;
}

// shard_of_block (fm_format.hpp) for the CPU tests of the shard map
extern "C" int fm_debug_shard_of_block(int64_t b, int64_t block_size, int64_t total_length, int nshards) {
  return fmb::shard_of_block(b, block_size, total_length, nshards);
}
