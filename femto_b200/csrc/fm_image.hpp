// fm_image.hpp -- the HBM-resident rank image derived from a femto index at load time.
//
// The on-disk format (big-endian, 8-byte aligned, RLE-gamma / raw 512-bit segments behind
// three dependent searches per wavelet-tree level) is hostile to SIMT, so fm_open() decodes it
// ONCE into the layout below; queries never touch the file bytes again.  The layout keeps the
// reference's structure -- one Huffman-shaped wavelet tree per bucket, one mark bit-vector and
// one sampled-SA array per (bucket, symbol) -- so every quantity the reference computes
// (Occ, L[row], mark bit, SA sample) has a one-to-one counterpart here:
//
//   rank block (128 B, 128-B aligned)   word 0      = ones in the node's bit sequence before this block
//                                        words 1..31 = 992 bits of the sequence, MSB-first per word
//     -> one rank = ONE 128-byte line; replaces bseq header + A0/A1 bsearch + AP + S scan +
//        64-B segment decode (reference src/main/wtree.c:635-763)
//   NodeRec  (16 B)  per internal wavelet-tree node: rank-block base and identity of both children
//     -> replaces the per-level bsearch of the node directory (wtree_funcs.h:583-626)
//   OccRec   (16 B)  per (bucket, symbol): C[ch]+block_occs[ch][blk]+bucket_occs[ch][bucket] folded
//                    into one int64, and the symbol's leaf code in this bucket (0 = not in use)
//     -> replaces get_C + get_block_occs + get_bucket_occs + cached Huffman code
//        (src/main/index.c:1538-1569, 1828-1843, 2078-2084)
//   MarkRec  (8 B)   per (bucket, symbol): rank-block base of the mark bit-vector and offset of its
//                    sampled-SA values (src/main/index.c:2102-2140)
//   BucketRec(16 B)  per bucket: root node, root rank-block base, base of the bucket's SA samples
//
// All indices are relative to the shard resident on this device.
#pragma once

#include <cstdint>
#include <vector_types.h>  // uint4 (CUDA toolkit header, host-safe)

namespace fmb {

// Rank block size is chosen at load time: 32, 16 or 8 words (128 / 64 / 32 bytes); word 0 is the
// header, the other block_words-1 words are payload.
constexpr int kDefaultBlockWords = 32;
constexpr uint32_t kChildLeaf = 0x80000000u;    // NodeRec::child_info flag
constexpr int kAlphaStride = 261;               // records per bucket in OccRec / MarkRec tables

struct alignas(16) NodeRec {
  uint32_t child_base[2];  // first rank block of child b (internal children only)
  uint32_t child_info[2];  // kChildLeaf | symbol   or   node index of the internal child
};

// ---- paired-level layout (optional, chosen at load time) -------------------------------------
// Backward search is bound by the NUMBER of dependent random HBM reads (one per wavelet-tree
// level), not by their size.  The paired layout answers TWO levels from one block: a block of an
// even-depth node X ("super node") carries kPairedSlicePos positions of X per 32-byte slice and,
// for the same positions, the bits those positions contribute to X's two children:
//
//   slice s (8 words):  h0 = ones of X before the block
//                       h1 = s even: ones of child 0's sequence before the bits stored here
//                            s odd : ones of child 1's sequence before the bits stored here
//                                    + all ones of the block's children region
//                       X0..X2  bits [96 s, 96 s + 96) of the block's stretch of X, MSB-first
//                       R0..R2  bits [96 s, 96 s + 96) of the block's children region
//   children region (96 * slices bits): child 0's bits for the zeros of the stretch, in order,
//   from the front; child 1's bits for the ones of the stretch, REVERSED, from the back.  So the
//   first j bits of either child stored here are a prefix (child 0) or a suffix (child 1) of
//   the region, and both ranks are single-ended popcounts:
//       rank1(child 0, j) = h1(even) + ones(region[0, j))
//       rank1(child 1, j) = h1(odd)  - ones(region[0, RB - j))
//   A leaf child contributes no bits.  Mark bit-vectors keep the plain one-level blocks.
constexpr int kPairedSliceWords = 8;
constexpr int kPairedSlicePos = 96;

struct alignas(16) SuperRec {
  uint32_t gc[4][2];       // grandchild 2*b1+b2: {first block, kChildLeaf | symbol  or  SuperRec index}
  uint32_t child_info[2];  // child b1: kChildLeaf | symbol, or 0 when it is an internal node
  uint32_t pad[2];
};

// ---- quad-level layout ------------------------------------------------------------------------
// The memory system serves a fixed number of independent random reads per second whatever their
// width up to a 128-byte line (profiles/r01_paired_level_sweep.md), so the quad layout spends the
// whole line on FOUR levels: a block of a node X at depth 0, 4, 8, ... ("quad node") carries
// kQuadPos positions of X and the bits those positions contribute to X's children,
// grandchildren and great-grandchildren (block levels 1..3; block-local heap ids 1 = X,
// 2..3, 4..7, 8..15).  Codes that end inside a block are extended with 0 bits, a leaf behaving
// like a node whose bits are all 0, so every query leaves a block through one of 16 exits.
//
//   words 0..15   H[q], q = the 4 path bits b1 b2 b3 b4:
//                   bits 0..23  E(q): positions routed to exit q by the node's earlier blocks
//                   bits 24..31 q even: anchor of the level-2 node on the path (b1 b2)
//                               q odd : anchor of the level-3 node on the path (b1 b2 b3)
//   words 16..31  four regions of kQuadPos bits, one per block level, MSB-first;
//                 word w of region l is block word 16 + 8 (w >> 1) + 2 l + (w & 1), so that a
//                 lane of a 2-lane group loads words (2 sub, 2 sub + 1) of all four regions
//                 with two 128-bit loads.
//   Region l is partitioned among the level-l nodes.  Node v with parent u occupying [LO, HI):
//   a 0-child stores its bits forward from LO, a 1-child stores them REVERSED, backward from
//   HI; the anchor of a node is that LO resp. HI.  (Level 1: 0 and kQuadPos, not stored.)  The first
//   j bits of a node are the range [anchor, anchor + j) resp. [anchor - j, anchor), their ones c
//   give the next j' = b ? c : j - c, and the result after four levels is E(q) + j''''.
//   Positions past the end of the node's sequence count as 0 bits.  Needs bucket_size < 2^24.
//
//   Root area.  The blocks of the ROOT quad nodes are not allocated with the rest of their bucket:
//   they fill the front of the block array, bucket g (local) at [g * root_stride, (g + 1) *
//   root_stride) with root_stride = ceil(bucket_size / kQuadPos).  The first block a backward-search
//   step reads is therefore a function of the row alone -- g * root_stride + (row in bucket) /
//   kQuadPos -- and its read is issued together with the (bucket, symbol) record instead of after it.
constexpr int kQuadPos = 128;
constexpr int kQuadBlockWords = 32;

struct alignas(16) QuadRec {
  // exit q: {first block, QuadRec index} of the quad node there, or {0, kChildLeaf | symbol} when
  // the path b1..b4 is a leaf's code extended with 0 bits
  uint32_t exit[16][2];
};

struct alignas(16) OccRec {
  int64_t occ_base;  // C[ch] + occurrences of ch before this bucket
  uint32_t leaf;     // wavelet-tree leaf id (1<<len | code) of ch in this bucket, 0 = absent
  uint32_t root_exit;  // quad layout: 16 * (root QuadRec index) + first four path bits of ch, i.e. the
                       // index of the exit entry {first block, QuadRec} the second block read needs;
                       // or kRootExitDirect | first block of that entry, for codes of 5..8 bits (the
                       // second block is their last, so the QuadRec behind it is never needed)
};
constexpr uint32_t kRootExitDirect = 0x80000000u;

struct alignas(8) MarkRec {
  uint32_t mark_base;    // first rank block of the mark bit-vector
  uint32_t markval_off;  // index of the first SA sample of (bucket, ch) within the bucket's samples
};

struct alignas(16) BucketRec {
  uint32_t root_base;
  uint32_t root_node;
  uint64_t markval_base;
};

// Device view handed to kernels by value.
struct DevImage {
  const uint4* blocks = nullptr;        // rank blocks, 8 x uint4 each
  const NodeRec* nodes = nullptr;       // plain layout
  const SuperRec* supers = nullptr;     // paired-level layout (then nodes == nullptr)
  const QuadRec* quads = nullptr;       // quad-level layout
  const OccRec* occ = nullptr;         // [nbuckets][kAlphaStride]
  const MarkRec* mark = nullptr;        // [nbuckets][kAlphaStride]
  const BucketRec* buckets = nullptr;   // [nbuckets]
  const int64_t* markvals = nullptr;    // sampled SA values
  const int64_t* C = nullptr;           // [262]; C[261] = total_length (get_C, index.c:1545)
  int64_t total_length = 0;
  int64_t first_row = 0, end_row = 0;   // rows resident here
  int64_t first_bucket = 0;             // global index of buckets[0]
  int32_t bucket_size = 0;
  int32_t bucket_shift = -1;            // log2(bucket_size) when it is a power of two, else -1
  int32_t block_words = kDefaultBlockWords;  // 32-bit words per rank block (32, 16 or 8)
  int32_t levels = 1;                        // wavelet-tree levels answered per block read: 1, 2 (paired) or 4 (quad)
  int64_t root_stride = 0;                   // quad layout: blocks per bucket in the root area
};

}  // namespace fmb
