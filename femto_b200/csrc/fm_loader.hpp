// fm_loader.hpp -- builds the host copy of the rank image (fm_image.hpp) from index files.
#pragma once

#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "fm_format.hpp"
#include "fm_image.hpp"

namespace fmb {

struct HostImage {
  // geometry
  BlockHeader hdr;              // header block's fields
  int64_t first_block = 0, end_block = 0;   // data blocks resident in this image
  int64_t first_bucket = 0, nbuckets = 0;   // global index of bucket 0, buckets resident
  int64_t first_row = 0, end_row = 0;
  int max_code_len = 0;
  // tables
  int block_words = kDefaultBlockWords;  // 32-bit words per rank block (32, 16 or 8)
  int levels = 1;                        // wavelet-tree levels per block: 1, 2 (paired) or 4 (quad); fm_image.hpp
  uint32_t* rank_words = nullptr;   // n_rank_blocks * block_words words (calloc'ed)
  int64_t n_rank_blocks = 0;
  int64_t root_stride = 0;          // quad layout: blocks per bucket in the root area at the front (fm_image.hpp)
  int64_t n_wtree_blocks = 0;       // of which wavelet-tree payload (the rest are mark bit-vectors)
  std::vector<NodeRec> nodes;       // plain layout
  std::vector<SuperRec> supers;     // paired-level layout
  std::vector<QuadRec> quads;       // quad-level layout
  std::vector<OccRec> occ;
  std::vector<MarkRec> mark;
  std::vector<BucketRec> buckets;
  std::vector<int64_t> markvals;
  std::vector<int64_t> C;           // 262 entries
  std::vector<int64_t> doc_ends, doc_eof_rows;
  std::vector<int64_t> doc_info_off;  // ndocs + 1 offsets into doc_info_bytes
  std::vector<uint8_t> doc_info_bytes;
  // document chunks (block_get_chunk, index.c:2147-2197): per resident bucket the bytes of its chunk
  // section as stored -- directory of (nchunks + 1) BE u32 offsets relative to the bucket start,
  // then per chunk the number of documents (chunk_docs_bits bits, MSB-first, byte-flushed) and the
  // gamma-coded ascending (document + 1) deltas.  Decoded on demand (chunk_documents).
  std::vector<uint8_t> chunk_bytes;
  std::vector<int64_t> chunk_off;     // nbuckets + 1 offsets into chunk_bytes
  std::vector<int32_t> chunk_count;   // chunks per bucket
  std::vector<uint32_t> chunk_dir_rel;  // offset of the directory from its bucket's start (to rebase entries)
  HostImage() = default;
  HostImage(const HostImage&) = delete;
  HostImage& operator=(const HostImage&) = delete;
  ~HostImage();
};

// shard/nshards select data blocks b with b*nshards/nblocks == shard (all blocks when nshards==1).
// levels: wavelet-tree levels answered per block read (fm_image.hpp): 1, 2 (paired) or 4 (quad);
// 0 = the process default (set_default_levels_per_block, else env FEMTO_B200_LEVELS_PER_BLOCK,
// else paired).
// block_words: 32, 16 or 8 (128/64/32-byte rank blocks); 0 = the process default
// (set_default_block_words, else env FEMTO_B200_BLOCK_BYTES, else 64-byte blocks for the paired
// layout and 128-byte blocks otherwise).  Paired needs >= 64 bytes, quad exactly 128.
std::unique_ptr<HostImage> build_host_image(const std::string& path, int shard, int nshards, int nthreads,
                                            int block_words = 0, int levels = 0);
int default_block_words(int levels);
bool set_default_block_words(int words);          // 0 = back to the built-in default
int default_levels_per_block();
bool set_default_levels_per_block(int levels);    // 0 = back to the built-in default

// Documents of one chunk: `row` (global) selects the chunk as block_chunk_request with
// BLOCK_CHUNK_FIND_NUMBER does (index.c:2200-2215); *first / *last receive the chunk's global row
// range, docs its ascending document numbers.  Works on the tables above (host side; the same
// function serves the C ABI and the CPU tests).  Throws Error (FM_ERR_MISSING: index built without
// chunks; FM_ERR_PARAM: row not resident; FM_ERR_FORMAT: malformed chunk).
void chunk_documents(const BlockHeader& hdr, int64_t first_row, int64_t end_row, int64_t first_bucket,
                     const std::vector<uint8_t>& chunk_bytes, const std::vector<int64_t>& chunk_off,
                     const std::vector<int32_t>& chunk_count, const std::vector<uint32_t>& chunk_dir_rel,
                     int64_t row, int64_t* first, int64_t* last, std::vector<int64_t>* docs);

// Host-side rank over the image (used by the loader's self-check and by unit tests of the
// image layout; NOT a query fallback -- the C ABI never calls it).
struct HostRank { uint32_t ones; uint32_t bit; };
HostRank host_rank(const uint32_t* rank_words, int block_words, uint32_t base_block, uint32_t index1);

// Two levels of a paired-level block (fm_image.hpp): level one as host_rank, taking child `follow`
// (0/1) or the child named by the bit at the position (follow < 0); then the rank of that child over
// its first `index1` bits.
struct HostPairedRank {
  uint32_t bit1;    // bit of the super node at the position
  uint32_t index1;  // 1-based index in the chosen child (0: none of its bits precede)
  uint32_t ones2;   // ones among the child's first index1 bits (internal child only)
  uint32_t bit2;    // the child's bit at index1-1 (internal child, index1 > 0)
};
HostPairedRank host_paired_rank(const uint32_t* rank_words, int block_words, uint32_t base_block, uint32_t index1,
                                int follow);

// All four levels of a quad-level block: follows `path` (b1 b2 b3 b4 as a number) or, when
// path < 0, the bits found at the position.
struct HostQuadRank {
  uint32_t exit;    // the 4 path bits taken
  uint32_t index1;  // 1-based index in the node at that exit
};
HostQuadRank host_quad_rank(const uint32_t* rank_words, uint32_t base_block, uint32_t index1, int path);

}  // namespace fmb
