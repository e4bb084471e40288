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
  uint32_t* rank_words = nullptr;   // n_rank_blocks * 32 words (calloc'ed)
  int64_t n_rank_blocks = 0;
  int64_t n_wtree_blocks = 0;       // of which wavelet-tree payload (the rest are mark bit-vectors)
  std::vector<NodeRec> nodes;
  std::vector<OccRec> occ;
  std::vector<MarkRec> mark;
  std::vector<BucketRec> buckets;
  std::vector<int64_t> markvals;
  std::vector<int64_t> C;           // 262 entries
  std::vector<int64_t> doc_ends, doc_eof_rows;
  HostImage() = default;
  HostImage(const HostImage&) = delete;
  HostImage& operator=(const HostImage&) = delete;
  ~HostImage();
};

// shard/nshards select data blocks b with b*nshards/nblocks == shard (all blocks when nshards==1).
std::unique_ptr<HostImage> build_host_image(const std::string& path, int shard, int nshards, int nthreads);

// Host-side rank over the image (used by the loader's self-check and by unit tests of the
// image layout; NOT a query fallback -- the C ABI never calls it).
struct HostRank { uint32_t ones; uint32_t bit; };
HostRank host_rank(const uint32_t* rank_words, uint32_t base_block, uint32_t index1);

}  // namespace fmb
