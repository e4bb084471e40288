// fm_format.hpp -- host-side reader of femto's on-disk index format (product code).
//
// The format is consumed verbatim; nothing here writes it.
// Reference definitions restated (paths relative to the reference tree):
//   block header (88 B)        src/main/index.c:817-868 (writer), :1348-1404 (reader)
//   header block tables        src/main/index.c:870-898
//   data block tables          src/main/index.c:900-910, 1073-1110
//   bucket layout              src/main/index.c:44-63, 490-726
//   map + Huffman bitstream    src/main/index.c:567-611 (writer), :1264-1314 (reader)
//   canonical code assignment  src/main/huffman.c:152-167, index.c:290-300
//   wavelet tree directory     src/main/wtree.c:907-1078, wtree_funcs.h:576-626
//   bseq                       src/main/wtree.c:510-591, wtree_funcs.h:294-511
//   containers                 src/main/block_storage.c:104-267, 464-588, index.c:2260-2365
#pragma once

#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/femto_b200.h"

namespace fmb {

constexpr uint32_t kMagicHeaderBlock = 0xb1177deaU;
constexpr uint32_t kMagicDataBlock = 0xb1501deaU;
constexpr uint32_t kMagicEndOfHeader = 0xe0ffff4dU;
constexpr uint32_t kMagicBucket = 0xb140bcc7U;
constexpr uint32_t kMagicFlattened = 0xb1497deaU;
constexpr uint32_t kFormatVersion = 6;
constexpr int kWtreeSettings = 31 + 0x1000 * 8;  // GROUP_SIZE + 0x1000*SEGMENT_WORDS
constexpr int kAlpha = FM_ALPHA_SIZE;
constexpr int kEscSeof = 2;
constexpr int kBlockHeaderBytes = 88;
constexpr int kSegsPerGroup = 31;
constexpr int kSegWords = 8;
constexpr int kSegBits = 512;
constexpr int kMaxCodeLen = 20;
constexpr uint32_t kEndOfBucketSym = 0x1ff;

// Shard (GPU) of data block b when an index of total_length rows is split over nshards GPUs by BWT row range
// at data-block granularity (the reference's partition unit, src/main/index.h:83-100): contiguous block
// ranges, a block going to the shard its MIDDLE row falls into when the rows are cut into nshards equal
// parts.  Balanced by rows: total_length is normally k * block_size + (a few), i.e. the last block is tiny,
// and counting blocks would give one GPU a whole block more than the others.
inline int shard_of_block(int64_t b, int64_t block_size, int64_t total_length, int nshards) {
  if (nshards <= 1 || total_length <= 0) return 0;
  const int64_t s = ((2 * b + 1) * block_size * nshards) / (2 * total_length);
  return static_cast<int>(s < nshards - 1 ? s : nshards - 1);
}

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

inline uint32_t be32(const uint8_t* p) {
  uint32_t v;
  std::memcpy(&v, p, 4);
  return __builtin_bswap32(v);
}
inline uint64_t be64(const uint8_t* p) {
  uint64_t v;
  std::memcpy(&v, p, 8);
  return __builtin_bswap64(v);
}
inline void put_be32(uint8_t* p, uint32_t v) {
  v = __builtin_bswap32(v);
  std::memcpy(p, &v, 4);
}
inline void put_be64(uint8_t* p, uint64_t v) {
  v = __builtin_bswap64(v);
  std::memcpy(p, &v, 8);
}
inline int num_bits64(uint64_t x) { return x ? 64 - __builtin_clzll(x) : 0; }

// A read-only mapping of one block's bytes.
class Blob {
 public:
  Blob() = default;
  Blob(const Blob&) = delete;
  Blob& operator=(const Blob&) = delete;
  Blob(Blob&& o) noexcept { *this = std::move(o); }
  Blob& operator=(Blob&& o) noexcept;
  ~Blob();
  static Blob map_file(const std::string& path, int64_t start, int64_t len);  // len<0: to EOF
  const uint8_t* data() const { return data_; }
  size_t size() const { return size_; }
  // bounds-checked pointer to [off, off+n)
  const uint8_t* at(size_t off, size_t n) const {
    if (off > size_ || n > size_ - off) throw Error(FM_ERR_FORMAT, "index block truncated");
    return data_ + off;
  }

 private:
  const uint8_t* data_ = nullptr;
  size_t size_ = 0;
  void* map_base_ = nullptr;
  size_t map_len_ = 0;
};

struct BlockHeader {
  uint32_t magic = 0, version = 0;
  int64_t block_number = 0, nblocks = 0, total_length = 0, ndocs = 0;
  int32_t num_buckets = 0, size = 0, var_block = 0, block_size = 0, bucket_size = 0, mark_period = 0,
          mark_type = 0, var_chunk = 0, chunk_size = 0, wtree_settings = 0, alpha_size = 0;
  uint32_t end_magic = 0;
};
BlockHeader parse_block_header(const Blob& b, uint32_t want_magic);

// What the reference's b_fault derives for a bucket.
struct BucketTables {
  int n_in_use = 0;
  uint8_t in_use[kAlpha] = {0};
  uint16_t seq_to_ch[kAlpha + 1] = {0};
  uint16_t ch_to_seq[kAlpha] = {0};
  uint8_t code_len[kAlpha + 1] = {0};
  uint32_t leaf[kAlpha + 1] = {0};  // canonical code | 1<<len; index n_in_use = end-of-bucket
  uint32_t off_bucket = 0, off_wtree = 0, off_marktab = 0, off_markarr = 0;  // absolute in block
  uint32_t off_end = 0;                                                      // end of this bucket
  int max_len = 0;
};

// An opened index: header block + data blocks (directory or flattened container).
class IndexFiles {
 public:
  static std::unique_ptr<IndexFiles> open(const std::string& path);
  const BlockHeader& header() const { return hdr_; }
  const Blob& header_blob() const { return header_; }
  int64_t nblocks() const { return hdr_.nblocks; }
  int buckets_per_block() const { return buckets_per_block_; }
  // Data blocks are mapped on demand so that a shard maps only what it loads.
  Blob map_block(int64_t b) const;
  bool flattened() const { return flattened_; }

  int64_t C(int ch) const;  // get_C: ch >= 261 -> total_length
  int64_t block_occs(int ch, int64_t blk) const;
  int64_t doc_end(int64_t doc) const;
  int64_t doc_eof_row(int64_t doc) const;
  // document info bytes (name / URL given at build time): document_info, index.c:1767-1784
  std::pair<const uint8_t*, int64_t> doc_info(int64_t doc) const;

 private:
  std::string path_;
  bool flattened_ = false;
  std::vector<int64_t> flat_offsets_;  // num_blocks+2 entries when flattened
  Blob header_;
  BlockHeader hdr_;
  int buckets_per_block_ = 0;
};

// Decode one bucket's tables; `bucket` is the index within the block.
void parse_bucket_tables(const Blob& blk, const BlockHeader& bh, int buckets_per_block, int bucket,
                         BucketTables* out);

// bucket_occs[ch][bucket] of a data block (char-major, stride = buckets present in the block).
inline uint32_t bucket_occs(const Blob& blk, const BlockHeader& bh, int buckets_per_block, int ch, int bucket) {
  size_t base = kBlockHeaderBytes + 4 * (size_t(buckets_per_block) + 1);
  return be32(blk.at(base + 4 * (size_t(ch) * size_t(bh.num_buckets) + size_t(bucket)), 4));
}

// ---- bseq decoding -------------------------------------------------------
struct BseqView {
  const uint8_t* z = nullptr;
  size_t avail = 0;  // bytes available from z to the end of the enclosing section
  int ngroups = 0;
  int total_words = 0;
  uint32_t d_off = 0;
  int64_t nsegs() const { return (int64_t(total_words) + kSegWords - 1) / kSegWords; }
};
BseqView open_bseq(const uint8_t* z, size_t avail);
// Total number of bits stored in the sequence (sum of all S entries); optionally the number of ones.
int64_t bseq_length(const BseqView& v, int64_t* ones_out = nullptr);
// Expand the whole sequence to `out` as MSB-first bits packed in 32-bit words:
// bit p is (out[p/32] >> (31 - p%32)) & 1.  out must hold ceil(nbits/32) zeroed words.
// Returns the number of bits written.
int64_t bseq_expand(const BseqView& v, uint32_t* out, int64_t out_bits_cap);

}  // namespace fmb
