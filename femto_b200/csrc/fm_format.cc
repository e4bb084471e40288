// fm_format.cc -- host-side reader of femto's on-disk index format (see fm_format.hpp).
#include "fm_format.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>

namespace fmb {

// ---------------------------------------------------------------- Blob
Blob& Blob::operator=(Blob&& o) noexcept {
  if (this != &o) {
    if (map_base_) munmap(map_base_, map_len_);
    data_ = o.data_; size_ = o.size_; map_base_ = o.map_base_; map_len_ = o.map_len_;
    o.data_ = nullptr; o.size_ = 0; o.map_base_ = nullptr; o.map_len_ = 0;
  }
  return *this;
}

Blob::~Blob() {
  if (map_base_) munmap(map_base_, map_len_);
}

Blob Blob::map_file(const std::string& path, int64_t start, int64_t len) {
  int fd = ::open(path.c_str(), O_RDONLY);
  if (fd < 0) throw Error(FM_ERR_IO, "cannot open " + path);
  struct stat st;
  if (fstat(fd, &st)) { ::close(fd); throw Error(FM_ERR_IO, "cannot stat " + path); }
  if (len < 0) len = st.st_size - start;
  if (start < 0 || len < 0 || start + len > st.st_size) {
    ::close(fd);
    throw Error(FM_ERR_FORMAT, "block region outside of file " + path);
  }
  Blob b;
  if (len == 0) { ::close(fd); return b; }
  const int64_t pg = sysconf(_SC_PAGESIZE);
  const int64_t astart = start - start % pg;
  const int64_t delta = start - astart;
  void* p = mmap(nullptr, size_t(len + delta), PROT_READ, MAP_SHARED, fd, astart);
  ::close(fd);
  if (p == MAP_FAILED) throw Error(FM_ERR_IO, "mmap failed for " + path);
  b.map_base_ = p;
  b.map_len_ = size_t(len + delta);
  b.data_ = static_cast<const uint8_t*>(p) + delta;
  b.size_ = size_t(len);
  return b;
}

// ---------------------------------------------------------------- block header
BlockHeader parse_block_header(const Blob& b, uint32_t want_magic) {
  BlockHeader h;
  const uint8_t* p = b.at(0, kBlockHeaderBytes);
  h.magic = be32(p);
  h.version = be32(p + 4);
  h.block_number = int64_t(be64(p + 8));
  h.nblocks = int64_t(be64(p + 16));
  h.total_length = int64_t(be64(p + 24));
  h.ndocs = int64_t(be64(p + 32));
  h.num_buckets = int32_t(be32(p + 40));
  h.size = int32_t(be32(p + 44));
  h.var_block = int32_t(be32(p + 48));
  h.block_size = int32_t(be32(p + 52));
  h.bucket_size = int32_t(be32(p + 56));
  h.mark_period = int32_t(be32(p + 60));
  h.mark_type = int32_t(be32(p + 64));
  h.var_chunk = int32_t(be32(p + 68));
  h.chunk_size = int32_t(be32(p + 72));
  h.wtree_settings = int32_t(be32(p + 76));
  h.alpha_size = int32_t(be32(p + 80));
  h.end_magic = be32(p + 84);
  if (h.magic != want_magic) throw Error(FM_ERR_FORMAT, "Invalid block start");
  if (h.version != kFormatVersion) throw Error(FM_ERR_FORMAT, "Wrong block version");
  if (h.block_size <= 0 || h.bucket_size <= 0) throw Error(FM_ERR_PARAM, "bad block/bucket size");
  if (h.block_size % h.bucket_size != 0) throw Error(FM_ERR_PARAM, "block_size not a multiple of bucket_size");
  if (h.chunk_size > 0 && h.bucket_size % h.chunk_size != 0)
    throw Error(FM_ERR_PARAM, "bucket_size not a multiple of chunk_size");
  if (h.wtree_settings != kWtreeSettings)
    throw Error(FM_ERR_FORMAT, "Wrong wavelet tree settings");
  if (h.alpha_size != kAlpha) throw Error(FM_ERR_FORMAT, "Wrong alphabet size");
  if (h.end_magic != kMagicEndOfHeader) throw Error(FM_ERR_FORMAT, "Bad end of header");
  if (h.var_block != 0) throw Error(FM_ERR_FORMAT, "variable block size is not supported");
  return h;
}

// ---------------------------------------------------------------- IndexFiles
static std::string block_file_name(const std::string& dir, int64_t block_id) {
  char buf[64];
  std::snprintf(buf, sizeof(buf), "/%02llx", static_cast<unsigned long long>(block_id));
  return dir + buf;
}

std::unique_ptr<IndexFiles> IndexFiles::open(const std::string& path) {
  struct stat st;
  if (stat(path.c_str(), &st)) throw Error(FM_ERR_IO, "stat failed: " + path);
  std::unique_ptr<IndexFiles> ix(new IndexFiles());
  ix->path_ = path;
  if (S_ISDIR(st.st_mode)) ix->flattened_ = false;
  else if (S_ISREG(st.st_mode)) ix->flattened_ = true;
  else throw Error(FM_ERR_IO, "index not file or directory: " + path);

  if (ix->flattened_) {
    Blob head = Blob::map_file(path, 0, 16);
    if (be32(head.at(0, 4)) != kMagicFlattened || be32(head.at(4, 4)) != kFormatVersion)
      throw Error(FM_ERR_FORMAT, "not a flattened femto index: " + path);
    const int64_t nb = int64_t(be64(head.at(8, 8)));  // includes the header block
    if (nb <= 0) throw Error(FM_ERR_FORMAT, "flattened index has no blocks");
    Blob table = Blob::map_file(path, 0, 16 + 8 * (nb + 1));
    ix->flat_offsets_.resize(size_t(nb + 1));
    for (int64_t i = 0; i <= nb; i++) ix->flat_offsets_[size_t(i)] = int64_t(be64(table.at(size_t(16 + 8 * i), 8)));
    ix->header_ = Blob::map_file(path, ix->flat_offsets_[0], ix->flat_offsets_[1] - ix->flat_offsets_[0]);
  } else {
    ix->header_ = Blob::map_file(block_file_name(path, 0), 0, -1);
  }
  ix->hdr_ = parse_block_header(ix->header_, kMagicHeaderBlock);
  const BlockHeader& h = ix->hdr_;
  if (h.nblocks < 0 || h.ndocs < 0 || h.total_length < 0) throw Error(FM_ERR_FORMAT, "negative header field");
  if (ix->flattened_ && int64_t(ix->flat_offsets_.size()) != h.nblocks + 2)
    throw Error(FM_ERR_FORMAT, "flattened block table does not match header");
  ix->buckets_per_block_ = (h.block_size + h.bucket_size - 1) / h.bucket_size;
  const size_t need = size_t(kBlockHeaderBytes) + 8 * size_t(kAlpha) + 8 * size_t(kAlpha) * size_t(h.nblocks) +
                      16 * size_t(h.ndocs);
  ix->header_.at(0, need);
  return ix;
}

Blob IndexFiles::map_block(int64_t b) const {
  if (b < 0 || b >= hdr_.nblocks) throw Error(FM_ERR_PARAM, "no such data block");
  if (flattened_)
    return Blob::map_file(path_, flat_offsets_[size_t(b + 1)], flat_offsets_[size_t(b + 2)] - flat_offsets_[size_t(b + 1)]);
  return Blob::map_file(block_file_name(path_, b + 1), 0, -1);
}

int64_t IndexFiles::C(int ch) const {
  if (ch >= kAlpha) return hdr_.total_length;
  return int64_t(be64(header_.at(size_t(kBlockHeaderBytes) + 8 * size_t(ch), 8)));
}

int64_t IndexFiles::block_occs(int ch, int64_t blk) const {
  const size_t off = size_t(kBlockHeaderBytes) + 8 * size_t(kAlpha) +
                     8 * (size_t(ch) * size_t(hdr_.nblocks) + size_t(blk));
  return int64_t(be64(header_.at(off, 8)));
}

int64_t IndexFiles::doc_end(int64_t doc) const {
  const size_t off = size_t(kBlockHeaderBytes) + 8 * size_t(kAlpha) + 8 * size_t(kAlpha) * size_t(hdr_.nblocks) +
                     8 * size_t(doc);
  return int64_t(be64(header_.at(off, 8)));
}

int64_t IndexFiles::doc_eof_row(int64_t doc) const {
  const size_t off = size_t(kBlockHeaderBytes) + 8 * size_t(kAlpha) + 8 * size_t(kAlpha) * size_t(hdr_.nblocks) +
                     8 * size_t(hdr_.ndocs) + 8 * size_t(doc);
  return int64_t(be64(header_.at(off, 8)));
}

std::pair<const uint8_t*, int64_t> IndexFiles::doc_info(int64_t doc) const {
  if (doc < 0 || doc >= hdr_.ndocs) throw Error(FM_ERR_PARAM, "no such document");
  // doc_info_off[ndocs+1] (absolute offsets in the header block) follows doc_eof_rows (index.c:882-898)
  const size_t tab = size_t(kBlockHeaderBytes) + 8 * size_t(kAlpha) + 8 * size_t(kAlpha) * size_t(hdr_.nblocks) +
                     16 * size_t(hdr_.ndocs) + 8 * size_t(doc);
  const int64_t start = int64_t(be64(header_.at(tab, 8))), end = int64_t(be64(header_.at(tab + 8, 8)));
  if (start < 0 || end < start) throw Error(FM_ERR_FORMAT, "bad document info offsets");
  if (end == start) return {nullptr, 0};
  return {header_.at(size_t(start), size_t(end - start)), end - start};
}

// ---------------------------------------------------------------- bucket tables
namespace {
struct MsbBitReader {
  const Blob& b;
  size_t bit;
  unsigned get(int n) {
    unsigned v = 0;
    while (n-- > 0) {
      const uint8_t byte = *b.at(bit >> 3, 1);
      v = (v << 1) | ((byte >> (7 - (bit & 7))) & 1u);
      bit++;
    }
    return v;
  }
};
}  // namespace

void parse_bucket_tables(const Blob& blk, const BlockHeader& bh, int buckets_per_block, int bucket,
                         BucketTables* t) {
  if (bucket < 0 || bucket >= bh.num_buckets || bucket >= buckets_per_block)
    throw Error(FM_ERR_PARAM, "no such bucket");
  *t = BucketTables();
  const uint32_t boff = be32(blk.at(size_t(kBlockHeaderBytes) + 4 * size_t(bucket), 4));
  const uint32_t bend = be32(blk.at(size_t(kBlockHeaderBytes) + 4 * size_t(bucket + 1), 4));
  const uint8_t* bh6 = blk.at(boff, 24);
  if (be32(bh6) != kMagicBucket) throw Error(FM_ERR_FORMAT, "bad bucket start");
  t->off_bucket = boff;
  t->off_end = bend;
  const uint32_t map_off = boff + be32(bh6 + 4);
  t->off_wtree = boff + be32(bh6 + 8);
  t->off_marktab = boff + be32(bh6 + 12);
  t->off_markarr = boff + be32(bh6 + 16);
  if ((t->off_bucket | t->off_wtree | t->off_marktab | t->off_markarr) & 7u)
    throw Error(FM_ERR_FORMAT, "misaligned bucket section");
  if (!(t->off_wtree <= t->off_marktab && t->off_marktab <= t->off_markarr && t->off_markarr <= blk.size()))
    throw Error(FM_ERR_FORMAT, "bucket sections out of order");

  MsbBitReader r{blk, size_t(map_off) * 8};
  constexpr int kGroups = (kAlpha + 15) / 16;
  bool group_used[kGroups];
  for (int i = 0; i < kGroups; i++) group_used[i] = r.get(1) != 0;
  for (int i = 0; i < kGroups; i++) {
    if (!group_used[i]) continue;
    for (int j = 0; j < 16; j++) {
      const unsigned bit = r.get(1);
      if (bit && i * 16 + j < kAlpha) t->in_use[i * 16 + j] = 1;
    }
  }
  int n = 0;
  for (int c = 0; c < kAlpha; c++) {
    if (t->in_use[c]) {
      t->seq_to_ch[n] = uint16_t(c);
      t->ch_to_seq[c] = uint16_t(n);
      n++;
    } else {
      t->ch_to_seq[c] = 0xffff;
    }
  }
  t->n_in_use = n;
  t->seq_to_ch[n] = uint16_t(kEndOfBucketSym);
  const int alpha = n + 1;  // the end-of-bucket symbol is coded too (index.c:404-406)

  int curr = int(r.get(5));
  int minl = 32, maxl = 0;
  for (int i = 0; i < alpha; i++) {
    for (;;) {
      if (curr < 1 || curr > kMaxCodeLen) throw Error(FM_ERR_BZ_DATA, "bad Huffman code length");
      if (r.get(1) == 0) break;
      if (r.get(1) == 0) curr++; else curr--;
    }
    t->code_len[i] = uint8_t(curr);
    minl = std::min(minl, curr);
    maxl = std::max(maxl, curr);
  }
  uint32_t vec = 0;
  for (int len = minl; len <= maxl; len++) {
    for (int i = 0; i < alpha; i++)
      if (t->code_len[i] == len) t->leaf[i] = (vec++) | (1u << len);
    vec <<= 1;
  }
  t->max_len = maxl;
}

// ---------------------------------------------------------------- bseq
BseqView open_bseq(const uint8_t* z, size_t avail) {
  BseqView v;
  if (avail < 16) throw Error(FM_ERR_FORMAT, "bseq truncated");
  v.z = z;
  v.avail = avail;
  if (be32(z) != 0) throw Error(FM_ERR_FORMAT, "bseq header word is not zero");
  v.ngroups = int(be32(z + 4));
  v.total_words = int(be32(z + 8));
  v.d_off = be32(z + 12);
  if (v.ngroups < 0 || v.total_words < 0) throw Error(FM_ERR_FORMAT, "bseq header negative");
  const size_t s_off = 16 + 12 * size_t(v.ngroups);
  if (s_off > avail || v.d_off > avail || size_t(v.d_off) + 8 * size_t(v.total_words) > avail || (v.d_off & 7u))
    throw Error(FM_ERR_FORMAT, "bseq sections out of bounds");
  if (v.nsegs() > int64_t(kSegsPerGroup) * v.ngroups || (v.ngroups > 0 && v.nsegs() <= int64_t(kSegsPerGroup) * (v.ngroups - 1)))
    throw Error(FM_ERR_FORMAT, "bseq group count does not match segment count");
  return v;
}

namespace {
// varbyte: 7 bits per byte LSB-first, the LAST byte has 0x80 set (wtree_funcs.h:437-479)
inline unsigned read_varbyte(const uint8_t*& p, const uint8_t* end) {
  unsigned v = 0;
  int sh = 0;
  for (;;) {
    if (p >= end || sh > 28) throw Error(FM_ERR_FORMAT, "bad varbyte in bseq S section");
    const uint8_t c = *p++;
    v |= unsigned(c & 0x7f) << sh;
    sh += 7;
    if (c & 0x80) return v;
  }
}

inline void set_ones(uint32_t* w, int64_t pos, int64_t len) {
  while (len > 0) {
    const int o = int(pos & 31);
    const int take = int(std::min<int64_t>(32 - o, len));
    const uint32_t mask = (take == 32) ? 0xffffffffu : (((1u << take) - 1u) << (32 - o - take));
    w[pos >> 5] |= mask;
    pos += take;
    len -= take;
  }
}

// append the top n (<=64) bits of v
inline void put_bits(uint32_t* w, int64_t pos, uint64_t v, int n) {
  while (n > 0) {
    const int o = int(pos & 31);
    const int take = std::min(32 - o, n);
    const uint32_t chunk = uint32_t(v >> (64 - take));
    w[pos >> 5] |= chunk << (32 - o - take);
    v = (take == 64) ? 0 : (v << take);
    pos += take;
    n -= take;
  }
}
}  // namespace

int64_t bseq_length(const BseqView& v, int64_t* ones_out) {
  const uint8_t* p = v.z + 16 + 12 * size_t(v.ngroups);
  const uint8_t* end = v.z + v.d_off;
  int64_t zeros = 0, ones = 0;
  const int64_t ns = v.nsegs();
  for (int64_t s = 0; s < ns; s++) {
    zeros += read_varbyte(p, end);
    ones += read_varbyte(p, end);
  }
  if (ones_out) *ones_out = ones;
  return zeros + ones;
}

int64_t bseq_expand(const BseqView& v, uint32_t* out, int64_t cap_bits) {
  const uint8_t* p = v.z + 16 + 12 * size_t(v.ngroups);
  const uint8_t* end = v.z + v.d_off;
  const uint8_t* D = v.z + v.d_off;
  int64_t pos = 0;
  const int64_t ns = v.nsegs();
  for (int64_t s = 0; s < ns; s++) {
    const unsigned s0 = read_varbyte(p, end);
    const unsigned s1 = read_varbyte(p, end);
    const int64_t n = int64_t(s0) + int64_t(s1);
    if (pos + n > cap_bits) throw Error(FM_ERR_FORMAT, "bseq longer than its S sums");
    uint64_t w[kSegWords];
    const int64_t first_word = s * kSegWords;
    int nw = int(std::min<int64_t>(kSegWords, v.total_words - first_word));
    for (int i = 0; i < nw; i++) w[i] = be64(D + 8 * size_t(first_word + i));
    for (int i = nw; i < kSegWords; i++) w[i] = 0;
    if (w[0] >> 63) {
      // RLE-gamma segment: bit 1 = value of the first run; codes = k zeros then a (k+1)-bit value
      int bit = int((w[0] >> 62) & 1);
      int bp = 2;
      int64_t done = 0;
      unsigned ones = 0;
      while (done < n) {
        uint64_t cur = 0;
        const int wi = bp >> 6, bo = bp & 63;
        if (wi < kSegWords) cur = w[wi] << bo;
        if (bo && wi + 1 < kSegWords) cur |= w[wi + 1] >> (64 - bo);
        if (cur == 0) throw Error(FM_ERR_FORMAT, "RLE segment ends before its S sums");
        const int k = 2 * __builtin_clzll(cur) + 1;
        if (k > 63) throw Error(FM_ERR_FORMAT, "bad gamma code");
        int64_t run = int64_t(cur >> (64 - k));
        bp += k;
        if (run > n - done) throw Error(FM_ERR_FORMAT, "RLE runs exceed the segment's S sums");
        if (bit) { set_ones(out, pos + done, run); ones += unsigned(run); }
        done += run;
        bit ^= 1;
      }
      if (ones != s1) throw Error(FM_ERR_FORMAT, "RLE segment ones do not match S");
    } else {
      // raw segment: payload bits follow the type bit
      if (n > kSegBits - 1) throw Error(FM_ERR_FORMAT, "raw segment too long");
      int bp = 1;
      int64_t left = n, q = pos;
      while (left > 0) {
        const int wi = bp >> 6, bo = bp & 63;
        const int take = int(std::min<int64_t>(64 - bo, left));
        put_bits(out, q, w[wi] << bo, take);
        bp += take;
        q += take;
        left -= take;
      }
    }
    pos += n;
  }
  return pos;
}

}  // namespace fmb
