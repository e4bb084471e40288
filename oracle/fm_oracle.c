/* oracle/fm_oracle.c -- TEST INFRASTRUCTURE ONLY (see fm_oracle.h for the parity status).
 *
 * Plain-C restatement of the reference's FM-index read path, written from the on-disk
 * format and the reference's semantics; each function cites the reference code it
 * restates (paths relative to /root/reference).  Single-threaded scalar code: this is
 * the checker for the CUDA path and the "port" CPU baseline, never the product.
 */
#define _GNU_SOURCE
#include "fm_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <fcntl.h>
#include <unistd.h>

/* ---- format constants (src/main/index.h:182-188, index_types.h:35-65,
 *      block_storage.h:125-126, wtree.c:48-53, wtree_funcs.h:33-35) ---- */
#define MAGIC_HEADER_BLOCK 0xb1177deaU
#define MAGIC_DATA_BLOCK   0xb1501deaU
#define MAGIC_END_OF_HDR   0xe0ffff4dU
#define MAGIC_BUCKET       0xb140bcc7U
#define MAGIC_FLATTENED    0xb1497deaU
#define FORMAT_VERSION     6U
#define WTREE_SETTINGS     0x801f       /* GROUP_SIZE 31 + 0x1000 * SEGMENT_WORDS 8 */
#define ALPHA              261          /* 5 escape codes + 256 bytes */
#define ESC_SEOF           2
#define HDR_BYTES          88
#define SEGS_PER_GROUP     31
#define SEG_WORDS          8
#define MAX_CODE_LEN       20

static inline uint32_t be32(const uint8_t* p)
{
  return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
}
static inline uint64_t be64(const uint8_t* p)
{
  return ((uint64_t)be32(p) << 32) | be32(p + 4);
}

typedef struct {
  const uint8_t* data;
  size_t len;
  void* map_base;      /* what to munmap */
  size_t map_len;
} blob_t;

/* what b_fault (src/main/index.c:1222-1342) derives once per bucket */
typedef struct {
  uint8_t ready;
  uint8_t in_use[ALPHA];
  int n_in_use;
  uint16_t seq_to_ch[ALPHA + 1];
  uint16_t ch_to_seq[ALPHA];
  uint8_t code_len[ALPHA + 1];
  uint32_t leaf[ALPHA + 1];      /* canonical code with a leading 1 (wavelet-tree leaf id) */
  uint32_t off_bucket, off_wtree, off_marktab, off_markarr; /* absolute in the block */
} bucket_tab_t;

typedef struct {
  blob_t blob;
  int64_t number;
  int32_t num_buckets;   /* buckets present in this block (hdr@40) */
  int32_t size;          /* rows in this block (hdr@44) */
  bucket_tab_t* tabs;    /* [num_buckets], filled lazily */
} dblock_t;

struct fmo_index {
  blob_t header;
  int64_t nblocks, total_length, ndocs;
  int32_t block_size, bucket_size, mark_period, chunk_size;
  int32_t buckets_per_block;
  int text_size_bits;
  dblock_t* blocks;
  /* instrumentation */
  int64_t ctr_bytes, ctr_occ, ctr_levels;
};

/* ------------------------------------------------------------------------- */
/* container: directory of %02x files, or flattened single file
 * (src/main/block_storage.c:104-240, 464-588; src/main/index.h:83-100)        */

static int map_file_region(const char* path, int64_t start, int64_t len, blob_t* out)
{
  int fd = open(path, O_RDONLY);
  long pg = sysconf(_SC_PAGESIZE);
  int64_t astart, delta;
  void* p;
  if (fd < 0) return FMO_ERR_IO;
  if (len < 0) {
    struct stat st;
    if (fstat(fd, &st)) { close(fd); return FMO_ERR_IO; }
    len = st.st_size - start;
  }
  astart = start - (start % pg);
  delta = start - astart;
  p = mmap(NULL, (size_t)(len + delta), PROT_READ, MAP_SHARED, fd, astart);
  close(fd);
  if (p == MAP_FAILED) return FMO_ERR_IO;
  out->map_base = p;
  out->map_len = (size_t)(len + delta);
  out->data = (const uint8_t*)p + delta;
  out->len = (size_t)len;
  return FMO_OK;
}

static void unmap_blob(blob_t* b)
{
  if (b->map_base) munmap(b->map_base, b->map_len);
  memset(b, 0, sizeof(*b));
}

/* the 88-byte block header (src/main/index.c:817-868 writer, :1348-1404 reader) */
typedef struct {
  uint32_t magic, version;
  int64_t block_number, nblocks, total_length, ndocs;
  int32_t num_buckets, size, var_block, block_size, bucket_size, mark_period, mark_type,
          var_chunk, chunk_size, wtree_settings, alpha_size;
  uint32_t end_magic;
} hdr88_t;

static int parse_hdr88(const blob_t* b, uint32_t want_magic, hdr88_t* h)
{
  const uint8_t* p = b->data;
  if (b->len < HDR_BYTES) return FMO_ERR_FORMAT;
  h->magic = be32(p);           h->version = be32(p + 4);
  h->block_number = (int64_t)be64(p + 8);
  h->nblocks = (int64_t)be64(p + 16);
  h->total_length = (int64_t)be64(p + 24);
  h->ndocs = (int64_t)be64(p + 32);
  h->num_buckets = (int32_t)be32(p + 40);  h->size = (int32_t)be32(p + 44);
  h->var_block = (int32_t)be32(p + 48);    h->block_size = (int32_t)be32(p + 52);
  h->bucket_size = (int32_t)be32(p + 56);  h->mark_period = (int32_t)be32(p + 60);
  h->mark_type = (int32_t)be32(p + 64);    h->var_chunk = (int32_t)be32(p + 68);
  h->chunk_size = (int32_t)be32(p + 72);   h->wtree_settings = (int32_t)be32(p + 76);
  h->alpha_size = (int32_t)be32(p + 80);   h->end_magic = be32(p + 84);
  if (h->magic != want_magic) return FMO_ERR_FORMAT;
  if (h->version != FORMAT_VERSION) return FMO_ERR_FORMAT;
  if (h->block_size <= 0 || h->bucket_size <= 0) return FMO_ERR_PARAM;
  if (h->block_size % h->bucket_size) return FMO_ERR_PARAM;        /* calculate_params, index.c:793-815 */
  if (h->chunk_size > 0 && h->bucket_size % h->chunk_size) return FMO_ERR_PARAM;
  if (h->wtree_settings != WTREE_SETTINGS) return FMO_ERR_FORMAT;
  if (h->alpha_size != ALPHA) return FMO_ERR_FORMAT;
  if (h->end_magic != MAGIC_END_OF_HDR) return FMO_ERR_FORMAT;
  return FMO_OK;
}

static int num_bits64(uint64_t x)   /* floor(log2 x) + 1, num_bits64(0)=0 (src/utils/bit_funcs.h) */
{
  int n = 0;
  while (x) { n++; x >>= 1; }
  return n;
}

fmo_index* fmo_open(const char* path, int* err_out)
{
  struct stat st;
  fmo_index* ix = NULL;
  int err = FMO_OK;
  int flattened;
  blob_t table;
  hdr88_t h;
  int64_t i;
  char fname[4096];

  memset(&table, 0, sizeof(table));
  if (stat(path, &st)) { err = FMO_ERR_IO; goto fail; }
  if (S_ISDIR(st.st_mode)) flattened = 0;
  else if (S_ISREG(st.st_mode)) flattened = 1;
  else { err = FMO_ERR_IO; goto fail; }

  ix = calloc(1, sizeof(*ix));
  if (!ix) { err = FMO_ERR_MEM; goto fail; }

  if (flattened) {
    int64_t nb;
    if (st.st_size < 16) { err = FMO_ERR_FORMAT; goto fail; }
    err = map_file_region(path, 0, 16, &table);
    if (err) goto fail;
    if (be32(table.data) != MAGIC_FLATTENED || be32(table.data + 4) != FORMAT_VERSION) {
      err = FMO_ERR_FORMAT; goto fail;
    }
    nb = (int64_t)be64(table.data + 8);      /* includes the header block */
    if (nb <= 0) { err = FMO_ERR_FORMAT; goto fail; }
    unmap_blob(&table);
    err = map_file_region(path, 0, 16 + 8 * (nb + 1), &table);
    if (err) goto fail;
    {
      int64_t s = (int64_t)be64(table.data + 16), e = (int64_t)be64(table.data + 24);
      err = map_file_region(path, s, e - s, &ix->header);
      if (err) goto fail;
    }
  } else {
    snprintf(fname, sizeof(fname), "%s/%02x", path, 0);
    err = map_file_region(fname, 0, -1, &ix->header);
    if (err) goto fail;
  }

  err = parse_hdr88(&ix->header, MAGIC_HEADER_BLOCK, &h);
  if (err) goto fail;
  ix->nblocks = h.nblocks;
  ix->total_length = h.total_length;
  ix->ndocs = h.ndocs;
  ix->block_size = h.block_size;
  ix->bucket_size = h.bucket_size;
  ix->mark_period = h.mark_period;
  ix->chunk_size = h.chunk_size;
  ix->buckets_per_block = (h.block_size + h.bucket_size - 1) / h.bucket_size;
  ix->text_size_bits = num_bits64((uint64_t)h.total_length);   /* index.c:1445 */
  if (ix->nblocks < 0 || (size_t)(HDR_BYTES + 8 * ALPHA + 8 * ALPHA * ix->nblocks + 16 * ix->ndocs) > ix->header.len) {
    err = FMO_ERR_FORMAT; goto fail;
  }

  ix->blocks = calloc((size_t)(ix->nblocks > 0 ? ix->nblocks : 1), sizeof(dblock_t));
  if (!ix->blocks) { err = FMO_ERR_MEM; goto fail; }
  for (i = 0; i < ix->nblocks; i++) {
    dblock_t* b = &ix->blocks[i];
    hdr88_t bh;
    if (flattened) {
      int64_t s = (int64_t)be64(table.data + 16 + 8 * (i + 1));
      int64_t e = (int64_t)be64(table.data + 16 + 8 * (i + 2));
      err = map_file_region(path, s, e - s, &b->blob);
    } else {
      snprintf(fname, sizeof(fname), "%s/%02llx", path, (unsigned long long)(i + 1));
      err = map_file_region(fname, 0, -1, &b->blob);
    }
    if (err) goto fail;
    err = parse_hdr88(&b->blob, MAGIC_DATA_BLOCK, &bh);
    if (err) goto fail;
    b->number = bh.block_number;
    b->num_buckets = bh.num_buckets;
    b->size = bh.size;
    b->tabs = calloc((size_t)(bh.num_buckets > 0 ? bh.num_buckets : 1), sizeof(bucket_tab_t));
    if (!b->tabs) { err = FMO_ERR_MEM; goto fail; }
  }
  unmap_blob(&table);
  if (err_out) *err_out = FMO_OK;
  return ix;

fail:
  unmap_blob(&table);
  if (ix) fmo_close(ix);
  if (err_out) *err_out = err;
  return NULL;
}

void fmo_close(fmo_index* ix)
{
  int64_t i;
  if (!ix) return;
  if (ix->blocks) {
    for (i = 0; i < ix->nblocks; i++) {
      free(ix->blocks[i].tabs);
      unmap_blob(&ix->blocks[i].blob);
    }
    free(ix->blocks);
  }
  unmap_blob(&ix->header);
  free(ix);
}

int fmo_header_info(const fmo_index* ix, int64_t* info)
{
  info[0] = ix->nblocks;      info[1] = ix->total_length; info[2] = ix->ndocs;
  info[3] = ix->block_size;   info[4] = ix->bucket_size;  info[5] = ix->mark_period;
  info[6] = ix->chunk_size;
  return FMO_OK;
}

/* ------------------------------------------------------------------------- */
/* header block tables: C at byte 88, block_occs char-major right after, then doc_ends,
 * doc_eof_rows (src/main/index.c:870-898, 1538-1569)                            */

static int64_t hdr_C(const fmo_index* ix, int ch)           /* get_C, index.c:1538-1554 */
{
  if (ch >= ALPHA) return ix->total_length;
  return (int64_t)be64(ix->header.data + HDR_BYTES + 8 * (size_t)ch);
}

static int64_t hdr_block_occs(const fmo_index* ix, int ch, int64_t blk) /* index.c:1556-1569 */
{
  size_t off = HDR_BYTES + 8 * ALPHA + 8 * ((size_t)ch * (size_t)ix->nblocks + (size_t)blk);
  return (int64_t)be64(ix->header.data + off);
}

static const uint8_t* hdr_doc_ends(const fmo_index* ix)
{
  return ix->header.data + HDR_BYTES + 8 * ALPHA + 8 * (size_t)ALPHA * (size_t)ix->nblocks;
}

int fmo_C(const fmo_index* ix, int ch, int64_t* out)
{
  if (ch < 0) return FMO_ERR_PARAM;
  *out = hdr_C(ix, ch);
  return FMO_OK;
}

int fmo_doc_info(const fmo_index* ix, int64_t doc, int64_t* doc_len, int64_t* eof_row)
{
  const uint8_t* ends = hdr_doc_ends(ix);          /* document_length, index.c:1668-1683 */
  const uint8_t* eofs = ends + 8 * (size_t)ix->ndocs;
  int64_t e;
  if (doc < 0 || doc >= ix->ndocs) return FMO_ERR_PARAM;
  e = (int64_t)be64(ends + 8 * doc);
  *doc_len = doc == 0 ? e : e - (int64_t)be64(ends + 8 * (doc - 1));
  *eof_row = (int64_t)be64(eofs + 8 * doc);        /* document_eof_row, index.c:1685-1696 */
  return FMO_OK;
}

/* document_info (index.c:1767-1784): after doc_eof_rows the header block holds ndocs + 1 offsets (from the
 * start of the block) delimiting every document's info bytes (written at index.c:882-898) */
int fmo_doc_name(const fmo_index* ix, int64_t doc, const unsigned char** info, int64_t* len)
{
  const uint8_t* dir = hdr_doc_ends(ix) + 16 * (size_t)ix->ndocs;
  int64_t start, end;
  if (doc < 0 || doc >= ix->ndocs) return FMO_ERR_PARAM;
  start = (int64_t)be64(dir + 8 * doc);
  end = (int64_t)be64(dir + 8 * (doc + 1));
  if (start < 0 || end < start || (size_t)end > ix->header.len) return FMO_ERR_FORMAT;
  *info = ix->header.data + start;
  *len = end - start;
  return FMO_OK;
}

/* resolve_location (index.c:1587-1611) over bsearch_int64_ntoh_arr (src/utils/util.c:346):
 * prev = last i with doc_ends[i] <= offset, or -1 */
int fmo_resolve(const fmo_index* ix, int64_t offset, int64_t* doc, int64_t* doc_off)
{
  const uint8_t* ends = hdr_doc_ends(ix);
  int64_t lo = -1, hi = ix->ndocs;   /* invariant: ends[lo] <= offset < ends[hi] */
  while (hi - lo > 1) {
    int64_t mid = lo + (hi - lo) / 2;
    if ((int64_t)be64(ends + 8 * mid) <= offset) lo = mid; else hi = mid;
  }
  if (lo < 0) { *doc = 0; *doc_off = offset; }
  else { *doc = lo + 1; *doc_off = offset - (int64_t)be64(ends + 8 * lo); }
  return FMO_OK;
}

/* ------------------------------------------------------------------------- */
/* MSB-first bit reader (bsR24, src/utils/buffer_funcs.h:136-152) */
typedef struct { const uint8_t* p; size_t bit; } bitrd_t;
static inline unsigned rd_bits(bitrd_t* r, int n)
{
  unsigned v = 0;
  while (n-- > 0) {
    v = (v << 1) | ((r->p[r->bit >> 3] >> (7 - (r->bit & 7))) & 1u);
    r->bit++;
  }
  return v;
}

/* b_fault: per-bucket map + Huffman lengths -> leaf codes (index.c:1222-1342;
 * canonical assignment BZ2_hbAssignCodes huffman.c:152-167; leaf = code | 1<<len, index.c:290-300) */
static int bucket_tables(const fmo_index* ix, dblock_t* b, int bucket, bucket_tab_t** out)
{
  bucket_tab_t* t;
  const uint8_t* blk = b->blob.data;
  uint32_t boff, map_off;
  bitrd_t r;
  uint8_t in16[(ALPHA + 15) / 16];
  int i, j, n, curr, minl = 32, maxl = 0, alpha;
  uint32_t vec;
  (void)ix;
  if (bucket < 0 || bucket >= b->num_buckets) return FMO_ERR_PARAM;
  t = &b->tabs[bucket];
  *out = t;
  if (t->ready) return FMO_OK;

  boff = be32(blk + HDR_BYTES + 4 * (size_t)bucket);
  if (boff + 24 > b->blob.len) return FMO_ERR_FORMAT;
  if (be32(blk + boff) != MAGIC_BUCKET) return FMO_ERR_FORMAT;
  t->off_bucket = boff;
  map_off = boff + be32(blk + boff + 4);
  t->off_wtree = boff + be32(blk + boff + 8);
  t->off_marktab = boff + be32(blk + boff + 12);
  t->off_markarr = boff + be32(blk + boff + 16);
  if ((t->off_bucket | t->off_wtree | t->off_marktab | t->off_markarr) & 7) return FMO_ERR_FORMAT;

  r.p = blk + map_off; r.bit = 0;
  for (i = 0; i < (ALPHA + 15) / 16; i++) in16[i] = (uint8_t)rd_bits(&r, 1);
  memset(t->in_use, 0, sizeof(t->in_use));
  for (i = 0; i < (ALPHA + 15) / 16; i++) {
    if (!in16[i]) continue;
    for (j = 0; j < 16; j++) {
      unsigned bit = rd_bits(&r, 1);
      if (bit && i * 16 + j < ALPHA) t->in_use[i * 16 + j] = 1;
    }
  }
  n = 0;
  for (i = 0; i < ALPHA; i++) {
    if (t->in_use[i]) { t->seq_to_ch[n] = (uint16_t)i; t->ch_to_seq[i] = (uint16_t)n; n++; }
    else t->ch_to_seq[i] = 0xffff;
  }
  t->n_in_use = n;
  alpha = n + 1;                      /* + end-of-bucket symbol */
  t->seq_to_ch[n] = 0xffff;

  curr = (int)rd_bits(&r, 5);
  for (i = 0; i < alpha; i++) {
    for (;;) {
      if (curr < 1 || curr > MAX_CODE_LEN) return FMO_ERR_BZ_DATA;
      if (rd_bits(&r, 1) == 0) break;
      if (rd_bits(&r, 1) == 0) curr++; else curr--;
    }
    t->code_len[i] = (uint8_t)curr;
    if (curr > maxl) maxl = curr;
    if (curr < minl) minl = curr;
  }
  vec = 0;
  for (j = minl; j <= maxl; j++) {
    for (i = 0; i < alpha; i++)
      if (t->code_len[i] == j) { t->leaf[i] = vec | (1u << j); vec++; }
    vec <<= 1;
  }
  t->ready = 1;
  return FMO_OK;
}

/* ------------------------------------------------------------------------- */
/* bseq_rank (src/main/wtree.c:635-763) with its helpers bsearch_A0A1 (:609-629),
 * decode_varbyte / bseq_segment (wtree_funcs.h:457-511), decode_gamma (:59-74).   */

typedef struct { int occ0, occ1, bit; int64_t bytes; } rank_out_t;

static void bseq_rank_impl(const uint8_t* z, unsigned index1, rank_out_t* o)
{
  unsigned x = index1 - 1;                       /* 0-based position */
  int ng = (int)be32(z + 4);
  int total_words = (int)be32(z + 8);
  uint32_t d_off = be32(z + 12);
  const uint8_t* A0 = z + 16;
  const uint8_t* A1 = A0 + 4 * (size_t)ng;
  const uint8_t* AP = A1 + 4 * (size_t)ng;
  const uint8_t* S = AP + 4 * (size_t)ng;
  int g, lo, hi, seg;
  unsigned o0, o1;
  const uint8_t* sp;
  uint64_t w[SEG_WORDS];
  int i, first_word, nwords;
  int64_t bytes = 16;

  /* group: last g with A0[g]+A1[g] <= x  (A0[0]+A1[0] == 0 always) */
#define GSUM(k) (be32(A0 + 4 * (size_t)(k)) + be32(A1 + 4 * (size_t)(k)))
  lo = 0; hi = ng - 1;
  bytes += 16;
  if (x >= GSUM(hi)) g = hi;
  else {
    while (hi - lo > 1) {
      int mid = (lo + hi) / 2;
      bytes += 8;
      if (x < GSUM(mid)) hi = mid; else lo = mid;
    }
    g = lo;
  }
#undef GSUM
  o0 = be32(A0 + 4 * (size_t)g);
  o1 = be32(A1 + 4 * (size_t)g);
  sp = S + be32(AP + 4 * (size_t)g);
  bytes += 4;

  /* segment within the group: pairs of LSB-first 7-bit varbytes, terminator has 0x80 */
  seg = 0;
  for (;;) {
    unsigned s0 = 0, s1 = 0;
    int sh = 0, k = 0;
    for (;;) { uint8_t c = sp[k++]; s0 |= (unsigned)(c & 0x7f) << sh; sh += 7; if (c & 0x80) break; }
    sh = 0;
    for (;;) { uint8_t c = sp[k++]; s1 |= (unsigned)(c & 0x7f) << sh; sh += 7; if (c & 0x80) break; }
    bytes += k;
    if (o0 + s0 + o1 + s1 <= x) { o0 += s0; o1 += s1; seg++; sp += k; }
    else break;
  }
  seg += SEGS_PER_GROUP * g;

  /* the 512-bit segment, zero-filled past total_segment_words */
  first_word = SEG_WORDS * seg;
  nwords = total_words - first_word;
  if (nwords > SEG_WORDS) nwords = SEG_WORDS;
  if (nwords < 0) nwords = 0;
  for (i = 0; i < nwords; i++) w[i] = be64(z + d_off + 8 * (size_t)(first_word + i));
  for (; i < SEG_WORDS; i++) w[i] = 0;
  bytes += 8 * nwords;

  if (w[0] >> 63) {
    /* RLE: bit 1 = value of the first run, then gamma codes (k zeros, k+1-bit value) */
    int pos = 2;
    int bit = (int)((w[0] >> 62) & 1);
    for (;;) {
      /* read 64 bits at 'pos' (zero past the end), as advance_segs_reader */
      uint64_t cur = 0;
      int wi = pos >> 6, bo = pos & 63, k;
      unsigned v;
      if (wi < SEG_WORDS) cur = w[wi] << bo;
      if (bo && wi + 1 < SEG_WORDS) cur |= w[wi + 1] >> (64 - bo);
      k = cur ? __builtin_clzll(cur) : 64;
      k = 2 * k + 1;
      v = k >= 64 ? 0 : (unsigned)(cur >> (64 - k));   /* k<64 for any valid code */
      pos += k;
      if (o0 + o1 + v <= x) {
        if (bit) o1 += v; else o0 += v;
        bit = !bit;
      } else {
        unsigned part = 1 + x - (o0 + o1);
        if (bit) o1 += part; else o0 += part;
        o->bit = bit;
        break;
      }
    }
  } else {
    /* raw: payload bits follow the type bit; count bits [1 .. 1+r] where r = x-(o0+o1) */
    unsigned r = x - (o0 + o1);
    unsigned nbits = r + 1;          /* payload bits to count */
    unsigned endpos = 1 + r;         /* position of the queried bit in the segment */
    unsigned ones = 0;
    unsigned full = (endpos + 1) / 64, rem = (endpos + 1) % 64;
    for (i = 0; i < (int)full; i++) ones += (unsigned)__builtin_popcountll(w[i]);
    if (rem) ones += (unsigned)__builtin_popcountll(w[full] >> (64 - rem));
    /* the type bit itself is 0 and contributes no ones */
    o1 += ones;
    o0 += nbits - ones;
    o->bit = (int)((w[endpos >> 6] >> (63 - (endpos & 63))) & 1);
  }
  o->occ0 = (int)o0;
  o->occ1 = (int)o1;
  o->bytes = bytes;
}

void fmo_bseq_rank(const unsigned char* zdata, int index1, int* occ0, int* occ1, int* bit)
{
  rank_out_t o;
  bseq_rank_impl(zdata, (unsigned)index1, &o);
  *occ0 = o.occ0; *occ1 = o.occ1; *bit = o.bit;
}

/* node directory lookup: wtree_bseq + stored_num_for_node_num (wtree_funcs.h:583-626).
 * Returns NULL when 'node' is not an internal node (i.e. it is a leaf). */
static const uint8_t* wt_node(const uint8_t* wt, unsigned node, int64_t* bytes)
{
  int n = (int)be32(wt);
  const uint8_t* dir = wt + 4;
  int lo = 0, hi = n - 1;
  *bytes += 4;
  while (lo <= hi) {
    int mid = (lo + hi) / 2;
    unsigned v = be32(dir + 8 * (size_t)mid);
    *bytes += 4;
    if (v == node) { *bytes += 4; return wt + be32(dir + 8 * (size_t)mid + 4); }
    if (v < node) lo = mid + 1; else hi = mid - 1;
  }
  return NULL;
}

/* wtree_occs (wtree.c:1081-1115): occurrences of 'leaf' among the first index1 symbols */
static int wt_occs(fmo_index* ix, const uint8_t* wt, uint32_t leaf, int index1)
{
  int L = 31 - __builtin_clz(leaf);
  unsigned node = 1;
  int idx = index1, i;
  for (i = 1; ; i++) {
    rank_out_t o;
    int64_t b = 0;
    const uint8_t* z = wt_node(wt, node, &b);
    ix->ctr_bytes += b;
    if (!z) break;
    bseq_rank_impl(z, (unsigned)idx, &o);
    ix->ctr_bytes += o.bytes;
    ix->ctr_levels++;
    node = leaf >> (L - i);
    idx -= (node & 1) ? o.occ0 : o.occ1;
    if (idx == 0) break;
  }
  return idx;
}

/* wtree_rank (wtree.c:1117-1148): the symbol at index1 and its occurrence number */
static void wt_rank(fmo_index* ix, const uint8_t* wt, int index1, uint32_t* leaf_out, int* count_out)
{
  unsigned node = 1;
  int idx = index1;
  for (;;) {
    rank_out_t o;
    int64_t b = 0;
    const uint8_t* z = wt_node(wt, node, &b);
    ix->ctr_bytes += b;
    if (!z) break;
    bseq_rank_impl(z, (unsigned)idx, &o);
    ix->ctr_bytes += o.bytes;
    ix->ctr_levels++;
    node = (node << 1) | (unsigned)o.bit;
    idx -= o.bit ? o.occ0 : o.occ1;
  }
  *leaf_out = node;
  *count_out = idx;
}

static inline uint32_t bucket_occs(const fmo_index* ix, const dblock_t* b, int ch, int bucket)
{
  /* get_bucket_occs, index.c:1828-1843: char-major, stride = num_buckets of THIS block */
  size_t base = HDR_BYTES + 4 * ((size_t)ix->buckets_per_block + 1);
  return be32(b->blob.data + base + 4 * ((size_t)ch * (size_t)b->num_buckets + (size_t)bucket));
}

/* block_request(BLOCK_REQUEST_OCCS) (index.c:1973-2100) */
static int block_occs_at(fmo_index* ix, dblock_t* b, int ch, int row_in_block, int64_t* out)
{
  bucket_tab_t* t;
  int bucket, rb, err;
  int64_t occ = 0;
  if (row_in_block < 0 || row_in_block >= b->size) return FMO_ERR_PARAM;
  bucket = row_in_block / ix->bucket_size;
  rb = row_in_block % ix->bucket_size;
  err = bucket_tables(ix, b, bucket, &t);
  if (err) return err;
  if (t->in_use[ch])
    occ += wt_occs(ix, b->blob.data + t->off_wtree, t->leaf[t->ch_to_seq[ch]], rb + 1);
  occ += bucket_occs(ix, b, ch, bucket);
  ix->ctr_bytes += 4;
  *out = occ;
  return FMO_OK;
}

/* header_occs_request(HDR_BSEARCH_BLOCK_ROWS|HDR_REQUEST_C|HDR_REQUEST_BLOCK_OCCS|HDR_BACK)
 * followed by block_request(OCCS): one half of a backward-search step (server.c:853-897) */
static int c_plus_occ(fmo_index* ix, int ch, int64_t row, int64_t* out)
{
  int64_t blk, row0, occ;
  int err;
  if (ch < 0 || ch >= ALPHA) return FMO_ERR_PARAM;
  if (row < 0 || row >= ix->total_length) return FMO_ERR_PARAM;
  blk = row / ix->block_size;                      /* bsearch_block_rows, index.c:1613-1617 */
  row0 = blk * (int64_t)ix->block_size;            /* get_block_row, index.c:1644-1648 */
  err = block_occs_at(ix, &ix->blocks[blk], ch, (int)(row - row0), &occ);
  if (err) return err;
  ix->ctr_bytes += 16;
  ix->ctr_occ++;
  *out = hdr_C(ix, ch) + hdr_block_occs(ix, ch, blk) + occ;
  return FMO_OK;
}

int fmo_occ(fmo_index* ix, int ch, int64_t row, int64_t* cpo, int64_t* occ_only)
{
  int err = c_plus_occ(ix, ch, row, cpo);
  if (err) return err;
  *occ_only = *cpo - hdr_C(ix, ch);
  return FMO_OK;
}

/* do_string_query (server.c:713-946) */
static int count_one(fmo_index* ix, int m, const uint16_t* pat, int64_t* first_out, int64_t* last_out)
{
  int64_t first, last;
  int i, err;
  if (m == 0) { *first_out = 0; *last_out = ix->total_length - 1; return FMO_OK; }
  first = hdr_C(ix, pat[m - 1]);
  last = hdr_C(ix, pat[m - 1] + 1) - 1;
  for (i = m - 1; i > 0 && first <= last; i--) {
    int ch = pat[i - 1];
    int64_t nf, nl;
    if (first == 0) nf = hdr_C(ix, ch);                       /* server.c:847-851, 889-891 */
    else { err = c_plus_occ(ix, ch, first - 1, &nf); if (err) return err; }
    err = c_plus_occ(ix, ch, last, &nl);
    if (err) return err;
    first = nf;
    last = nl - 1;
  }
  *first_out = first;
  *last_out = last;
  return FMO_OK;
}

int fmo_count(fmo_index* ix, int npats, const int32_t* plen, const uint16_t* flat,
              const int64_t* offs, int64_t* first, int64_t* last)
{
  int i;
  for (i = 0; i < npats; i++) {
    int64_t f, l;
    int err = count_one(ix, plen[i], flat + offs[i], &f, &l);
    if (err) return err;
    if (last) { first[i] = f; last[i] = l; }
    else first[i] = l - f + 1;                                /* femto.c:313-318 */
  }
  return FMO_OK;
}

/* do_back_query (server.c:2228-2359) over block_request(CHAR|OCCS|LOCATION) (index.c:2037-2140) */
int fmo_back_step(fmo_index* ix, int64_t row, int* ch_out, int64_t* next_row, int64_t* offset)
{
  int64_t blk, row0;
  dblock_t* b;
  bucket_tab_t* t;
  int bucket, rb, err, count, seq, ch, len, i;
  uint32_t leaf;
  const uint8_t* blkdata;
  rank_out_t mo;
  if (row < 0 || row >= ix->total_length) return FMO_ERR_PARAM;
  blk = row / ix->block_size;
  row0 = blk * (int64_t)ix->block_size;
  b = &ix->blocks[blk];
  blkdata = b->blob.data;
  if (row - row0 >= b->size) return FMO_ERR_PARAM;
  bucket = (int)((row - row0) / ix->bucket_size);
  rb = (int)((row - row0) % ix->bucket_size);
  err = bucket_tables(ix, b, bucket, &t);
  if (err) return err;
  wt_rank(ix, blkdata + t->off_wtree, rb + 1, &leaf, &count);
  /* leaf -> symbol: equivalent of huff_perm/base/limit decode (index.c:2046-2062) */
  len = 31 - __builtin_clz(leaf);
  if (len > MAX_CODE_LEN) return FMO_ERR_BZ_DATA;
  seq = -1;
  for (i = 0; i <= t->n_in_use; i++) if (t->leaf[i] == leaf) { seq = i; break; }
  if (seq < 0 || seq >= t->n_in_use) return FMO_ERR_INVALID;
  ch = t->seq_to_ch[seq];
  *ch_out = ch;

  /* mark test: bseq_rank over this symbol's mark table at its occurrence number */
  {
    uint32_t toff = be32(blkdata + t->off_marktab + 4 * (size_t)seq);
    if (toff & 7) return FMO_ERR_FORMAT;
    bseq_rank_impl(blkdata + t->off_marktab + toff, (unsigned)count, &mo);
    ix->ctr_bytes += mo.bytes + 4;
  }
  if (mo.bit) {
    uint32_t aoff = be32(blkdata + t->off_markarr + 4 * (size_t)seq);
    const uint8_t* arr = blkdata + t->off_markarr + aoff;
    bitrd_t r;
    int64_t v = 0;
    int k;
    r.p = arr; r.bit = (size_t)ix->text_size_bits * (size_t)(mo.occ1 - 1);
    for (k = 0; k < ix->text_size_bits; k++) v = (v << 1) | rd_bits(&r, 1);
    *offset = v;
    ix->ctr_bytes += 4 + (ix->text_size_bits + 7) / 8;
  } else *offset = -1;

  if (ch <= ESC_SEOF) *next_row = -1;                        /* server.c:2341-2346 */
  else *next_row = hdr_C(ix, ch) + hdr_block_occs(ix, ch, blk)
                 + bucket_occs(ix, b, ch, bucket) + count - 1;
  ix->ctr_bytes += 20;
  return FMO_OK;
}

/* what do_context_query with LOCATE_STRONG and no context computes for one row
 * (server.c:2627-2795): SA[row].  The reference walks forward and backward at once;
 * the backward half alone reaches a mark within mark_period-1 steps and before a
 * document start (should_mark, index_types.h:134-144), and SA[row] = mark + steps. */
static int locate_row(fmo_index* ix, int64_t row, int64_t* out)
{
  int64_t steps = 0;
  for (;;) {
    int ch, err;
    int64_t next, off;
    err = fmo_back_step(ix, row, &ch, &next, &off);
    if (err) return err;
    if (off >= 0) { *out = off + steps; return FMO_OK; }
    if (next < 0) return FMO_ERR_INVALID;     /* unmarked document start: malformed index */
    row = next;
    steps++;
  }
}

int fmo_locate_range(fmo_index* ix, int64_t first, int64_t last, int64_t* offsets)
{
  int64_t r;
  for (r = first; r <= last; r++) {
    int err = locate_row(ix, r, &offsets[r - first]);
    if (err) return err;
  }
  return FMO_OK;
}

/* parallel_locate (femto.c:331-399) over do_locate_query (server.c:4373-4436):
 * note the '>' clip at :4411 -- up to max_occs+1 rows when count == max_occs+1 */
int fmo_locate(fmo_index* ix, int npats, const int32_t* plen, const uint16_t* flat,
               const int64_t* offs, int max_occs_each, int32_t* noccs, int64_t* out_start,
               int64_t* out, int64_t out_cap)
{
  int i;
  int64_t pos = 0;
  for (i = 0; i < npats; i++) {
    int64_t f, l, r;
    int err = count_one(ix, plen[i], flat + offs[i], &f, &l);
    if (err) return err;
    out_start[i] = pos;
    if (f > l) { noccs[i] = 0; continue; }
    if (l - f > (int64_t)max_occs_each) l = f + (int64_t)max_occs_each - 1;
    noccs[i] = (int32_t)(l - f + 1);
    if (pos + noccs[i] > out_cap) return FMO_ERR_PARAM;
    for (r = f; r <= l; r++) {
      err = locate_row(ix, r, &out[pos++]);
      if (err) return err;
    }
  }
  return FMO_OK;
}

/* do_extract_document_query (server.c:6364-6437): doc_len-1 LF steps from the document's
 * EOF row, emitting L right to left (context query with beforeCtxLen = doc_len-1) */
int fmo_extract(fmo_index* ix, int64_t doc, uint16_t* out, int64_t out_cap, int64_t* out_len)
{
  int64_t len, row, t;
  int err = fmo_doc_info(ix, doc, &len, &row);
  if (err) return err;
  *out_len = len - 1;
  if (len - 1 > out_cap) return FMO_ERR_PARAM;
  for (t = 0; t < len - 1; t++) {
    int ch;
    int64_t next, off;
    err = fmo_back_step(ix, row, &ch, &next, &off);
    if (err) return err;
    out[len - 2 - t] = (uint16_t)ch;
    row = next;
    if (row < 0 && t + 1 < len - 1) return FMO_ERR_INVALID;
  }
  return FMO_OK;
}

void fmo_counters(const fmo_index* ix, int64_t* bytes, int64_t* occ_calls, int64_t* levels)
{
  *bytes = ix->ctr_bytes; *occ_calls = ix->ctr_occ; *levels = ix->ctr_levels;
}
void fmo_reset_counters(fmo_index* ix) { ix->ctr_bytes = ix->ctr_occ = ix->ctr_levels = 0; }
