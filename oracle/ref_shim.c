/* oracle/ref_shim.c -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * A thin flat-C wrapper around the UNMODIFIED reference (femto-dev/femto,
 * compiled where it lies under /root/reference by oracle/Makefile into
 * oracle/_ref/libfemto_ref.so).  Everything that computes anything here is the
 * reference's own code; this file only marshals flat arrays in and out so that
 * tests (and bench.py's cpu_baseline / --impl reference legs) can drive it via
 * ctypes.  Nothing under femto_b200/ may link or load this.
 *
 * Reference entry points used (all cited file:line are under /root/reference):
 *   parallel_count / parallel_locate / parallel_locate_range  src/main/femto.c:275,331,481
 *   femto_start_server_err / femto_stop_server / femto_loc_for_path_err  src/main/femto.c:54,77,269
 *   header_occs_request / block_request (leaf interface)      src/main/index.c:1698,1973
 *   open_header_block / open_data_block                       src/main/index.c:1482,1419
 *   init_prepared_text / count_file / append_file_mem         src/main/bwt_prepare.c:123,193,231
 *   save_prepared_bwt                                         src/main/bwt_creator.c:37
 *   index_documents                                           src/main/construct.c:572
 *   bseq_construct_forcetype / bseq_rank                      src/main/wtree.c:364,635
 *   compress_bucket (via constructor_*)                       src/main/index.c:309
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <unistd.h>

#include "index_types.h"
#include "error.h"
#include "femto.h"
#include "femto_internal.h"
#include "server.h"
#include "index.h"
#include "wtree.h"
#include "wtree_funcs.h"
#include "bwt_prepare.h"
#include "bwt_creator.h"
#include "bwt_reader.h"
#include "construct.h"
#include "block_storage.h"
#include "timing.h"

typedef struct {
  femto_server_t srv;
  index_locator_t loc;
  /* leaf-level access (opened lazily) */
  path_translator_t trans;
  int trans_inited;
  index_locator_t leaf_loc;
  header_block_t hdr;
  int hdr_open;
  data_block_t blk;
  int64_t blk_num; /* -1 = none */
} ref_handle_t;

static int code_of(error_t err)
{
  int c;
  if (!err) return 0;
  c = (int) err_code(err);
  if (getenv("FEMTO_REF_VERBOSE")) warn_if_err(err);
  return c ? c : -1;
}

void* ref_open(const char* index_path)
{
  error_t err;
  ref_handle_t* h = calloc(1, sizeof(ref_handle_t));
  if (!h) return NULL;
  h->blk_num = -1;
  err = femto_start_server_err(&h->srv, 0);
  if (err) { code_of(err); free(h); return NULL; }
  err = femto_loc_for_path_err(&h->srv, index_path, &h->loc);
  if (err) { code_of(err); femto_stop_server(&h->srv); free(h); return NULL; }
  err = path_translator_init(&h->trans);
  if (err) { code_of(err); femto_stop_server(&h->srv); free(h); return NULL; }
  h->trans_inited = 1;
  err = path_translator_id_for_path(&h->trans, index_path, &h->leaf_loc);
  if (err) { code_of(err); femto_stop_server(&h->srv); free(h); return NULL; }
  return h;
}

void ref_close(void* hv)
{
  ref_handle_t* h = hv;
  if (!h) return;
  if (h->blk_num >= 0) close_data_block(&h->blk);
  if (h->hdr_open) close_header_block(&h->hdr);
  if (h->trans_inited) path_translator_destroy(&h->trans);
  femto_stop_server(&h->srv);
  free(h);
}

/* patterns: flat alpha_t buffer; pattern i = flat[offs[i] .. offs[i]+plen[i]) */
int ref_count(void* hv, int npats, const int* plen, const uint16_t* flat,
              const int64_t* offs, int64_t* first, int64_t* last)
{
  ref_handle_t* h = hv;
  error_t err;
  int i;
  alpha_t** pats = malloc(sizeof(alpha_t*) * (npats > 0 ? npats : 1));
  if (!pats) return ERR_CODE_MEM;
  for (i = 0; i < npats; i++) pats[i] = (alpha_t*) &flat[offs[i]];
  err = parallel_count(&h->srv, h->loc, npats, (int*) plen, pats, first, last);
  free(pats);
  return code_of(err);
}

/* Locate: results are appended to out[] in pattern order; out_start[i] is the
 * index of pattern i's first offset, noccs[i] the number returned.  If the
 * total exceeds out_cap the call fails with ERR_CODE_PARAM after filling noccs. */
int ref_locate(void* hv, int npats, const int* plen, const uint16_t* flat,
               const int64_t* offs, int max_occs_each,
               int* noccs, int64_t* out_start, int64_t* out, int64_t out_cap)
{
  ref_handle_t* h = hv;
  error_t err;
  int i, rc = 0;
  int64_t pos = 0;
  alpha_t** pats = malloc(sizeof(alpha_t*) * (npats > 0 ? npats : 1));
  int64_t** res = calloc(npats > 0 ? npats : 1, sizeof(int64_t*));
  if (!pats || !res) { free(pats); free(res); return ERR_CODE_MEM; }
  for (i = 0; i < npats; i++) pats[i] = (alpha_t*) &flat[offs[i]];
  err = parallel_locate(&h->srv, h->loc, npats, (int*) plen, pats, max_occs_each, noccs, res);
  if (err) { rc = code_of(err); goto done; }
  for (i = 0; i < npats; i++) {
    out_start[i] = pos;
    if (pos + noccs[i] > out_cap) { rc = ERR_CODE_PARAM; }
    else if (noccs[i] > 0) memcpy(&out[pos], res[i], sizeof(int64_t) * noccs[i]);
    pos += noccs[i];
  }
done:
  for (i = 0; i < npats; i++) free(res[i]);
  free(res);
  free(pats);
  return rc;
}

int ref_locate_range(void* hv, int64_t first, int64_t last, int64_t* offsets)
{
  ref_handle_t* h = hv;
  return code_of(parallel_locate_range(&h->srv, h->loc, first, last, offsets));
}

/* ---- leaf interface: header_occs_request + block_request, as the state
 *      machines in server.c call them ---- */
static int leaf_get(ref_handle_t* h, int64_t blk)
{
  error_t err;
  if (!h->hdr_open) {
    err = open_header_block(&h->hdr, &h->trans, h->leaf_loc);
    if (err) return code_of(err);
    h->hdr_open = 1;
  }
  if (blk >= 0 && h->blk_num != blk) {
    if (h->blk_num >= 0) { close_data_block(&h->blk); h->blk_num = -1; }
    err = open_data_block(&h->blk, &h->trans, h->leaf_loc, blk, 8);
    if (err) return code_of(err);
    h->blk_num = blk;
  }
  return 0;
}

/* header fields: out[0]=nblocks out[1]=total_length out[2]=ndocs out[3]=block_size
 * out[4]=bucket_size out[5]=mark_period out[6]=chunk_size */
int ref_header_info(void* hv, int64_t* out)
{
  ref_handle_t* h = hv;
  int rc = leaf_get(h, -1);
  if (rc) return rc;
  out[0] = h->hdr.hdr.number_of_blocks;
  out[1] = h->hdr.hdr.total_length;
  out[2] = h->hdr.hdr.number_of_documents;
  out[3] = h->hdr.hdr.param.block_size;
  out[4] = h->hdr.hdr.param.b_size;
  out[5] = h->hdr.hdr.param.mark_period;
  out[6] = h->hdr.hdr.param.chunk_size;
  return 0;
}

/* C[ch] exactly as get_C (ch may be ALPHA_SIZE -> total_length). */
int ref_C(void* hv, int ch, int64_t* out)
{
  ref_handle_t* h = hv;
  header_occs_request_t r;
  int rc = leaf_get(h, -1);
  if (rc) return rc;
  memset(&r, 0, sizeof(r));
  r.ch = ch;
  rc = code_of(header_occs_request(&h->hdr, HDR_REQUEST_C, &r));
  *out = r.occs;
  return rc;
}

/* out = C[ch] + Occ(ch,row): the HDR_BACK header request followed by the
 * BLOCK_REQUEST_OCCS block request, i.e. one half of a backward-search step
 * (server.c:853-897).  occ_only gets Occ(ch,row) without C[ch]. */
int ref_occ(void* hv, int ch, int64_t row, int64_t* c_plus_occ, int64_t* occ_only)
{
  ref_handle_t* h = hv;
  header_occs_request_t r;
  block_request_t b;
  int64_t C;
  int rc = leaf_get(h, -1);
  if (rc) return rc;
  memset(&r, 0, sizeof(r));
  r.ch = ch; r.row = row;
  rc = code_of(header_occs_request(&h->hdr,
        HDR_BSEARCH_BLOCK_ROWS | HDR_REQUEST_C | HDR_REQUEST_BLOCK_OCCS | HDR_BACK, &r));
  if (rc) return rc;
  rc = leaf_get(h, r.block_num);
  if (rc) return rc;
  memset(&b, 0, sizeof(b));
  b.ch = ch; b.row_in_block = (int) (row - r.row);
  rc = code_of(block_request(&h->blk, BLOCK_REQUEST_OCCS, &b));
  if (rc) return rc;
  *c_plus_occ = r.occs + b.occs_in_block;
  rc = ref_C(hv, ch, &C);
  if (rc) return rc;
  *occ_only = *c_plus_occ - C;
  return 0;
}

/* One LF step with mark test, as do_back_query (server.c:2228-2359):
 * ch = L[row], next = LF(row) (or -1 when ch <= SEOF), offset = SA[row] or -1. */
int ref_back_step(void* hv, int64_t row, int* ch, int64_t* next_row, int64_t* offset)
{
  ref_handle_t* h = hv;
  header_occs_request_t r;
  block_request_t b;
  int rc = leaf_get(h, -1);
  if (rc) return rc;
  memset(&r, 0, sizeof(r));
  r.row = row;
  rc = code_of(header_occs_request(&h->hdr, HDR_BSEARCH_BLOCK_ROWS, &r));
  if (rc) return rc;
  rc = leaf_get(h, r.block_num);
  if (rc) return rc;
  memset(&b, 0, sizeof(b));
  b.ch = INVALID_ALPHA; b.row_in_block = (int) (row - r.row);
  rc = code_of(block_request(&h->blk,
        BLOCK_REQUEST_CHAR | BLOCK_REQUEST_OCCS | BLOCK_REQUEST_LOCATION, &b));
  if (rc) return rc;
  *ch = b.ch;
  *offset = b.offset;
  r.ch = b.ch; r.occs = 0; r.row = 0;
  rc = code_of(header_occs_request(&h->hdr,
        HDR_REQUEST_C | HDR_REQUEST_BLOCK_OCCS | HDR_BACK, &r));
  if (rc) return rc;
  *next_row = (int64_t) b.occs_in_block - 1 + r.occs;
  if (b.ch <= ESCAPE_CODE_SEOF) *next_row = -1;
  return 0;
}

/* doc tables from the header: out_end = doc_ends[doc], out_eof_row = doc_eof_rows[doc] */
/* The document chunk that holds `row`: block_chunk_request with BLOCK_CHUNK_FIND_NUMBER |
 * BLOCK_CHUNK_REQUEST_DOCUMENTS (src/main/index.c:2200-2236), its results read back with
 * results_reader_next (src/main/results.c:356-371).  first/last = global rows of the chunk. */
int ref_chunk_documents(void* hv, int64_t row, int64_t* first, int64_t* last, int64_t* docs, int64_t cap,
                        int64_t* ndocs)
{
  ref_handle_t* h = hv;
  header_occs_request_t r;
  block_chunk_request_t c;
  results_reader_t rd;
  int64_t doc, off, n = 0;
  int rc = leaf_get(h, -1);
  if (rc) return rc;
  memset(&r, 0, sizeof(r));
  r.row = row;
  rc = code_of(header_occs_request(&h->hdr, HDR_BSEARCH_BLOCK_ROWS, &r));
  if (rc) return rc;
  rc = leaf_get(h, r.block_num);
  if (rc) return rc;
  memset(&c, 0, sizeof(c));
  c.first = (int) (row - r.row);
  rc = code_of(block_chunk_request(&h->blk, BLOCK_CHUNK_FIND_NUMBER | BLOCK_CHUNK_REQUEST_DOCUMENTS, &c));
  if (rc) return rc;
  *first = r.row + c.first;
  *last = r.row + c.last;
  rc = code_of(results_reader_create(&rd, &c.results));
  if (!rc) {
    while (results_reader_next(&rd, &doc, &off)) {
      if (n < cap) docs[n] = doc;
      n++;
    }
    results_reader_destroy(&rd);
  }
  results_destroy(&c.results);
  *ndocs = n;
  return rc;
}

int ref_doc_info(void* hv, int64_t doc, int64_t* out_len, int64_t* out_eof_row)
{
  ref_handle_t* h = hv;
  header_loc_request_t r;
  int rc = leaf_get(h, -1);
  if (rc) return rc;
  memset(&r, 0, sizeof(r));
  r.loc.doc = doc;
  rc = code_of(header_loc_request(&h->hdr, HDR_LOC_REQUEST_DOC_LEN | HDR_LOC_REQUEST_DOC_EOF_ROW, &r));
  *out_len = r.doc_len;
  *out_eof_row = r.offset;
  return rc;
}

/* resolve a logical offset to (doc, in-doc offset) as resolve_location (index.c:1587) */
int ref_resolve(void* hv, int64_t offset, int64_t* doc, int64_t* doc_off)
{
  ref_handle_t* h = hv;
  header_loc_request_t r;
  int rc = leaf_get(h, -1);
  if (rc) return rc;
  memset(&r, 0, sizeof(r));
  r.offset = offset;
  rc = code_of(header_loc_request(&h->hdr, HDR_LOC_RESOLVE_LOCATION, &r));
  *doc = r.loc.doc;
  *doc_off = r.loc.offset;
  return rc;
}

/* ---- index construction through the reference's in-memory builder ---- */
int ref_build_index(int ndocs, const int64_t* doc_lens, const unsigned char* const* docs,
                    const char* index_path, const char* scratch_dir,
                    int block_size, int bucket_size, int chunk_size, int mark_period)
{
  prepared_text_t p;
  index_block_param_t param;
  bwt_reader_t bwt;
  FILE* bwt_f = NULL;
  error_t err;
  char info_path[4096];
  int i;

  set_default_param(&param);
  if (block_size > 0) param.block_size = block_size;
  if (bucket_size > 0) param.b_size = bucket_size;
  /* No document map is passed to index_documents (as the reference's own test_construct does,
   * src/main/index_test.c:566): its map branch (construct.c:663-679) never advances the map reader
   * and indexes chunks[] by block, not bucket.  Document chunks are pinned through femto_index
   * (oracle/_ref/femto_index) instead.  chunk_size is therefore ignored here. */
  (void) chunk_size;
  param.chunk_size = 0;
  if (mark_period >= 0) param.mark_period = mark_period;
  err = calculate_params(&param);
  if (err) return code_of(err);

  snprintf(info_path, sizeof(info_path), "%s/ref_doc_info.%ld", scratch_dir, (long) getpid());
  memset(&p, 0, sizeof(p));
  err = init_prepared_text(&p, info_path);
  if (err) return code_of(err);
  for (i = 0; i < ndocs; i++) {
    char name[64];
    int n = snprintf(name, sizeof(name), "doc%d", i);
    err = count_file(&p, doc_lens[i], 0, NULL, n, (unsigned char*) name);
    if (err) goto fail;
  }
  for (i = 0; i < ndocs; i++) {
    char name[64];
    int n = snprintf(name, sizeof(name), "doc%d", i);
    err = append_file_mem(&p, doc_lens[i], (unsigned char*) docs[i], 0, NULL, NULL,
                          n, (unsigned char*) name);
    if (err) goto fail;
  }
  bwt_f = tmpfile();
  if (!bwt_f) { err = ERR_IO_UNK; goto fail; }
  start_clock(); /* save_prepared_bwt ends with one stop_clock() more than it starts (bwt_creator.c:135) */
  err = save_prepared_bwt(&p, param.mark_period, bwt_f, 0, NULL, 0);
  if (err) goto fail;
  rewind(bwt_f);
  err = bwt_reader_open(&bwt, bwt_f);
  if (err) goto fail;
  err = index_documents(&bwt, NULL, &p.info_reader, &param, index_path, NULL);
  bwt_reader_close(&bwt);
  if (err) goto fail;
  err = free_prepared_text(&p);
  unlink(info_path);
  if (bwt_f) fclose(bwt_f);
  return code_of(err);
fail:
  {
    int rc = code_of(err);
    free_prepared_text(&p);
    unlink(info_path);
    if (bwt_f) fclose(bwt_f);
    return rc;
  }
}

/* ---- bseq level (wtree.c) for builder / decoder parity tests ---- */
int ref_bseq_construct(int bitlen, const unsigned char* data, int type,
                       unsigned char** zdata, int* zlen)
{
  return code_of(bseq_construct_forcetype(zlen, zdata, bitlen, (unsigned char*) data, NULL, type));
}

void ref_bseq_rank(const unsigned char* zdata, int index1, int* occ0, int* occ1, int* bit)
{
  bseq_query_t q;
  memset(&q, 0, sizeof(q));
  q.index = index1;
  bseq_rank(zdata, &q);
  *occ0 = q.occs[0]; *occ1 = q.occs[1]; *bit = q.bit;
}

/* femto's generic request interface, start to finish (femto.h:75-149): *response is malloc()ed by femto */
int ref_generic_request(void* hv, const char* index_path, const char* request, char** response)
{
  ref_handle_t* h = hv;
  femto_request_t* req = NULL;
  int rc;
  *response = NULL;
  rc = femto_create_generic_request(&req, &h->srv, index_path, request);
  if (rc) return rc;
  rc = femto_begin_request(&h->srv, req);
  if (!rc) rc = femto_wait_request(&h->srv, req);
  if (!rc) rc = femto_response_for_generic_request(req, &h->srv, response);
  femto_destroy_request(req);
  return rc;
}

void ref_free(void* p) { free(p); }

int ref_wtree_settings(void) { return wtree_settings_number(); }
