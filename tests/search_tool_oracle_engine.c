/* search_tool_oracle_engine.c -- TEST INFRASTRUCTURE, never shipped: the handful of fm_* entry points that
 * integration/femto_search_b200.c calls, answered by the CPU oracle (oracle/fm_oracle.c), so that the tool's
 * whole main() -- option parsing, the queries it makes, sorting and grouping of the results, the report --
 * can be replayed against the reference tool's golden output on a machine without a GPU
 * (tests/test_search_format.py).  The product binary oracle/_ref/femto_search_b200 links libfemto_b200.so
 * and nothing else; its own run on the GPU is tests/test_gpu_zz_search_tool.py.
 */
#include <stdlib.h>
#include <string.h>

#include "../include/femto_b200.h"
#include "../oracle/fm_oracle.h"

struct fm_index {
  fmo_index* o;
};

const char* fm_last_error(void) { return "oracle stand-in"; }

int fm_open(const char* path, int device, fm_index_t** out) {
  int err = 0;
  fm_index_t* ix = (fm_index_t*)calloc(1, sizeof *ix);
  (void)device;
  ix->o = fmo_open(path, &err);
  if (!ix->o) return err ? err : FM_ERR_IO;
  *out = ix;
  return FM_OK;
}

void fm_close(fm_index_t* ix) {
  fmo_close(ix->o);
  free(ix);
}

int fm_info(const fm_index_t* ix, fm_info_t* out) {
  int64_t info[16] = {0};
  int rc = fmo_header_info(ix->o, info);
  memset(out, 0, sizeof *out);
  out->num_blocks = info[0];
  out->total_length = info[1];
  out->num_documents = info[2];
  return rc;
}

int fm_count(fm_index_t* ix, int npats, const int* plen, const uint16_t* const* pats, int64_t* first, int64_t* last) {
  for (int i = 0; i < npats; i++) {
    const int64_t off = 0;
    const int32_t len = plen[i];
    const int rc = fmo_count(ix->o, 1, &len, pats[i], &off, &first[i], &last[i]);
    if (rc) return rc;
  }
  return FM_OK;
}

int fm_locate_range(fm_index_t* ix, int64_t first, int64_t last, int64_t* offsets) {
  return fmo_locate_range(ix->o, first, last, offsets);
}

int fm_resolve(const fm_index_t* ix, int64_t n, const int64_t* offsets, int64_t* doc, int64_t* doc_off) {
  for (int64_t i = 0; i < n; i++) {
    const int rc = fmo_resolve(ix->o, offsets[i], &doc[i], &doc_off[i]);
    if (rc) return rc;
  }
  return FM_OK;
}

int fm_doc_name(const fm_index_t* ix, int64_t doc, void* out, int64_t out_cap, int64_t* out_len) {
  const unsigned char* info = NULL;
  const int rc = fmo_doc_name(ix->o, doc, &info, out_len);
  if (rc) return rc;
  if (*out_len > out_cap) return FM_ERR_FULL;
  if (*out_len) memcpy(out, info, (size_t)*out_len);
  return FM_OK;
}

/* ascending documents of the rows first..last, by locating every row */
int fm_range_documents(fm_index_t* ix, int64_t first, int64_t last, int64_t* docs, int64_t docs_cap, int64_t* ndocs) {
  const int64_t n = last - first + 1;
  int64_t info[16] = {0};
  int64_t* offs = (int64_t*)malloc((size_t)(n > 0 ? n : 1) * sizeof(int64_t));
  int rc = fmo_header_info(ix->o, info);
  char* seen = (char*)calloc((size_t)info[2] + 1, 1);
  if (!rc) rc = fmo_locate_range(ix->o, first, last, offs);
  for (int64_t i = 0; !rc && i < n; i++) {
    int64_t d, k;
    rc = fmo_resolve(ix->o, offs[i], &d, &k);
    if (!rc) seen[d] = 1;
  }
  *ndocs = 0;
  for (int64_t d = 0; !rc && d < info[2]; d++)
    if (seen[d]) {
      if (*ndocs >= docs_cap) rc = FM_ERR_FULL;
      else docs[(*ndocs)++] = d;
    }
  free(offs);
  free(seen);
  return rc;
}
