/* search_format_check.c -- TEST HARNESS: prints a femto_search report through integration/femto_search_format.h
 * from results given on stdin, so that the formatting can be compared with the reference tool's output on a
 * machine without a GPU (tests/test_search_format.py computes the results with the oracle / host suffix array).
 *
 * stdin (whitespace separated):
 *   count offsets json sep            four integers (sep = 10 or 0)
 *   plen s0 s1 ...                    the pattern as alpha_t symbols
 *   nindexes
 *   per index:  first last ndocs
 *     per doc:  info_len  b0 b1 ... (info bytes as integers)  noffsets  o0 o1 ...
 */
#include "../integration/femto_search_format.h"

static long long rd(void) {
  long long v;
  if (scanf("%lld", &v) != 1) {
    fprintf(stderr, "search_format_check: malformed input\n");
    exit(2);
  }
  return v;
}

int main(void) {
  fs_options o;
  o.count = (int)rd();
  o.offsets = (int)rd();
  o.json = (int)rd();
  o.sep = (char)rd();
  const int plen = (int)rd();
  uint16_t* pat = (uint16_t*)malloc(((size_t)plen + 1) * sizeof(uint16_t));
  for (int i = 0; i < plen; i++) pat[i] = (uint16_t)rd();
  const int nidx = (int)rd();
  fs_index_results* r = (fs_index_results*)calloc((size_t)nidx + 1, sizeof(fs_index_results));
  for (int x = 0; x < nidx; x++) {
    r[x].first = rd();
    r[x].last = rd();
    const int64_t nd = r[x].ndocs = rd();
    unsigned char** info = (unsigned char**)calloc((size_t)nd + 1, sizeof(unsigned char*));
    int64_t* info_len = (int64_t*)calloc((size_t)nd + 1, sizeof(int64_t));
    int64_t* start = (int64_t*)calloc((size_t)nd + 2, sizeof(int64_t));
    int64_t cap = 16, n = 0;
    int64_t* off = (int64_t*)malloc((size_t)cap * sizeof(int64_t));
    for (int64_t k = 0; k < nd; k++) {
      info_len[k] = rd();
      info[k] = (unsigned char*)malloc((size_t)info_len[k] + 1);
      for (int64_t b = 0; b < info_len[k]; b++) info[k][b] = (unsigned char)rd();
      const int64_t no = rd();
      start[k] = n;
      for (int64_t j = 0; j < no; j++) {
        if (n == cap) off = (int64_t*)realloc(off, (size_t)(cap *= 2) * sizeof(int64_t));
        off[n++] = rd();
      }
    }
    start[nd] = n;
    r[x].info = (const unsigned char* const*)info;
    r[x].info_len = info_len;
    r[x].off_start = start;
    r[x].off = off;
  }
  fs_print_report(stdout, &o, plen, pat, nidx, r);
  return 0;
}
