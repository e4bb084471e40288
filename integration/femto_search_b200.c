/* femto_search_b200.c -- femto_search's string-pattern queries on the GPU engine.
 *
 * The reference's search front end (src/main_cc/search_tool.cc) parses a query language with a
 * flex/bison parser; for a LITERAL pattern it offers --raw-pattern / --raw-pattern-from (:629-645),
 * which builds a plain string query without the parser.  That subset is this tool: the same command
 * line, the same report byte for byte (--count, documents, --offsets; --json; --null; --output;
 * several indexes), computed with libfemto_b200's count / locate / document calls:
 *
 *   femto_search_b200 [--count | --offsets] [--json] [--null] [--output FILE] [--device N]
 *                     <index_path> [<index_path>...] --raw-pattern PATTERN | --raw-pattern-from FILE
 *
 * Query forms that need femto's parser or NFA (a pattern argument in the query language, --matches,
 * --suggest, --icase, --grep, --filter-results) are refused: they are outside the backward-search path.
 * Like the reference (default_chunk_size, :546), a query reports the first 1 Mi rows of the pattern's range.
 */
#include <errno.h>
#include <sys/stat.h>

#include "femto_b200.h"
#include "femto_search_format.h"

static void usage(const char* name) {
  printf("Usage: %s [options] <index_path> [<index_path>...] --raw-pattern <pattern>\n", name);
  printf("  --count                     number of occurrences only\n");
  printf("  --offsets                   occurrences as offsets inside their documents (default: the documents)\n");
  printf("  --json                      report as JSON\n");
  printf("  --null                      end lines with a 0 byte instead of a newline\n");
  printf("  --output <file>             write the report there instead of stdout\n");
  printf("  --raw-pattern <text>        the literal pattern\n");
  printf("  --raw-pattern-from <file>   the literal pattern = the file's bytes\n");
  printf("  --device <n>                CUDA device (default 0)\n");
  exit(-1);
}

static void die(const char* what, int rc) {
  fprintf(stderr, "%s: error %d: %s\n", what, rc, fm_last_error());
  exit(-1);
}

typedef struct {
  int64_t doc, off;
} doc_off;

static int cmp_doc_off(const void* a, const void* b) {
  const doc_off* x = (const doc_off*)a;
  const doc_off* y = (const doc_off*)b;
  if (x->doc != y->doc) return x->doc < y->doc ? -1 : 1;
  return x->off < y->off ? -1 : x->off > y->off;
}

/* Everything fs_index_results points to, for one index. */
typedef struct {
  unsigned char** info;
  int64_t* info_len;
  int64_t* off_start;
  int64_t* off;
} owned_results;

static void fetch_infos(fm_index_t* ix, int64_t ndocs, const int64_t* docs, owned_results* w) {
  w->info = (unsigned char**)calloc((size_t)ndocs + 1, sizeof(unsigned char*));
  w->info_len = (int64_t*)calloc((size_t)ndocs + 1, sizeof(int64_t));
  for (int64_t k = 0; k < ndocs; k++) {
    int64_t len = 0;
    int rc = fm_doc_name(ix, docs[k], NULL, 0, &len);
    if (rc != FM_OK && rc != FM_ERR_FULL) die("fm_doc_name", rc);
    w->info[k] = (unsigned char*)malloc((size_t)len + 1);
    if (len && (rc = fm_doc_name(ix, docs[k], w->info[k], len, &len)) != FM_OK) die("fm_doc_name", rc);
    w->info_len[k] = len;
  }
}

static void query_index(const char* path, int device, const fs_options* o, int plen, const uint16_t* pat,
                        fs_index_results* r, owned_results* w) {
  fm_index_t* ix = NULL;
  int rc = fm_open(path, device, &ix);
  if (rc != FM_OK) die(path, rc);
  const uint16_t* pats[1] = {pat};
  int lens[1] = {plen};
  memset(r, 0, sizeof *r);
  memset(w, 0, sizeof *w);
  if ((rc = fm_count(ix, 1, lens, pats, &r->first, &r->last)) != FM_OK) die("fm_count", rc);
  if (!o->count && r->last >= r->first) {
    /* one query = the first FS_ROWS_PER_QUERY rows of the range (chunk_size of create_generic_ast_query) */
    int64_t last = r->last;
    if (last - r->first + 1 > FS_ROWS_PER_QUERY) last = r->first + FS_ROWS_PER_QUERY - 1;
    const int64_t nrows = last - r->first + 1;
    int64_t* docs = NULL;
    if (o->offsets) {
      int64_t* offs = (int64_t*)malloc((size_t)nrows * sizeof(int64_t));
      int64_t* doc = (int64_t*)malloc((size_t)nrows * sizeof(int64_t));
      int64_t* doff = (int64_t*)malloc((size_t)nrows * sizeof(int64_t));
      doc_off* both = (doc_off*)malloc((size_t)nrows * sizeof(doc_off));
      if (!offs || !doc || !doff || !both) die("out of memory", FM_ERR_MEM);
      if ((rc = fm_locate_range(ix, r->first, last, offs)) != FM_OK) die("fm_locate_range", rc);
      if ((rc = fm_resolve(ix, nrows, offs, doc, doff)) != FM_OK) die("fm_resolve", rc);
      for (int64_t i = 0; i < nrows; i++) {
        both[i].doc = doc[i];
        both[i].off = doff[i];
      }
      qsort(both, (size_t)nrows, sizeof(doc_off), cmp_doc_off); /* results are ordered by document, then offset */
      docs = (int64_t*)malloc((size_t)nrows * sizeof(int64_t));
      w->off_start = (int64_t*)malloc(((size_t)nrows + 1) * sizeof(int64_t));
      w->off = doff; /* reused for the sorted offsets */
      for (int64_t i = 0; i < nrows; i++) {
        if (i == 0 || both[i].doc != both[i - 1].doc) {
          docs[r->ndocs] = both[i].doc;
          w->off_start[r->ndocs++] = i;
        }
        w->off[i] = both[i].off;
      }
      w->off_start[r->ndocs] = nrows;
      free(offs);
      free(doc);
      free(both);
    } else {
      int64_t cap = nrows;
      fm_info_t info;
      if ((rc = fm_info(ix, &info)) != FM_OK) die("fm_info", rc);
      if (cap > info.num_documents) cap = info.num_documents;
      docs = (int64_t*)malloc(((size_t)cap + 1) * sizeof(int64_t));
      if ((rc = fm_range_documents(ix, r->first, last, docs, cap, &r->ndocs)) != FM_OK) die("fm_range_documents", rc);
    }
    fetch_infos(ix, r->ndocs, docs, w);
    free(docs);
    r->info = (const unsigned char* const*)w->info;
    r->info_len = w->info_len;
    r->off_start = w->off_start;
    r->off = w->off;
  }
  fm_close(ix);
}

int main(int argc, char** argv) {
  fs_options o = {0, 0, 0, '\n'};
  const char** index_paths = (const char**)malloc((size_t)argc * sizeof(char*));
  int num_indexes = 0, device = 0;
  const unsigned char* raw = NULL;
  size_t raw_len = 0;
  const char* output_fname = NULL;
  FILE* out = stdout;

  for (int i = 1; i < argc; i++) {
    const char* a = argv[i];
    const int has_value = i + 1 < argc;
    if (!strcmp(a, "--offsets")) o.offsets = 1;
    else if (!strcmp(a, "--count")) o.count = 1;
    else if (!strcmp(a, "--null")) o.sep = '\0';
    else if (!strcmp(a, "--json")) o.json = 1;
    else if (!strcmp(a, "--output") && has_value) output_fname = argv[++i];
    else if (!strcmp(a, "--device") && has_value) device = atoi(argv[++i]);
    else if (!strcmp(a, "--raw-pattern") && has_value) {
      raw = (const unsigned char*)argv[++i];
      raw_len = strlen(argv[i]);
    } else if (!strcmp(a, "--raw-pattern-from") && has_value) {
      FILE* f = fopen(argv[++i], "rb");
      long n = -1;
      unsigned char* buf = NULL;
      if (f && !fseek(f, 0, SEEK_END) && (n = ftell(f)) >= 0 && !fseek(f, 0, SEEK_SET)) {
        buf = (unsigned char*)malloc((size_t)n + 1);
        if (buf && fread(buf, 1, (size_t)n, f) != (size_t)n) n = -1;
      }
      if (!f || n < 0 || !buf) {
        printf("Could not read pattern from %s\n", argv[i]);
        exit(-1);
      }
      fclose(f);
      raw = buf;
      raw_len = (size_t)n;
    } else if (!strcmp(a, "--pattern") || !strcmp(a, "--pattern-from") || !strcmp(a, "--matches") ||
               !strcmp(a, "--suggest") || !strcmp(a, "--icase") || !strcmp(a, "--grep") || !strcmp(a, "--multigrep") ||
               !strcmp(a, "--filter-results") || !strcmp(a, "--max_results")) {
      fprintf(stderr, "%s needs femto's query parser / NFA, which this front end does not carry; use --raw-pattern\n", a);
      exit(-1);
    } else if (a[0] == '-') {
      printf("Unknown option %s\n", a);
      usage(argv[0]);
    } else {
      index_paths[num_indexes++] = a;
    }
  }
  if (o.json) o.sep = '\n';
  if (!raw) {
    fprintf(stderr, "a pattern in the query language needs femto's parser; pass the literal pattern with --raw-pattern\n");
    exit(-1);
  }
  if (num_indexes <= 0) usage(argv[0]);
  for (int i = 0; i < num_indexes; i++) {
    struct stat st;
    if (stat(index_paths[i], &st) != 0) {
      printf("Could not open index at %s\n", index_paths[i]);
      exit(-1);
    }
  }
  if (output_fname) {
    out = fopen(output_fname, "w");
    if (!out) {
      perror("Could not fopen");
      fprintf(stderr, "Could not open output filename '%s' for writing\n", output_fname);
      exit(-1);
    }
  }

  /* the pattern as alpha_t symbols (construct_buf_string: CHARACTER_OFFSET + byte) */
  uint16_t* pat = (uint16_t*)malloc((raw_len + 1) * sizeof(uint16_t));
  for (size_t i = 0; i < raw_len; i++) pat[i] = (uint16_t)(FS_CHARACTER_OFFSET + raw[i]);

  fs_index_results* res = (fs_index_results*)calloc((size_t)num_indexes, sizeof(fs_index_results));
  owned_results* own = (owned_results*)calloc((size_t)num_indexes, sizeof(owned_results));
  for (int i = 0; i < num_indexes; i++) query_index(index_paths[i], device, &o, (int)raw_len, pat, &res[i], &own[i]);
  fs_print_report(out, &o, (int)raw_len, pat, num_indexes, res);
  if (out != stdout) fclose(out);
  return 0;
}
