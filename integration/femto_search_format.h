/* femto_search_format.h -- the report femto_search prints for a plain-string query, as text.
 *
 * The reference's search front end (src/main_cc/search_tool.cc) prints three kinds of report for a
 * string pattern: --count (one row per matching pattern and a total, :1018-1110), documents (the
 * default) and --offsets (print_matches, :352-530), each as plain lines or as --json, with '\n' or
 * --null separators.  This header restates those formats over plain arrays, with no engine behind it,
 * so that integration/femto_search_b200.c (results from the GPU engine) and the CPU check in
 * tests/test_search_format.py (results from the oracle) print through the same code, and both are
 * compared byte for byte with the reference tool's own output (tests/golden).
 */
#ifndef FEMTO_SEARCH_FORMAT_H
#define FEMTO_SEARCH_FORMAT_H

#include <ctype.h>
#include <inttypes.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define FS_CHARACTER_OFFSET 5 /* alpha_t = 5 + byte, src/main/index_types.h:64-68 */
#define FS_GLOM_CHAR '|'      /* separates glommed document names in an info string, search_tool.cc:72 */
#define FS_ROWS_PER_QUERY (1024 * 1024) /* default_chunk_size: the rows of a range one query reports, :546 */

typedef struct {
  int count;   /* --count */
  int offsets; /* --offsets (else documents) */
  int json;    /* --json */
  char sep;    /* '\n', or 0 with --null */
} fs_options;

/* Results of one index for the pattern. */
typedef struct {
  int64_t first, last;              /* the pattern's BWT rows; first > last = no match */
  int64_t ndocs;                    /* documents holding reported rows, ascending document number */
  const unsigned char* const* info; /* [ndocs] the documents' info bytes */
  const int64_t* info_len;          /* [ndocs] */
  const int64_t* off_start;         /* --offsets: offsets of document k = off[off_start[k] .. off_start[k+1]), */
  const int64_t* off;               /*            ascending, relative to the document's start */
} fs_index_results;

/* alphatos, src/main/index_types.h:103-119; returns the number of characters written (<= 6) */
static inline int fs_alphatos(char* dst, unsigned alpha) {
  int ch = (int)alpha - FS_CHARACTER_OFFSET;
  if (ch < 0) return sprintf(dst, "\\x-%02x", -ch);
  if (ch == '\\' || ch == '"') return sprintf(dst, "\\%c", ch);
  if (isprint(ch)) return sprintf(dst, "%c", ch);
  return sprintf(dst, "\\x%02x", ch);
}

/* encode_ch_json, src/main/json.c:34-60 (ch as the caller's type delivered it: fprint_cstr_json passes a
 * plain char, so a byte >= 0x80 arrives negative there and prints as ￿ffxx) */
static inline void fs_json_ch(FILE* f, int ch) {
  if (ch == '"') fputs("\\\"", f);
  else if (ch == '\\') fputs("\\\\", f);
  else if (ch >= 0 && isprint(ch)) fputc(ch, f);
  else fprintf(f, "\\u%04x", ch);
}

static inline void fs_print_alpha(FILE* f, int len, const uint16_t* pat) { /* fprint_alpha, index_types.h:121-131 */
  char buf[8];
  for (int i = 0; i < len; i++) {
    fs_alphatos(buf, pat[i]);
    fputs(buf, f);
  }
}

static inline void fs_print_alpha_json(FILE* f, int len, const uint16_t* pat) { /* fprint_alpha_json, json.c:81-90 */
  char buf[8];
  for (int i = 0; i < len; i++) {
    fs_alphatos(buf, pat[i]);
    for (const char* p = buf; *p; p++) fs_json_ch(f, *p);
  }
}

/* ast_to_string(string node, 0, usequotes = 1), src/main/ast.c:878-899, 1048-1069: the pattern between double
 * quotes, '"' escaped, other bytes as they are when graphic or a space, else \xNN.  malloc()ed. */
static inline char* fs_echo_query(int len, const uint16_t* pat) {
  char* s = (char*)malloc(5 * (size_t)len + 4);
  size_t n = 0;
  if (!s) return NULL;
  s[n++] = '"';
  for (int i = 0; i < len; i++) {
    int chr = (int)pat[i] - FS_CHARACTER_OFFSET;
    if (chr == '"') {
      s[n++] = '\\';
      s[n++] = (char)chr;
    } else if (chr > 0 && (isgraph(chr) || chr == ' ')) {
      s[n++] = (char)chr;
    } else if (chr < 0) {
      n += (size_t)sprintf(s + n, "\\x-%02x", -chr);
    } else {
      n += (size_t)sprintf(s + n, "\\x%02x", chr);
    }
  }
  s[n++] = '"';
  s[n] = 0;
  return s;
}

/* print_document_group_json, search_tool.cc:213-226 */
static inline void fs_print_document_group_json(FILE* out, int64_t len, const unsigned char* info) {
  fputc('"', out);
  for (int64_t k = 0; k < len; k++) {
    if (info[k] == FS_GLOM_CHAR) fputs("\",\"", out);
    else fs_json_ch(out, info[k]);
  }
  fputc('"', out);
}

/* print_matches without grep, search_tool.cc:352-530, for one index; *first as the reference's first_match */
static inline void fs_print_matches(FILE* out, const fs_options* o, const fs_index_results* r, int* first) {
  int first_keep = 1;
  int64_t printed = 0;
  for (int64_t k = 0; k < r->ndocs; k++) {
    const int64_t n_off = o->offsets ? r->off_start[k + 1] - r->off_start[k] : 1; /* documents: one result each */
    if (n_off <= 0) continue;
    if (o->json) {
      if (!first_keep) fputs("] ],\n   ", out);
      else if (!*first) fputs(",\n   ", out);
      *first = 0;
      fputs("[ [", out);
      fs_print_document_group_json(out, r->info_len[k], r->info[k]);
      fputs("], [", out);
    } else {
      if (!first_keep) fputc(o->sep, out);
      fwrite(r->info[k], 1, (size_t)r->info_len[k], out);
      if (o->offsets) fprintf(out, "%c\t", o->sep);
    }
    first_keep = 0;
    if (o->offsets) {
      for (int64_t j = 0; j < n_off; j++) {
        const int64_t off = r->off[r->off_start[k] + j];
        if (o->json) fprintf(out, j ? ", %" PRIi64 : "%" PRIi64, off);
        else fprintf(out, " %" PRIi64, off);
      }
    }
    printed += n_off;
  }
  if (printed) {
    if (o->json) fputs(" ] ] ", out);
    else fputc(o->sep, out);
  }
}

/* The whole report of femto_search for ONE string pattern over nindexes indexes (search_tool.cc:889-1112). */
static inline void fs_print_report(FILE* out, const fs_options* o, int plen, const uint16_t* pat, int nindexes,
                                   const fs_index_results* r) {
  int first_match = 1;
  if (o->json) {
    char* echo = fs_echo_query(plen, pat);
    fputs("{\n \"pattern\":\"", out);
    for (const char* p = echo ? echo : ""; *p; p++) fs_json_ch(out, *p); /* fprint_cstr_json: plain char */
    fputs("\",\n \"results\":[\n   ", out);
    free(echo);
  }
  if (o->count) {
    /* the same pattern in every index: one row with the matches added up (matchcmp groups them, :1018-1040) */
    int64_t total = 0;
    for (int i = 0; i < nindexes; i++)
      if (r[i].last >= r[i].first) total += r[i].last - r[i].first + 1;
    if (total > 0) {
      if (o->json) {
        fputs("[\"", out);
        fs_print_alpha_json(out, plen, pat);
        fprintf(out, "\", %" PRIi64 "]", total);
      } else {
        fprintf(out, "% 4" PRIi64 " \"", total);
        fs_print_alpha(out, plen, pat);
        fprintf(out, "\"%c", o->sep);
      }
    }
    if (o->json) fprintf(out, " ],\n \"total\":%" PRIi64, total);
    else fprintf(out, "% 4" PRIi64 " total matches%c", total, o->sep);
  } else {
    for (int i = 0; i < nindexes; i++) fs_print_matches(out, o, &r[i], &first_match);
    if (o->json) fputs(" ]", out);
  }
  if (o->json) fputs("\n}\n", out);
}

#endif /* FEMTO_SEARCH_FORMAT_H */
