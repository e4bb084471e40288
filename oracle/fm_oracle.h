/* oracle/fm_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the reference's (femto-dev/femto) FM-index read path:
 * index open, C[] / block_occs / bucket_occs lookup, Huffman-shaped wavelet-tree
 * Occ / rank over RLE-gamma or raw 512-bit segments, mark-table test, sampled-SA
 * read, backward search (count), backward LF walk (locate) and document extract.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product (femto_b200/) never does.
 *
 * PARITY STATUS: pinned.  tests/test_oracle_pin.py checks every function here against
 * (a) the reference's own golden values from src/main/index_test.c:595-725 and
 * (b) the unmodified reference compiled into oracle/_ref/libfemto_ref.so, on indexes
 * built by the reference's own builder (exhaustive Occ for every row x symbol, count,
 * locate, LF walk), plus committed fixtures under tests/golden/.
 */
#ifndef FM_ORACLE_H
#define FM_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fmo_index fmo_index;

/* error codes follow the reference's err_code_t numbering (src/utils/error.h:25-39) */
enum { FMO_OK = 0, FMO_ERR_MEM = 1, FMO_ERR_IO = 2, FMO_ERR_PARAM = 3, FMO_ERR_FORMAT = 4,
       FMO_ERR_BZ_DATA = 5, FMO_ERR_INVALID = 6 };

fmo_index* fmo_open(const char* path, int* err_out);
void fmo_close(fmo_index* ix);

/* info[0..6] = nblocks, total_length, ndocs, block_size, bucket_size, mark_period, chunk_size */
int fmo_header_info(const fmo_index* ix, int64_t* info);

int fmo_C(const fmo_index* ix, int ch, int64_t* out);
/* C[ch] + Occ(ch,row) and Occ(ch,row) alone (row is a 0-based BWT row, counted inclusively) */
int fmo_occ(fmo_index* ix, int ch, int64_t row, int64_t* c_plus_occ, int64_t* occ_only);
/* one LF step with mark test: ch=L[row], next=LF(row) or -1 if ch<=SEOF, offset=SA[row] or -1 */
int fmo_back_step(fmo_index* ix, int64_t row, int* ch, int64_t* next_row, int64_t* offset);

int fmo_count(fmo_index* ix, int npats, const int32_t* plen, const uint16_t* flat,
              const int64_t* offs, int64_t* first, int64_t* last);
int fmo_locate(fmo_index* ix, int npats, const int32_t* plen, const uint16_t* flat,
               const int64_t* offs, int max_occs_each, int32_t* noccs, int64_t* out_start,
               int64_t* out, int64_t out_cap);
int fmo_locate_range(fmo_index* ix, int64_t first, int64_t last, int64_t* offsets);
int fmo_doc_info(const fmo_index* ix, int64_t doc, int64_t* doc_len, int64_t* eof_row);
int fmo_resolve(const fmo_index* ix, int64_t offset, int64_t* doc, int64_t* doc_off);
/* the info bytes stored with a document (its name); *info points into the index, valid until fmo_close */
int fmo_doc_name(const fmo_index* ix, int64_t doc, const unsigned char** info, int64_t* len);
/* out must hold doc_len-1 symbols (alphabet values, i.e. 5+byte); see server.c:6364-6437 */
int fmo_extract(fmo_index* ix, int64_t doc, uint16_t* out, int64_t out_cap, int64_t* out_len);

/* bseq level (wtree.c:635): 1-based index; returns zeros/ones at or before index and the bit */
void fmo_bseq_rank(const unsigned char* zdata, int index1, int* occ0, int* occ1, int* bit);

/* instrumentation: bytes of the on-disk layout dereferenced by the calls so far
 * (SURVEY.md section 8d "algorithmic bytes"), and number of Occ evaluations */
void fmo_counters(const fmo_index* ix, int64_t* bytes, int64_t* occ_calls, int64_t* levels);
void fmo_reset_counters(fmo_index* ix);

#ifdef __cplusplus
}
#endif
#endif
