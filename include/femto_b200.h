/* femto_b200.h -- C ABI of the B200-native FM-index query engine.
 *
 * Drop-in boundary for the backward-search hot path of femto-dev/femto: the batch
 * functions of the reference's src/main/femto_internal.h:63-74 (parallel_count,
 * parallel_locate, parallel_locate_range), the server/index lifecycle around them
 * (src/main/femto.c:54-81, 269-272) and document extraction
 * (src/main/server.c:6364-6437), over the reference's UNCHANGED on-disk index
 * (directory of block files or flattened single file).
 *
 * Plain pointers and sizes only: no C++, CUDA or torch types cross this boundary.
 * Every entry point returns an fm_err_t whose numbering equals the reference's
 * err_code_t (src/utils/error.h:25-39); fm_last_error() returns a thread-local
 * message (the reference's error ring, src/utils/error.c:30-57, is process-global
 * and not thread-safe, so it is not reproduced).
 *
 * Symbols are alpha_t-compatible: uint16_t holding 5+byte for text bytes, 1..4 for
 * the escape codes (src/main/index_types.h:42-68).  Rows and offsets are int64_t.
 * "No match" is first > last, not an error (src/main/server.c:832-841).
 *
 * The library has NO CPU fallback for the query path: without a usable CUDA device
 * fm_open() fails with FM_ERR_IO and says why.
 */
#ifndef FEMTO_B200_H
#define FEMTO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FM_ALPHA_SIZE 261        /* src/main/index_types.h:64-65 */
#define FM_CHARACTER_OFFSET 5

typedef enum {                   /* == err_code_t, src/utils/error.h:25-39 */
  FM_OK = 0,
  FM_ERR_MEM = 1,
  FM_ERR_IO = 2,
  FM_ERR_PARAM = 3,
  FM_ERR_FORMAT = 4,
  FM_ERR_BZ_DATA = 5,
  FM_ERR_INVALID = 6,
  FM_ERR_PTHREADS = 7,
  FM_ERR_MISSING = 8,
  FM_ERR_CANCELED = 9,
  FM_ERR_FULL = 10,
  FM_ERR_OVERWORKED = 11,
  FM_ERR_UNKNOWN = 12
} fm_err_t;

typedef struct fm_index fm_index_t;   /* replaces (femto_server_t, index_locator_t) */

typedef struct {
  int64_t total_length;      /* rows of the BWT = indexed symbols incl. one SEOF per document */
  int64_t num_documents;
  int64_t num_blocks;        /* data blocks in the index (all of them, not only this shard's) */
  int32_t block_size;        /* rows per block   (index_block_param_t, src/main/index.h:103-121) */
  int32_t bucket_size;       /* rows per bucket */
  int32_t mark_period;
  int32_t chunk_size;
  int64_t first_row;         /* rows [first_row, end_row) are resident on this device */
  int64_t end_row;
  int64_t hbm_bytes;         /* device memory held by the index image */
  int64_t rank_block_bytes;  /* of which wavelet-tree rank blocks */
  int32_t device;            /* CUDA device ordinal */
  int32_t max_code_len;      /* deepest wavelet-tree leaf over all resident buckets */
  int32_t rank_block_size;   /* bytes per rank block of the HBM image: 128, 64 or 32 */
  int32_t levels_per_block;  /* wavelet-tree levels one rank block read answers: 1, 2 or 4 */
} fm_info_t;

/* --------------------------------------------------------------------------
 * Lifecycle.  fm_open == femto_start_server_err + femto_loc_for_path_err
 * (src/main/femto.c:54-69, 269-272): parses the index at `path`, validates the
 * block headers exactly as read_block_header (src/main/index.c:1348-1404), decodes
 * every bucket's map/Huffman tables (b_fault, index.c:1222-1342) and wavelet-tree
 * segments once, and uploads the rank image to `device`'s HBM.
 * fm_close == femto_stop_server (femto.c:77-81).
 * fm_open_shard loads one of nshards contiguous ranges of data blocks (BWT row range sharding at the
 * reference's partition unit): block b belongs to the shard its middle row falls
 * into when the rows are cut into nshards equal parts, i.e. shard = min(nshards - 1,
 * (2 b + 1) * block_size * nshards / (2 * total_length)) -- balanced by rows, not by block count (the
 * last block of an index is usually a few rows long). */
int fm_open(const char* path, int device, fm_index_t** out);
int fm_open_shard(const char* path, int device, int shard, int nshards, fm_index_t** out);
void fm_close(fm_index_t* ix);
int fm_info(const fm_index_t* ix, fm_info_t* out);
const char* fm_last_error(void);

/* --------------------------------------------------------------------------
 * count: backward search.  Mirrors parallel_count (src/main/femto.c:275-329):
 * for pattern i, [first[i], last[i]] is the BWT row range of its occurrences
 * (first > last when there is none); if last == NULL, first[i] receives the count
 * (femto.c:313-318).  plen[i] == 0 yields [0, total_length-1] (server.c:782-808).
 * Host buffers; the call blocks until results are in first/last.  Re-entrant. */
int fm_count(fm_index_t* ix, int npats, const int* plen, const uint16_t* const* pats,
             int64_t* first, int64_t* last);

/* Same operation with the patterns in one flat buffer: pattern i is
 * flat[offs[i] .. offs[i]+plen[i]).  This is the form the engine works on;
 * fm_count() gathers into it.
 * Batches of >= 128 Ki patterns that lie in `flat` in batch order are STREAMED: the kernel is
 * launched at once and takes patterns from its queue as the copy stream delivers them, the first
 * results travel back while the last patterns are searched; a batch of equal-length, densely
 * packed patterns travels without plen / offs.  Pinned buffers (fm_host_alloc) get the full
 * PCIe rate.  Environment: FEMTO_B200_NO_STREAM=1 keeps copies and kernel in series (what a
 * serialising profiler needs; the call also falls back to that by itself when the kernel sees no
 * data arrive for ~0.1 s), FEMTO_B200_TRACE=1 prints the timeline of each call to stderr. */
int fm_count_flat(fm_index_t* ix, int64_t npats, const int32_t* plen, const uint16_t* flat,
                  const int64_t* offs, int64_t* first, int64_t* last);

/* The same call for patterns given as raw text bytes -- what the reference's batch tool reads from its
 * pattern file before widening every byte to an alpha_t (read_queries, src/main/query_tool.c:48-98;
 * strtoalpha, src/main/index_types.h:85-97): pattern i is text[offs[i] .. offs[i]+plen[i]), its symbols
 * are FM_CHARACTER_OFFSET + byte.  Half the host->device bytes of fm_count_flat; the count kernel adds
 * the offset as it reads (quad image, default schedule; otherwise the bytes are widened on the host).
 * Escape codes cannot be expressed in this form -- use fm_count_flat for those. */
int fm_count_bytes(fm_index_t* ix, int64_t npats, const int32_t* plen, const uint8_t* text, const int64_t* offs,
                   int64_t* first, int64_t* last);

/* Device-resident form: all pointers are device pointers on ix's device, `stream`
 * is a cudaStream_t (NULL = default stream).  Asynchronous: returns after enqueueing.
 * flat_len = number of symbols in d_flat. */
int fm_count_device(fm_index_t* ix, int64_t npats, const int32_t* d_plen, const uint16_t* d_flat,
                    const int64_t* d_offs, int64_t* d_first, int64_t* d_last, void* stream);

/* Range-sharded count (index opened with fm_open_shard).  A batch is a set
 * of pattern STATES that travel between the GPUs owning the BWT rows they need next; this call
 * advances every state while its rows are resident on ix's device.  d_state holds nstates rows of
 * 6 int64 {pattern id, first, last, i, C[c]+Occ(c,first-1) when known, phase | home_rank<<4};
 * phase 3 = new (initialised here from the pattern), 0 = needs Occ(c,first-1), 1 = needs
 * Occ(c,last), 2 = finished.  d_dest[k] receives the rank that must see state k next (its home
 * rank once finished).  The whole pattern batch (d_plen/d_flat/d_offs, indexed by pattern id) is
 * replicated on every rank.  All pointers are device pointers; asynchronous on `stream`.
 * femto_b200/sharded.py drives the exchange between ranks with NCCL all-to-all. */
int fm_count_shard_step(fm_index_t* ix, int64_t nstates, int64_t* d_state, const int32_t* d_plen,
                        const uint16_t* d_flat, const int64_t* d_offs, int32_t* d_dest, int nshards,
                        void* stream);

/* Range-sharded locate: the sampled-SA walk of do_back_query / do_context_query
 * (src/main/server.c:2228-2359, 2627-2795) over an index opened with fm_open_shard.  The LF
 * mapping sends a row to an arbitrary shard, so -- as for count -- the walk's STATE travels:
 * d_state holds nstates rows of 4 int64 {result slot at the home rank, BWT row (the text offset
 * once finished, -1 on a malformed index), LF steps taken so far, phase | home_rank<<4}; phase
 * 0 = walking, 2 = finished.  This call follows LF from every state while its row is resident on
 * ix's device and a mark has not been reached; d_dest[k] receives the rank that must see state k
 * next (the owner of its row, or its home rank once finished).  Device pointers; asynchronous on
 * `stream`.  femto_b200/sharded.py (sharded_locate_rows / sharded_locate) drives the exchange. */
int fm_locate_shard_step(fm_index_t* ix, int64_t nstates, int64_t* d_state, int32_t* d_dest, int nshards,
                         void* stream);

/* --------------------------------------------------------------------------
 * Device-initiated exchange for the range-sharded index ("mesh").  One fm_mesh_t per rank (GPU),
 * bound to an index opened with fm_open_shard(path, dev, rank, world).  Every rank runs ONE
 * persistent kernel per batch; a pattern's 32-byte state is stored straight into the inbox of the
 * GPU owning the BWT row its next Occ needs, over NVLink peer memory, and returns to its home
 * rank with [first, last] -- no host round trip and no collective inside a batch
 * (femto_b200/csrc/fm_mesh.cuh).  Partition unit = the reference's data block
 * (src/main/index.h:83-100); the computation per state is do_string_query's (src/main/server.c:713-946).
 *
 * Set-up (collective): fm_mesh_create on every rank with the same world / window / cap_log2;
 * exchange the FM_MESH_HANDLE_BYTES handles of fm_mesh_export between the processes (any transport;
 * femto_b200/sharded.py uses torch.distributed) and pass all of them, in rank order, to
 * fm_mesh_connect -- or, for ranks living in ONE process, fm_mesh_connect_local with the peers'
 * fm_mesh_t.  window = patterns of its own batch a rank keeps in flight (0 = default); cap_log2 = 0
 * sizes the inbox rings from it.
 *
 * Per batch (collective, asynchronous on `stream`): the pattern batch of ALL ranks is replicated on
 * every rank (d_plen / d_flat / d_offs indexed by global pattern id; uniform_len > 0: pattern p is
 * d_flat[p * uniform_len ...) and d_plen / d_offs are not read); this rank's patterns are ids
 * [pid_lo, pid_lo + n_mine), their results go to d_first / d_last [0, n_mine) (d_last NULL: counts).
 * A rank must not launch batch k+1 before every rank has finished batch k-1 (the all-gather that
 * replicates the patterns provides that).  fm_mesh_finish waits for the stream and reports the
 * kernel's status (FM_ERR_CANCELED: it gave up waiting for the other ranks) and counters
 * {states sent, received, evaluation rounds, Occ pairs, single Occ, 0 (unused), patterns
 * injected, 0}. */
typedef struct fm_mesh fm_mesh_t;
#define FM_MESH_HANDLE_BYTES 64
int fm_mesh_create(fm_index_t* ix, int rank, int world, int64_t window, int cap_log2, fm_mesh_t** out);
void fm_mesh_destroy(fm_mesh_t* m);
int fm_mesh_export(fm_mesh_t* m, void* handle, int64_t handle_bytes);
int fm_mesh_connect(fm_mesh_t* m, const void* handles, int64_t handle_stride);
int fm_mesh_connect_local(fm_mesh_t* m, fm_mesh_t* const* peers);
/* max_ctas > 0 bounds the kernel's grid (several meshes sharing one GPU must all be resident at
 * once); timeout_seconds > 0 replaces the 10 s a kernel waits for the other ranks before giving up. */
int fm_mesh_set_limits(fm_mesh_t* m, int max_ctas, double timeout_seconds);
int fm_mesh_count(fm_mesh_t* m, const int32_t* d_plen, const uint16_t* d_flat, const int64_t* d_offs,
                  int uniform_len, int64_t pid_lo, int64_t n_mine, int64_t* d_first, int64_t* d_last,
                  void* stream);
/* Sampled-SA walks (parallel_locate_range's per-row work, src/main/femto.c:481-536) for this rank's
 * nrows global BWT rows; d_offsets[k] = SA[d_rows[k]].  Collective like fm_mesh_count. */
int fm_mesh_locate_rows(fm_mesh_t* m, int64_t nrows, const int64_t* d_rows, int64_t* d_offsets, void* stream);
int fm_mesh_finish(fm_mesh_t* m, void* stream, int* status, uint64_t* stats8);

/* --------------------------------------------------------------------------
 * locate.  Mirrors parallel_locate (src/main/femto.c:331-399): for pattern i,
 * noccs[i] offsets are returned in BWT row order first..; offsets[i] is malloc()ed
 * by the callee and freed by the caller with free() (NULL when noccs[i]==0).
 * The max_occs clip reproduces do_locate_query (src/main/server.c:4411-4415):
 * all rows when last-first <= max_occs, else the first max_occs rows. */
int fm_locate(fm_index_t* ix, int npats, const int* plen, const uint16_t* const* pats,
              int max_occs_each, int* noccs, int64_t** offsets);

/* Flat form: results of pattern i are out[out_start[i] .. out_start[i]+noccs[i]).
 * Fails with FM_ERR_FULL (noccs/out_start still filled) when out_cap is too small. */
int fm_locate_flat(fm_index_t* ix, int64_t npats, const int32_t* plen, const uint16_t* flat,
                   const int64_t* offs, int max_occs_each, int32_t* noccs, int64_t* out_start,
                   int64_t* out, int64_t out_cap);

/* parallel_locate_range (src/main/femto.c:481-536): text offsets SA[first..last];
 * offsets must have room for last-first+1 entries. */
int fm_locate_range(fm_index_t* ix, int64_t first, int64_t last, int64_t* offsets);

/* Arbitrary rows (host buffers): offsets[i] = SA[rows[i]]. */
int fm_locate_rows(fm_index_t* ix, int64_t nrows, const int64_t* rows, int64_t* offsets);
int fm_locate_rows_device(fm_index_t* ix, int64_t nrows, const int64_t* d_rows, int64_t* d_offsets,
                          void* stream);

/* Status of the caller-stream walk launches (fm_locate_rows_device, fm_locate_shard_step) enqueued on
 * `stream` so far: waits for the stream, returns 0 or the code of a malformed walk (1 row not
 * resident / out of range, 2 symbol or occurrence number out of range, 3 unmarked document start)
 * and clears it.  The host-buffer calls keep their own status word and report through their return
 * value. */
int fm_take_status(fm_index_t* ix, void* stream, int* status);

/* --------------------------------------------------------------------------
 * Single LF step with mark test = do_back_query (src/main/server.c:2228-2359):
 * ch[i] = L[rows[i]], next[i] = LF(rows[i]) or -1 when ch <= ESCAPE_CODE_SEOF,
 * offset[i] = SA[rows[i]] when the row is marked, else -1.  Host buffers. */
int fm_back_step(fm_index_t* ix, int64_t nrows, const int64_t* rows, int32_t* ch, int64_t* next,
                 int64_t* offset);

/* C[ch] + Occ(ch,row) for each (ch,row): one half of a backward-search step, i.e.
 * header_occs_request(HDR_BACK...) + block_request(BLOCK_REQUEST_OCCS)
 * (src/main/index.c:1698-1765, 1973-2100).  Host buffers. */
int fm_occ(fm_index_t* ix, int64_t n, const uint16_t* ch, const int64_t* rows, int64_t* c_plus_occ);

/* One backward-search step for a batch of (range, symbol) triples -- the reference's
 * backward_search_query (src/main/server.c:948-1122), the primitive its regexp / approximate-match
 * NFA simulation issues per frontier edge (server.h:518-564):
 *   new_first[i] = C[ch] + Occ(ch, first[i]-1)   (C[ch] when first[i]==0)
 *   new_last[i]  = C[ch] + Occ(ch, last[i]) - 1
 * Host buffers; ranges must satisfy 0 <= first <= last+1, last < total_length. */
int fm_backward_step(fm_index_t* ix, int64_t n, const int64_t* first, const int64_t* last,
                     const uint16_t* ch, int64_t* new_first, int64_t* new_last);

/* --------------------------------------------------------------------------
 * Documents.  fm_doc_info: length (incl. its SEOF) and EOF row (header tables,
 * src/main/index.c:1668-1696).  fm_resolve: text offset -> (document, offset in
 * document) as resolve_location (index.c:1587-1611).  fm_extract: the doc_len-1
 * symbols of the document, as do_extract_document_query's backward context
 * (src/main/server.c:6364-6437). */
int fm_doc_info(const fm_index_t* ix, int64_t doc, int64_t* doc_len, int64_t* eof_row);
int fm_resolve(const fm_index_t* ix, int64_t n, const int64_t* offsets, int64_t* doc, int64_t* doc_off);
/* The info bytes stored with a document at build time (its name / URL): document_info,
 * src/main/index.c:1767-1784 (header_loc_request with HDR_LOC_REQUEST_DOC_INFO).  *out_len receives
 * the length; FM_ERR_FULL when it exceeds out_cap. */
int fm_doc_name(const fm_index_t* ix, int64_t doc, void* out, int64_t out_cap, int64_t* out_len);
/* The document chunk that holds BWT row `row`: block_chunk_request with BLOCK_CHUNK_FIND_NUMBER |
 * BLOCK_CHUNK_REQUEST_DOCUMENTS (src/main/index.c:2147-2236).  chunk_first / chunk_last receive the
 * rows the chunk covers (chunk_size rows, fewer at the end of a data block), docs the ascending
 * numbers of the documents that hold the suffix of at least one of them -- read from the index's
 * own chunk section, decoded on the host, no kernel involved.  FM_ERR_MISSING when the index was
 * built without chunks, FM_ERR_FULL when there are more than docs_cap documents (*ndocs says how many). */
int fm_chunk_documents(const fm_index_t* ix, int64_t row, int64_t* chunk_first, int64_t* chunk_last, int64_t* docs,
                       int64_t docs_cap, int64_t* ndocs);
/* Documents that contain the suffixes of BWT rows first..last, ascending and unique: what the
 * reference's range_to_results query delivers for RESULT_TYPE_DOCUMENTS (src/main/server.c:4549-4889),
 * computed the same way: the stored document lists of the chunks that lie inside the range, plus the
 * rows of the (at most two) chunks that stick out of it located on the GPU and resolved with the
 * header's document table; an index built without chunks locates every row.  *ndocs
 * receives the number of documents; FM_ERR_FULL when it exceeds docs_cap. */
int fm_range_documents(fm_index_t* ix, int64_t first, int64_t last, int64_t* docs, int64_t docs_cap, int64_t* ndocs);
int fm_extract(fm_index_t* ix, int64_t doc, uint16_t* out, int64_t out_cap, int64_t* out_len);
/* Several documents in ONE launch (a document is a strictly sequential chain of LF steps, so throughput
 * comes from extracting many side by side): the symbols of docs[0], docs[1], ... back to back in out;
 * out_start (ndocs + 1 entries) receives where each document begins, out_start[ndocs] the total.
 * FM_ERR_FULL (out_start filled) when the total exceeds out_cap. */
int fm_extract_batch(fm_index_t* ix, int64_t ndocs, const int64_t* docs, uint16_t* out, int64_t out_cap,
                     int64_t* out_start);

/* --------------------------------------------------------------------------
 * Generic requests: femto_create_generic_request + femto_begin_request + femto_wait_request +
 * femto_response_for_generic_request (src/main/femto.h:75-149, src/main/femto.c:566-1000) in one blocking
 * call, for the requests that are batches of backward searches:
 *   "string_rows B B ..."        -> {"range":[first,last]}
 *   "string_rows_left B B ..."   -> the ranges of c+pattern for every symbol c of the alphabet
 *   "string_rows_right B B ..."  -> the ranges of pattern+c
 *   "string_rows_all B B ..."    -> both
 * (B = byte values as integers).  *response is malloc()ed, freed by the caller, and byte-identical to the
 * reference's answer.  find_strings / find_docs / docs_for_range need femto's query parser and result
 * encoder and are refused with FM_ERR_INVALID.  integration/femto_request_b200.c is the reference's
 * femto_handle_request tool (src/main/handle_request.c) on top of it. */
int fm_generic_request(fm_index_t* ix, const char* request, char** response);

/* --------------------------------------------------------------------------
 * Pinned host memory for callers that want zero-staging transfers. */
void* fm_host_alloc(size_t bytes);
void fm_host_free(void* p);

/* Counters of the engine's own kernel launches since fm_open (all entry points). */
int64_t fm_kernel_launches(const fm_index_t* ix);
/* Bytes the most recent fm_count / fm_count_flat call copied host->device and device->host
 * (a batch of equal-length, densely packed patterns travels without its length / offset arrays). */
int fm_last_transfer(const fm_index_t* ix, int64_t* h2d_bytes, int64_t* d2h_bytes);

/* Instrumented count (not a timed path): runs the same batch through a counter-carrying variant
 * of the count kernel and returns stats8 = { rank blocks requested, distinct rank blocks per
 * step and level (the two Occ of a step often share a block), Occ evaluations, backward-search
 * steps, lane groups present in descent iterations, of which had a block to evaluate (quad image:
 * the warp-divergence measure of mixed-length batches and deep codes), 0, 0 }.  bench.py derives the
 * kernel's algorithmic HBM bytes from these. */
int fm_count_stats(fm_index_t* ix, int64_t npats, const int32_t* plen, const uint16_t* flat,
                   const int64_t* offs, uint64_t* stats8);

/* Instrumented sampled-SA walks (not a timed path): SA[rows[i]] for nrows host rows through the walk
 * kernel with its counters on; stats4 = { LF steps, wavelet-tree rank blocks read, mark bit-vector
 * blocks read, SA samples read }.  bench.py derives the locate leg's roofline from these. */
int fm_walk_stats(fm_index_t* ix, int64_t nrows, const int64_t* rows, uint64_t* stats4);

/* Measurement aid: the ceiling the count kernel is compared with in bench.py.  Runs `steps`
 * rounds of dependent, uniformly random, naturally aligned reads of bytes_per_access (32, 64 or
 * 128) bytes over the resident rank blocks with every SM filled, and reports how many reads were
 * made and the CUDA-event time of the launch.  Touches no query state. */
int fm_probe_random_reads(fm_index_t* ix, int bytes_per_access, int steps, int64_t* accesses, double* ms);

/* Tuning knob: lanes cooperating on one rank query in the locate/extract/occ kernels (1, 2, 4 or
 * 8; default 4).  The value is clamped to what the image's block layout offers: 4 or 8 lanes on
 * 128-byte blocks, 2 or 4 on 64-byte, 1 or 2 on 32-byte; paired-level blocks half of that. */
int fm_set_lanes_per_query(fm_index_t* ix, int lanes);

/* Tuning knob of the count kernel.  merged != 0 (default): one group of `lanes` lanes per pattern
 * advances both Occ of a step together and reads a shared rank block once; merged == 0: two
 * sub-groups of `lanes` lanes per pattern, one per Occ.  Lane counts available: 128-byte rank
 * blocks 2/4/8 (merged) 4/8 (pair); 64-byte 1/2/4 and 2/4; 32-byte 1/2 and 2.  Paired-level images
 * (below) always run the merged schedule, with 1/2/4 lanes on 128-byte and 1/2 on 64-byte blocks;
 * quad-level images with 2 lanes. */
int fm_set_count_schedule(fm_index_t* ix, int merged, int lanes);

/* Wavelet-tree block layout of the HBM image built by subsequent fm_open calls: how many tree
 * levels one rank block read answers.  Backward search is a chain of dependent random HBM reads
 * and a B200 serves a fixed number of those per second whatever their width up to 128 bytes, so
 * levels per read is what sets the speed.
 *   1: one level per block (blocks of 128, 64 or 32 bytes).
 *   2: paired levels.  A block holds a stretch of an even-depth node together with the matching
 *      bits of both children (blocks of 64 or 128 bytes).
 *   4: quad levels.  A 128-byte block holds 128 positions of a node at depth 0, 4, 8, ... and the
 *      matching bits of its children, grandchildren and great-grandchildren
 *      (needs bucket_size < 2^24 rows).
 *   0: back to the default (environment FEMTO_B200_LEVELS_PER_BLOCK, else the built-in choice). */
int fm_set_default_levels_per_block(int levels);

/* Rank block size in bytes of the HBM image built by subsequent fm_open calls: 128, 64 or 32
 * with one level per block, 128 or 64 with paired levels (quad levels always use 128);
 * 0 = back to the default (environment FEMTO_B200_BLOCK_BYTES, else 64 for paired levels and
 * 128 otherwise). */
int fm_set_default_block_bytes(int bytes);

/* --------------------------------------------------------------------------
 * Index construction (host side; "next" row f-1 of the scope table).  Emits an
 * index byte-identical to what the reference's compress_bucket / constructor_*
 * path (src/main/index.c:309-738, src/main/construct.c:146-566) writes for the
 * same BWT rows.  Input per row r of the BWT, in row order:
 *   L[r]       alpha_t symbol preceding the suffix (SEOF for a document start)
 *   sa[r]      suffix array value, i.e. text offset of the suffix
 * and the document ends (exclusive prefix sums of document lengths incl. SEOF).
 * The marking rule should_mark() (src/main/index_types.h:134-144) is applied
 * here.  chunk_size <= 0 writes no document chunks (header chunk_size = -1, as
 * index_documents with map == NULL, construct.c:606).
 * out_dir is created if needed and receives "00", "01", ... and "_femto_index". */
typedef struct fm_builder fm_builder_t;
int fm_builder_create(const char* out_dir, int64_t total_length, int64_t ndocs, const int64_t* doc_ends,
                      int32_t block_size, int32_t bucket_size, int32_t chunk_size, int32_t mark_period,
                      int nthreads, fm_builder_t** out);
/* Append the next `nrows` BWT rows (any granularity; rows must arrive in order). */
int fm_builder_append(fm_builder_t* b, int64_t nrows, const uint16_t* L, const int64_t* sa);
/* Document information strings stored in the header block (may be NULL => "doc<i>"). */
int fm_builder_set_doc_info(fm_builder_t* b, int64_t doc, const void* info, int64_t len);
int fm_builder_finish(fm_builder_t* b);       /* writes the header block; frees b */
void fm_builder_abort(fm_builder_t* b);
/* Builders working side by side (one per GPU / process; the reference's partition unit is the data
 * block, src/main/index.h:83-100, and its constructor emits blocks one after the other,
 * src/main/construct.c:293-566): a range builder writes only the data blocks
 * [first_block, first_block + range_blocks) into out_dir and is fed exactly the BWT rows of those blocks,
 * first_block * block_size onwards, in order.  fm_builder_finish_range hands back what the header needs from
 * this range -- block_counts[k * 261 + c] = occurrences of symbol c inside the k-th block of the range, and
 * eof_rows[d] = row of document d's end when it lies in the range, else -1 -- and frees b.  Once every range
 * is done, one caller gathers the counts of all blocks in block order, merges eof_rows (max) and writes
 * the header block with fm_builder_write_header (doc_info / doc_info_len: NULL, or one string per document,
 * a NULL entry = "doc<i>").  The files are byte-identical to those of one builder fed all rows. */
int fm_builder_create_range(const char* out_dir, int64_t total_length, int64_t ndocs, const int64_t* doc_ends,
                            int32_t block_size, int32_t bucket_size, int32_t chunk_size, int32_t mark_period,
                            int nthreads, int64_t first_block, int64_t range_blocks, fm_builder_t** out);
int fm_builder_finish_range(fm_builder_t* b, int64_t* block_counts, int64_t* eof_rows);
int fm_builder_write_header(const char* out_dir, int64_t total_length, int64_t ndocs, const int64_t* doc_ends,
                            int32_t block_size, int32_t bucket_size, int32_t chunk_size, int32_t mark_period,
                            const int64_t* block_counts, const int64_t* eof_rows, const void* const* doc_info,
                            const int64_t* doc_info_len);
/* Convert a directory index into the flattened single-file form (flatten_index,
 * src/main/index.c:2260-2365). */
int fm_flatten(const char* index_dir, const char* out_file);

/* Suffix array of a prepared text (symbols 1..260, no zeros) by prefix doubling on the
 * host; for tests and small corpora.  sa must hold n entries. */
int fm_suffix_sort_host(const uint16_t* text, int64_t n, int64_t* sa);

#ifdef __cplusplus
}
#endif
#endif /* FEMTO_B200_H */
