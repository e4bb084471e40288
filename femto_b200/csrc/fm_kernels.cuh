// fm_kernels.cuh -- launch interface of the sm_100a query kernels (fm_kernels.cu).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

#include "fm_image.hpp"

namespace fmb {

// Work-queue counters live in device memory owned by the caller (one uint64 per launch).
struct CountArgs {
  int64_t npats;
  const int32_t* plen;
  const uint16_t* flat;
  const int64_t* offs;
  int64_t* first;
  int64_t* last;   // may be NULL: first[i] receives the count
  // Streamed batches (host-buffer calls): the kernel is launched while the patterns are still being
  // copied in; *avail (device memory, written by the copy stream after each chunk) = number of
  // leading patterns whose plen / offs / symbols have arrived.  NULL = everything is there.
  const unsigned long long* avail;
  int32_t* stalled;  // set to 1 if a pattern did not arrive within ~0.1 s (the launch then ends; results invalid)
  // > 0: every pattern has this length and pattern i starts at flat + i * uniform_len; plen and offs
  // are not read (and need not be copied in).
  int32_t uniform_len;
  // != 0: flat holds raw text bytes, one per symbol (symbol = byte + 5, strtoalpha); offs / uniform_len
  // count bytes.  Halves the host->device traffic of a batch; default quad schedule only.
  int32_t sym8;
};

enum WalkMode : int {
  kWalkLocate = 0,  // follow LF until a marked row; out_offset[i] = SA[rows[i]]
  kWalkStep = 1,    // one LF step: out_ch, out_next, out_offset (mark of the row itself or -1)
  kWalkExtract = 2, // nsteps[i] LF steps writing L right-to-left into out_sym[sym_off[i] + nsteps[i]-1-t]
  kWalkShard = 3    // range-sharded locate: walk state[i] while its row is resident (see WalkArgs::state)
};

struct WalkArgs {
  int64_t nrows;
  const int64_t* rows;
  int64_t* out_offset;   // locate / step
  int32_t* out_ch;       // step
  int64_t* out_next;     // step
  const int64_t* nsteps; // extract
  const int64_t* sym_off;
  uint16_t* out_sym;
  int32_t* status;       // device int: set non-zero on malformed walk (unmarked document start, bad row)
  // kWalkShard (index opened as a BWT row range): nrows states of kWalkStateWords int64
  //   {slot at the home rank, row -- or the text offset once finished --, LF steps so far,
  //    phase | home_rank << 4}, phase 0 = walking, 2 = finished.
  // The kernel follows LF while the row is resident; dest[i] = the rank that must see state i
  // next: the owner of its row, or its home rank once finished.
  int64_t* state;
  int32_t* dest;
  int32_t nshards;
  int32_t block_size;    // rows per data block; owner(row) = (row / block_size) * nshards / nblocks
  int64_t nblocks;
  // optional (4 x uint64, zeroed by the caller): LF steps, wavelet-tree rank blocks read (quad image),
  // mark bit-vector blocks read, SA samples read -- what the locate roofline is computed from
  unsigned long long* stats;
};
constexpr int kWalkStateWords = 4;

struct OccArgs {
  int64_t n;
  const uint16_t* ch;
  const int64_t* rows;
  int64_t* out;  // C[ch] + Occ(ch,row)
};

// Range-sharded count: pattern states travel between the GPUs that own the
// BWT rows they need.  One state = 6 int64: pid, first, last, i, obA (C[c]+Occ(c,first-1) once
// known), meta = phase | home_rank << 4.  phase: 3 new (initialise from the pattern), 0 needs
// Occ(c,first-1), 1 needs Occ(c,last), 2 finished (travel home).  The kernel advances every state
// while the rows it needs are resident and writes the rank that must see it next into dest[].
constexpr int kShardStateWords = 6;
struct ShardArgs {
  int64_t n;
  int64_t* state;        // [n][kShardStateWords]
  const int32_t* plen;   // all patterns of the batch are replicated on every rank
  const uint16_t* flat;
  const int64_t* offs;
  int32_t* dest;         // out: rank owning the next row needed (or the home rank when finished)
  int32_t nshards;
  int32_t block_size;    // rows per data block; shard(row) = (row / block_size) * nshards / nblocks
  int64_t nblocks;
};

// lanes_per_query: 4 or 8.  Each launch bumps *launch_counter (host) by the number of kernels launched.
// d_stats (optional, 4 x uint64 zeroed by the caller) selects the instrumented kernel variant:
// [0] rank blocks requested, [1] distinct rank blocks per step and level, [2] Occ evaluations,
// [3] backward-search steps.
cudaError_t launch_count(const DevImage& im, const CountArgs& a, unsigned long long* d_work_counter,
                         int lanes_per_query, int sm_count, cudaStream_t stream, int64_t* launch_counter,
                         unsigned long long* d_stats = nullptr);
cudaError_t launch_walk(const DevImage& im, const WalkArgs& a, WalkMode mode, unsigned long long* d_work_counter,
                        int lanes_per_query, int sm_count, cudaStream_t stream, int64_t* launch_counter);
cudaError_t launch_occ(const DevImage& im, const OccArgs& a, unsigned long long* d_work_counter,
                       int lanes_per_query, int sm_count, cudaStream_t stream, int64_t* launch_counter);
cudaError_t launch_count_shard(const DevImage& im, const ShardArgs& a, int lanes_per_query, int sm_count,
                               cudaStream_t stream, int64_t* launch_counter);

// Ranges -> rows on the device (parallel_locate between its count and its walks,
// src/main/server.c:4407-4415): noccs[i] = rows of pattern i after the reference's clip
// (`last - first > max_occs` cuts to max_occs rows), out_start = exclusive prefix sum of noccs,
// *d_total = their sum.  scratch: device bytes for the scan, at least expand_scratch_bytes(npats).
size_t expand_scratch_bytes(int64_t npats);
cudaError_t launch_clip_and_scan(int64_t npats, const int64_t* d_first, const int64_t* d_last, int max_occs,
                                 int32_t* d_noccs, int64_t* d_out_start, int64_t* d_total, void* d_scratch,
                                 size_t scratch_bytes, cudaStream_t stream, int64_t* launch_counter);
// rows[out_start[i] + j] = first[i] + j for j < noccs[i]; total = number of rows
cudaError_t launch_expand_rows(int64_t npats, int64_t total, const int64_t* d_first, const int64_t* d_out_start,
                               int64_t* d_rows, cudaStream_t stream, int64_t* launch_counter);

// Measurement aid (the ceiling bench.py quotes for the count kernel): `steps` rounds of dependent
// uniformly random reads of `bytes_per_access` (32, 64 or 128, naturally aligned) over `n_units`
// such units starting at `base`; one 128-bit load per lane, two independent chains per lane
// group, every SM filled.  *accesses receives the number of reads the launch performs.
cudaError_t launch_probe(const uint4* base, uint64_t n_units, int bytes_per_access, int steps, int sm_count,
                         cudaStream_t stream, unsigned long long* d_sink, int64_t* accesses);

}  // namespace fmb
