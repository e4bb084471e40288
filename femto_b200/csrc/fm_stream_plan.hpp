// fm_stream_plan.hpp -- how a streamed host-buffer count (fm_api.cu, count_host) cuts a batch.
//
// The batch runs as two kernels (patterns [0, mid) and [mid, npats)); its patterns travel in chunks
// that grow from 8 Ki to 128 Ki patterns, each followed by the arrival mark of its kernel.  Two
// alignment rules keep the kernels from ever seeing a cache line that a later chunk will still
// change: pattern boundaries of chunks are multiples of 32 (128 bytes of plen, 256 of offs) and the
// symbol range of a chunk ends at a multiple of 64 symbols (128 bytes) at or after the end of its
// last pattern.  Host-only, header-only; exercised on CPU by tests/test_stream_plan.py through
// fm_debug_stream_plan.
#pragma once

#include <algorithm>
#include <cstdint>

namespace fmb {

constexpr int64_t kStreamChunk = 1 << 16;          // reference chunk: grows 1/8x .. 2x of this
constexpr int64_t kStreamMinBatch = 2 * kStreamChunk;

// first pattern of the second kernel: the last quarter, on a 32-pattern boundary
inline int64_t stream_split(int64_t npats) { return (npats - npats / 4) & ~int64_t(31); }

// patterns in the chunk that starts at absolute pattern index `lo`
inline int64_t stream_chunk_at(int64_t lo) {
  return std::min<int64_t>(kStreamChunk * 2, std::max<int64_t>(kStreamChunk / 8, lo));
}

// end (exclusive, in symbols) of the symbol copy of a chunk whose last pattern is hi - 1, for an
// in-order batch: the end of that pattern rounded up to a 128-byte line, the whole buffer for the
// last chunk
// (sym_bytes: 2 for alpha_t symbols, 1 for raw text bytes -- 64 resp. 128 symbols per line)
inline int64_t stream_symbol_cut(const int32_t* plen, const int64_t* offs, int64_t hi, int64_t npats,
                                 int64_t flat_len, int sym_bytes = 2) {
  if (hi >= npats) return flat_len;
  const int64_t end = offs[hi - 1] + plen[hi - 1];
  const int64_t per_line = 128 / sym_bytes;
  return std::min(flat_len, (end + per_line - 1) & ~(per_line - 1));
}

}  // namespace fmb
