// fm_api.cu -- the C ABI of include/femto_b200.h: index lifecycle, host<->device staging and the
// entry points that mirror the reference's parallel_count / parallel_locate / parallel_locate_range
// (src/main/femto.c:275-536).  Query work is done only by the kernels in fm_kernels.cu; there is no
// CPU fallback on any query path.
#include <cuda_runtime.h>

#include <algorithm>
#include <functional>
#include <atomic>
#include <chrono>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/femto_b200.h"
#include "fm_format.hpp"
#include "fm_image.hpp"
#include "fm_kernels.cuh"
#include "fm_loader.hpp"
#include "fm_mesh.cuh"
#include "fm_stream_plan.hpp"

using namespace fmb;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

struct CudaFail : std::runtime_error {
  explicit CudaFail(const std::string& m) : std::runtime_error(m) {}
};

#define CK(expr)                                                                                        \
  do {                                                                                                  \
    cudaError_t e_ = (expr);                                                                            \
    if (e_ != cudaSuccess)                                                                              \
      throw CudaFail(std::string(#expr) + ": " + cudaGetErrorName(e_) + " (" + cudaGetErrorString(e_) + ")"); \
  } while (0)

// A growable device buffer (freed with the index).
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  void* get(size_t bytes) {
    if (bytes > cap) {
      if (p) CK(cudaFree(p));
      p = nullptr;
      cap = 0;
      const size_t want = std::max<size_t>(bytes + bytes / 4, 4096);
      CK(cudaMalloc(&p, want));
      cap = want;
    }
    return p;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

struct HostBuf {  // pinned staging
  void* p = nullptr;
  size_t cap = 0;
  void* get(size_t bytes) {
    if (bytes > cap) {
      if (p) CK(cudaFreeHost(p));
      p = nullptr;
      cap = 0;
      const size_t want = std::max<size_t>(bytes + bytes / 4, 4096);
      CK(cudaMallocHost(&p, want));
      cap = want;
    }
    return p;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

}  // namespace

struct fm_index {
  int device = 0;
  int sm_count = 148;
  int lanes_per_query = 4;   // walk / occ kernels
  int count_sched = 1024;    // launch_count schedule code; set from default_count_sched() at open
  fm_info_t info{};
  DevImage im;
  // device allocations of the image
  void* d_blocks = nullptr;
  void* d_nodes = nullptr;
  void* d_occ = nullptr;
  void* d_mark = nullptr;
  void* d_buckets = nullptr;
  void* d_markvals = nullptr;
  void* d_C = nullptr;
  unsigned long long* d_work = nullptr;  // work-queue counters (one per concurrent launch slot)
  int32_t* d_status = nullptr;
  // host-side header tables
  std::vector<int64_t> doc_ends, doc_eof_rows, C_host, doc_info_off;
  std::vector<uint8_t> doc_info_bytes;
  // document chunks as stored (HostImage::chunk_*), decoded on demand
  std::vector<uint8_t> chunk_bytes;
  std::vector<int64_t> chunk_off;
  std::vector<int32_t> chunk_count;
  std::vector<uint32_t> chunk_dir_rel;
  BlockHeader hdr;
  // per-call scratch, serialised by mu
  std::mutex mu;
  cudaStream_t stream = nullptr, stream2 = nullptr, stream3 = nullptr;
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  DevBuf d_in[4], d_out[4];
  HostBuf h_stage[2], h_marks, h_ranges;
  std::vector<int64_t> gather_offs;   // fm_count (pointer array): offsets of the patterns in the gathered buffer
  int64_t launches = 0;
  int64_t last_h2d = 0, last_d2h = 0;  // bytes copied by the most recent host-buffer count call
  bool stream_ok = true;               // cleared when a streamed batch stalled (e.g. under a serialising profiler)
  int dev_slot = 0;
};

// One rank's end of the device-initiated exchange (fm_mesh.cuh).
struct fm_mesh {
  fm_index* ix = nullptr;
  int rank = 0, world = 1, cap_shift = 0;
  unsigned long long window = 0;
  size_t region_bytes = 0;
  void* region = nullptr;                    // MeshCtl + inbox rings, one cudaMalloc (exported through cudaIpc)
  void* peer_region[kMeshMaxRanks] = {};     // every rank's region in this process's address space
  bool peer_ipc[kMeshMaxRanks] = {};         // opened with cudaIpcOpenMemHandle
  bool connected = false;
  unsigned long long epoch = 0;
  int max_ctas = 0;
  double timeout_s = 10.0;
  int clock_khz = 1000000;
};

namespace {

// count schedule codes understood by launch_count: pair = lanes per Occ query;
// sync = 1000 + 10*lanes + resident blocks per SM the variant is compiled for.
int sync_sched(int block_words, int lanes) {
  // register budgets picked from the sweep in profiles/r01_count_schedule_sweep.md
  const int minb = block_words == 32 ? (lanes == 8 ? 5 : lanes == 4 ? 4 : 3)
                 : block_words == 16 ? (lanes == 4 ? 6 : 4)
                                     : (lanes == 2 ? 6 : 4);
  return 1000 + 10 * lanes + minb;
}

int paired_sched(int block_words, int lanes) {
  // profiles/r01_paired_level_sweep.md
  const int minb = block_words == 32 ? (lanes == 4 ? 6 : lanes == 2 ? 4 : 3) : 4;
  return 1000 + 10 * lanes + minb;
}

// quad-level blocks: 1000 + 10 * 2 + CTAs per SM = sync schedule (two lanes per pattern evaluate both
// positions of a step together; default); 1000 + 60 + v = split schedule (one lane per Occ, selectable
// with fm_set_count_schedule(.., 1)); 1000 + 70 + 10 * EXP + CTAs per SM = measurement variants of the
// sync kernel (fm_kernels.cu, profiles/r01_quad_schedules.md)
constexpr int kQuadSched = 1000 + 10 * 2 + 4;  // 64 registers, no spills, 4 CTAs x 8 warps per SM
constexpr int kQuadSchedSplit = 1000 + 67;

int default_count_sched(int block_words, int levels) {
  int sched = levels == 4 ? kQuadSched
            : levels == 2 ? paired_sched(block_words, 2)
                          : sync_sched(block_words, block_words == 32 ? 4 : block_words == 16 ? 2 : 1);
  if (const char* e = std::getenv("FEMTO_B200_COUNT_SCHED")) {  // tuning experiments only
    const int v = std::atoi(e);
    if (v > 0) sched = v;
  }
  return sched;
}

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    CK(cudaSetDevice(dev));
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

template <typename T>
void upload(void** dptr, const T* src, size_t count, int64_t* total) {
  const size_t bytes = std::max<size_t>(count * sizeof(T), 16);
  CK(cudaMalloc(dptr, bytes));
  if (count) CK(cudaMemcpy(*dptr, src, count * sizeof(T), cudaMemcpyHostToDevice));
  *total += int64_t(bytes);
}

void destroy(fm_index* ix) {
  if (!ix) return;
  cudaSetDevice(ix->device);
  for (void* p : {ix->d_blocks, ix->d_nodes, ix->d_occ, ix->d_mark, ix->d_buckets, ix->d_markvals, ix->d_C,
                  static_cast<void*>(ix->d_work), static_cast<void*>(ix->d_status)})
    if (p) cudaFree(p);
  for (auto& b : ix->d_in) b.release();
  for (auto& b : ix->d_out) b.release();
  for (auto& b : ix->h_stage) b.release();
  ix->h_marks.release();
  ix->h_ranges.release();
  for (cudaEvent_t e : ix->ev) if (e) cudaEventDestroy(e);
  if (ix->stream) cudaStreamDestroy(ix->stream);
  if (ix->stream2) cudaStreamDestroy(ix->stream2);
  if (ix->stream3) cudaStreamDestroy(ix->stream3);
  delete ix;
}

int open_impl(const char* path, int device, int shard, int nshards, fm_index_t** out) {
  if (!path || !out) return fail(FM_ERR_PARAM, "fm_open: null argument");
  *out = nullptr;
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev <= 0)
    return fail(FM_ERR_IO, std::string("fm_open: no usable CUDA device (") + cudaGetErrorString(ce) +
                               "); this library has no CPU query path");
  if (device < 0 || device >= ndev) return fail(FM_ERR_PARAM, "fm_open: no such CUDA device");
  std::unique_ptr<HostImage> host;
  try {
    host = build_host_image(path, shard, nshards, 0);
  } catch (const Error& e) {
    return fail(e.code, std::string("fm_open: ") + e.what());
  } catch (const std::bad_alloc&) {
    return fail(FM_ERR_MEM, "fm_open: out of host memory");
  }
  fm_index* ix = new fm_index();
  try {
    ix->device = device;
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    ix->sm_count = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&ix->stream2, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&ix->stream3, cudaStreamNonBlocking));
    for (cudaEvent_t& e : ix->ev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    int64_t total = 0;
    upload(&ix->d_blocks, host->rank_words, size_t(host->n_rank_blocks) * size_t(host->block_words), &total);
    if (host->levels == 4) upload(&ix->d_nodes, host->quads.data(), host->quads.size(), &total);
    else if (host->levels == 2) upload(&ix->d_nodes, host->supers.data(), host->supers.size(), &total);
    else upload(&ix->d_nodes, host->nodes.data(), host->nodes.size(), &total);
    upload(&ix->d_occ, host->occ.data(), host->occ.size(), &total);
    upload(&ix->d_mark, host->mark.data(), host->mark.size(), &total);
    upload(&ix->d_buckets, host->buckets.data(), host->buckets.size(), &total);
    upload(&ix->d_markvals, host->markvals.data(), host->markvals.size(), &total);
    upload(&ix->d_C, host->C.data(), host->C.size(), &total);
    CK(cudaMalloc(reinterpret_cast<void**>(&ix->d_work), 16 * sizeof(unsigned long long)));  // queue slots:
    // 0 host-buffer calls, 1..7 caller-stream launches, 8..9 the two kernels of a streamed
    // host-buffer call, 12..13 their arrival counters (CountArgs::avail)
    CK(cudaMalloc(reinterpret_cast<void**>(&ix->d_status), 64));
    CK(cudaMemset(ix->d_status, 0, 64));

    const BlockHeader& h = host->hdr;
    ix->im.blocks = static_cast<const uint4*>(ix->d_blocks);
    if (host->levels == 4) ix->im.quads = static_cast<const QuadRec*>(ix->d_nodes);
    else if (host->levels == 2) ix->im.supers = static_cast<const SuperRec*>(ix->d_nodes);
    else ix->im.nodes = static_cast<const NodeRec*>(ix->d_nodes);
    ix->im.levels = host->levels;
    ix->im.root_stride = host->root_stride;
    ix->info.levels_per_block = host->levels;
    ix->im.occ = static_cast<const OccRec*>(ix->d_occ);
    ix->im.mark = static_cast<const MarkRec*>(ix->d_mark);
    ix->im.buckets = static_cast<const BucketRec*>(ix->d_buckets);
    ix->im.markvals = static_cast<const int64_t*>(ix->d_markvals);
    ix->im.C = static_cast<const int64_t*>(ix->d_C);
    ix->im.total_length = h.total_length;
    ix->im.first_row = host->first_row;
    ix->im.end_row = host->end_row;
    ix->im.first_bucket = host->first_bucket;
    ix->im.bucket_size = h.bucket_size;
    ix->im.bucket_shift = (h.bucket_size & (h.bucket_size - 1)) == 0 ? __builtin_ctz(unsigned(h.bucket_size)) : -1;

    ix->info.total_length = h.total_length;
    ix->info.num_documents = h.ndocs;
    ix->info.num_blocks = h.nblocks;
    ix->info.block_size = h.block_size;
    ix->info.bucket_size = h.bucket_size;
    ix->info.mark_period = h.mark_period;
    ix->info.chunk_size = h.chunk_size;
    ix->info.first_row = host->first_row;
    ix->info.end_row = host->end_row;
    ix->info.hbm_bytes = total;
    ix->info.rank_block_bytes = host->n_rank_blocks * int64_t(host->block_words) * 4;
    ix->im.block_words = host->block_words;
    ix->count_sched = default_count_sched(host->block_words, host->levels);
    ix->info.device = device;
    ix->info.max_code_len = host->max_code_len;
    ix->info.rank_block_size = host->block_words * 4;
    ix->C_host = host->C;
    ix->doc_ends = std::move(host->doc_ends);
    ix->doc_eof_rows = std::move(host->doc_eof_rows);
    ix->doc_info_off = std::move(host->doc_info_off);
    ix->doc_info_bytes = std::move(host->doc_info_bytes);
    ix->chunk_bytes = std::move(host->chunk_bytes);
    ix->chunk_off = std::move(host->chunk_off);
    ix->chunk_count = std::move(host->chunk_count);
    ix->chunk_dir_rel = std::move(host->chunk_dir_rel);
    ix->hdr = host->hdr;
  } catch (const CudaFail& e) {
    destroy(ix);
    return fail(FM_ERR_IO, std::string("fm_open: ") + e.what());
  }
  *out = ix;
  return FM_OK;
}

// Run `body` with the handle locked, its device current, CUDA failures mapped to FM_ERR_IO.
template <typename F>
int guarded(fm_index* ix, const char* what, F&& body) {
  if (!ix) return fail(FM_ERR_PARAM, std::string(what) + ": null index");
  std::lock_guard<std::mutex> lock(ix->mu);
  // A failure may leave asynchronous copies from / to the caller's buffers in flight: drain the
  // handle's streams before the error is reported, so that the caller may free them.
  auto drain = [&]() {
    for (cudaStream_t st : {ix->stream, ix->stream2, ix->stream3})
      if (st) cudaStreamSynchronize(st);
  };
  try {
    DeviceGuard guard(ix->device);
    return body();
  } catch (const CudaFail& e) {
    drain();
    return fail(FM_ERR_IO, std::string(what) + ": " + e.what());
  } catch (const Error& e) {
    drain();
    return fail(e.code, std::string(what) + ": " + e.what());
  } catch (const std::bad_alloc&) {
    drain();
    return fail(FM_ERR_MEM, std::string(what) + ": out of memory");
  }
}

// Validate a flat pattern batch on the host and return the number of symbols referenced.
// `ordered` (optional) reports whether the patterns lie in the flat buffer in batch order without
// overlap -- the layout that allows chunked, pipelined transfers.
int check_patterns(int64_t npats, const int32_t* plen, const int64_t* offs, int64_t* flat_len,
                   bool* ordered = nullptr) {
  int64_t end = 0;
  bool ord = true;
  for (int64_t i = 0; i < npats; i++) {
    if (plen[i] < 0 || offs[i] < 0) return FM_ERR_PARAM;
    ord = ord && offs[i] >= end;
    end = std::max(end, offs[i] + plen[i]);
  }
  *flat_len = end;
  if (ordered) *ordered = ord;
  return FM_OK;
}

// One pass over plen / offs of a large batch, split over a few host threads (the pass is on the
// critical path of every host-buffer call: 12 bytes per pattern).  Reports
//   bad      a negative length or offset,
//   dense    pattern i starts where pattern i-1 ends (then the batch is in order and flat_len is the
//            end of the last pattern),
//   uniform  dense, starting at 0, all of one length m > 0: pattern i starts at i * m and neither plen
//            nor offs has to travel to the device (CountArgs::uniform_len).
struct BatchShape {
  bool bad = false, dense = true;
  int uniform = 0;
  int64_t flat_len = 0;
};

BatchShape scan_batch(int64_t npats, const int32_t* plen, const int64_t* offs) {
  BatchShape r;
  if (npats == 0) return r;
  const int nt = int(std::min<int64_t>(8, std::max<int64_t>(1, npats >> 17)));
  std::vector<char> bad(size_t(nt), 0), dense(size_t(nt), 1), same(size_t(nt), 1);
  const int32_t m0 = plen[0];
  auto work = [&](int t) {
    const int64_t lo = npats * t / nt, hi = npats * (t + 1) / nt;
    bool b = false, d = true, sm = true;
    for (int64_t i = lo; i < hi; i++) {
      b |= (plen[i] < 0) | (offs[i] < 0);
      sm &= plen[i] == m0;
      if (i) d &= offs[i] == offs[i - 1] + plen[i - 1];
    }
    bad[size_t(t)] = b; dense[size_t(t)] = d; same[size_t(t)] = sm;
  };
  if (nt == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (int t = 1; t < nt; t++) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
  }
  bool sm = true;
  for (int t = 0; t < nt; t++) { r.bad |= bad[size_t(t)] != 0; r.dense &= dense[size_t(t)] != 0; sm &= same[size_t(t)] != 0; }
  if (r.bad) return r;
  if (r.dense) {
    r.flat_len = offs[npats - 1] + plen[npats - 1];
    if (sm && m0 > 0 && offs[0] == 0) r.uniform = m0;
  }
  return r;
}

int walk_status(fm_index* ix) {
  int32_t st = 0;
  CK(cudaMemcpyAsync(&st, ix->d_status, sizeof(st), cudaMemcpyDeviceToHost, ix->stream));
  CK(cudaStreamSynchronize(ix->stream));
  if (st) CK(cudaMemsetAsync(ix->d_status, 0, sizeof(int32_t), ix->stream));
  return st;
}

// Arrival marks of a streamed batch are written with the driver's stream memory operation
// (cuStreamWriteValue64: ordered behind the stream's earlier copies, and the operation CUDA defines
// for signalling a kernel that polls device memory); resolved at run time so that the library has no
// link-time dependency on libcuda and still loads on hosts without a driver.  NULL: not available,
// the mark then travels as an 8-byte copy on the same stream.
using StreamWrite64 = int (*)(cudaStream_t, unsigned long long, unsigned long long, unsigned int);
StreamWrite64 stream_write64() {
  static const StreamWrite64 fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (std::getenv("FEMTO_B200_MARK_MEMCPY") ||
        cudaGetDriverEntryPoint("cuStreamWriteValue64", &f, cudaEnableDefault, &qr) != cudaSuccess ||
        qr != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<StreamWrite64>(f);
  }();
  return fn;
}

constexpr int kRetryValidated = -1000;  // count_host(lazy): the batch is not what its first / last pattern claimed

// count on host buffers; leaves first/last on the host.  flat_len symbols in flat.
// sym_bytes: 2 = alpha_t symbols, 1 = raw text bytes (fm_count_bytes; flat is then a byte buffer).
int count_host(fm_index* ix, int64_t npats, const int32_t* plen, const uint16_t* flat16, int64_t flat_len,
               const int64_t* offs, int64_t* first, int64_t* last, bool in_order = false, int uniform_len = 0,
               bool to_host = true, bool lazy = false,
               const std::function<void(int64_t)>* ready_upto = nullptr, int sym_bytes = 2) {
  const unsigned char* flat = reinterpret_cast<const unsigned char*>(flat16);
  const size_t sb = size_t(sym_bytes);
  // ready_upto (fm_count): the flat buffer is being filled by gather threads while this function
  //   runs; (*ready_upto)(hi) returns once the symbols of patterns [0, hi) are in place.
  // to_host == false: first/last stay in ix->d_out[0] / d_out[1] for a kernel that follows (locate)
  // lazy == true (streamed batches only): the caller has NOT validated plen / offs; in_order,
  //   uniform_len and flat_len are its claims, derived from the first and last pattern.  Every chunk
  //   is checked just before its copy is enqueued -- behind the kernel that is already running, so
  //   the 12 bytes per pattern of validation leave the critical path.  A chunk that breaks the
  //   claim aborts the kernels and the call returns kRetryValidated: nothing was read out of
  //   bounds, the caller validates the whole batch and calls again.
  if (npats == 0) return FM_OK;
  const auto t_call = std::chrono::steady_clock::now();
  cudaStream_t s = ix->stream;
  int32_t* d_plen = static_cast<int32_t*>(ix->d_in[0].get(size_t(npats) * 4));
  unsigned char* d_flat = static_cast<unsigned char*>(ix->d_in[1].get(size_t(std::max<int64_t>(flat_len, 1)) * sb));
  auto dflat16 = [&](int64_t sym) { return reinterpret_cast<const uint16_t*>(d_flat + size_t(sym) * sb); };
  int64_t* d_offs = static_cast<int64_t*>(ix->d_in[2].get(size_t(npats) * 8));
  int64_t* d_first = static_cast<int64_t*>(ix->d_out[0].get(size_t(npats) * 8));
  int64_t* d_last = (last || !to_host) ? static_cast<int64_t*>(ix->d_out[1].get(size_t(npats) * 8)) : nullptr;

  // Large batches whose patterns lie in the flat buffer in order are STREAMED: the count kernel is
  // launched at once and pulls patterns from its queue as the copy stream delivers them (a device
  // counter written after every chunk gates the queue, CountArgs::avail), so the host->device
  // transfer overlaps the search instead of preceding it.  The batch runs as two kernels, one per
  // half, so that the results of the first half travel back while the second half is searched.
  // (chunk sizes, the split and the symbol cuts: fm_stream_plan.hpp)
  static const bool no_stream = std::getenv("FEMTO_B200_NO_STREAM") != nullptr;
  const bool streamed = in_order && npats >= kStreamMinBatch && ix->stream2 && ix->stream3 && ix->stream_ok && !no_stream;
  if (lazy && !streamed) return kRetryValidated;
  if (streamed) {
    const int m = uniform_len;
    // the second kernel takes the last quarter: its results are the only transfer nothing overlaps
    const int64_t mid = stream_split(npats);
    const int64_t half_lo[2] = {0, mid}, half_hi[2] = {mid, npats};
    // copy chunks grow from 8 Ki to 128 Ki patterns: the first patterns arrive after a few microseconds,
    // the bulk travels in few, large transfers
    auto chunk_at = [](int64_t lo) { return stream_chunk_at(lo); };
    int64_t nmarks = 0;
    for (int h = 0; h < 2; h++)
      for (int64_t lo = half_lo[h]; lo < half_hi[h]; lo += chunk_at(lo)) nmarks++;
    unsigned long long* marks = static_cast<unsigned long long*>(ix->h_marks.get(size_t(nmarks + 2) * 8));
    unsigned long long* d_avail = ix->d_work + 12;
    cudaStream_t sk = ix->stream, sc = ix->stream2, sr = ix->stream3;
    // FEMTO_B200_TRACE=1: print when (ms after the start) the copies, each kernel and the results finish
    static const bool trace = std::getenv("FEMTO_B200_TRACE") != nullptr;
    cudaEvent_t tev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    if (trace) for (cudaEvent_t& e : tev) CK(cudaEventCreate(&e));
    CK(cudaMemsetAsync(d_avail, 0, 2 * sizeof(unsigned long long), sk));
    CK(cudaMemsetAsync(ix->d_status + 1, 0, sizeof(int32_t), sk));  // a call that failed half-way may have left it set
    if (trace) CK(cudaEventRecord(tev[0], sk));
    CK(cudaEventRecord(ix->ev[0], sk));
    CK(cudaStreamWaitEvent(sc, ix->ev[0], 0));
    CK(cudaStreamWaitEvent(sr, ix->ev[0], 0));
    // kernels first: they spin on their arrival counters until the copies below land.  The second
    // kernel sits on its own stream: its CTAs take the SM slots the first kernel's CTAs free one by
    // one, so the first kernel's tail is not idle time.
    for (int h = 0; h < 2; h++) {
      const int64_t lo = half_lo[h], n = half_hi[h] - lo;
      cudaStream_t ks = h == 0 ? sk : sr;
      CountArgs a{n, d_plen + lo, dflat16(m ? lo * m : 0), d_offs + lo, d_first + lo,
                  d_last ? d_last + lo : nullptr, d_avail + h, ix->d_status + 1, m, sym_bytes == 1};
      CK(launch_count(ix->im, a, ix->d_work + 8 + h, ix->count_sched, ix->sm_count, ks, &ix->launches));
      if (trace) CK(cudaEventRecord(tev[1 + h], ks));
    }
    // copies: chunk after chunk, each followed by its arrival mark.  Symbol ranges are cut at
    // 128-byte boundaries of the device buffer so that no cache line is shared by two chunks (a
    // line read for an early pattern must not hold not-yet-copied symbols of a later one).
    int64_t k = 0, fdone = 0;
    ix->last_h2d = (m ? 0 : npats * 12) + flat_len * int64_t(sb) + nmarks * 8;
    ix->last_d2h = to_host ? npats * (last ? 16 : 8) : 0;
    for (int h = 0; h < 2; h++) {
      for (int64_t lo = half_lo[h]; lo < half_hi[h]; lo += chunk_at(lo), k++) {
        const int64_t hi = std::min(half_hi[h], lo + chunk_at(lo));
        if (ready_upto) {
          // the symbol copy below is cut at a 128-byte line, i.e. up to 63 symbols into the patterns that
          // follow: wait for every pattern that starts before that cut
          const int64_t cut = stream_symbol_cut(plen, offs, hi, npats, flat_len, sym_bytes);
          const int64_t j = hi == npats ? npats : std::lower_bound(offs + hi, offs + npats, cut) - offs;
          (*ready_upto)(j);
        }
        if (lazy) {  // does this chunk keep the claim?  (uniform: length m at i * m; else: densely packed)
          bool ok = true;
          if (m) {
            for (int64_t i = lo; i < hi; i++) ok &= (plen[i] == m) & (offs[i] == i * int64_t(m));
          } else {
            // every pattern must also END inside the claimed buffer: the claim comes from the last
            // pattern alone, which may point back into a dense prefix (the kernel would otherwise read
            // symbols the device buffer, sized from the claim, does not hold)
            if (lo == 0) ok = plen[0] >= 0 && offs[0] >= 0 && offs[0] + int64_t(plen[0]) <= flat_len;
            for (int64_t i = std::max<int64_t>(lo, 1); i < hi; i++)
              ok &= (plen[i] >= 0) & (offs[i] == offs[i - 1] + plen[i - 1]) & (offs[i] + int64_t(plen[i]) <= flat_len);
          }
          if (!ok) {  // stop the kernels (they give up on the flag), drain, hand the batch back
            int32_t* one = reinterpret_cast<int32_t*>(marks + nmarks);
            *one = 1;
            CK(cudaMemcpyAsync(ix->d_status + 1, one, sizeof(int32_t), cudaMemcpyHostToDevice, sc));
            CK(cudaStreamSynchronize(sc));
            CK(cudaStreamSynchronize(sk));
            CK(cudaStreamSynchronize(sr));
            CK(cudaMemsetAsync(ix->d_status + 1, 0, sizeof(int32_t), sk));
            CK(cudaStreamSynchronize(sk));
            if (trace) for (cudaEvent_t e : tev) cudaEventDestroy(e);
            return kRetryValidated;
          }
        }
        if (!m) {
          CK(cudaMemcpyAsync(d_plen + lo, plen + lo, size_t(hi - lo) * 4, cudaMemcpyHostToDevice, sc));
          CK(cudaMemcpyAsync(d_offs + lo, offs + lo, size_t(hi - lo) * 8, cudaMemcpyHostToDevice, sc));
        }
        const int64_t fend = stream_symbol_cut(plen, offs, hi, npats, flat_len, sym_bytes);
        if (fend > fdone) {
          CK(cudaMemcpyAsync(d_flat + size_t(fdone) * sb, flat + size_t(fdone) * sb, size_t(fend - fdone) * sb,
                             cudaMemcpyHostToDevice, sc));
          fdone = fend;
        }
        marks[k] = static_cast<unsigned long long>(hi - half_lo[h]);
        if (StreamWrite64 wv = stream_write64()) {
          if (wv(sc, reinterpret_cast<unsigned long long>(d_avail + h), marks[k], 0) != 0)
            throw CudaFail("cuStreamWriteValue64 failed");
        } else {
          CK(cudaMemcpyAsync(d_avail + h, marks + k, sizeof(unsigned long long), cudaMemcpyHostToDevice, sc));
        }
      }
    }
    if (trace) CK(cudaEventRecord(tev[3], sc));
    // results: each part behind its kernel, on that kernel's stream
    if (to_host) {
      CK(cudaMemcpyAsync(first, d_first, size_t(mid) * 8, cudaMemcpyDeviceToHost, sk));
      if (last) CK(cudaMemcpyAsync(last, d_last, size_t(mid) * 8, cudaMemcpyDeviceToHost, sk));
      CK(cudaMemcpyAsync(first + mid, d_first + mid, size_t(npats - mid) * 8, cudaMemcpyDeviceToHost, sr));
      if (last) CK(cudaMemcpyAsync(last + mid, d_last + mid, size_t(npats - mid) * 8, cudaMemcpyDeviceToHost, sr));
    }
    if (trace) {
      CK(cudaEventRecord(tev[4], sk));
      CK(cudaEventRecord(tev[5], sr));
    }
    int32_t* h_stalled = reinterpret_cast<int32_t*>(marks + nmarks + 1);
    const auto t_enq = std::chrono::steady_clock::now();
    CK(cudaStreamSynchronize(sc));
    CK(cudaStreamSynchronize(sk));
    CK(cudaMemcpyAsync(h_stalled, ix->d_status + 1, sizeof(int32_t), cudaMemcpyDeviceToHost, sr));  // after both kernels
    CK(cudaStreamSynchronize(sr));
    const int32_t stalled = *h_stalled;
    if (trace) {
      float t[6] = {0, 0, 0, 0, 0, 0};
      for (int i = 1; i < 6; i++) cudaEventElapsedTime(&t[i], tev[0], tev[i]);
      const double enq_ms = std::chrono::duration<double, std::milli>(t_enq - t_call).count();
      const double all_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call).count();
      std::fprintf(stderr, "[femto_b200 trace] count %lld patterns: kernel A done %.3f ms, kernel B done %.3f, copies in done "
                   "%.3f, results A out %.3f, results B out %.3f | host: enqueue %.3f ms, whole call %.3f\n",
                   (long long)npats, t[1], t[2], t[3], t[4], t[5], enq_ms, all_ms);
      for (cudaEvent_t e : tev) cudaEventDestroy(e);
    }
    if (!stalled) return FM_OK;
    // The kernels did not see their patterns arrive (streams serialised by a profiler, or a copy
    // stream that made no progress for 0.1 s): repeat the batch the plain way, and stay there.
    CK(cudaMemsetAsync(ix->d_status + 1, 0, sizeof(int32_t), sk));
    ix->stream_ok = false;
    if (lazy) return kRetryValidated;  // the plain path below needs a validated batch
  }
  if (ready_upto) (*ready_upto)(npats);
  ix->last_h2d = npats * 12 + flat_len * int64_t(sb);
  ix->last_d2h = npats * (last ? 16 : 8);
  CK(cudaMemcpyAsync(d_plen, plen, size_t(npats) * 4, cudaMemcpyHostToDevice, s));
  if (flat_len) CK(cudaMemcpyAsync(d_flat, flat, size_t(flat_len) * sb, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d_offs, offs, size_t(npats) * 8, cudaMemcpyHostToDevice, s));
  CountArgs a{npats, d_plen, dflat16(0), d_offs, d_first, d_last};
  a.sym8 = sym_bytes == 1;
  CK(launch_count(ix->im, a, ix->d_work, ix->count_sched, ix->sm_count, s, &ix->launches));
  if (!to_host) {  // results stay on the device; the caller's next kernel runs on the same stream
    ix->last_d2h = 0;
    return FM_OK;
  }
  CK(cudaMemcpyAsync(first, d_first, size_t(npats) * 8, cudaMemcpyDeviceToHost, s));
  if (last) CK(cudaMemcpyAsync(last, d_last, size_t(npats) * 8, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return FM_OK;
}

int locate_rows_host(fm_index* ix, int64_t nrows, const int64_t* rows, int64_t* offsets) {
  if (nrows == 0) return FM_OK;
  cudaStream_t s = ix->stream;
  int64_t* d_rows = static_cast<int64_t*>(ix->d_in[3].get(size_t(nrows) * 8));
  int64_t* d_off = static_cast<int64_t*>(ix->d_out[2].get(size_t(nrows) * 8));
  CK(cudaMemcpyAsync(d_rows, rows, size_t(nrows) * 8, cudaMemcpyHostToDevice, s));
  WalkArgs w{};
  w.nrows = nrows;
  w.rows = d_rows;
  w.out_offset = d_off;
  w.status = ix->d_status;
  CK(cudaMemsetAsync(ix->d_status, 0, sizeof(int32_t), s));  // word 0 belongs to the host-buffer calls (serialised by mu)
  CK(launch_walk(ix->im, w, kWalkLocate, ix->d_work, ix->lanes_per_query, ix->sm_count, s, &ix->launches));
  CK(cudaMemcpyAsync(offsets, d_off, size_t(nrows) * 8, cudaMemcpyDeviceToHost, s));
  const int st = walk_status(ix);
  if (st) return fail(st == 1 ? FM_ERR_PARAM : FM_ERR_INVALID, "locate: malformed walk (status " + std::to_string(st) + ")");
  return FM_OK;
}

// A few long-lived host threads for the pointer-array calls (creating threads per call costs more than the
// work they do at 2 ms per batch).  run() hands every worker the same job and returns at once; wait() blocks
// until all of them have returned from it.  One job at a time (callers hold the handle's mutex; the pool
// itself is shared by all handles and serialises jobs).
class WorkerPool {
 public:
  static WorkerPool& instance() {
    // never destroyed: its threads wait on the condition variables for as long as the process lives, and
    // destroying a condition variable that has waiters (at exit) blocks forever
    static WorkerPool* pool = new WorkerPool();
    return *pool;
  }
  int size() const { return int(threads_.size()); }
  void run(std::function<void(int)> job) {
    std::unique_lock<std::mutex> lk(mu_);
    cv_done_.wait(lk, [this] { return running_ == 0 && !job_; });
    job_ = std::move(job);
    running_ = size();
    generation_++;
    cv_work_.notify_all();
  }
  void wait() {
    std::unique_lock<std::mutex> lk(mu_);
    cv_done_.wait(lk, [this] { return running_ == 0; });
    job_ = nullptr;
    cv_done_.notify_all();
  }

 private:
  WorkerPool() {
    const unsigned hw = std::max(2u, std::thread::hardware_concurrency());
    const int n = int(std::min(16u, std::max(2u, hw / 2)));
    for (int t = 0; t < n; t++) threads_.emplace_back([this, t] { loop(t); });
    for (auto& th : threads_) th.detach();  // live as long as the process
  }
  void loop(int t) {
    uint64_t seen = 0;
    for (;;) {
      std::function<void(int)> job;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_work_.wait(lk, [&] { return generation_ != seen; });
        seen = generation_;
        job = job_;
      }
      if (job) job(t);
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (--running_ == 0) cv_done_.notify_all();
      }
    }
  }
  std::mutex mu_;
  std::condition_variable cv_work_, cv_done_;
  std::function<void(int)> job_;
  std::vector<std::thread> threads_;
  uint64_t generation_ = 0;
  int running_ = 0;
};

// Gathers the reference-style pointer array (one pointer per pattern) into one pinned flat buffer
// with the worker pool, IN ORDER and in the background: the streamed count enqueues the copy of a
// chunk as soon as its patterns are in place (wait_upto), so the gather -- the most expensive host
// step of the pointer-array call -- overlaps the search.  The offsets of the patterns in the flat
// buffer are computed by the same workers: for a batch of equal lengths (probed at three places,
// verified by every worker on its share) they are i * m and need no serial pass.
class PatternGatherer {
 public:
  PatternGatherer(fm_index* ix, int64_t npats, const int* plen, const uint16_t* const* pats)
      : npats_(npats), plen_(plen), pats_(pats) {
    if (ix->gather_offs.size() < size_t(npats) + 1) ix->gather_offs.resize(size_t(npats) + 1 + size_t(npats) / 4);
    offs_ = ix->gather_offs.data();  // kept with the handle: a fresh 8 MB vector per call costs more than the gather's share
    nblocks_ = (npats + kBlock - 1) / kBlock;
    done_.reset(new std::atomic<unsigned char>[size_t(std::max<int64_t>(nblocks_, 1))]);
    for (int64_t b = 0; b < nblocks_; b++) done_[size_t(b)].store(0, std::memory_order_relaxed);
    const int m = npats > 0 ? plen[0] : 0;
    const bool big = npats >= 8 * kBlock;
    bool uniform = big && m > 0 && plen[npats / 2] == m && plen[npats - 1] == m;
    WorkerPool* pool = big ? &WorkerPool::instance() : nullptr;
    if (uniform) {  // offsets i * m, every worker checking its share of the lengths
      std::atomic<bool> ok{true};
      std::atomic<int64_t> next{0};
      pool->run([&](int) {
        for (;;) {
          const int64_t b = next.fetch_add(1);
          if (b >= nblocks_) return;
          const int64_t lo = b * kBlock, hi = std::min(npats_, lo + kBlock);
          bool good = true;
          for (int64_t i = lo; i < hi; i++) {
            good &= plen_[i] == m;
            offs_[size_t(i)] = i * int64_t(m);
          }
          if (!good) ok.store(false, std::memory_order_relaxed);
        }
      });
      pool->wait();
      uniform = ok.load();
      offs_[size_t(npats)] = npats * int64_t(m);
    }
    if (!uniform) {
      int64_t total = 0;
      for (int64_t i = 0; i < npats; i++) {
        if (plen[i] < 0) throw Error(FM_ERR_PARAM, "negative pattern length");
        offs_[size_t(i)] = total;
        total += plen[i];
      }
      offs_[size_t(npats)] = total;
    }
    uniform_ = uniform ? m : 0;
    if (!uniform && npats > 0) {  // small batches: equal lengths found by the serial pass
      bool same = plen[0] > 0;
      for (int64_t i = 1; i < npats && same; i++) same = plen[i] == plen[0];
      if (same) uniform_ = plen[0];
    }
    flat_len_ = offs_[size_t(npats)];
    dst_ = static_cast<uint16_t*>(ix->h_stage[0].get(size_t(std::max<int64_t>(flat_len_, 1)) * 2));
    if (pool) {
      pool_ = pool;
      pool->run([this](int) { work(); });  // returns at once; joined in the destructor
    } else {
      work();  // small batch: gather inline
    }
  }
  ~PatternGatherer() {
    if (pool_) pool_->wait();
  }
  // returns once the symbols of patterns [0, hi) are in the flat buffer
  void wait_upto(int64_t hi) {
    const int64_t need = (std::min(hi, npats_) + kBlock - 1) / kBlock;
    while (ready_blocks_ < need) {
      if (done_[size_t(ready_blocks_)].load(std::memory_order_acquire)) ready_blocks_++;
      else std::this_thread::yield();
    }
  }
  const uint16_t* flat() const { return dst_; }
  int64_t flat_len() const { return flat_len_; }
  const int64_t* offs() const { return offs_; }
  int uniform() const { return uniform_; }

 private:
  static constexpr int64_t kBlock = 8192;  // patterns per unit of work; = the first copy chunk
  void work() {
    for (;;) {
      const int64_t b = next_.fetch_add(1);
      if (b >= nblocks_) return;
      const int64_t lo = b * kBlock, hi = std::min(npats_, lo + kBlock);
      for (int64_t i = lo; i < hi; i++)
        if (plen_[i]) std::memcpy(dst_ + offs_[size_t(i)], pats_[i], size_t(plen_[i]) * 2);
      done_[size_t(b)].store(1, std::memory_order_release);
    }
  }
  int64_t npats_;
  const int* plen_;
  const uint16_t* const* pats_;
  int64_t* offs_ = nullptr;
  uint16_t* dst_ = nullptr;
  int64_t flat_len_ = 0, nblocks_ = 0, ready_blocks_ = 0;
  int uniform_ = 0;
  WorkerPool* pool_ = nullptr;
  std::unique_ptr<std::atomic<unsigned char>[]> done_;
  std::atomic<int64_t> next_{0};
};

}  // namespace

// =============================================================================================
extern "C" {

const char* fm_last_error(void) { return g_last_error.c_str(); }

int fm_open(const char* path, int device, fm_index_t** out) { return open_impl(path, device, 0, 1, out); }

int fm_open_shard(const char* path, int device, int shard, int nshards, fm_index_t** out) {
  return open_impl(path, device, shard, nshards, out);
}

void fm_close(fm_index_t* ix) { destroy(ix); }

int fm_info(const fm_index_t* ix, fm_info_t* out) {
  if (!ix || !out) return fail(FM_ERR_PARAM, "fm_info: null argument");
  *out = ix->info;
  return FM_OK;
}

int fm_last_transfer(const fm_index_t* ix, int64_t* h2d_bytes, int64_t* d2h_bytes) {
  if (!ix) return fail(FM_ERR_PARAM, "fm_last_transfer: null index");
  if (h2d_bytes) *h2d_bytes = ix->last_h2d;
  if (d2h_bytes) *d2h_bytes = ix->last_d2h;
  return FM_OK;
}

int64_t fm_kernel_launches(const fm_index_t* ix) { return ix ? ix->launches : 0; }

int fm_set_lanes_per_query(fm_index_t* ix, int lanes) {
  if (!ix || (lanes != 1 && lanes != 2 && lanes != 4 && lanes != 8))
    return fail(FM_ERR_PARAM, "fm_set_lanes_per_query: lanes must be 1, 2, 4 or 8");
  ix->lanes_per_query = lanes;
  return FM_OK;
}

int fm_set_count_schedule(fm_index_t* ix, int merged, int lanes) {
  if (!ix) return fail(FM_ERR_PARAM, "fm_set_count_schedule: null index");
  const int bw = ix->im.block_words;
  if (ix->im.levels == 4) {
    if (lanes != 1 && lanes != 2)
      return fail(FM_ERR_PARAM, "fm_set_count_schedule: quad-level blocks run with 1 or 2 lanes per pattern");
    ix->count_sched = lanes == 1 ? kQuadSchedSplit : kQuadSched;
    return FM_OK;
  }
  if (ix->im.levels == 2) {  // paired-level blocks: merged schedule only, a lane owns whole 32-byte slices
    if (!((bw == 32 && (lanes == 1 || lanes == 2 || lanes == 4)) || (bw == 16 && (lanes == 1 || lanes == 2))))
      return fail(FM_ERR_PARAM, "fm_set_count_schedule: lane count not available for paired-level blocks");
    ix->count_sched = paired_sched(bw, lanes);
    return FM_OK;
  }
  // each lane must own at least 2 words of the block (sync) / 4 words (pair)
  const bool ok = merged ? ((bw == 32 && (lanes == 2 || lanes == 4 || lanes == 8)) ||
                            (bw == 16 && (lanes == 1 || lanes == 2 || lanes == 4)) ||
                            (bw == 8 && (lanes == 1 || lanes == 2)))
                         : ((bw == 32 && (lanes == 4 || lanes == 8)) || (bw == 16 && (lanes == 2 || lanes == 4)) ||
                            (bw == 8 && lanes == 2));
  if (!ok) return fail(FM_ERR_PARAM, "fm_set_count_schedule: lane count not available for this rank block size");
  ix->count_sched = merged ? sync_sched(bw, lanes) : lanes;
  return FM_OK;
}

int fm_set_default_block_bytes(int bytes) {
  if (bytes < 0 || bytes % 4 || !set_default_block_words(bytes / 4))
    return fail(FM_ERR_PARAM, "fm_set_default_block_bytes: 32, 64, 128 or 0");
  return FM_OK;
}

int fm_set_default_levels_per_block(int levels) {
  if (!set_default_levels_per_block(levels)) return fail(FM_ERR_PARAM, "fm_set_default_levels_per_block: 1, 2, 4 or 0");
  return FM_OK;
}

void* fm_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, std::max<size_t>(bytes, 1)) != cudaSuccess) return nullptr;
  return p;
}

void fm_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

namespace {
// A flat host batch, handle already locked: large ordered batches streamed (shape claimed from the first and
// last pattern, validated chunk by chunk behind the running kernel), the others validated first.
// to_host == false: the ranges stay in ix->d_out[0] / d_out[1] (locate).
int count_batch_locked(fm_index_t* ix, int64_t npats, const int32_t* plen, const uint16_t* flat, const int64_t* offs,
                       int64_t* first, int64_t* last, int sym_bytes, bool to_host) {
  {
    if (npats < 0 || (npats && (!plen || !offs || (to_host && !first)))) return fail(FM_ERR_PARAM, "fm_count_flat: bad argument");
    int64_t flat_len = 0;
    bool ordered = false;
    if (ix->info.first_row != 0 || ix->info.end_row != ix->info.total_length)
      return fail(FM_ERR_MISSING, "fm_count_flat: index is a shard; use the sharded driver");
    if (npats >= (1 << 17) && flat) {
      // Large batch: take its shape from the first and the last pattern, start searching at once
      // and validate chunk by chunk behind the running kernel (count_host, lazy).
      const int64_t claim_len = offs[npats - 1] + int64_t(plen[npats - 1]);
      int m = (plen[0] > 0 && offs[0] == 0 && plen[1] == plen[0] && offs[1] == plen[0]) ? plen[0] : 0;
      if (m && claim_len != npats * int64_t(m)) m = 0;
      // (a claim of more than 4 Ki symbols per pattern on average is not believed without a full check:
      // the device buffers are sized from it)
      if (claim_len >= 0 && offs[npats - 1] >= 0 && plen[npats - 1] >= 0 && claim_len <= npats * int64_t(4096)) {
        const int rc = count_host(ix, npats, plen, flat, claim_len, offs, first, last, /*in_order=*/true, m,
                                  to_host, /*lazy=*/true, nullptr, sym_bytes);
        if (rc != kRetryValidated) return rc;
      }
    }
    const auto t_scan = std::chrono::steady_clock::now();
    const BatchShape shape = scan_batch(npats, plen, offs);
    if (std::getenv("FEMTO_B200_TRACE"))
      std::fprintf(stderr, "[femto_b200 trace] batch validation %.3f ms\n",
                   std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_scan).count());
    if (shape.bad) return fail(FM_ERR_PARAM, "fm_count_flat: negative length/offset");
    if (shape.dense) {
      flat_len = shape.flat_len;
      ordered = true;
    } else if (check_patterns(npats, plen, offs, &flat_len, &ordered)) {
      return fail(FM_ERR_PARAM, "fm_count_flat: negative length/offset");
    }
    if (flat_len && !flat) return fail(FM_ERR_PARAM, "fm_count_flat: null pattern buffer");
    return count_host(ix, npats, plen, flat, flat_len, offs, first, last, ordered, shape.uniform, to_host, false, nullptr,
                      sym_bytes);
  }
}

int count_flat_impl(fm_index_t* ix, int64_t npats, const int32_t* plen, const uint16_t* flat, const int64_t* offs,
                    int64_t* first, int64_t* last, int sym_bytes) {
  return guarded(ix, "fm_count_flat", [&]() -> int {
    return count_batch_locked(ix, npats, plen, flat, offs, first, last, sym_bytes, /*to_host=*/true);
  });
}
}  // namespace

int fm_count_flat(fm_index_t* ix, int64_t npats, const int32_t* plen, const uint16_t* flat, const int64_t* offs,
                  int64_t* first, int64_t* last) {
  return count_flat_impl(ix, npats, plen, flat, offs, first, last, 2);
}

int fm_count_bytes(fm_index_t* ix, int64_t npats, const int32_t* plen, const uint8_t* text, const int64_t* offs,
                   int64_t* first, int64_t* last) {
  if (!ix) return fail(FM_ERR_PARAM, "fm_count_bytes: null index");
  if (ix->im.levels == 4 && ix->count_sched == kQuadSched)  // the kernel reads the bytes itself
    return count_flat_impl(ix, npats, plen, reinterpret_cast<const uint16_t*>(text), offs, first, last, 1);
  // other layouts / schedules: widen on the host (strtoalpha, src/main/index_types.h:85-97)
  if (npats < 0 || (npats && (!plen || !offs || !first))) return fail(FM_ERR_PARAM, "fm_count_bytes: bad argument");
  int64_t flat_len = 0;
  if (check_patterns(npats, plen, offs, &flat_len)) return fail(FM_ERR_PARAM, "fm_count_bytes: negative length/offset");
  if (flat_len && !text) return fail(FM_ERR_PARAM, "fm_count_bytes: null pattern buffer");
  std::vector<uint16_t> wide(size_t(std::max<int64_t>(flat_len, 1)));
  for (int64_t i = 0; i < flat_len; i++) wide[size_t(i)] = uint16_t(text[i]) + FM_CHARACTER_OFFSET;
  return count_flat_impl(ix, npats, plen, wide.data(), offs, first, last, 2);
}

int fm_count_stats(fm_index_t* ix, int64_t npats, const int32_t* plen, const uint16_t* flat, const int64_t* offs,
                   uint64_t* stats8) {
  uint64_t* stats4 = stats8;
  return guarded(ix, "fm_count_stats", [&]() -> int {
    if (npats < 0 || !stats4 || (npats && (!plen || !offs))) return fail(FM_ERR_PARAM, "fm_count_stats: bad argument");
    int64_t flat_len = 0;
    if (check_patterns(npats, plen, offs, &flat_len)) return fail(FM_ERR_PARAM, "fm_count_stats: negative length/offset");
    std::memset(stats4, 0, 8 * sizeof(uint64_t));
    if (npats == 0) return FM_OK;
    cudaStream_t s = ix->stream;
    int32_t* d_plen = static_cast<int32_t*>(ix->d_in[0].get(size_t(npats) * 4));
    uint16_t* d_flat = static_cast<uint16_t*>(ix->d_in[1].get(size_t(std::max<int64_t>(flat_len, 1)) * 2));
    int64_t* d_offs = static_cast<int64_t*>(ix->d_in[2].get(size_t(npats) * 8));
    int64_t* d_first = static_cast<int64_t*>(ix->d_out[0].get(size_t(npats) * 8));
    int64_t* d_last = static_cast<int64_t*>(ix->d_out[1].get(size_t(npats) * 8));
    unsigned long long* d_stats = static_cast<unsigned long long*>(ix->d_out[3].get(64));
    CK(cudaMemcpyAsync(d_plen, plen, size_t(npats) * 4, cudaMemcpyHostToDevice, s));
    if (flat_len) CK(cudaMemcpyAsync(d_flat, flat, size_t(flat_len) * 2, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_offs, offs, size_t(npats) * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(d_stats, 0, 64, s));
    CountArgs a{npats, d_plen, d_flat, d_offs, d_first, d_last};
    CK(launch_count(ix->im, a, ix->d_work, ix->count_sched, ix->sm_count, s, &ix->launches, d_stats));
    CK(cudaMemcpyAsync(stats4, d_stats, 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return FM_OK;
  });
}

int fm_probe_random_reads(fm_index_t* ix, int bytes_per_access, int steps, int64_t* accesses, double* ms) {
  return guarded(ix, "fm_probe_random_reads", [&]() -> int {
    if (!accesses || !ms || steps <= 0) return fail(FM_ERR_PARAM, "fm_probe_random_reads: bad argument");
    const uint64_t units = uint64_t(ix->info.rank_block_bytes) / uint64_t(std::max(bytes_per_access, 1));
    if (units == 0) return fail(FM_ERR_PARAM, "fm_probe_random_reads: image smaller than one access");
    cudaStream_t s = ix->stream;
    unsigned long long* d_sink = static_cast<unsigned long long*>(ix->d_out[3].get(64));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    cudaError_t le = launch_probe(ix->im.blocks, units, bytes_per_access, std::min(steps, 64), ix->sm_count, s,
                                  d_sink, accesses);  // warm-up
    if (le == cudaSuccess) {
      CK(cudaEventRecord(e0, s));
      le = launch_probe(ix->im.blocks, units, bytes_per_access, steps, ix->sm_count, s, d_sink, accesses);
      CK(cudaEventRecord(e1, s));
    }
    if (le != cudaSuccess) {
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
      return fail(FM_ERR_PARAM, "fm_probe_random_reads: bytes_per_access must be 32, 64 or 128");
    }
    ix->launches += 2;
    CK(cudaStreamSynchronize(s));
    float t = 0;
    CK(cudaEventElapsedTime(&t, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ms = double(t);
    return FM_OK;
  });
}

int fm_count(fm_index_t* ix, int npats, const int* plen, const uint16_t* const* pats, int64_t* first,
             int64_t* last) {
  return guarded(ix, "fm_count", [&]() -> int {
    if (npats < 0 || (npats && (!plen || !pats || !first))) return fail(FM_ERR_PARAM, "fm_count: bad argument");
    if (ix->info.first_row != 0 || ix->info.end_row != ix->info.total_length)
      return fail(FM_ERR_MISSING, "fm_count: index is a shard; use the sharded driver");
    if (npats == 0) return FM_OK;
    PatternGatherer g(ix, npats, plen, pats);  // gathers in the background; joined when it goes out of scope
    const std::function<void(int64_t)> ready = [&g](int64_t hi) { g.wait_upto(hi); };
    return count_host(ix, npats, reinterpret_cast<const int32_t*>(plen), g.flat(), g.flat_len(), g.offs(), first, last,
                      /*in_order=*/true, g.uniform(), /*to_host=*/true, /*lazy=*/false, &ready);
  });
}

int fm_count_device(fm_index_t* ix, int64_t npats, const int32_t* d_plen, const uint16_t* d_flat,
                    const int64_t* d_offs, int64_t* d_first, int64_t* d_last, void* stream) {
  return guarded(ix, "fm_count_device", [&]() -> int {
    if (npats < 0) return fail(FM_ERR_PARAM, "fm_count_device: negative npats");
    CountArgs a{npats, d_plen, d_flat, d_offs, d_first, d_last};
    // caller-stream launches rotate over work-queue slots 1..7 (slot 0 belongs to the host-buffer calls),
    // so up to 7 of them may be in flight on different streams
    ix->dev_slot = ix->dev_slot % 7 + 1;
    CK(launch_count(ix->im, a, ix->d_work + ix->dev_slot, ix->count_sched, ix->sm_count,
                    static_cast<cudaStream_t>(stream), &ix->launches));
    return FM_OK;
  });
}

int fm_count_shard_step(fm_index_t* ix, int64_t nstates, int64_t* d_state, const int32_t* d_plen,
                        const uint16_t* d_flat, const int64_t* d_offs, int32_t* d_dest, int nshards, void* stream) {
  return guarded(ix, "fm_count_shard_step", [&]() -> int {
    if (nstates < 0 || nshards < 1) return fail(FM_ERR_PARAM, "fm_count_shard_step: bad argument");
    ShardArgs a{};
    a.n = nstates;
    a.state = d_state;
    a.plen = d_plen;
    a.flat = d_flat;
    a.offs = d_offs;
    a.dest = d_dest;
    a.nshards = nshards;
    a.block_size = ix->info.block_size;
    a.nblocks = ix->info.num_blocks;
    CK(launch_count_shard(ix->im, a, ix->lanes_per_query, ix->sm_count, static_cast<cudaStream_t>(stream),
                          &ix->launches));
    return FM_OK;
  });
}

int fm_locate_rows(fm_index_t* ix, int64_t nrows, const int64_t* rows, int64_t* offsets) {
  return guarded(ix, "fm_locate_rows", [&]() -> int {
    if (nrows < 0 || (nrows && (!rows || !offsets))) return fail(FM_ERR_PARAM, "fm_locate_rows: bad argument");
    return locate_rows_host(ix, nrows, rows, offsets);
  });
}

int fm_walk_stats(fm_index_t* ix, int64_t nrows, const int64_t* rows, uint64_t* stats4) {
  return guarded(ix, "fm_walk_stats", [&]() -> int {
    if (nrows < 0 || !stats4 || (nrows && !rows)) return fail(FM_ERR_PARAM, "fm_walk_stats: bad argument");
    std::memset(stats4, 0, 4 * sizeof(uint64_t));
    if (nrows == 0) return FM_OK;
    cudaStream_t s = ix->stream;
    int64_t* d_rows = static_cast<int64_t*>(ix->d_in[3].get(size_t(nrows) * 8));
    int64_t* d_off = static_cast<int64_t*>(ix->d_out[2].get(size_t(nrows) * 8));
    unsigned long long* d_stats = static_cast<unsigned long long*>(ix->d_out[3].get(64));
    CK(cudaMemcpyAsync(d_rows, rows, size_t(nrows) * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(d_stats, 0, 64, s));
    WalkArgs w{};
    w.nrows = nrows;
    w.rows = d_rows;
    w.out_offset = d_off;
    w.status = ix->d_status;
    w.stats = d_stats;
    CK(cudaMemsetAsync(ix->d_status, 0, sizeof(int32_t), s));
    CK(launch_walk(ix->im, w, kWalkLocate, ix->d_work, ix->lanes_per_query, ix->sm_count, s, &ix->launches));
    CK(cudaMemcpyAsync(stats4, d_stats, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    const int st = walk_status(ix);
    if (st) return fail(st == 1 ? FM_ERR_PARAM : FM_ERR_INVALID, "fm_walk_stats: malformed walk (status " + std::to_string(st) + ")");
    return FM_OK;
  });
}

int fm_locate_rows_device(fm_index_t* ix, int64_t nrows, const int64_t* d_rows, int64_t* d_offsets, void* stream) {
  return guarded(ix, "fm_locate_rows_device", [&]() -> int {
    if (nrows < 0) return fail(FM_ERR_PARAM, "fm_locate_rows_device: negative nrows");
    WalkArgs w{};
    w.nrows = nrows;
    w.rows = d_rows;
    w.out_offset = d_offsets;
    w.status = ix->d_status + 2;  // caller-stream launches: read and cleared by fm_take_status
    ix->dev_slot = ix->dev_slot % 7 + 1;
    CK(launch_walk(ix->im, w, kWalkLocate, ix->d_work + ix->dev_slot, ix->lanes_per_query, ix->sm_count,
                   static_cast<cudaStream_t>(stream), &ix->launches));
    return FM_OK;
  });
}

int fm_locate_shard_step(fm_index_t* ix, int64_t nstates, int64_t* d_state, int32_t* d_dest, int nshards,
                         void* stream) {
  return guarded(ix, "fm_locate_shard_step", [&]() -> int {
    if (nstates < 0 || nshards < 1 || (nstates && (!d_state || !d_dest)))
      return fail(FM_ERR_PARAM, "fm_locate_shard_step: bad argument");
    if (nstates == 0) return FM_OK;
    WalkArgs w{};
    w.nrows = nstates;
    w.status = ix->d_status + 2;
    w.state = d_state;
    w.dest = d_dest;
    w.nshards = nshards;
    w.block_size = ix->info.block_size;
    w.nblocks = ix->info.num_blocks;
    ix->dev_slot = ix->dev_slot % 7 + 1;
    CK(launch_walk(ix->im, w, kWalkShard, ix->d_work + ix->dev_slot, ix->lanes_per_query, ix->sm_count,
                   static_cast<cudaStream_t>(stream), &ix->launches));
    return FM_OK;
  });
}

int fm_take_status(fm_index_t* ix, void* stream, int* status) {
  return guarded(ix, "fm_take_status", [&]() -> int {
    if (!status) return fail(FM_ERR_PARAM, "fm_take_status: null argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int32_t st = 0;
    CK(cudaMemcpyAsync(&st, ix->d_status + 2, sizeof(st), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (st) CK(cudaMemsetAsync(ix->d_status + 2, 0, sizeof(int32_t), s));
    *status = st;
    return FM_OK;
  });
}

// ---- device-initiated exchange over a BWT-range-sharded index (fm_mesh.cuh) --------------------------
int fm_mesh_create(fm_index_t* ix, int rank, int world, int64_t window, int cap_log2, fm_mesh_t** out) {
  if (!out) return fail(FM_ERR_PARAM, "fm_mesh_create: null argument");
  *out = nullptr;
  return guarded(ix, "fm_mesh_create", [&]() -> int {
    if (world < 1 || world > kMeshMaxRanks || rank < 0 || rank >= world || window < 0)
      return fail(FM_ERR_PARAM, "fm_mesh_create: bad rank / world / window");
    if (ix->im.levels != 4) return fail(FM_ERR_PARAM, "fm_mesh_create: needs the quad-level image (the default)");
    if (window == 0) window = 256 << 10;  // measured: throughput levels off from about 256 Ki own patterns in flight
    // a ring must hold every state in flight: world * (window + what the warps of one rank may claim
    // beyond the window between its check and their claims)
    // per warp (SMs x 4 CTAs x 8 warps at most): a chunk of 32 ids it may have claimed past the window, and
    // the three blocks of 16 ring indices it owns per ring
    const int64_t slack = int64_t(ix->sm_count) * 4 * 8 * (32 + 48);
    int shift = 10;
    while ((int64_t(1) << shift) < int64_t(world) * (window + slack)) shift++;
    if (cap_log2 > 0) {  // tests: a small ring that wraps many times (the caller also bounds the grid: fm_mesh_set_limits)
      if (cap_log2 < 8 || cap_log2 > 30) return fail(FM_ERR_PARAM, "fm_mesh_create: cap_log2 out of range");
      shift = cap_log2;
    }
    if (shift > 30) return fail(FM_ERR_PARAM, "fm_mesh_create: window too large");
    std::unique_ptr<fm_mesh> m(new fm_mesh());
    m->ix = ix;
    m->rank = rank;
    m->world = world;
    m->cap_shift = shift;
    m->window = static_cast<unsigned long long>(window);
    m->region_bytes = sizeof(MeshCtl) + (size_t(world) << shift) * 32;
    CK(cudaMalloc(&m->region, m->region_bytes));
    CK(cudaMemset(m->region, 0, m->region_bytes));
    CK(cudaDeviceSynchronize());
    int khz = 0;
    if (cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ix->device) == cudaSuccess && khz > 0) m->clock_khz = khz;
    m->peer_region[rank] = m->region;
    if (world == 1) m->connected = true;
    *out = m.release();
    return FM_OK;
  });
}

void fm_mesh_destroy(fm_mesh_t* m) {
  if (!m) return;
  cudaSetDevice(m->ix->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < m->world; r++)
    if (m->peer_ipc[r] && m->peer_region[r]) cudaIpcCloseMemHandle(m->peer_region[r]);
  if (m->region) cudaFree(m->region);
  delete m;
}

int fm_mesh_export(fm_mesh_t* m, void* handle, int64_t handle_bytes) {
  if (!m || !handle || handle_bytes < int64_t(sizeof(cudaIpcMemHandle_t)))
    return fail(FM_ERR_PARAM, "fm_mesh_export: needs FM_MESH_HANDLE_BYTES bytes");
  return guarded(m->ix, "fm_mesh_export", [&]() -> int {
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, m->region));
    std::memcpy(handle, &h, sizeof(h));
    return FM_OK;
  });
}

int fm_mesh_connect(fm_mesh_t* m, const void* handles, int64_t handle_stride) {
  if (!m || !handles || handle_stride < int64_t(sizeof(cudaIpcMemHandle_t)))
    return fail(FM_ERR_PARAM, "fm_mesh_connect: bad argument");
  return guarded(m->ix, "fm_mesh_connect", [&]() -> int {
    for (int r = 0; r < m->world; r++) {
      if (r == m->rank || m->peer_region[r]) continue;
      cudaIpcMemHandle_t h;
      std::memcpy(&h, static_cast<const char*>(handles) + size_t(r) * size_t(handle_stride), sizeof(h));
      void* p = nullptr;
      CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
      m->peer_region[r] = p;
      m->peer_ipc[r] = true;
    }
    m->connected = true;
    return FM_OK;
  });
}

int fm_mesh_connect_local(fm_mesh_t* m, fm_mesh_t* const* peers) {
  if (!m || !peers) return fail(FM_ERR_PARAM, "fm_mesh_connect_local: bad argument");
  return guarded(m->ix, "fm_mesh_connect_local", [&]() -> int {
    for (int r = 0; r < m->world; r++) {
      if (r == m->rank) continue;
      const fm_mesh* p = peers[r];
      if (!p || p->world != m->world || p->rank != r || p->cap_shift != m->cap_shift)
        return fail(FM_ERR_PARAM, "fm_mesh_connect_local: peer list does not match");
      if (p->ix->device != m->ix->device) {
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, m->ix->device, p->ix->device));
        if (!can) return fail(FM_ERR_IO, "fm_mesh_connect_local: no peer access between the devices");
        const cudaError_t e = cudaDeviceEnablePeerAccess(p->ix->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
        (void)cudaGetLastError();
      }
      m->peer_region[r] = p->region;
    }
    m->connected = true;
    return FM_OK;
  });
}

int fm_mesh_set_limits(fm_mesh_t* m, int max_ctas, double timeout_seconds) {
  if (!m || max_ctas < 0 || timeout_seconds < 0) return fail(FM_ERR_PARAM, "fm_mesh_set_limits: bad argument");
  m->max_ctas = max_ctas;
  if (timeout_seconds > 0) m->timeout_s = timeout_seconds;
  return FM_OK;
}

namespace {
int mesh_launch(fm_mesh_t* m, bool walk, const int32_t* d_plen, const uint16_t* d_flat, const int64_t* d_offs,
                int uniform_len, int64_t pid_lo, int64_t n_mine, int64_t* d_first, int64_t* d_last,
                const int64_t* d_rows, int64_t* d_offsets, void* stream) {
  if (!m) return fail(FM_ERR_PARAM, "fm_mesh: null mesh");
  fm_index* ix = m->ix;
  return guarded(ix, "fm_mesh", [&]() -> int {
    if (!m->connected) return fail(FM_ERR_PARAM, "fm_mesh: not connected");
    if (n_mine < 0 || pid_lo < 0 || pid_lo + n_mine > (int64_t(1) << 32))
      return fail(FM_ERR_PARAM, "fm_mesh: pattern ids must fit 32 bits");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    {  // the ring must hold every state in flight (fm_mesh.cuh): world * (window + one claim per group of the grid)
      const int64_t ctas = m->max_ctas > 0 ? m->max_ctas : int64_t(ix->sm_count) * 4;
      if ((int64_t(1) << m->cap_shift) < int64_t(m->world) * (int64_t(m->window) + ctas * 8 * (32 + 48)))
        return fail(FM_ERR_PARAM, "fm_mesh: ring too small for this window and grid (cap_log2 / fm_mesh_set_limits)");
    }
    m->epoch++;
    MeshArgs a{};
    char* base = static_cast<char*>(m->region);
    a.ctl = reinterpret_cast<MeshCtl*>(base);
    a.ring = reinterpret_cast<uint4*>(base + sizeof(MeshCtl));
    for (int r = 0; r < m->world; r++) {
      char* pb = static_cast<char*>(m->peer_region[r]);
      a.peer_ctl[r] = reinterpret_cast<MeshCtl*>(pb);
      a.peer_ring[r] = reinterpret_cast<uint4*>(pb + sizeof(MeshCtl));
    }
    a.rank = m->rank;
    a.world = m->world;
    a.cap_shift = m->cap_shift;
    // 8 bits of the batch number (1..255) travel in the tag of every message word; the rings are cleared
    // before that field repeats, so nothing 255 batches old can look current (callers keep batches apart:
    // see fm_mesh_count in the header)
    a.epoch = m->epoch;
    if (m->epoch % 255 == 0) CK(cudaMemsetAsync(base + sizeof(MeshCtl), 0, m->region_bytes - sizeof(MeshCtl), s));
    a.window = m->window;
    a.cap_mask = (1u << m->cap_shift) - 1u;
    a.eptag = (m->epoch % 255ull + 1ull) << 56;
    a.timeout_polls = static_cast<unsigned>(std::min(m->timeout_s * 1e6, 4e9));
    a.plen = d_plen;
    a.flat = d_flat;
    a.offs = d_offs;
    a.uniform_len = uniform_len;
    a.pid_lo = pid_lo;
    a.n_mine = n_mine;
    a.first = d_first;
    a.last = d_last;
    a.rows = d_rows;
    a.out_offset = d_offsets;
    a.block_size = ix->info.block_size;
    a.nblocks = ix->info.num_blocks;
    a.block_shift = ((a.block_size & (a.block_size - 1)) == 0 && a.nblocks <= kMeshOwnerTab)
                        ? __builtin_ctzll(static_cast<unsigned long long>(a.block_size)) : -1;
    {  // shard r holds the rows from shard_start[r] on: the first row of its first block (shard_of_block)
      int64_t b = 0;
      for (int r = 0; r <= kMeshMaxRanks; r++) {
        while (r < m->world && b < a.nblocks && shard_of_block(b, a.block_size, ix->info.total_length, m->world) < r) b++;
        if (r >= m->world) b = a.nblocks;
        a.shard_start[r] = std::min<int64_t>(b * a.block_size, ix->info.total_length);
      }
    }
    // the kernel's own counters start from zero; rank_done (written by the peers) is never reset
    CK(cudaMemsetAsync(base, 0, offsetof(MeshCtl, rank_done), s));
    if (walk) CK(launch_mesh_walk(ix->im, a, ix->sm_count, m->max_ctas, s, &ix->launches));
    else CK(launch_mesh_count(ix->im, a, ix->sm_count, m->max_ctas, s, &ix->launches));
    return FM_OK;
  });
}
}  // namespace

int fm_mesh_count(fm_mesh_t* m, const int32_t* d_plen, const uint16_t* d_flat, const int64_t* d_offs, int uniform_len,
                  int64_t pid_lo, int64_t n_mine, int64_t* d_first, int64_t* d_last, void* stream) {
  if (n_mine > 0 && (!d_flat || !d_first || (uniform_len <= 0 && (!d_plen || !d_offs))))
    return fail(FM_ERR_PARAM, "fm_mesh_count: null argument");
  return mesh_launch(m, false, d_plen, d_flat, d_offs, uniform_len, pid_lo, n_mine, d_first, d_last, nullptr, nullptr,
                     stream);
}

int fm_mesh_locate_rows(fm_mesh_t* m, int64_t nrows, const int64_t* d_rows, int64_t* d_offsets, void* stream) {
  if (nrows > 0 && (!d_rows || !d_offsets)) return fail(FM_ERR_PARAM, "fm_mesh_locate_rows: null argument");
  return mesh_launch(m, true, nullptr, nullptr, nullptr, 0, 0, nrows, nullptr, nullptr, d_rows, d_offsets, stream);
}

int fm_mesh_finish(fm_mesh_t* m, void* stream, int* status, uint64_t* stats8) {
  if (!m) return fail(FM_ERR_PARAM, "fm_mesh_finish: null mesh");
  return guarded(m->ix, "fm_mesh_finish", [&]() -> int {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    MeshCtl h;
    CK(cudaMemcpyAsync(&h, m->region, sizeof(MeshCtl), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (status) *status = h.status;
    if (stats8) std::memcpy(stats8, h.stats, sizeof(h.stats));
    if (h.status)
      return fail(h.status == 1 ? FM_ERR_CANCELED : FM_ERR_INVALID,
                  h.status == 1 ? "fm_mesh: a kernel waited too long for the other ranks (timed out)"
                                : "fm_mesh: malformed state");
    return FM_OK;
  });
}

int fm_locate_range(fm_index_t* ix, int64_t first, int64_t last, int64_t* offsets) {
  return guarded(ix, "fm_locate_range", [&]() -> int {
    if (last < first) return FM_OK;
    if (first < 0 || last >= ix->info.total_length || !offsets)
      return fail(FM_ERR_PARAM, "fm_locate_range: rows out of range");
    const int64_t n = last - first + 1;
    int64_t* rows = static_cast<int64_t*>(ix->h_stage[1].get(size_t(n) * 8));
    for (int64_t i = 0; i < n; i++) rows[i] = first + i;
    return locate_rows_host(ix, n, rows, offsets);
  });
}

namespace {
// fm_locate_flat / fm_locate: count -> ranges -> rows -> walks.  `provide(total)` is called once the
// number of offsets is known and returns the buffer they go to (NULL: it cannot hold them, FM_ERR_FULL
// with noccs / out_start filled) -- fm_locate allocates there, so it needs a single pass.
int locate_flat_impl(fm_index_t* ix, int64_t npats, const int32_t* plen, const uint16_t* flat, const int64_t* offs,
                     int max_occs_each, int32_t* noccs, int64_t* out_start,
                     const std::function<int64_t*(int64_t)>& provide) {
  return guarded(ix, "fm_locate_flat", [&]() -> int {
    if (npats < 0 || (npats && (!plen || !offs || !noccs || !out_start)))
      return fail(FM_ERR_PARAM, "fm_locate_flat: bad argument");
    if (npats == 0) return FM_OK;
    if (npats > INT32_MAX) return fail(FM_ERR_PARAM, "fm_locate_flat: too many patterns in one call");
    static const bool trace = std::getenv("FEMTO_B200_TRACE") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    // count (streamed like fm_count_flat when the batch is large and in order); the ranges stay in HBM
    int rc = count_batch_locked(ix, npats, plen, flat, offs, nullptr, nullptr, 2, /*to_host=*/false);
    if (rc) return rc;
    cudaStream_t s = ix->stream;
    const int64_t* d_first = static_cast<const int64_t*>(ix->d_out[0].get(size_t(npats) * 8));
    const int64_t* d_last = static_cast<const int64_t*>(ix->d_out[1].get(size_t(npats) * 8));
    // ranges -> clipped sizes -> row starts, on the device (do_locate_query's clip, server.c:4407-4415)
    int32_t* d_noccs = static_cast<int32_t*>(ix->d_in[0].get(size_t(npats) * 4));   // the pattern lengths are done with
    int64_t* d_start = static_cast<int64_t*>(ix->d_in[2].get(size_t(npats) * 8));   // so are their offsets
    const size_t scratch = expand_scratch_bytes(npats);
    char* d_scr = static_cast<char*>(ix->d_out[3].get(scratch + 128));
    int64_t* d_total = reinterpret_cast<int64_t*>(d_scr + ((scratch + 63) & ~size_t(63)));
    CK(launch_clip_and_scan(npats, d_first, d_last, max_occs_each, d_noccs, d_start, d_total, d_scr, scratch, s,
                            &ix->launches));
    int64_t* h_total = static_cast<int64_t*>(ix->h_ranges.get(64));
    CK(cudaMemcpyAsync(h_total, d_total, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(noccs, d_noccs, size_t(npats) * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(out_start, d_start, size_t(npats) * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const int64_t total = *h_total;
    const auto t1 = std::chrono::steady_clock::now();
    if (total == 0) return FM_OK;
    int64_t* out = provide(total);
    if (!out) return fail(FM_ERR_FULL, "fm_locate_flat: output buffer too small");
    // rows first[i] .. first[i] + noccs[i] - 1 of every pattern, then the sampled-SA walks
    int64_t* d_rows = static_cast<int64_t*>(ix->d_in[3].get(size_t(total) * 8));
    int64_t* d_off = static_cast<int64_t*>(ix->d_out[2].get(size_t(total) * 8));
    CK(launch_expand_rows(npats, total, d_first, d_start, d_rows, s, &ix->launches));
    WalkArgs w{};
    w.nrows = total;
    w.rows = d_rows;
    w.out_offset = d_off;
    w.status = ix->d_status;
    CK(cudaMemsetAsync(ix->d_status, 0, sizeof(int32_t), s));
    CK(launch_walk(ix->im, w, kWalkLocate, ix->d_work, ix->lanes_per_query, ix->sm_count, s, &ix->launches));
    CK(cudaMemcpyAsync(out, d_off, size_t(total) * 8, cudaMemcpyDeviceToHost, s));
    const int st = walk_status(ix);
    if (trace) {
      const auto t2 = std::chrono::steady_clock::now();
      auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
      std::fprintf(stderr, "[femto_b200 trace] locate %lld patterns, %lld rows: count + ranges %.3f ms, rows + walks %.3f\n",
                   (long long)npats, (long long)total, ms(t0, t1), ms(t1, t2));
    }
    if (st) return fail(st == 1 ? FM_ERR_PARAM : FM_ERR_INVALID, "locate: malformed walk (status " + std::to_string(st) + ")");
    return FM_OK;
  });
}

}  // namespace

int fm_locate_flat(fm_index_t* ix, int64_t npats, const int32_t* plen, const uint16_t* flat, const int64_t* offs,
                   int max_occs_each, int32_t* noccs, int64_t* out_start, int64_t* out, int64_t out_cap) {
  return locate_flat_impl(ix, npats, plen, flat, offs, max_occs_each, noccs, out_start,
                          [&](int64_t total) -> int64_t* { return (out && total <= out_cap) ? out : nullptr; });
}

int fm_locate(fm_index_t* ix, int npats, const int* plen, const uint16_t* const* pats, int max_occs_each,
              int* noccs, int64_t** offsets) {
  if (!ix) return fail(FM_ERR_PARAM, "fm_locate: null index");
  if (npats < 0 || (npats && (!plen || !pats || !noccs || !offsets))) return fail(FM_ERR_PARAM, "fm_locate: bad argument");
  std::vector<int64_t> offs(size_t(npats) + 1), start(size_t(npats) + 1);
  std::vector<uint16_t> flat;
  int64_t total = 0;
  for (int i = 0; i < npats; i++) {
    if (plen[i] < 0) return fail(FM_ERR_PARAM, "fm_locate: negative pattern length");
    offs[size_t(i)] = total;
    total += plen[i];
  }
  flat.resize(size_t(std::max<int64_t>(total, 1)));
  for (int i = 0; i < npats; i++)
    if (plen[i]) std::memcpy(flat.data() + offs[size_t(i)], pats[i], size_t(plen[i]) * 2);
  // one pass: the result buffer is allocated when the number of offsets is known
  std::vector<int64_t> outv;
  bool oom = false;
  const int rc = locate_flat_impl(ix, npats, reinterpret_cast<const int32_t*>(plen), flat.data(), offs.data(),
                                  max_occs_each, noccs, start.data(), [&](int64_t total) -> int64_t* {
                                    try {
                                      outv.resize(size_t(total));
                                    } catch (const std::bad_alloc&) {
                                      oom = true;
                                      return nullptr;
                                    }
                                    return outv.data();
                                  });
  if (oom) return fail(FM_ERR_MEM, "fm_locate: out of memory");
  if (rc) return rc;
  for (int i = 0; i < npats; i++) {
    offsets[i] = nullptr;
    if (noccs[i] > 0) {
      offsets[i] = static_cast<int64_t*>(std::malloc(sizeof(int64_t) * size_t(noccs[i])));
      if (!offsets[i]) {
        for (int j = 0; j < i; j++) { std::free(offsets[j]); offsets[j] = nullptr; }
        return fail(FM_ERR_MEM, "fm_locate: out of memory");
      }
      std::memcpy(offsets[i], outv.data() + start[size_t(i)], sizeof(int64_t) * size_t(noccs[i]));
    }
  }
  return FM_OK;
}

int fm_back_step(fm_index_t* ix, int64_t nrows, const int64_t* rows, int32_t* ch, int64_t* next, int64_t* offset) {
  return guarded(ix, "fm_back_step", [&]() -> int {
    if (nrows < 0 || (nrows && (!rows || !ch || !next || !offset))) return fail(FM_ERR_PARAM, "fm_back_step: bad argument");
    if (nrows == 0) return FM_OK;
    cudaStream_t s = ix->stream;
    int64_t* d_rows = static_cast<int64_t*>(ix->d_in[3].get(size_t(nrows) * 8));
    int64_t* d_off = static_cast<int64_t*>(ix->d_out[2].get(size_t(nrows) * 8));
    int64_t* d_next = static_cast<int64_t*>(ix->d_out[3].get(size_t(nrows) * 8));
    int32_t* d_ch = static_cast<int32_t*>(ix->d_out[1].get(size_t(nrows) * 4));
    CK(cudaMemcpyAsync(d_rows, rows, size_t(nrows) * 8, cudaMemcpyHostToDevice, s));
    WalkArgs w{};
    w.nrows = nrows;
    w.rows = d_rows;
    w.out_offset = d_off;
    w.out_next = d_next;
    w.out_ch = d_ch;
    w.status = ix->d_status;
    CK(cudaMemsetAsync(ix->d_status, 0, sizeof(int32_t), s));
    CK(launch_walk(ix->im, w, kWalkStep, ix->d_work, ix->lanes_per_query, ix->sm_count, s, &ix->launches));
    CK(cudaMemcpyAsync(offset, d_off, size_t(nrows) * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(next, d_next, size_t(nrows) * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(ch, d_ch, size_t(nrows) * 4, cudaMemcpyDeviceToHost, s));
    const int st = walk_status(ix);
    if (st) return fail(st == 1 ? FM_ERR_PARAM : FM_ERR_INVALID, "fm_back_step: bad row (status " + std::to_string(st) + ")");
    return FM_OK;
  });
}

int fm_occ(fm_index_t* ix, int64_t n, const uint16_t* ch, const int64_t* rows, int64_t* c_plus_occ) {
  return guarded(ix, "fm_occ", [&]() -> int {
    if (n < 0 || (n && (!ch || !rows || !c_plus_occ))) return fail(FM_ERR_PARAM, "fm_occ: bad argument");
    if (n == 0) return FM_OK;
    cudaStream_t s = ix->stream;
    uint16_t* d_ch = static_cast<uint16_t*>(ix->d_in[1].get(size_t(n) * 2));
    int64_t* d_rows = static_cast<int64_t*>(ix->d_in[3].get(size_t(n) * 8));
    int64_t* d_out = static_cast<int64_t*>(ix->d_out[2].get(size_t(n) * 8));
    CK(cudaMemcpyAsync(d_ch, ch, size_t(n) * 2, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_rows, rows, size_t(n) * 8, cudaMemcpyHostToDevice, s));
    OccArgs a{n, d_ch, d_rows, d_out};
    CK(launch_occ(ix->im, a, ix->d_work, ix->lanes_per_query, ix->sm_count, s, &ix->launches));
    CK(cudaMemcpyAsync(c_plus_occ, d_out, size_t(n) * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    for (int64_t i = 0; i < n; i++)
      if (c_plus_occ[i] < 0) return fail(FM_ERR_PARAM, "fm_occ: symbol or row out of range");
    return FM_OK;
  });
}

int fm_backward_step(fm_index_t* ix, int64_t n, const int64_t* first, const int64_t* last, const uint16_t* ch,
                     int64_t* new_first, int64_t* new_last) {
  return guarded(ix, "fm_backward_step", [&]() -> int {
    if (n < 0 || (n && (!first || !last || !ch || !new_first || !new_last)))
      return fail(FM_ERR_PARAM, "fm_backward_step: bad argument");
    if (n == 0) return FM_OK;
    // two Occ evaluations per range: rows first-1 and last (do_backward_search_query, server.c:980-1112)
    std::vector<uint16_t> qc(size_t(2 * n));
    std::vector<int64_t> qr(size_t(2 * n)), res(size_t(2 * n));
    for (int64_t i = 0; i < n; i++) {
      if (ch[i] >= kAlpha || last[i] < 0 || last[i] >= ix->info.total_length || first[i] < 0 || first[i] > last[i] + 1)
        return fail(FM_ERR_PARAM, "fm_backward_step: symbol or range out of bounds");
      qc[size_t(2 * i)] = qc[size_t(2 * i + 1)] = ch[i];
      qr[size_t(2 * i)] = first[i] > 0 ? first[i] - 1 : last[i];  // first == 0: Occ(c,-1) = 0, slot unused
      qr[size_t(2 * i + 1)] = last[i];
    }
    cudaStream_t s = ix->stream;
    uint16_t* d_ch = static_cast<uint16_t*>(ix->d_in[1].get(size_t(2 * n) * 2));
    int64_t* d_rows = static_cast<int64_t*>(ix->d_in[3].get(size_t(2 * n) * 8));
    int64_t* d_out = static_cast<int64_t*>(ix->d_out[2].get(size_t(2 * n) * 8));
    CK(cudaMemcpyAsync(d_ch, qc.data(), size_t(2 * n) * 2, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_rows, qr.data(), size_t(2 * n) * 8, cudaMemcpyHostToDevice, s));
    OccArgs a{2 * n, d_ch, d_rows, d_out};
    CK(launch_occ(ix->im, a, ix->d_work, ix->lanes_per_query, ix->sm_count, s, &ix->launches));
    CK(cudaMemcpyAsync(res.data(), d_out, size_t(2 * n) * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    for (int64_t i = 0; i < n; i++) {
      if (res[size_t(2 * i + 1)] < 0 || (first[i] > 0 && res[size_t(2 * i)] < 0))
        return fail(FM_ERR_MISSING, "fm_backward_step: row not resident (sharded index)");
      new_first[i] = first[i] > 0 ? res[size_t(2 * i)] : ix->C_host[ch[i]];
      new_last[i] = res[size_t(2 * i + 1)] - 1;
    }
    return FM_OK;
  });
}

int fm_doc_info(const fm_index_t* ix, int64_t doc, int64_t* doc_len, int64_t* eof_row) {
  if (!ix || !doc_len || !eof_row) return fail(FM_ERR_PARAM, "fm_doc_info: null argument");
  if (doc < 0 || doc >= ix->info.num_documents) return fail(FM_ERR_PARAM, "fm_doc_info: no such document");
  const int64_t e = ix->doc_ends[size_t(doc)];
  *doc_len = doc == 0 ? e : e - ix->doc_ends[size_t(doc - 1)];
  *eof_row = ix->doc_eof_rows[size_t(doc)];
  return FM_OK;
}

int fm_resolve(const fm_index_t* ix, int64_t n, const int64_t* offsets, int64_t* doc, int64_t* doc_off) {
  if (!ix || n < 0 || (n && (!offsets || !doc || !doc_off))) return fail(FM_ERR_PARAM, "fm_resolve: bad argument");
  for (int64_t i = 0; i < n; i++) {
    // previous document = last d with doc_ends[d] <= offset (bsearch_int64_ntoh_arr, src/utils/util.c:346)
    const auto it = std::upper_bound(ix->doc_ends.begin(), ix->doc_ends.end(), offsets[i]);
    const int64_t prev = int64_t(it - ix->doc_ends.begin()) - 1;
    if (prev < 0) { doc[i] = 0; doc_off[i] = offsets[i]; }
    else { doc[i] = prev + 1; doc_off[i] = offsets[i] - ix->doc_ends[size_t(prev)]; }
  }
  return FM_OK;
}

int fm_doc_name(const fm_index_t* ix, int64_t doc, void* out, int64_t out_cap, int64_t* out_len) {
  if (!ix || !out_len) return fail(FM_ERR_PARAM, "fm_doc_name: null argument");
  if (doc < 0 || doc >= ix->info.num_documents) return fail(FM_ERR_PARAM, "fm_doc_name: no such document");
  const int64_t lo = ix->doc_info_off[size_t(doc)], len = ix->doc_info_off[size_t(doc) + 1] - lo;
  *out_len = len;
  if (len > out_cap) return fail(FM_ERR_FULL, "fm_doc_name: output buffer too small");
  if (len > 0) {
    if (!out) return fail(FM_ERR_PARAM, "fm_doc_name: null output");
    std::memcpy(out, ix->doc_info_bytes.data() + lo, size_t(len));
  }
  return FM_OK;
}

int fm_chunk_documents(const fm_index_t* ix, int64_t row, int64_t* chunk_first, int64_t* chunk_last, int64_t* docs,
                       int64_t docs_cap, int64_t* ndocs) {
  if (!ix || !chunk_first || !chunk_last || !ndocs) return fail(FM_ERR_PARAM, "fm_chunk_documents: null argument");
  *ndocs = 0;
  try {
    std::vector<int64_t> d;
    chunk_documents(ix->hdr, ix->info.first_row, ix->info.end_row, ix->im.first_bucket, ix->chunk_bytes, ix->chunk_off,
                    ix->chunk_count, ix->chunk_dir_rel, row, chunk_first, chunk_last, &d);
    *ndocs = int64_t(d.size());
    if (*ndocs > docs_cap) return fail(FM_ERR_FULL, "fm_chunk_documents: output buffer too small");
    if (*ndocs && !docs) return fail(FM_ERR_PARAM, "fm_chunk_documents: null output");
    std::copy(d.begin(), d.end(), docs);
    return FM_OK;
  } catch (const Error& e) {
    return fail(e.code, std::string("fm_chunk_documents: ") + e.what());
  } catch (const std::bad_alloc&) {
    return fail(FM_ERR_MEM, "fm_chunk_documents: out of memory");
  }
}

int fm_range_documents(fm_index_t* ix, int64_t first, int64_t last, int64_t* docs, int64_t docs_cap, int64_t* ndocs) {
  if (!ix || !ndocs) return fail(FM_ERR_PARAM, "fm_range_documents: null argument");
  *ndocs = 0;
  if (last < first) return FM_OK;
  if (first < 0 || last >= ix->info.total_length) return fail(FM_ERR_PARAM, "fm_range_documents: rows out of range");
  std::vector<int64_t> found, rows;
  try {
    // As range_to_results for documents (src/main/server.c:4549-4889): walk the range chunk by chunk; a chunk
    // that lies inside the range contributes its stored document list, the rows of a chunk that sticks out
    // of the range (at most one at each end) are located one by one.  An index built without chunks
    // locates every row.
    bool chunks = ix->info.chunk_size > 0 && !ix->chunk_bytes.empty();
    for (int64_t i = first; i <= last;) {
      int64_t cf = i, cl = last;
      std::vector<int64_t> d;
      if (chunks) {
        try {
          chunk_documents(ix->hdr, ix->info.first_row, ix->info.end_row, ix->im.first_bucket, ix->chunk_bytes,
                          ix->chunk_off, ix->chunk_count, ix->chunk_dir_rel, i, &cf, &cl, &d);
        } catch (const Error& e) {
          if (e.code != FM_ERR_MISSING) throw;
          chunks = false;
          cf = i;
          cl = last;
        }
      }
      if (chunks && cf >= first && cl <= last) {
        found.insert(found.end(), d.begin(), d.end());
      } else {
        const int64_t lo = std::max(cf, first), hi = std::min(cl, last);
        for (int64_t r = lo; r <= hi; r++) rows.push_back(r);
      }
      i = std::min(cl, last) + 1;
    }
    if (!rows.empty()) {
      std::vector<int64_t> offs(rows.size());
      const int rc = fm_locate_rows(ix, int64_t(rows.size()), rows.data(), offs.data());  // SA[row] on the GPU
      if (rc) return rc;
      // resolve_location (index.c:1587-1611) per offset
      for (int64_t o : offs)
        found.push_back(int64_t(std::upper_bound(ix->doc_ends.begin(), ix->doc_ends.end(), o) - ix->doc_ends.begin()));
    }
  } catch (const Error& e) {
    return fail(e.code, std::string("fm_range_documents: ") + e.what());
  } catch (const std::bad_alloc&) {
    return fail(FM_ERR_MEM, "fm_range_documents: out of memory");
  }
  std::sort(found.begin(), found.end());
  found.erase(std::unique(found.begin(), found.end()), found.end());
  *ndocs = int64_t(found.size());
  if (*ndocs > docs_cap) return fail(FM_ERR_FULL, "fm_range_documents: output buffer too small");
  if (*ndocs && !docs) return fail(FM_ERR_PARAM, "fm_range_documents: null output");
  std::copy(found.begin(), found.end(), docs);
  return FM_OK;
}

int fm_extract(fm_index_t* ix, int64_t doc, uint16_t* out, int64_t out_cap, int64_t* out_len) {
  return guarded(ix, "fm_extract", [&]() -> int {
    if (!out_len) return fail(FM_ERR_PARAM, "fm_extract: null argument");
    if (doc < 0 || doc >= ix->info.num_documents) return fail(FM_ERR_PARAM, "fm_extract: no such document");
    const int64_t e = ix->doc_ends[size_t(doc)];
    const int64_t len = (doc == 0 ? e : e - ix->doc_ends[size_t(doc - 1)]) - 1;
    *out_len = len;
    if (len > out_cap) return fail(FM_ERR_FULL, "fm_extract: output buffer too small");
    if (len <= 0) return FM_OK;
    if (!out) return fail(FM_ERR_PARAM, "fm_extract: null output");
    cudaStream_t s = ix->stream;
    int64_t hrow[3] = {ix->doc_eof_rows[size_t(doc)], len, 0};
    int64_t* d_par = static_cast<int64_t*>(ix->d_in[3].get(3 * 8));
    uint16_t* d_sym = static_cast<uint16_t*>(ix->d_out[0].get(size_t(len) * 2));
    CK(cudaMemcpyAsync(d_par, hrow, sizeof(hrow), cudaMemcpyHostToDevice, s));
    WalkArgs w{};
    w.nrows = 1;
    w.rows = d_par;
    w.nsteps = d_par + 1;
    w.sym_off = d_par + 2;
    w.out_sym = d_sym;
    w.status = ix->d_status;
    CK(cudaMemsetAsync(ix->d_status, 0, sizeof(int32_t), s));
    CK(launch_walk(ix->im, w, kWalkExtract, ix->d_work, ix->lanes_per_query, ix->sm_count, s, &ix->launches));
    CK(cudaMemcpyAsync(out, d_sym, size_t(len) * 2, cudaMemcpyDeviceToHost, s));
    const int st = walk_status(ix);
    if (st) return fail(FM_ERR_INVALID, "fm_extract: malformed walk (status " + std::to_string(st) + ")");
    return FM_OK;
  });
}

// ---- generic requests (femto.h:75-149) ---------------------------------------------------------------
int fm_generic_request(fm_index_t* ix, const char* request, char** response) {
  if (!ix || !request || !response) return fail(FM_ERR_PARAM, "fm_generic_request: null argument");
  *response = nullptr;
  // the request word, as femto_create_generic_request_err tells them apart (src/main/femto.c:595-611)
  enum { kRows = 1, kAll, kLeft, kRight } type;
  const char* p = request;
  auto starts = [&](const char* word) {
    const size_t n = std::strlen(word);
    if (std::strncmp(request, word, n) != 0) return false;
    p = request + n;
    return true;
  };
  if (starts("string_rows_left")) type = kLeft;
  else if (starts("string_rows_right")) type = kRight;
  else if (starts("string_rows_all")) type = kAll;
  else if (starts("string_rows")) type = kRows;
  else if (starts("find_strings") || starts("docs_for_range") || starts("find_docs"))
    return fail(FM_ERR_INVALID, "fm_generic_request: this request needs femto's query parser / result encoder; "
                                "only the string_rows* requests run on this engine");
  else return fail(FM_ERR_INVALID, "Bad request");
  // the pattern: integers (any base sscanf's %i takes), each a byte value: symbol = CHARACTER_OFFSET + value
  std::vector<uint16_t> pat;
  const char* end = request + std::strlen(request);
  while (p != end) {
    while (*p == ' ') p++;
    if (p == end) break;
    int num = 0, used = 0;
    if (std::sscanf(p, "%i%n", &num, &used) <= 0) return fail(FM_ERR_INVALID, "Could not scan");
    const int ch = FM_CHARACTER_OFFSET + num;
    if (ch >= FM_ALPHA_SIZE || ch < 0) return fail(FM_ERR_INVALID, "Bad request character");
    pat.push_back(uint16_t(ch));
    p += used;
  }
  // the batch: the pattern itself, or the pattern with every symbol of the alphabet (0 .. ALPHA_SIZE-1, escape
  // codes included) put in front of it ("left") and / or behind it ("right"): setup_string_rows_*_query,
  // src/main/server.c:4194-4320
  const int m = int(pat.size());
  const int per = FM_ALPHA_SIZE;
  const int n = type == kRows ? 1 : type == kAll ? 2 * per : per;
  const int mlen = type == kRows ? m : m + 1;
  const size_t nn = static_cast<size_t>(n);
  std::vector<int32_t> plen(nn, mlen);
  std::vector<int64_t> offs(nn), first(nn), last(nn);
  std::vector<uint16_t> flat(static_cast<size_t>(std::max(1, n * mlen)));
  for (int i = 0; i < n; i++) {
    offs[size_t(i)] = int64_t(i) * mlen;
    uint16_t* dst = flat.data() + size_t(i) * size_t(mlen);
    if (type == kRows) {
      std::copy(pat.begin(), pat.end(), dst);
    } else {
      const bool left = type == kLeft || (type == kAll && i < per);
      const uint16_t c = uint16_t(i % per);
      if (left) { dst[0] = c; std::copy(pat.begin(), pat.end(), dst + 1); }
      else { std::copy(pat.begin(), pat.end(), dst); dst[m] = c; }
    }
  }
  const int rc = fm_count_flat(ix, n, plen.data(), flat.data(), offs.data(), first.data(), last.data());
  if (rc) return rc;
  // the answer, character for character what femto_response_for_generic_request_err prints (femto.c:919-995)
  std::string out;
  char line[128];
  if (type == kRows) {
    std::snprintf(line, sizeof(line), "{\"range\":[%lld,%lld]}\n", (long long)first[0], (long long)last[0]);
    out = line;
  } else {
    bool on_right = type == kRight, first_entry = true;
    out += std::string("{\"") + (on_right ? "right" : "left") + "\":[\n";
    for (int i = 0; i < n; i++) {
      if (i == per) {
        on_right = true;
        out += "\n ],\n \"right\":[\n";
        first_entry = true;
      }
      const int ch = (i >= per ? i - per : i) - FM_CHARACTER_OFFSET;
      if (first[size_t(i)] <= last[size_t(i)]) {
        if (!first_entry) out += ",\n";
        first_entry = false;
        std::snprintf(line, sizeof(line), "  {\"ch\":%i, \"range\":[%lld,%lld]}", ch, (long long)first[size_t(i)],
                      (long long)last[size_t(i)]);
        out += line;
      }
    }
    out += "\n ]\n}\n";
  }
  char* buf = static_cast<char*>(std::malloc(out.size() + 1));
  if (!buf) return fail(FM_ERR_MEM, "fm_generic_request: out of memory");
  std::memcpy(buf, out.c_str(), out.size() + 1);
  *response = buf;
  return FM_OK;
}

int fm_extract_batch(fm_index_t* ix, int64_t ndocs, const int64_t* docs, uint16_t* out, int64_t out_cap,
                     int64_t* out_start) {
  return guarded(ix, "fm_extract_batch", [&]() -> int {
    if (ndocs < 0 || !out_start || (ndocs && !docs)) return fail(FM_ERR_PARAM, "fm_extract_batch: bad argument");
    // per document: its EOF row, doc_len - 1 LF steps, where its symbols go (do_extract_document_query,
    // src/main/server.c:6364-6437: strictly sequential per document, so the documents run side by side)
    std::vector<int64_t> par(size_t(3 * std::max<int64_t>(ndocs, 1)));
    int64_t total = 0;
    for (int64_t k = 0; k < ndocs; k++) {
      const int64_t d = docs[k];
      if (d < 0 || d >= ix->info.num_documents) return fail(FM_ERR_PARAM, "fm_extract_batch: no such document");
      const int64_t e = ix->doc_ends[size_t(d)];
      const int64_t len = (d == 0 ? e : e - ix->doc_ends[size_t(d - 1)]) - 1;
      out_start[k] = total;
      par[size_t(k)] = ix->doc_eof_rows[size_t(d)];
      par[size_t(ndocs + k)] = len;
      par[size_t(2 * ndocs + k)] = total;
      total += len;
    }
    out_start[ndocs] = total;
    if (total > out_cap) return fail(FM_ERR_FULL, "fm_extract_batch: output buffer too small");
    if (total == 0) return FM_OK;
    if (!out) return fail(FM_ERR_PARAM, "fm_extract_batch: null output");
    cudaStream_t s = ix->stream;
    int64_t* d_par = static_cast<int64_t*>(ix->d_in[3].get(size_t(3 * ndocs) * 8));
    uint16_t* d_sym = static_cast<uint16_t*>(ix->d_out[0].get(size_t(total) * 2));
    CK(cudaMemcpyAsync(d_par, par.data(), size_t(3 * ndocs) * 8, cudaMemcpyHostToDevice, s));
    WalkArgs w{};
    w.nrows = ndocs;
    w.rows = d_par;
    w.nsteps = d_par + ndocs;
    w.sym_off = d_par + 2 * ndocs;
    w.out_sym = d_sym;
    w.status = ix->d_status;
    CK(cudaMemsetAsync(ix->d_status, 0, sizeof(int32_t), s));
    CK(launch_walk(ix->im, w, kWalkExtract, ix->d_work, ix->lanes_per_query, ix->sm_count, s, &ix->launches));
    CK(cudaMemcpyAsync(out, d_sym, size_t(total) * 2, cudaMemcpyDeviceToHost, s));
    const int st = walk_status(ix);
    if (st) return fail(FM_ERR_INVALID, "fm_extract_batch: malformed walk (status " + std::to_string(st) + ")");
    return FM_OK;
  });
}

}  // extern "C"
