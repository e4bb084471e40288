// fm_mesh.cuh -- device-initiated state exchange for the BWT-range-sharded index (SURVEY.md
// section 8e, second case: the index exceeds one GPU's HBM and is split by data block,
// reference partition unit src/main/index.h:83-100, src/main/block_storage.c:257-267).
//
// The LF mapping scatters the rows a pattern needs over the whole BWT, so the STATE of a
// backward search (or of a sampled-SA walk) travels to the GPU that owns the row it needs next.
// Here the GPUs do that themselves: every rank runs ONE persistent kernel per batch; a lane group
// that finds its next row on another shard stores its 32-byte state straight into the owner's
// inbox over NVLink peer memory and takes the next state from its own inbox.  No host round trip,
// no collective inside a batch; termination by counters the kernels exchange the same way.
//
// Inbox of a rank: one ring per SOURCE rank, so a slot index is claimed with an atomic in the
// sender's own memory (one atomic per destination and warp) -- the transfer itself is a posted
// store, nothing comes back over the link.  On the consumer side a warp owns up to two BLOCKS of 16
// consecutive indices per ring, handed out in index order by a ticket counter in the rank's own
// memory to whichever warp has just used one up; the ticket is requested in one round and used in
// the next.  A warp never waits for a block: whenever it has idle lane groups it looks at the blocks
// it owns, whose slots were copied to shared memory (cp.async, L2 only) at the end of the round
// before.  A message is four 64-bit words that each carry a tag = (batch, lap of the ring) next to
// 48 bits of payload; a slot is consumed when all four tags are the expected one -- no flag, no
// fence, no assumption about the order in which NVLink delivers the words.
//
// Capacity: a state is in exactly one place, so the unconsumed messages of a ring are at most the
// states in flight; blocks are handed out in index order to whichever warp is free, so they span
// at most that many indices plus two (partly filled) blocks per warp.  Every rank injects new
// patterns only while fewer than `window` of its own are unfinished, and the rings hold
// world * (window + slack) slots, so a producer can never lap an unconsumed slot.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

#include "fm_image.hpp"

namespace fmb {

constexpr int kMeshMaxRanks = 16;
constexpr int kMeshBlock = 16;  // consecutive ring indices owned by one consumer warp
constexpr int kMeshOwnerTab = 2048;  // data blocks the kernels' block -> owner table holds

// Control block at the start of every rank's exported region.
struct MeshCtl {
  // written by the owner's kernel only
  unsigned long long out_tail[kMeshMaxRanks];  // next index in the ring (me -> dst) at dst
  unsigned long long head_block[kMeshMaxRanks];  // next block of 16 indices of my ring src that no warp owns yet
  unsigned long long injected;                 // patterns of my batch taken so far
  unsigned long long done_count;               // patterns of my batch delivered
  unsigned long long inflight;                 // of my batch: taken and not yet delivered
  unsigned long long stats[8];                 // sent, received, rounds, occ pairs, occ singles, - , injected, -
  int status;                                  // 0 ok, 1 timed out, 2 malformed message
  int pad0;
  unsigned long long pad1[4];
  // written by the peers (rank r writes entry r): the last epoch for which rank r holds all its results
  unsigned long long rank_done[kMeshMaxRanks];
};
static_assert(sizeof(MeshCtl) % 128 == 0, "control block keeps the rings 128-byte aligned");

struct MeshArgs {
  MeshCtl* ctl;                        // mine
  uint4* ring;                         // my inbox: ring[src][cap][2]
  MeshCtl* peer_ctl[kMeshMaxRanks];    // every rank's control block (mine included)
  uint4* peer_ring[kMeshMaxRanks];     // every rank's inbox
  int rank, world;
  int cap_shift;                       // slots per ring = 1 << cap_shift
  unsigned cap_mask;                   // (1 << cap_shift) - 1
  unsigned long long epoch;            // batch number, starting at 1
  unsigned long long eptag;            // the batch's part of every message tag, in place: (epoch % 255 + 1) << 56
  unsigned long long window;           // own patterns in flight at most
  unsigned timeout_polls;              // a warp idle for more polls (~1 us each) gives up (status 1)
  // the batch: patterns of ALL ranks (replicated), indexed by global pattern id
  const int32_t* plen;
  const uint16_t* flat;
  const int64_t* offs;
  int uniform_len;                     // > 0: pattern p is flat[p * uniform_len ...), plen / offs not read
  int64_t pid_lo, n_mine;              // my patterns: ids [pid_lo, pid_lo + n_mine)
  int64_t* first;                      // results of my patterns, indexed by id - pid_lo
  int64_t* last;                       // NULL: first receives the count
  // sampled-SA walks (mesh_walk_kernel): my rows [0, n_mine) and their offsets
  const int64_t* rows;
  int64_t* out_offset;
  // owner(row) = (row / block_size) * world / nblocks: shard r holds the rows from shard_start[r] on
  // (first row of block ceil(r * nblocks / world)); shard_start[world] = total_length
  int64_t block_size, nblocks;
  int64_t shard_start[kMeshMaxRanks + 1];
  int block_shift;                     // log2(block_size) when it is a power of two and nblocks <= kMeshOwnerTab, else -1
};

// max_ctas: 0 = fill the device; else an upper bound (several meshes sharing one GPU in a test).
cudaError_t launch_mesh_count(const DevImage& im, const MeshArgs& a, int sm_count, int max_ctas, cudaStream_t stream,
                              int64_t* launch_counter);
cudaError_t launch_mesh_walk(const DevImage& im, const MeshArgs& a, int sm_count, int max_ctas, cudaStream_t stream,
                             int64_t* launch_counter);

}  // namespace fmb
