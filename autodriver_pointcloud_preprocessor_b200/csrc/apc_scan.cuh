// Single-pass order-preserving stream compaction building blocks:
// warp-ballot ranks inside a tile + a two-level cross-tile scan.
//
// Scan state (uint64 words, never cleared between launches):
//   [63:34] epoch of the API call that wrote it  [33:32] status  [31:0] value
// A word whose epoch differs from the current one reads as "not ready", so the state array
// needs no memset per launch (one less node on the latency-critical path).
//   words [0, APC_SCAN_GROUPS)       group aggregates: sum of the totals of 32 consecutive tiles
//   words [APC_SCAN_GROUPS + t]      total of tile t
#pragma once
#include "apc_common.cuh"

#define APC_TILE_THREADS 256
#define APC_TILE_ITEMS 4
#define APC_TILE_POINTS (APC_TILE_THREADS * APC_TILE_ITEMS)  // 1024 points per CTA

enum { APC_ST_INVALID = 0u, APC_ST_VALID = 1u };
#define APC_SCAN_GROUP 32u          // tiles per group
#define APC_SCAN_GROUPS 1024u       // group words reserved at the front of a state array (32M points)

__device__ __forceinline__ uint64_t scan_pack(uint32_t epoch, uint32_t status, uint32_t value) {
  return ((uint64_t)(epoch & 0x3fffffffu) << 34) | ((uint64_t)status << 32) | value;
}
// spins until the word carries the current epoch; returns its value
__device__ __forceinline__ uint32_t scan_wait(const uint64_t* word, uint32_t ep) {
  uint64_t w = ld_volatile_u64(word);
  while ((uint32_t)(w >> 34) != ep) w = ld_volatile_u64(word);
  return (uint32_t)w;
}

// Exclusive prefix of this tile's `total` over all lower-numbered tiles, two levels deep:
//   1. the tile publishes its total;
//   2. warp 0 reads the totals of the earlier tiles of its own 32-tile group (one coalesced load,
//      lane i <- tile i of the group); the group's last tile also publishes the group aggregate;
//   3. warp 0 reads the aggregates of all earlier groups (lane g <- group g).
// Two dependent round trips after the publish, O(1) polled words per tile.  Measured on a
// 256-tile launch (profiles/cta_trace.py): ~4 us, against ~8 us for a 256-wide look-back window
// (65k polling threads queue up on 16 L2 lines) and ~7 us for reading all earlier totals directly
// (n^2/2 polled words): what costs is the number of threads polling the same lines.
// Called by all threads of the CTA; relies on CTAs being dispatched in blockIdx order (lower
// tiles are resident or finished whenever a tile waits).  smem: one uint32_t.
// Step 1 alone (one thread): lets a kernel put independent work between publishing its total and
// waiting for its predecessors', so that the successors are not held up by that work.
__device__ __forceinline__ void scan_publish(uint64_t* __restrict__ state, uint32_t tile, uint32_t epoch, uint32_t total) {
  st_volatile_u64(&state[APC_SCAN_GROUPS + tile], scan_pack(epoch & 0x3fffffffu, APC_ST_VALID, total));
}
template <bool PUBLISHED>
__device__ __forceinline__ uint32_t scan_two_level_impl(uint64_t* __restrict__ state, uint32_t tile, uint32_t n_tiles,
                                                        uint32_t epoch, uint32_t total, uint32_t* smem1) {
  const uint32_t ep = epoch & 0x3fffffffu;
  if (threadIdx.x < 32) {
    const uint32_t lane = threadIdx.x;
    const uint32_t group = tile / APC_SCAN_GROUP, r = tile % APC_SCAN_GROUP;
    uint64_t* totals = state + APC_SCAN_GROUPS;
    if (!PUBLISHED && lane == 0) st_volatile_u64(&totals[tile], scan_pack(ep, APC_ST_VALID, total));
    uint32_t v = 0;
    if (lane < r) v = scan_wait(&totals[group * APC_SCAN_GROUP + lane], ep);
    const uint32_t in_group = warp_sum_u32(v);
    const uint32_t last = min(group * APC_SCAN_GROUP + APC_SCAN_GROUP - 1u, n_tiles - 1u);
    if (tile == last && lane == 0) st_volatile_u64(&state[group], scan_pack(ep, APC_ST_VALID, in_group + total));
    uint32_t before = 0;
    for (uint32_t g0 = 0; g0 < group; g0 += 32) {
      uint32_t u = 0;
      if (g0 + lane < group) u = scan_wait(&state[g0 + lane], ep);
      before += warp_sum_u32(u);
    }
    if (lane == 0) *smem1 = before + in_group;
  }
  __syncthreads();
  return *smem1;
}
__device__ __forceinline__ uint32_t scan_two_level(uint64_t* __restrict__ state, uint32_t tile, uint32_t n_tiles,
                                                   uint32_t epoch, uint32_t total, uint32_t* smem1) {
  return scan_two_level_impl<false>(state, tile, n_tiles, epoch, total, smem1);
}

// Order-preserving ranks for a striped tile: item j of thread t is tile element j*256+t.
// keep[j] in, rank[j] out (exclusive rank inside the tile, in element order); returns the
// tile total.  Needs a __shared__ uint32_t warp_tot[ITEMS*8 + 1] scratch.
__device__ __forceinline__ uint32_t tile_ranks(const bool (&keep)[APC_TILE_ITEMS],
                                               uint32_t (&rank)[APC_TILE_ITEMS], uint32_t* warp_tot) {
  const uint32_t lane = lane_id();
  const uint32_t warp = threadIdx.x >> 5;
  constexpr uint32_t NW = APC_TILE_THREADS / 32;
  uint32_t below[APC_TILE_ITEMS];
#pragma unroll
  for (int j = 0; j < APC_TILE_ITEMS; ++j) {
    const uint32_t b = __ballot_sync(0xffffffffu, keep[j]);
    below[j] = __popc(b & ((1u << lane) - 1u));
    if (lane == 0) warp_tot[j * NW + warp] = __popc(b);
  }
  __syncthreads();
  if (warp == 0) {  // exclusive scan of the ITEMS*NW (=32) warp totals in element order
    const uint32_t v = warp_tot[lane];
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= (uint32_t)o) incl += n;
    }
    warp_tot[lane] = incl - v;
    if (lane == 31) warp_tot[32] = incl;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < APC_TILE_ITEMS; ++j) rank[j] = warp_tot[j * NW + warp] + below[j];
  return warp_tot[32];
}
static_assert(APC_TILE_ITEMS * (APC_TILE_THREADS / 32) == 32, "tile_ranks scans exactly 32 warp totals");

// Full tile step: ranks + cross-tile scan.  Returns the global exclusive offset of this tile in
// `tile_base` (all threads) and leaves per-item ranks in rank[].  smem: uint32_t[34].
__device__ __forceinline__ uint32_t tile_compact_offsets(const bool (&keep)[APC_TILE_ITEMS],
                                                         uint32_t (&rank)[APC_TILE_ITEMS],
                                                         uint32_t* smem34, uint64_t* state,
                                                         uint32_t tile, uint32_t epoch,
                                                         uint32_t* total_out, uint32_t n_tiles) {
  const uint32_t total = tile_ranks(keep, rank, smem34);
  const uint32_t excl = scan_two_level(state, tile, n_tiles, epoch, total, &smem34[33]);
  if (threadIdx.x == 0 && tile == n_tiles - 1 && total_out) *total_out = excl + total;
  return excl;
}
