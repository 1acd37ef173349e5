// Single-pass order-preserving stream compaction building blocks:
// warp-ballot ranks inside a tile + decoupled look-back across tiles.
//
// Tile state word (one uint64 per tile, never cleared between launches):
//   [63:34] epoch of the API call that wrote it  [33:32] status  [31:0] value
// A word whose epoch differs from the current one reads as "not ready", so the state array
// needs no memset per launch (one less node on the latency-critical path).
#pragma once
#include "apc_common.cuh"

#define APC_TILE_THREADS 256
#define APC_TILE_ITEMS 4
#define APC_TILE_POINTS (APC_TILE_THREADS * APC_TILE_ITEMS)  // 1024 points per CTA

enum { APC_ST_INVALID = 0u, APC_ST_AGGREGATE = 1u, APC_ST_PREFIX = 2u };

__device__ __forceinline__ uint64_t scan_pack(uint32_t epoch, uint32_t status, uint32_t value) {
  return ((uint64_t)(epoch & 0x3fffffffu) << 34) | ((uint64_t)status << 32) | value;
}

// Called by warp 0 of the CTA (all 32 lanes).  Publishes this tile's aggregate, walks back
// over predecessor tiles 32 at a time and returns the exclusive prefix (same value in all
// lanes).  Relies on CTAs being dispatched in blockIdx order (as CUB's scan does).
__device__ __forceinline__ uint32_t scan_lookback(uint64_t* __restrict__ state, uint32_t tile,
                                                  uint32_t epoch, uint32_t aggregate) {
  const uint32_t lane = lane_id();
  const uint32_t ep = epoch & 0x3fffffffu;
  if (tile == 0) {
    if (lane == 0) st_volatile_u64(&state[0], scan_pack(ep, APC_ST_PREFIX, aggregate));
    return 0;
  }
  if (lane == 0) st_volatile_u64(&state[tile], scan_pack(ep, APC_ST_AGGREGATE, aggregate));
  uint32_t exclusive = 0;
  int32_t base = (int32_t)tile - 1;
  while (true) {
    const int32_t idx = base - (int32_t)lane;
    uint32_t status, value;
    do {
      if (idx >= 0) {
        const uint64_t w = ld_volatile_u64(&state[idx]);
        status = ((uint32_t)(w >> 34) == ep) ? (uint32_t)((w >> 32) & 3u) : APC_ST_INVALID;
        value = (uint32_t)w;
      } else {  // before tile 0: behaves as a prefix of zero
        status = APC_ST_PREFIX;
        value = 0;
      }
    } while (__any_sync(0xffffffffu, status == APC_ST_INVALID));
    const uint32_t pmask = __ballot_sync(0xffffffffu, status == APC_ST_PREFIX);
    if (pmask) {
      const uint32_t first = __ffs(pmask) - 1;  // nearest predecessor holding a full prefix
      exclusive += warp_sum_u32(lane <= first ? value : 0u);
      break;
    }
    exclusive += warp_sum_u32(value);
    base -= 32;
  }
  if (lane == 0) st_volatile_u64(&state[tile], scan_pack(ep, APC_ST_PREFIX, exclusive + aggregate));
  return exclusive;
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

// CTA-wide decoupled look-back: the whole CTA inspects a window of 256 predecessor tiles at
// once (thread i looks at tile base-i), so a scan of <= 256 tiles - every per-scan launch at
// 262k points - resolves in ONE round of independent loads instead of up to 8 dependent
// 32-wide rounds by a single warp.  Called by all 256 threads; smem: uint32_t[34] (shared with
// tile_ranks: slots 0..15 are reused here after the ranks have been read).
__device__ __forceinline__ uint32_t scan_lookback_cta(uint64_t* __restrict__ state, uint32_t tile, uint32_t epoch,
                                                      uint32_t aggregate, uint32_t* smem34) {
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  const uint32_t ep = epoch & 0x3fffffffu;
  if (tile == 0) {
    if (threadIdx.x == 0) st_volatile_u64(&state[0], scan_pack(ep, APC_ST_PREFIX, aggregate));
    return 0;
  }
  if (threadIdx.x == 0) st_volatile_u64(&state[tile], scan_pack(ep, APC_ST_AGGREGATE, aggregate));
  __syncthreads();  // every thread has finished reading the rank scratch in smem34
  uint32_t exclusive = 0;
  int32_t base = (int32_t)tile - 1;
  while (true) {
    const int32_t idx = base - (int32_t)threadIdx.x;
    uint32_t status, value;
    do {
      if (idx >= 0) {
        const uint64_t w = ld_volatile_u64(&state[idx]);
        status = ((uint32_t)(w >> 34) == ep) ? (uint32_t)((w >> 32) & 3u) : APC_ST_INVALID;
        value = (uint32_t)w;
      } else {
        status = APC_ST_PREFIX;
        value = 0;
      }
    } while (__any_sync(0xffffffffu, status == APC_ST_INVALID));
    // per warp: sum of the values up to and including its nearest full prefix (or all 32)
    const uint32_t pmask = __ballot_sync(0xffffffffu, status == APC_ST_PREFIX);
    const uint32_t first = pmask ? (uint32_t)__ffs(pmask) - 1u : 31u;
    const uint32_t wsum = warp_sum_u32(lane <= first ? value : 0u);
    if (lane == 0) {
      smem34[warp] = wsum;
      smem34[8 + warp] = pmask ? 1u : 0u;
    }
    __syncthreads();
    bool found = false;
#pragma unroll
    for (int w = 0; w < APC_TILE_THREADS / 32; ++w) {  // warps in order of increasing distance
      if (!found) {
        exclusive += smem34[w];
        found = smem34[8 + w] != 0u;
      }
    }
    __syncthreads();
    if (found) break;
    base -= APC_TILE_THREADS;
  }
  if (threadIdx.x == 0) st_volatile_u64(&state[tile], scan_pack(ep, APC_ST_PREFIX, exclusive + aggregate));
  return exclusive;
}

// Full tile step: ranks + look-back.  Returns the global exclusive offset of this tile in
// `tile_base` (all threads) and leaves per-item ranks in rank[].  smem: uint32_t[34].
__device__ __forceinline__ uint32_t tile_compact_offsets(const bool (&keep)[APC_TILE_ITEMS],
                                                         uint32_t (&rank)[APC_TILE_ITEMS],
                                                         uint32_t* smem34, uint64_t* state,
                                                         uint32_t tile, uint32_t epoch,
                                                         uint32_t* total_out, uint32_t n_tiles) {
  const uint32_t total = tile_ranks(keep, rank, smem34);
  const uint32_t excl = scan_lookback_cta(state, tile, epoch, total, smem34);
  if (threadIdx.x == 0 && tile == n_tiles - 1 && total_out) *total_out = excl + total;
  return excl;
}
