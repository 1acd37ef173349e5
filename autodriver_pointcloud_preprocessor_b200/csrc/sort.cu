// Sorted-unique rows on the device: the numpy and torch back ends of remove_duplicates.
//
// Reference call sites replaced: utils.py:532-534
//   np.unique(points, axis=0, return_index=True, sorted=False) -> select_by_index(first_index)
// and utils.py:538-542
//   torch.unique(points, dim=0, return_inverse=True, sorted=False) -> select_by_index(inverse)
// Both sort the rows lexicographically (x-major) whatever `sorted` says, so the survivors come
// out in sorted order.  Here: an order-preserving map float32 -> uint32, a stable least-
// significant-digit radix sort of {kx, ky, kz, index} records on the 96-bit key (12 passes of 8
// bits, "onesweep" style: all twelve digit histograms in one upfront pass, then one kernel per
// pass that ranks a tile, chains its digit counts to the earlier tiles by decoupled look-back
// and scatters), and a head-flag compaction over the sorted records.
//
// Ordering rules restated from numpy (pinned by tests/golden/dedup_*.npz and by
// tests/test_oracle_golden.py::test_sort_key_model): a row compares field by field with
// float `<`; -0.0 == +0.0; NaN sorts after +inf and all NaNs tie in the sort, but a row that
// holds a NaN never equals its neighbour (`!=` on NaN), so NaN rows are never merged.  The sort
// is stable, so the representative of a group is its lowest input index (numpy uses a stable
// mergesort when return_index is set).
#include "apc_scan.cuh"

#define SORT_THREADS 256
#define SORT_ITEMS 8
#define SORT_TILE (SORT_THREADS * SORT_ITEMS)   // 2048 records per CTA
#define SORT_PASSES 12
#define SORT_WARPS (SORT_THREADS / 32)
#define SORT_KEY_NAN 0xffffffffu

__device__ __forceinline__ uint32_t sort_key_f32(float f) {
  const uint32_t b = __float_as_uint(f);
  if ((b & 0x7fffffffu) > 0x7f800000u) return SORT_KEY_NAN;   // NaN: after +inf (0xff800000), all NaNs tie
  if ((b & 0x7fffffffu) == 0u) return 0x80000000u;            // -0.0 == +0.0
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// digit of pass p: bytes of z (passes 0-3), then y (4-7), then x (8-11), least significant first
__device__ __forceinline__ uint32_t sort_digit(const uint4& r, uint32_t pass) {
  const uint32_t c = pass < 4 ? r.z : (pass < 8 ? r.y : r.x);
  return (c >> ((pass & 3u) * 8u)) & 255u;
}

// voxel index of one coordinate, biased to [0, 2^21): the same arithmetic as voxel.cu voxel_key
__device__ __forceinline__ bool sort_voxel_coord(float v, float vs, uint32_t& u) {
  if (!(fabsf(v) < 65536.0f)) return false;
  const float q = floorf(__fdiv_rn(v, vs));
  if (!(q >= -1048576.0f && q < 1048576.0f)) return false;
  u = (uint32_t)((int32_t)q + 1048576);
  return true;
}

// ---- keys + the twelve digit histograms in one pass over the points ---------------------------
// VOXEL = false: order-preserving float keys of x, y, z (duplicate removal, numpy / torch back ends)
// VOXEL = true:  biased voxel indices floor(x / voxel_size) (sort-based voxel grid, apc_voxel_downsample_sorted)
template <bool VOXEL>
__global__ void __launch_bounds__(256)
k_sort_keys(const float4* __restrict__ pts, uint32_t n_max, const uint32_t* n_dev, uint4* __restrict__ keys,
            uint32_t* __restrict__ hist, float vs, ApcCtrl* ctrl) {
  __shared__ uint32_t sh[SORT_PASSES * 256];
  for (uint32_t i = threadIdx.x; i < SORT_PASSES * 256; i += blockDim.x) sh[i] = 0u;
  __syncthreads();
  const uint32_t n = apc_count(n_dev, n_max);
  for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t vm = __ballot_sync(0xffffffffu, i < n);
    if (i >= n) continue;
    const float4 p = pts[i];
    uint4 r = make_uint4(sort_key_f32(p.x), sort_key_f32(p.y), sort_key_f32(p.z), i);
    if (VOXEL) {
      uint32_t ux = 0x1fffffu, uy = 0x1fffffu, uz = 0x1fffffu;
      if (!(sort_voxel_coord(p.x, vs, ux) && sort_voxel_coord(p.y, vs, uy) && sort_voxel_coord(p.z, vs, uz)))
        atomicOr(&ctrl->err, APC_DEVERR_KEY_RANGE);
      r = make_uint4(ux, uy, uz, i);
    }
    keys[i] = r;
#pragma unroll
    for (uint32_t pass = 0; pass < SORT_PASSES; ++pass) {
      // neighbouring points share their high bytes: count a warp's equal digits with one atomic
      const uint32_t d = sort_digit(r, pass);
      const uint32_t peers = __match_any_sync(vm, d);
      if ((peers & ((1u << lane_id()) - 1u)) == 0u) atomicAdd(&sh[pass * 256 + d], (uint32_t)__popc(peers));
    }
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < SORT_PASSES * 256; i += blockDim.x)
    if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

// ---- one radix pass ---------------------------------------------------------------------------
// status words: [63:34] tag = epoch*16 + pass, [33:32] 1 = tile count, 2 = inclusive prefix, [31:0] value
__global__ void __launch_bounds__(SORT_THREADS)
k_sort_pass(const uint4* __restrict__ in, uint4* __restrict__ out, uint32_t n_max, const uint32_t* n_dev,
            const uint32_t* __restrict__ hist, uint64_t* __restrict__ status, uint32_t pass, const ApcCtrl* ctrl) {
  __shared__ uint32_t whist[SORT_WARPS][256];
  __shared__ uint32_t s_lstart[256];   // first slot of digit d inside the tile's sorted order
  __shared__ uint32_t s_gbase[256];    // global row of local slot 0 as seen by digit d
  __shared__ uint64_t s_wtot[SORT_WARPS];
  __shared__ __align__(16) uint4 xch[SORT_TILE];
  const uint32_t n = apc_count(n_dev, n_max);
  const uint32_t tile = blockIdx.x;
  const uint32_t first = tile * SORT_TILE;
  if (first >= n) return;
  const uint32_t in_tile = min((uint32_t)SORT_TILE, n - first);
  const uint32_t t = threadIdx.x, lane = t & 31u, w = t >> 5;
  const uint32_t d_total = hist[pass * 256 + t];
  // every record has the same digit: the pass is the identity permutation, just move the tile
  if (__syncthreads_or(d_total == n)) {
    for (uint32_t e = t; e < in_tile; e += SORT_THREADS) out[first + e] = in[first + e];
    return;
  }
  uint4 rec[SORT_ITEMS];
  uint32_t dig[SORT_ITEMS], rank[SORT_ITEMS];
  bool valid[SORT_ITEMS];
#pragma unroll
  for (int j = 0; j < SORT_ITEMS; ++j) {   // warp w owns 256 consecutive records, item j = 32 of them
    const uint32_t e = w * (SORT_ITEMS * 32) + j * 32 + lane;
    valid[j] = e < in_tile;
    rec[j] = valid[j] ? in[first + e] : make_uint4(0u, 0u, 0u, 0u);
    dig[j] = sort_digit(rec[j], pass);
  }
#pragma unroll
  for (int k = 0; k < SORT_WARPS; ++k) whist[k][t] = 0u;
  __syncthreads();
  // stable rank inside the warp's 256 records: per-warp digit counters, one writer per digit per item
#pragma unroll
  for (int j = 0; j < SORT_ITEMS; ++j) {
    const uint32_t vm = __ballot_sync(0xffffffffu, valid[j]);
    if (valid[j]) {
      const uint32_t peers = __match_any_sync(vm, dig[j]);
      const uint32_t leader = __ffs(peers) - 1u;
      uint32_t prev = 0u;
      if (lane == leader) {
        prev = whist[w][dig[j]];
        whist[w][dig[j]] = prev + __popc(peers);
      }
      prev = __shfl_sync(peers, prev, leader);
      rank[j] = prev + __popc(peers & ((1u << lane) - 1u));
    }
    __syncwarp();
  }
  __syncthreads();
  // thread t = digit t: warp counts -> exclusive offsets across the warps; tile count
  uint32_t tile_count = 0u;
#pragma unroll
  for (int k = 0; k < SORT_WARPS; ++k) {
    const uint32_t c = whist[k][t];
    whist[k][t] = tile_count;
    tile_count += c;
  }
  // publish, then look back over the earlier tiles of this digit
  const uint32_t tag = (ctrl->epoch * 16u + pass) & 0x3fffffffu;
  uint64_t* mine = status + (size_t)tile * 256u + t;
  uint32_t before = 0u;
  if (tile == 0) {
    st_volatile_u64(mine, scan_pack(tag, 2u, tile_count));
  } else {
    st_volatile_u64(mine, scan_pack(tag, 1u, tile_count));
    for (uint32_t p = tile; p-- > 0;) {
      const uint64_t* word = status + (size_t)p * 256u + t;
      uint64_t v = ld_volatile_u64(word);
      while ((uint32_t)(v >> 34) != tag) v = ld_volatile_u64(word);
      before += (uint32_t)v;
      if (((uint32_t)(v >> 32) & 3u) == 2u) break;
    }
    st_volatile_u64(mine, scan_pack(tag, 2u, before + tile_count));
  }
  // exclusive scans over the digits: tile counts (low half) and whole-array counts (high half)
  const uint64_t mine64 = ((uint64_t)d_total << 32) | tile_count;
  uint64_t incl = mine64;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint64_t up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= (uint32_t)o) incl += up;
  }
  if (lane == 31) s_wtot[w] = incl;
  __syncthreads();
  uint64_t warp_base = 0;
#pragma unroll
  for (int k = 0; k < SORT_WARPS; ++k)
    if ((uint32_t)k < w) warp_base += s_wtot[k];
  const uint64_t excl = warp_base + incl - mine64;
  const uint32_t lstart = (uint32_t)excl, gdigit = (uint32_t)(excl >> 32);
  s_lstart[t] = lstart;
  s_gbase[t] = gdigit + before - lstart;
  __syncthreads();
  // exchange through shared memory so that the global writes are runs of consecutive rows
#pragma unroll
  for (int j = 0; j < SORT_ITEMS; ++j)
    if (valid[j]) xch[s_lstart[dig[j]] + whist[w][dig[j]] + rank[j]] = rec[j];
  __syncthreads();
#pragma unroll
  for (int j = 0; j < SORT_ITEMS; ++j) {
    const uint32_t p = j * SORT_THREADS + t;
    if (p < in_tile) {
      const uint4 r = xch[p];
      out[s_gbase[sort_digit(r, pass)] + p] = r;
    }
  }
}

// ---- heads of the groups of equal rows -> first index / inverse map ------------------------------
__global__ void __launch_bounds__(APC_TILE_THREADS)
k_unique_heads(const uint4* __restrict__ sorted, uint32_t n_max, const uint32_t* n_dev, uint32_t* __restrict__ first_idx,
               uint32_t* __restrict__ inverse, uint32_t* out_count, uint64_t* scan_state, const ApcCtrl* ctrl,
               uint32_t n_tiles) {
  __shared__ uint32_t sm_scan[34];
  const uint32_t n = apc_count(n_dev, n_max);
  const uint32_t epoch = ctrl->epoch;
  const uint32_t tile = blockIdx.x;
  bool head[APC_TILE_ITEMS];
  uint32_t src[APC_TILE_ITEMS];
#pragma unroll
  for (int j = 0; j < APC_TILE_ITEMS; ++j) {
    const uint32_t i = tile * APC_TILE_POINTS + j * APC_TILE_THREADS + threadIdx.x;
    head[j] = false;
    src[j] = 0u;
    if (i < n) {
      const uint4 c = sorted[i];
      src[j] = c.w;
      if (i == 0) {
        head[j] = true;
      } else {
        const uint4 p = sorted[i - 1];
        const bool has_nan = c.x == SORT_KEY_NAN || c.y == SORT_KEY_NAN || c.z == SORT_KEY_NAN;
        head[j] = has_nan || c.x != p.x || c.y != p.y || c.z != p.z;
      }
    }
  }
  uint32_t rank[APC_TILE_ITEMS];
  const uint32_t base = tile_compact_offsets(head, rank, sm_scan, scan_state, tile, epoch, out_count, n_tiles);
#pragma unroll
  for (int j = 0; j < APC_TILE_ITEMS; ++j) {
    const uint32_t i = tile * APC_TILE_POINTS + j * APC_TILE_THREADS + threadIdx.x;
    if (i < n) {
      const uint32_t row = base + rank[j] + (head[j] ? 1u : 0u) - 1u;   // group of record i
      if (head[j] && first_idx) first_idx[row] = src[j];
      if (inverse) inverse[src[j]] = row;
    }
  }
}

// ---- host side ------------------------------------------------------------------------------------
int apc_sort_prepare(apc_ctx* ctx) {
  if (ctx->sort_a) return APC_OK;
  const size_t M = ctx->max_points;
  const size_t tiles = apc_div_up(ctx->max_points, SORT_TILE);
  APC_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->sort_a), M * sizeof(uint4)));
  APC_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->sort_b), M * sizeof(uint4)));
  APC_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->sort_hist), SORT_PASSES * 256 * sizeof(uint32_t)));
  APC_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->sort_status), tiles * 256 * sizeof(uint64_t)));
  APC_CUDA(ctx, cudaMemset(ctx->sort_status, 0, tiles * 256 * sizeof(uint64_t)));
  APC_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->sort_idx), M * sizeof(uint32_t)));
  return APC_OK;
}

void apc_sort_release(apc_ctx* ctx) {
  void* ptrs[] = {ctx->sort_a, ctx->sort_b, ctx->sort_hist, ctx->sort_status, ctx->sort_idx};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  ctx->sort_a = ctx->sort_b = nullptr;
  ctx->sort_hist = nullptr;
  ctx->sort_status = nullptr;
  ctx->sort_idx = nullptr;
}

// Internal: unique rows without the epoch bump.  Call apc_sort_prepare first (it allocates, which a
// stream capture must not see).
int apc_unique_rows_nobegin(apc_ctx* ctx, const float* xyzi, uint32_t n_max, const uint32_t* n_dev,
                            uint32_t* out_first_idx, uint32_t* out_inverse, uint32_t* out_count_dev, int scan_slot,
                            cudaStream_t s) {
  APC_REQUIRE(ctx, out_count_dev, "out_count_dev is NULL");
  if (n_max == 0) {
    APC_CUDA(ctx, cudaMemsetAsync(out_count_dev, 0, sizeof(uint32_t), s));
    return APC_OK;
  }
  APC_REQUIRE(ctx, xyzi, "NULL pointer");
  APC_REQUIRE(ctx, n_max <= ctx->max_points, "more points than the context was created for");
  APC_REQUIRE(ctx, ctx->sort_a, "sort scratch not prepared");
  const float4* pts = reinterpret_cast<const float4*>(xyzi);
  APC_CUDA(ctx, cudaMemsetAsync(ctx->sort_hist, 0, SORT_PASSES * 256 * sizeof(uint32_t), s));
  {
    APC_PROF(ctx, "k_sort_keys", s);
    const uint32_t blocks = min(apc_div_up(n_max, 256), (uint32_t)APC_SM_COUNT * 4);
    k_sort_keys<false><<<blocks, 256, 0, s>>>(pts, n_max, n_dev, ctx->sort_a, ctx->sort_hist, 0.0f, ctx->ctrl);
    APC_LAUNCH_CHECK(ctx, "k_sort_keys");
  }
  uint4 *a = ctx->sort_a, *b = ctx->sort_b;
  const uint32_t sort_tiles = apc_div_up(n_max, SORT_TILE);
  for (uint32_t pass = 0; pass < SORT_PASSES; ++pass) {
    APC_PROF(ctx, "k_sort_pass", s);
    k_sort_pass<<<sort_tiles, SORT_THREADS, 0, s>>>(a, b, n_max, n_dev, ctx->sort_hist, ctx->sort_status, pass, ctx->ctrl);
    APC_LAUNCH_CHECK(ctx, "k_sort_pass");
    uint4* tmp = a;
    a = b;
    b = tmp;
  }
  const uint32_t n_tiles = apc_div_up(n_max, APC_TILE_POINTS);
  APC_REQUIRE(ctx, n_tiles <= ctx->max_tiles, "more points than the context was created for");
  APC_PROF(ctx, "k_unique_heads", s);
  k_unique_heads<<<n_tiles, APC_TILE_THREADS, 0, s>>>(a, n_max, n_dev, out_first_idx, out_inverse, out_count_dev,
                                                      ctx->scan_state[scan_slot], ctx->ctrl, n_tiles);
  APC_LAUNCH_CHECK(ctx, "k_unique_heads");
  return APC_OK;
}

// ---- sort-based voxel grid: the alternative north_star (3) names, kept for the A/B against the hash ----
// One thread per sorted record; the head of a run (key differs from its predecessor) walks the run -
// stable sort, so in input order - and accumulates the same fixed-point sums as voxel.cu (order
// independent, so the centroids are bit-identical to the hash path's), then writes the voxel at its
// rank among the heads.  Output order: ascending (ix, iy, iz) - the other canonical order SURVEY.md
// section 8(c) allows - not first occurrence.
__device__ __forceinline__ unsigned long long seg_fixed(float v, double scale) {
  return (unsigned long long)__double2ll_rn((double)v * scale);
}
__global__ void __launch_bounds__(APC_TILE_THREADS)
k_voxel_segreduce(const uint4* __restrict__ sorted, const float4* __restrict__ pts, uint32_t n_max, const uint32_t* n_dev,
                  float4* __restrict__ out, uint32_t* __restrict__ out_counts, uint32_t* out_count, uint64_t* scan_state,
                  const ApcCtrl* ctrl, uint32_t n_tiles) {
  __shared__ uint32_t sm_scan[34];
  const uint32_t n = apc_count(n_dev, n_max);
  const uint32_t epoch = ctrl->epoch;
  const uint32_t tile = blockIdx.x;
  bool head[APC_TILE_ITEMS];
  float4 cen[APC_TILE_ITEMS];
  uint32_t cnt[APC_TILE_ITEMS];
#pragma unroll
  for (int j = 0; j < APC_TILE_ITEMS; ++j) {
    const uint32_t i = tile * APC_TILE_POINTS + j * APC_TILE_THREADS + threadIdx.x;
    head[j] = false;
    cnt[j] = 0u;
    if (i < n) {
      const uint4 c = sorted[i];
      if (i == 0) head[j] = true;
      else {
        const uint4 p = sorted[i - 1];
        head[j] = c.x != p.x || c.y != p.y || c.z != p.z;
      }
      if (head[j]) {
        unsigned long long sx = 0, sy = 0, sz = 0, sw = 0;
        uint32_t e = i;
        uint4 r = c;
        do {
          const float4 q = pts[r.w];
          sx += seg_fixed(q.x, 16777216.0); sy += seg_fixed(q.y, 16777216.0); sz += seg_fixed(q.z, 16777216.0);
          sw += seg_fixed(q.w, 1048576.0);
          ++cnt[j];
          if (++e >= n) break;
          r = sorted[e];
        } while (r.x == c.x && r.y == c.y && r.z == c.z);
        const double dc = (double)cnt[j];
        cen[j] = make_float4(
            __double2float_rn(__dmul_rn(__ddiv_rn(__ll2double_rn((long long)sx), dc), 1.0 / 16777216.0)),
            __double2float_rn(__dmul_rn(__ddiv_rn(__ll2double_rn((long long)sy), dc), 1.0 / 16777216.0)),
            __double2float_rn(__dmul_rn(__ddiv_rn(__ll2double_rn((long long)sz), dc), 1.0 / 16777216.0)),
            __double2float_rn(__dmul_rn(__ddiv_rn(__ll2double_rn((long long)sw), dc), 1.0 / 1048576.0)));
      }
    }
  }
  uint32_t rank[APC_TILE_ITEMS];
  const uint32_t base = tile_compact_offsets(head, rank, sm_scan, scan_state, tile, epoch, out_count, n_tiles);
#pragma unroll
  for (int j = 0; j < APC_TILE_ITEMS; ++j)
    if (head[j]) {
      out[base + rank[j]] = cen[j];
      if (out_counts) out_counts[base + rank[j]] = cnt[j];
    }
}

extern "C" int apc_voxel_downsample_sorted(apc_ctx* ctx, const float* xyzi, uint32_t n_max, const uint32_t* n_dev,
                                           float voxel_size, float* out_xyzi, uint32_t* out_voxel_counts,
                                           uint32_t* out_count_dev, void* stream) {
  if (!ctx) return APC_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  APC_REQUIRE(ctx, out_count_dev, "out_count_dev is NULL");
  APC_REQUIRE(ctx, voxel_size > 0.0f, "voxel_size must be > 0");
  APC_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = apc_sort_prepare(ctx);
  if (rc) return rc;
  rc = apc_begin(ctx, s);
  if (rc) return rc;
  if (n_max == 0) {
    APC_CUDA(ctx, cudaMemsetAsync(out_count_dev, 0, sizeof(uint32_t), s));
    return APC_OK;
  }
  APC_REQUIRE(ctx, xyzi && out_xyzi, "NULL pointer");
  APC_REQUIRE(ctx, n_max <= ctx->max_points, "more points than the context was created for");
  const float4* pts = reinterpret_cast<const float4*>(xyzi);
  APC_CUDA(ctx, cudaMemsetAsync(ctx->sort_hist, 0, SORT_PASSES * 256 * sizeof(uint32_t), s));
  {
    APC_PROF(ctx, "k_sort_keys", s);
    const uint32_t blocks = min(apc_div_up(n_max, 256), (uint32_t)APC_SM_COUNT * 4);
    k_sort_keys<true><<<blocks, 256, 0, s>>>(pts, n_max, n_dev, ctx->sort_a, ctx->sort_hist, voxel_size, ctx->ctrl);
    APC_LAUNCH_CHECK(ctx, "k_sort_keys");
  }
  uint4 *a = ctx->sort_a, *b = ctx->sort_b;
  const uint32_t sort_tiles = apc_div_up(n_max, SORT_TILE);
  for (uint32_t pass = 0; pass < SORT_PASSES; ++pass) {
    if ((pass & 3u) == 3u) continue;       // voxel indices are 21 bits: the top byte of every component is zero
    APC_PROF(ctx, "k_sort_pass", s);
    k_sort_pass<<<sort_tiles, SORT_THREADS, 0, s>>>(a, b, n_max, n_dev, ctx->sort_hist, ctx->sort_status, pass, ctx->ctrl);
    APC_LAUNCH_CHECK(ctx, "k_sort_pass");
    uint4* tmp = a;
    a = b;
    b = tmp;
  }
  const uint32_t n_tiles = apc_div_up(n_max, APC_TILE_POINTS);
  APC_REQUIRE(ctx, n_tiles <= ctx->max_tiles, "more points than the context was created for");
  APC_PROF(ctx, "k_voxel_segreduce", s);
  k_voxel_segreduce<<<n_tiles, APC_TILE_THREADS, 0, s>>>(a, pts, n_max, n_dev, reinterpret_cast<float4*>(out_xyzi),
                                                         out_voxel_counts, out_count_dev, ctx->scan_state[1], ctx->ctrl, n_tiles);
  APC_LAUNCH_CHECK(ctx, "k_voxel_segreduce");
  return APC_OK;
}

extern "C" int apc_unique_rows(apc_ctx* ctx, const float* xyzi, uint32_t n_max, const uint32_t* n_dev,
                               uint32_t* out_first_idx, uint32_t* out_inverse, uint32_t* out_count_dev, void* stream) {
  if (!ctx) return APC_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  APC_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = apc_sort_prepare(ctx);
  if (rc) return rc;
  rc = apc_begin(ctx, s);
  if (rc) return rc;
  return apc_unique_rows_nobegin(ctx, xyzi, n_max, n_dev, out_first_idx, out_inverse, out_count_dev, 0, s);
}
