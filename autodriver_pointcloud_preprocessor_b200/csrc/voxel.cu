// Voxel-grid mean downsampling on packed 63-bit voxel keys.
//
// Replaces Open3D t.PointCloud.voxel_down_sample (pp.py:509-512).
//   key       = floor(float32(x) / float32(voxel_size)) per axis, 21 bits each (+2^20 bias)
//   table     = open addressing, linear probing, 64-bit atomicCAS on the key word
//   centroid  = order-independent fixed-point sums (rint(x*2^24), 64-bit integer atomics) so
//               the result is deterministic and bit-identical to oracle/voxel.py
//               centroids_fixed whatever order the atomics land in
//   order     = first-occurrence: a point is "first" when it holds the lowest index of its
//               slot; an order-preserving scan over the first-flags numbers the voxels
// The table is self-cleaning: the thread that finalises a voxel resets its slot, so no
// per-frame memset of the (capacity x 52 B) table sits on the critical path.
#include "apc_scan.cuh"
APC_TRACE_EXPORT(voxel)

#define VOX_EMPTY 0xffffffffffffffffull
#define VOX_NOSLOT 0xffffffffu

__device__ __forceinline__ bool voxel_key(float4 p, float vs, uint64_t& key) {
  const float lim = 65536.0f;
  if (!(fabsf(p.x) < lim && fabsf(p.y) < lim && fabsf(p.z) < lim)) return false;  // also rejects NaN
  const float qx = floorf(__fdiv_rn(p.x, vs)), qy = floorf(__fdiv_rn(p.y, vs)), qz = floorf(__fdiv_rn(p.z, vs));
  const float h = 1048576.0f;
  if (!(qx >= -h && qx < h && qy >= -h && qy < h && qz >= -h && qz < h)) return false;
  const uint64_t ux = (uint64_t)((int64_t)qx + 1048576), uy = (uint64_t)((int64_t)qy + 1048576),
                 uz = (uint64_t)((int64_t)qz + 1048576);
  key = (ux << 42) | (uy << 21) | uz;
  return true;
}

// VOX_ILP = points per thread per round.  Measured on B200 at 262k points: 1 -> 25 us,
// 2 -> 28 us, 4 -> 35 us (more points in flight per thread = fewer resident warps to hide the
// dependent atomic chain), so the kernel runs with 1.
#define VOX_ILP 1
__global__ void __launch_bounds__(256)
k_voxel_insert(const float4* __restrict__ pts, uint32_t n_max, const uint32_t* n_dev, float vs,
               VoxSlot* __restrict__ slots, uint32_t cap_mask, uint32_t* __restrict__ p2slot, ApcCtrl* ctrl) {
  const uint32_t n = apc_count(n_dev, n_max);
  const uint32_t stride = gridDim.x * blockDim.x;
  APC_STAMP(0, 0);
  for (uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += stride * VOX_ILP) {
    float4 p[VOX_ILP];
    uint64_t key[VOX_ILP];
    uint32_t slot[VOX_ILP];
    unsigned long long old[VOX_ILP];
    bool ok[VOX_ILP];
#pragma unroll
    for (int u = 0; u < VOX_ILP; ++u) {
      const uint32_t i = i0 + u * stride;
      ok[u] = i < n;
      if (ok[u]) {
        p[u] = pts[i];
        ok[u] = voxel_key(p[u], vs, key[u]);
        if (!ok[u]) {
          atomicOr(&ctrl->err, APC_DEVERR_KEY_RANGE);
          p2slot[i] = VOX_NOSLOT;
        }
      }
      slot[u] = ok[u] ? ((uint32_t)mix64(key[u]) & cap_mask) : 0u;
    }
#pragma unroll
    for (int u = 0; u < VOX_ILP; ++u)  // independent first probes, all in flight together
      if (ok[u]) old[u] = atomicCAS(&slots[slot[u]].key, VOX_EMPTY, (unsigned long long)key[u]);
#pragma unroll
    for (int u = 0; u < VOX_ILP; ++u) {
      if (!ok[u]) continue;
      const uint32_t i = i0 + u * stride;
      bool found = (old[u] == VOX_EMPTY || old[u] == key[u]);
      for (uint32_t probe = 1; !found && probe <= cap_mask; ++probe) {  // rare: linear probing
        slot[u] = (slot[u] + 1) & cap_mask;
        const unsigned long long o = atomicCAS(&slots[slot[u]].key, VOX_EMPTY, (unsigned long long)key[u]);
        found = (o == VOX_EMPTY || o == key[u]);
      }
      if (!found) {
        atomicOr(&ctrl->err, APC_DEVERR_CAPACITY);
        p2slot[i] = VOX_NOSLOT;
        continue;
      }
      VoxSlot* s = &slots[slot[u]];
      p2slot[i] = slot[u];
      atomicMin(&s->first, i);
      atomicAdd(&s->cnt, 1u);
      // rint(x * 2^24): exact product in float64, round-half-even like numpy.rint
      atomicAdd(&s->acc[0], (unsigned long long)__double2ll_rn((double)p[u].x * 16777216.0));
      atomicAdd(&s->acc[1], (unsigned long long)__double2ll_rn((double)p[u].y * 16777216.0));
      atomicAdd(&s->acc[2], (unsigned long long)__double2ll_rn((double)p[u].z * 16777216.0));
      long long qi = 0;
      if (fabsf(p[u].w) < 1048576.0f) qi = __double2ll_rn((double)p[u].w * 1048576.0);
      else atomicOr(&ctrl->err, APC_DEVERR_KEY_RANGE);
      atomicAdd(&s->acc[3], (unsigned long long)qi);
    }
  }
  APC_STAMP(0, 1);
}

__device__ __forceinline__ float fixed_mean(unsigned long long sum, double cnt, double inv_scale) {
  return __double2float_rn(__dmul_rn(__ddiv_rn(__ll2double_rn((long long)sum), cnt), inv_scale));
}

__global__ void __launch_bounds__(APC_TILE_THREADS)
k_voxel_finalize(uint32_t n_max, const uint32_t* n_dev, const uint32_t* __restrict__ p2slot,
                 VoxSlot* __restrict__ slots, uint32_t* __restrict__ rank_of_slot,
                 float4* __restrict__ out, uint32_t* __restrict__ out_counts, uint32_t* out_count,
                 uint64_t* scan_state, const ApcCtrl* ctrl, uint32_t n_tiles) {
  __shared__ uint32_t sm_scan[34];
  const uint32_t n = apc_count(n_dev, n_max);
  const uint32_t epoch = ctrl->epoch;
  const uint32_t tile = blockIdx.x;
  bool is_first[APC_TILE_ITEMS];
  uint32_t slot[APC_TILE_ITEMS];
  APC_STAMP(1, 0);
#pragma unroll
  for (int j = 0; j < APC_TILE_ITEMS; ++j) {
    const uint32_t i = tile * APC_TILE_POINTS + j * APC_TILE_THREADS + threadIdx.x;
    is_first[j] = false;
    slot[j] = VOX_NOSLOT;
    if (i < n) {
      slot[j] = p2slot[i];
      if (slot[j] != VOX_NOSLOT) is_first[j] = (slots[slot[j]].first == i);
    }
  }
  uint32_t rank[APC_TILE_ITEMS];
  APC_STAMP(1, 1);
  const uint32_t base = tile_compact_offsets(is_first, rank, sm_scan, scan_state, tile, epoch, out_count, n_tiles);
  APC_STAMP(1, 2);
#pragma unroll
  for (int j = 0; j < APC_TILE_ITEMS; ++j) {
    if (is_first[j]) {
      const uint32_t s = slot[j];
      const uint32_t r = base + rank[j];
      // the slot is two 16-byte + one 32-byte aligned pieces of one 64-byte half line
      uint4* raw = reinterpret_cast<uint4*>(&slots[s]);
      const uint4 head = raw[0];                                    // key lo/hi, first, cnt
      const ulonglong2 a01 = *reinterpret_cast<const ulonglong2*>(&raw[1]);
      const ulonglong2 a23 = *reinterpret_cast<const ulonglong2*>(&raw[2]);
      const uint32_t c = head.w;
      const double dc = (double)c;
      out[r] = make_float4(fixed_mean(a01.x, dc, 1.0 / 16777216.0), fixed_mean(a01.y, dc, 1.0 / 16777216.0),
                           fixed_mean(a23.x, dc, 1.0 / 16777216.0), fixed_mean(a23.y, dc, 1.0 / 1048576.0));
      if (out_counts) out_counts[r] = c;
      rank_of_slot[s] = r;
      // self-clean the slot for the next frame: {key = empty, first = max, cnt = 0, acc = 0}
      raw[0] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0u);
      raw[1] = make_uint4(0u, 0u, 0u, 0u);
      raw[2] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  APC_STAMP(1, 3);
}

__global__ void k_voxel_p2v(uint32_t n_max, const uint32_t* n_dev, const uint32_t* __restrict__ p2slot,
                            const uint32_t* __restrict__ rank_of_slot, int32_t* __restrict__ p2v) {
  const uint32_t n = apc_count(n_dev, n_max);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t s = p2slot[i];
    p2v[i] = (s == VOX_NOSLOT) ? -1 : (int32_t)rank_of_slot[s];
  }
}

// Whole-table reset (context creation and error recovery only).
__global__ void k_voxel_reset(VoxSlot* slots, uint32_t cap) {
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < cap; s += gridDim.x * blockDim.x) {
    uint4* raw = reinterpret_cast<uint4*>(&slots[s]);
    raw[0] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0u);
    raw[1] = make_uint4(0u, 0u, 0u, 0u);
    raw[2] = make_uint4(0u, 0u, 0u, 0u);
    raw[3] = make_uint4(0u, 0u, 0u, 0u);
  }
}

int apc_voxel_reset(apc_ctx* ctx, cudaStream_t s) {
  k_voxel_reset<<<APC_SM_COUNT * 4, 256, 0, s>>>(ctx->vox_slots, ctx->hash_cap);
  APC_LAUNCH_CHECK(ctx, "k_voxel_reset");
  return APC_OK;
}

int apc_voxel_nobegin(apc_ctx* ctx, const float* xyzi, uint32_t n_max, const uint32_t* n_dev, float voxel_size,
                      float* out_xyzi, int32_t* out_p2v, uint32_t* out_voxel_counts, uint32_t* out_count_dev,
                      int scan_slot, cudaStream_t s) {
  APC_REQUIRE(ctx, out_count_dev, "out_count_dev is NULL");
  APC_REQUIRE(ctx, voxel_size > 0.0f, "voxel_size must be > 0");
  if (n_max == 0) {
    APC_CUDA(ctx, cudaMemsetAsync(out_count_dev, 0, sizeof(uint32_t), s));
    return APC_OK;
  }
  APC_REQUIRE(ctx, xyzi && out_xyzi, "NULL pointer");
  APC_REQUIRE(ctx, n_max <= ctx->max_points, "more points than the context was created for");
  const uint32_t blocks = min(apc_div_up(n_max, 256), (uint32_t)APC_SM_COUNT * 8);
  {
    APC_PROF(ctx, "k_voxel_insert", s);
    const uint32_t ib = min(apc_div_up(n_max, 256 * VOX_ILP), (uint32_t)APC_SM_COUNT * 8);
    k_voxel_insert<<<ib, 256, 0, s>>>(reinterpret_cast<const float4*>(xyzi), n_max, n_dev, voxel_size, ctx->vox_slots,
                                      ctx->hash_cap - 1, ctx->p2slot, ctx->ctrl);
  }
  APC_LAUNCH_CHECK(ctx, "k_voxel_insert");
  const uint32_t n_tiles = apc_div_up(n_max, APC_TILE_POINTS);
  APC_PROF(ctx, "k_voxel_finalize", s);
  k_voxel_finalize<<<n_tiles, APC_TILE_THREADS, 0, s>>>(n_max, n_dev, ctx->p2slot, ctx->vox_slots, ctx->vox_rank,
                                                        reinterpret_cast<float4*>(out_xyzi), out_voxel_counts,
                                                        out_count_dev, ctx->scan_state[scan_slot], ctx->ctrl, n_tiles);
  APC_LAUNCH_CHECK(ctx, "k_voxel_finalize");
  if (out_p2v) {
    k_voxel_p2v<<<blocks, 256, 0, s>>>(n_max, n_dev, ctx->p2slot, ctx->vox_rank, out_p2v);
    APC_LAUNCH_CHECK(ctx, "k_voxel_p2v");
  }
  return APC_OK;
}

extern "C" int apc_voxel_downsample(apc_ctx* ctx, const float* xyzi, uint32_t n_max, const uint32_t* n_dev,
                                    float voxel_size, float* out_xyzi, int32_t* out_p2v,
                                    uint32_t* out_voxel_counts, uint32_t* out_count_dev, void* stream) {
  if (!ctx) return APC_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  int rc = apc_begin(ctx, s);
  if (rc) return rc;
  return apc_voxel_nobegin(ctx, xyzi, n_max, n_dev, voxel_size, out_xyzi, out_p2v, out_voxel_counts, out_count_dev, 1, s);
}

// Per-attribute voxel mean, Open3D style: float32 sums (atomics), then sum / count in float32.
__global__ void k_attr_zero(uint32_t n_max, const uint32_t* n_vox, float* sum, float* cnt) {
  const uint32_t n = apc_count(n_vox, n_max);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { sum[i] = 0.f; cnt[i] = 0.f; }
}
__global__ void k_attr_add(const float* __restrict__ attr, const int32_t* __restrict__ p2v, uint32_t n_max,
                           const uint32_t* n_dev, float* sum, float* cnt) {
  const uint32_t n = apc_count(n_dev, n_max);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int32_t v = p2v[i];
    if (v >= 0) { atomicAdd(&sum[v], attr[i]); atomicAdd(&cnt[v], 1.0f); }
  }
}
__global__ void k_attr_div(uint32_t n_max, const uint32_t* n_vox, const float* __restrict__ sum,
                           const float* __restrict__ cnt, float* __restrict__ out) {
  const uint32_t n = apc_count(n_vox, n_max);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = __fdiv_rn(sum[i], cnt[i]);
}

extern "C" int apc_voxel_mean_attr(apc_ctx* ctx, const float* attr, const int32_t* p2v, uint32_t n_max,
                                   const uint32_t* n_dev, const uint32_t* n_voxels_dev, float* out_attr, void* stream) {
  if (!ctx) return APC_ERR_BAD_ARG;
  if (n_max == 0) return APC_OK;
  APC_REQUIRE(ctx, attr && p2v && out_attr, "NULL pointer");
  APC_REQUIRE(ctx, n_max <= ctx->max_points, "more points than the context was created for");
  cudaStream_t s = (cudaStream_t)stream;
  float* sum = ctx->knn_avg;                               // scratch reuse: [max_points] floats each
  float* cnt = reinterpret_cast<float*>(ctx->nb_count);
  const uint32_t blocks = min(apc_div_up(n_max, 256), (uint32_t)APC_SM_COUNT * 8);
  k_attr_zero<<<blocks, 256, 0, s>>>(n_max, n_voxels_dev, sum, cnt);
  k_attr_add<<<blocks, 256, 0, s>>>(attr, p2v, n_max, n_dev, sum, cnt);
  k_attr_div<<<blocks, 256, 0, s>>>(n_max, n_voxels_dev, sum, cnt, out_attr);
  APC_LAUNCH_CHECK(ctx, "k_attr_*");
  return APC_OK;
}
