// RANSAC ground-plane segmentation with batched hypothesis scoring.
//
// Replaces Open3D segment_plane (pp.py:533-543), float64 like the legacy
// implementation the tensor API converts to.  Contract shared with oracle/ransac.py:
//   - hypotheses from the counter-based splitmix64 stream (or an explicit sample table),
//     fitted with the "fast plane fit" in a fixed operation order (bit-exact vs the oracle);
//   - one scoring pass scores all hypotheses: each CTA stages a chunk of 16 planes in shared
//     memory and streams a tile of points past them; inlier counts and the integer error
//     term floor(dist^2 * 2^32/thr^2) are order-independent, so warp shuffles + atomics
//     give deterministic totals;
//   - a single-thread epilogue applies Open3D's sequential selection / early-stop rule;
//   - final pass: inlier mask against the winning hypothesis + moment sums for the
//     least-squares refit, reduced in a fixed order.
#include "apc_scan.cuh"

#define RS_CHUNK 16
#define RS_MAX_N 16

__device__ __forceinline__ uint64_t splitmix64_dev(uint64_t x) {
  uint64_t z = x + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__device__ __forceinline__ void plane_from_moments(double cx, double cy, double cz, double xx, double xy, double xz,
                                                   double yy, double yz, double zz, double* out) {
  const double det_x = __dsub_rn(__dmul_rn(yy, zz), __dmul_rn(yz, yz));
  const double det_y = __dsub_rn(__dmul_rn(xx, zz), __dmul_rn(xz, xz));
  const double det_z = __dsub_rn(__dmul_rn(xx, yy), __dmul_rn(xy, xy));
  double nx, ny, nz;
  if (det_x >= det_y && det_x >= det_z) {
    nx = det_x;
    ny = __dsub_rn(__dmul_rn(xz, yz), __dmul_rn(xy, zz));
    nz = __dsub_rn(__dmul_rn(xy, yz), __dmul_rn(xz, yy));
  } else if (det_y >= det_z) {
    nx = __dsub_rn(__dmul_rn(xz, yz), __dmul_rn(xy, zz));
    ny = det_y;
    nz = __dsub_rn(__dmul_rn(xy, xz), __dmul_rn(yz, xx));
  } else {
    nx = __dsub_rn(__dmul_rn(xy, yz), __dmul_rn(xz, yy));
    ny = __dsub_rn(__dmul_rn(xy, xz), __dmul_rn(yz, xx));
    nz = det_z;
  }
  const double norm = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(nx, nx), __dmul_rn(ny, ny)), __dmul_rn(nz, nz)));
  if (!(norm > 0.0)) { out[0] = out[1] = out[2] = out[3] = 0.0; return; }
  nx = __ddiv_rn(nx, norm); ny = __ddiv_rn(ny, norm); nz = __ddiv_rn(nz, norm);
  out[0] = nx; out[1] = ny; out[2] = nz;
  out[3] = -__dadd_rn(__dadd_rn(__dmul_rn(nx, cx), __dmul_rn(ny, cy)), __dmul_rn(nz, cz));
}

__global__ void k_rs_hypotheses(const float4* __restrict__ pts, uint32_t n_max, const uint32_t* n_dev, uint32_t ransac_n,
                                uint32_t iters, uint64_t seed, const int32_t* __restrict__ table,
                                double* __restrict__ planes, unsigned long long* __restrict__ scores) {
  const uint32_t P = apc_count(n_dev, n_max);
  const uint32_t it = blockIdx.x * blockDim.x + threadIdx.x;
  if (it >= iters) return;
  scores[2 * (size_t)it] = 0ull;       // tallies of the scoring pass start from zero
  scores[2 * (size_t)it + 1] = 0ull;
  double* out = planes + 4 * (size_t)it;
  if (P < ransac_n) { out[0] = out[1] = out[2] = out[3] = 0.0; return; }
  uint32_t idx[RS_MAX_N];
  if (table) {
    for (uint32_t j = 0; j < ransac_n; ++j) idx[j] = min((uint32_t)table[(size_t)it * ransac_n + j], P - 1);
  } else {
    uint32_t got = 0;
    for (uint64_t c = 0; got < ransac_n; ++c) {
      const uint64_t z = splitmix64_dev(seed + ((uint64_t)it << 32) + c);
      const uint32_t cand = (uint32_t)(((z >> 32) * (uint64_t)P) >> 32);
      bool dup = false;
      for (uint32_t j = 0; j < got; ++j) dup |= (idx[j] == cand);
      if (!dup) idx[got++] = cand;
    }
  }
  if (ransac_n == 3) {  // plane through three points
    const float4 a = pts[idx[0]], b = pts[idx[1]], c = pts[idx[2]];
    const double e1x = __dsub_rn((double)b.x, (double)a.x), e1y = __dsub_rn((double)b.y, (double)a.y), e1z = __dsub_rn((double)b.z, (double)a.z);
    const double e2x = __dsub_rn((double)c.x, (double)a.x), e2y = __dsub_rn((double)c.y, (double)a.y), e2z = __dsub_rn((double)c.z, (double)a.z);
    double nx = __dsub_rn(__dmul_rn(e1y, e2z), __dmul_rn(e1z, e2y));
    double ny = __dsub_rn(__dmul_rn(e1z, e2x), __dmul_rn(e1x, e2z));
    double nz = __dsub_rn(__dmul_rn(e1x, e2y), __dmul_rn(e1y, e2x));
    const double norm = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(nx, nx), __dmul_rn(ny, ny)), __dmul_rn(nz, nz)));
    if (!(norm > 0.0)) { out[0] = out[1] = out[2] = out[3] = 0.0; return; }
    nx = __ddiv_rn(nx, norm); ny = __ddiv_rn(ny, norm); nz = __ddiv_rn(nz, norm);
    out[0] = nx; out[1] = ny; out[2] = nz;
    out[3] = -__dadd_rn(__dadd_rn(__dmul_rn(nx, (double)a.x), __dmul_rn(ny, (double)a.y)), __dmul_rn(nz, (double)a.z));
    return;
  }
  double cx = 0.0, cy = 0.0, cz = 0.0;
  for (uint32_t j = 0; j < ransac_n; ++j) {
    const float4 p = pts[idx[j]];
    cx = __dadd_rn(cx, (double)p.x); cy = __dadd_rn(cy, (double)p.y); cz = __dadd_rn(cz, (double)p.z);
  }
  const double dn = (double)ransac_n;
  cx = __ddiv_rn(cx, dn); cy = __ddiv_rn(cy, dn); cz = __ddiv_rn(cz, dn);
  double xx = 0, xy = 0, xz = 0, yy = 0, yz = 0, zz = 0;
  for (uint32_t j = 0; j < ransac_n; ++j) {
    const float4 p = pts[idx[j]];
    const double rx = __dsub_rn((double)p.x, cx), ry = __dsub_rn((double)p.y, cy), rz = __dsub_rn((double)p.z, cz);
    xx = __dadd_rn(xx, __dmul_rn(rx, rx)); xy = __dadd_rn(xy, __dmul_rn(rx, ry)); xz = __dadd_rn(xz, __dmul_rn(rx, rz));
    yy = __dadd_rn(yy, __dmul_rn(ry, ry)); yz = __dadd_rn(yz, __dmul_rn(ry, rz)); zz = __dadd_rn(zz, __dmul_rn(rz, rz));
  }
  plane_from_moments(cx, cy, cz, xx, xy, xz, yy, yz, zz, out);
}

__device__ __forceinline__ double plane_dist(const double* pl, double x, double y, double z) {
  return fabs(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(pl[0], x), __dmul_rn(pl[1], y)), __dmul_rn(pl[2], z)), pl[3]));
}

#define RS_SELECT_CHUNK 512
#define CTR_RS_SCORE_TICKET 21  // ctrl->counters slot: retired CTAs of k_rs_score
__device__ void rs_select_cta(const double* __restrict__ planes, const unsigned long long* scores_in, uint32_t P,
                              uint32_t ransac_n, uint32_t iters, double prob, double* __restrict__ plane8,
                              uint32_t* __restrict__ info);

// grid = (persistent CTAs striding over point tiles, hypothesis chunks);
// scores[h] = {inlier count, sum floor(d^2 * scale)}.  Every CTA keeps its chunk of planes in
// shared memory and its per-hypothesis tallies in registers across all the tiles it visits,
// so the plane staging, the warp reduction and the atomics are paid once per CTA, not per tile.
__global__ void __launch_bounds__(APC_TILE_THREADS, 2)
k_rs_score(const float4* __restrict__ pts, uint32_t n_max, const uint32_t* n_dev, const double* __restrict__ planes,
           uint32_t iters, double thr, double scale, unsigned long long* scores, uint32_t ransac_n, double prob,
           double* __restrict__ plane8, uint32_t* __restrict__ info, ApcCtrl* ctrl) {
  __shared__ double s_pl[RS_CHUNK][4];
  __shared__ unsigned long long s_cnt[RS_CHUNK], s_err[RS_CHUNK];
  const uint32_t P = apc_count(n_dev, n_max);
  const uint32_t h0 = blockIdx.y * RS_CHUNK;
  const uint32_t nh = min((uint32_t)RS_CHUNK, iters - h0);
  if (threadIdx.x < RS_CHUNK * 4) {
    const uint32_t h = threadIdx.x >> 2;
    s_pl[h][threadIdx.x & 3] = h < nh ? planes[4 * (size_t)(h0 + h) + (threadIdx.x & 3)] : 0.0;
  }
  if (threadIdx.x < RS_CHUNK) { s_cnt[threadIdx.x] = 0; s_err[threadIdx.x] = 0; }
  __syncthreads();
  uint32_t cnt[RS_CHUNK];
  unsigned long long err[RS_CHUNK];
#pragma unroll
  for (int h = 0; h < RS_CHUNK; ++h) { cnt[h] = 0; err[h] = 0; }
  for (uint32_t first = blockIdx.x * APC_TILE_POINTS; first < P; first += gridDim.x * APC_TILE_POINTS) {
    float4 p[APC_TILE_ITEMS];
#pragma unroll
    for (int j = 0; j < APC_TILE_ITEMS; ++j) {   // all loads of the tile in flight before the math
      const uint32_t i = first + j * APC_TILE_THREADS + threadIdx.x;
      p[j] = i < P ? pts[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < APC_TILE_ITEMS; ++j) {
      const uint32_t i = first + j * APC_TILE_THREADS + threadIdx.x;
      if (i < P) {
        const double x = p[j].x, y = p[j].y, z = p[j].z;
#pragma unroll
        for (int h = 0; h < RS_CHUNK; ++h) {
          const double d = plane_dist(s_pl[h], x, y, z);
          if (d < thr) {
            cnt[h] += 1u;
            // d < thr => d^2 * 2^32/thr^2 < 2^32 (saturating at 2^32-1 in the 1-ulp corner):
            // a single native F2I.U32.F64 instead of the emulated 64-bit conversion
            err[h] += (unsigned long long)__double2uint_rd(__dmul_rn(__dmul_rn(d, d), scale));
          }
        }
      }
    }
  }
  // warp totals -> CTA totals in shared memory -> one pair of global atomics per hypothesis
#pragma unroll
  for (int h = 0; h < RS_CHUNK; ++h) {
    const uint32_t c = __reduce_add_sync(0xffffffffu, cnt[h]);
    unsigned long long e = err[h];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if (lane_id() == 0 && c) {
      atomicAdd(&s_cnt[h], (unsigned long long)c);
      atomicAdd(&s_err[h], e);
    }
  }
  __syncthreads();
  if (threadIdx.x < nh && s_cnt[threadIdx.x]) {
    atomicAdd(&scores[2 * (size_t)(h0 + threadIdx.x)], s_cnt[threadIdx.x]);
    atomicAdd(&scores[2 * (size_t)(h0 + threadIdx.x) + 1], s_err[threadIdx.x]);
  }
  // the last CTA to retire (ticket) runs the sequential selection: one launch and one
  // inter-kernel dependency less on the per-scan critical path
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&ctrl->counters[CTR_RS_SCORE_TICKET], 1u) == gridDim.x * gridDim.y - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  rs_select_cta(planes, scores, P, ransac_n, iters, prob, plane8, info);
}

// Open3D's sequential selection + early-stop rule over the batched scores, run by one CTA.
// info = {best iteration | 0xffffffff, n_inliers (filled later), 0, 0}; plane8[4..7] = winner.
__device__ void rs_select_cta(const double* __restrict__ planes, const unsigned long long* scores_in, uint32_t P,
                              uint32_t ransac_n, uint32_t iters, double prob, double* __restrict__ plane8,
                              uint32_t* __restrict__ info) {
  const volatile unsigned long long* scores = scores_in;  // written by other CTAs' atomics: read through L2
  // the scan is inherently sequential, its loads are not: the CTA stages the scores (and a
  // validity flag per hypothesis) in shared memory, then thread 0 walks them
  __shared__ unsigned long long s_inl[RS_SELECT_CHUNK], s_err[RS_SELECT_CHUNK];
  __shared__ uint8_t s_valid[RS_SELECT_CHUNK];
  __shared__ unsigned long long sb_inl, sb_err;
  __shared__ uint32_t sb_it;
  __shared__ double sb_break;
  const double log1mp = prob < 1.0 ? log(1.0 - prob) : -__longlong_as_double(0x7ff0000000000000ll);
  if (threadIdx.x == 0) { sb_inl = 0; sb_err = 0; sb_it = 0xffffffffu; sb_break = (double)iters; }
  for (uint32_t c0 = 0; c0 < iters; c0 += RS_SELECT_CHUNK) {
    __syncthreads();
    for (uint32_t t = threadIdx.x; t < RS_SELECT_CHUNK && c0 + t < iters; t += blockDim.x) {
      const double* pl = planes + 4 * (size_t)(c0 + t);
      s_valid[t] = !(pl[0] == 0.0 && pl[1] == 0.0 && pl[2] == 0.0 && pl[3] == 0.0);
      s_inl[t] = scores[2 * (size_t)(c0 + t)];
      s_err[t] = scores[2 * (size_t)(c0 + t) + 1];
    }
    __syncthreads();
    if (threadIdx.x != 0) continue;
    unsigned long long best_inl = sb_inl, best_err = sb_err;
    uint32_t best_it = sb_it;
    double break_it = sb_break;
    const uint32_t cend = min(iters, c0 + RS_SELECT_CHUNK);
    for (uint32_t it = c0; it < cend; ++it) {
      if ((double)it > break_it) continue;
      if (!s_valid[it - c0]) continue;
      const unsigned long long inl = s_inl[it - c0], err = s_err[it - c0];
      if (inl > best_inl || (inl == best_inl && inl > 0 && err < best_err)) {
        best_inl = inl; best_err = err; best_it = it;
        if (inl >= P) {
          break_it = 0.0;
        } else {
          const double fitness = __ddiv_rn((double)inl, (double)P);
          double fn = fitness;
          for (uint32_t j = 1; j < ransac_n; ++j) fn = __dmul_rn(fn, fitness);
          const double denom = log(__dsub_rn(1.0, fn));
          const double cand = (denom == 0.0) ? __longlong_as_double(0x7ff0000000000000ll) : __ddiv_rn(log1mp, denom);
          break_it = cand < (double)iters ? cand : (double)iters;
        }
      }
    }
    sb_inl = best_inl; sb_err = best_err; sb_it = best_it; sb_break = break_it;
  }
  __syncthreads();
  if (threadIdx.x == 0) { info[0] = sb_it; info[1] = 0; info[2] = 0; info[3] = 0; }
  if (threadIdx.x < 4) plane8[4 + threadIdx.x] = sb_it == 0xffffffffu ? 0.0 : planes[4 * (size_t)sb_it + threadIdx.x];
}

#define CTR_RS_TICKET 20  // ctrl->counters slot: CTAs of k_rs_final that have published their partials

// Final inliers against the winning hypothesis + moment sums for the least-squares refit.
// Each CTA (1024 points) writes one partial; the last CTA to finish (ticket counter) sums the
// partials in index order - a fixed reduction order, so the refit is deterministic - and
// solves for the plane.
__global__ void __launch_bounds__(APC_TILE_THREADS)
k_rs_final(const float4* __restrict__ pts, uint32_t n_max, const uint32_t* n_dev, double* __restrict__ plane8,
           uint32_t* __restrict__ info, double thr, uint8_t* __restrict__ mask, double* __restrict__ partials,
           ApcCtrl* ctrl) {
  __shared__ double s_red[8][10];
  __shared__ bool s_last;
  const uint32_t P = apc_count(n_dev, n_max);
  const bool have = info[0] != 0xffffffffu;
  const double pl[4] = {plane8[4], plane8[5], plane8[6], plane8[7]};
  double acc[10];
#pragma unroll
  for (int k = 0; k < 10; ++k) acc[k] = 0.0;
#pragma unroll
  for (int j = 0; j < APC_TILE_ITEMS; ++j) {
    const uint32_t i = blockIdx.x * APC_TILE_POINTS + j * APC_TILE_THREADS + threadIdx.x;
    if (i < P) {
      const float4 p = pts[i];
      const double x = p.x, y = p.y, z = p.z;
      const bool inl = have && plane_dist(pl, x, y, z) < thr;
      mask[i] = inl ? 1 : 0;
      if (inl) {
        acc[0] += 1.0; acc[1] += x; acc[2] += y; acc[3] += z;
        acc[4] += x * x; acc[5] += x * y; acc[6] += x * z; acc[7] += y * y; acc[8] += y * z; acc[9] += z * z;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 10; ++k) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    if (lane_id() == 0) s_red[threadIdx.x >> 5][k] = acc[k];
  }
  __syncthreads();
  if (threadIdx.x < 10) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += s_red[w][threadIdx.x];
    partials[(size_t)blockIdx.x * 10 + threadIdx.x] = s;
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&ctrl->counters[CTR_RS_TICKET], 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // last CTA: stage 256 partial rows at a time in shared memory (parallel, coalesced loads),
  // then thread k < 10 adds moment k in CTA order
  __shared__ double s_part[256 * 10];
  __shared__ double s_tot[10];
  double run = 0.0;
  const volatile double* vp = partials;
  for (uint32_t b0 = 0; b0 < gridDim.x; b0 += 256) {
    const uint32_t rows = min(256u, gridDim.x - b0);
    __syncthreads();
    for (uint32_t e = threadIdx.x; e < rows * 10; e += blockDim.x) s_part[e] = vp[(size_t)b0 * 10 + e];
    __syncthreads();
    if (threadIdx.x < 10)
      for (uint32_t b = 0; b < rows; ++b) run += s_part[b * 10 + threadIdx.x];
  }
  if (threadIdx.x < 10) s_tot[threadIdx.x] = run;
  __syncthreads();
  if (threadIdx.x == 0) {
    const double n = s_tot[0];
    info[1] = (uint32_t)n;
    if (n < 1.0) { plane8[0] = plane8[1] = plane8[2] = plane8[3] = 0.0; return; }
    const double cx = s_tot[1] / n, cy = s_tot[2] / n, cz = s_tot[3] / n;
    const double xx = s_tot[4] - n * cx * cx, xy = s_tot[5] - n * cx * cy, xz = s_tot[6] - n * cx * cz;
    const double yy = s_tot[7] - n * cy * cy, yz = s_tot[8] - n * cy * cz, zz = s_tot[9] - n * cz * cz;
    plane_from_moments(cx, cy, cz, xx, xy, xz, yy, yz, zz, plane8);
  }
}

int apc_segment_plane_nobegin(apc_ctx* ctx, const float* xyzi, uint32_t n_max, const uint32_t* n_dev, double thr,
                              int ransac_n, int iters, double prob, uint64_t seed, const int32_t* table,
                              double* out_plane, uint8_t* out_mask, uint32_t* out_info, cudaStream_t s) {
  APC_REQUIRE(ctx, out_plane && out_mask && out_info, "NULL output pointer");
  APC_REQUIRE(ctx, prob > 0.0 && prob <= 1.0, "probability must be in (0, 1]");
  APC_REQUIRE(ctx, ransac_n >= 3 && ransac_n <= RS_MAX_N, "ransac_n must be in 3..16");
  APC_REQUIRE(ctx, iters >= 1 && (uint32_t)iters <= ctx->rs_max_iters, "num_iterations must be in 1..4096");
  APC_REQUIRE(ctx, thr > 0.0, "distance_threshold must be > 0");
  APC_REQUIRE(ctx, n_max <= ctx->max_points, "more points than the context was created for");
  if (!n_dev && n_max < (uint32_t)ransac_n) return apc_set_error(ctx, APC_ERR_TOO_FEW, "fewer points than ransac_n");
  if (n_max == 0) {
    APC_CUDA(ctx, cudaMemsetAsync(out_plane, 0, 8 * sizeof(double), s));
    APC_CUDA(ctx, cudaMemsetAsync(out_info, 0xff, sizeof(uint32_t), s));
    APC_CUDA(ctx, cudaMemsetAsync(out_info + 1, 0, 3 * sizeof(uint32_t), s));
    return APC_OK;
  }
  APC_REQUIRE(ctx, xyzi, "NULL pointer");
  const float4* pts = reinterpret_cast<const float4*>(xyzi);
  const double scale = 4294967296.0 / (thr * thr);
  {
    APC_PROF(ctx, "k_rs_hypotheses", s);
    k_rs_hypotheses<<<apc_div_up(iters, 32), 32, 0, s>>>(pts, n_max, n_dev, ransac_n, iters, seed, table, ctx->rs_planes,
                                                         ctx->rs_scores);
  }
  // ~2 resident CTAs per SM in total (register-limited), each striding over point tiles
  const uint32_t n_chunks = apc_div_up(iters, RS_CHUNK);
  const uint32_t gx = min(apc_div_up(n_max, APC_TILE_POINTS), max(1u, (uint32_t)(APC_SM_COUNT * 2) / n_chunks));
  const dim3 grid(gx, n_chunks);
  {
    APC_PROF(ctx, "k_rs_score", s);
    k_rs_score<<<grid, APC_TILE_THREADS, 0, s>>>(pts, n_max, n_dev, ctx->rs_planes, iters, thr, scale, ctx->rs_scores,
                                                 ransac_n, prob, out_plane, out_info, ctx->ctrl);
  }
  APC_PROF(ctx, "k_rs_final", s);
  k_rs_final<<<apc_div_up(n_max, APC_TILE_POINTS), APC_TILE_THREADS, 0, s>>>(pts, n_max, n_dev, out_plane, out_info, thr,
                                                                             out_mask, ctx->rs_partials, ctx->ctrl);
  APC_LAUNCH_CHECK(ctx, "segment_plane");
  return APC_OK;
}

extern "C" int apc_segment_plane(apc_ctx* ctx, const float* xyzi, uint32_t n_max, const uint32_t* n_dev,
                                 double distance_threshold, int ransac_n, int num_iterations, double probability,
                                 uint64_t seed, const int32_t* sample_table_dev, double* out_plane_dev,
                                 uint8_t* out_inlier_mask, uint32_t* out_info_dev, void* stream) {
  if (!ctx) return APC_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  int rc = apc_begin(ctx, s);
  if (rc) return rc;
  return apc_segment_plane_nobegin(ctx, xyzi, n_max, n_dev, distance_threshold, ransac_n, num_iterations, probability,
                                   seed, sample_table_dev, out_plane_dev, out_inlier_mask, out_info_dev, s);
}
