// SoA cloud -> packed PointCloud2 byte records (inverse of the unpack).
//
// Replaces set_fields / prepare_pointcloud / copy_fields / create_cloud (pp.py:546-625,
// pp.py:790-812, pp.py:769): the reference copies every attribute device->host separately,
// fills a structured array field by field and copies it again into array('B').  Here one
// kernel assembles the output records in shared memory (zeros where the reference leaves
// zeros) and ships each tile to HBM with a single TMA bulk store, so the host needs one D2H.
#include "apc_load.cuh"

#define RP_MAX_FIELDS 64   // the reference has no limit; 64 fields of a <= 192-byte record (1.5 KB of kernel parameters)

struct RepackField {
  int32_t offset, datatype, source, attr_datatype;
  const void* attr;
};
struct RepackParams {
  RepackField f[RP_MAX_FIELDS];
  uint32_t n_fields, step;
};

__device__ __forceinline__ double attr_as_f64(const void* a, int dt, uint32_t i) {
  switch (dt) {
    case APC_FLOAT32: return (double)reinterpret_cast<const float*>(a)[i];
    case APC_FLOAT64: return reinterpret_cast<const double*>(a)[i];
    case APC_INT8: return (double)reinterpret_cast<const int8_t*>(a)[i];
    case APC_UINT8: return (double)reinterpret_cast<const uint8_t*>(a)[i];
    case APC_INT16: return (double)reinterpret_cast<const int16_t*>(a)[i];
    case APC_UINT16: return (double)reinterpret_cast<const uint16_t*>(a)[i];
    case APC_INT32: return (double)reinterpret_cast<const int32_t*>(a)[i];
    case APC_UINT32: return (double)reinterpret_cast<const uint32_t*>(a)[i];
    default: return 0.0;
  }
}

// numpy ``.astype(dst)`` of a float value, written little-endian byte by byte (records may
// be unaligned).  float -> integer casts truncate toward zero like the C cast numpy uses.
__device__ __forceinline__ void store_as(uint8_t* p, int dt, double v) {
  uint64_t bits = 0;
  int size = 0;
  switch (dt) {
    case APC_FLOAT32: bits = __float_as_uint(__double2float_rn(v)); size = 4; break;
    case APC_FLOAT64: bits = (uint64_t)__double_as_longlong(v); size = 8; break;
    case APC_INT8: bits = (uint64_t)(int64_t)(int8_t)(int32_t)v; size = 1; break;
    case APC_UINT8: bits = (uint64_t)(uint8_t)(int64_t)v; size = 1; break;
    case APC_INT16: bits = (uint64_t)(int64_t)(int16_t)(int32_t)v; size = 2; break;
    case APC_UINT16: bits = (uint64_t)(uint16_t)(int64_t)v; size = 2; break;
    case APC_INT32: bits = (uint64_t)(int64_t)(int32_t)v; size = 4; break;
    case APC_UINT32: bits = (uint64_t)(uint32_t)(int64_t)v; size = 4; break;
    default: break;
  }
  for (int b = 0; b < size; ++b) p[b] = (uint8_t)(bits >> (8 * b));
}

__global__ void __launch_bounds__(APC_TILE_THREADS)
k_repack(const float4* __restrict__ pts, uint32_t n_max, const uint32_t* n_dev, const __grid_constant__ RepackParams prm,
         uint8_t* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t stage[];
  const uint32_t n = apc_count(n_dev, n_max);
  const uint32_t first = blockIdx.x * APC_TILE_POINTS;
  if (first >= n) return;
  const uint32_t in_tile = min((uint32_t)APC_TILE_POINTS, n - first);
  const uint32_t bytes = in_tile * prm.step;
  for (uint32_t b = threadIdx.x * 4; b < ((bytes + 3u) & ~3u); b += APC_TILE_THREADS * 4)
    *reinterpret_cast<uint32_t*>(stage + b) = 0u;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < APC_TILE_ITEMS; ++j) {
    const uint32_t e = j * APC_TILE_THREADS + threadIdx.x;
    if (e < in_tile) {
      const float4 p = pts[first + e];
      uint8_t* rec = stage + (size_t)e * prm.step;
      for (uint32_t f = 0; f < prm.n_fields; ++f) {
        const RepackField& fd = prm.f[f];
        double v;
        switch (fd.source) {
          case 1: v = p.x; break;
          case 2: v = p.y; break;
          case 3: v = p.z; break;
          case 4: v = p.w; break;
          case 5: v = attr_as_f64(fd.attr, fd.attr_datatype, first + e); break;
          default: continue;  // zeros
        }
        store_as(rec + fd.offset, fd.datatype, v);
      }
    }
  }
  // generic-proxy writes must be visible to the async proxy before the bulk store reads them
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  uint8_t* gdst = out + (size_t)first * prm.step;
  const bool aligned = ((((uintptr_t)out) & 15u) == 0);
  const uint32_t bulk = aligned ? (bytes & ~15u) : 0u;
  if (bulk && threadIdx.x == 0) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(stage)),
                 "r"(bulk)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  for (uint32_t b = bulk + threadIdx.x; b < bytes; b += APC_TILE_THREADS) gdst[b] = stage[b];
  if (bulk && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

static int dt_size(int dt) {
  switch (dt) {
    case APC_INT8: case APC_UINT8: return 1;
    case APC_INT16: case APC_UINT16: return 2;
    case APC_INT32: case APC_UINT32: case APC_FLOAT32: return 4;
    case APC_FLOAT64: return 8;
    default: return 0;
  }
}

extern "C" int apc_repack(apc_ctx* ctx, const float* xyzi, uint32_t n_max, const uint32_t* n_dev,
                          const apc_out_field* fields, uint32_t n_fields, uint32_t point_step, uint8_t* out_bytes,
                          void* stream) {
  if (!ctx) return APC_ERR_BAD_ARG;
  if (n_max == 0) return APC_OK;
  APC_REQUIRE(ctx, xyzi && out_bytes && fields, "NULL pointer");
  APC_REQUIRE(ctx, n_fields >= 1 && n_fields <= RP_MAX_FIELDS, "n_fields must be in 1..64");
  APC_REQUIRE(ctx, point_step >= 1 && point_step <= 192, "point_step must be 1..192");
  RepackParams prm;
  memset(&prm, 0, sizeof(prm));
  for (uint32_t f = 0; f < n_fields; ++f) {
    const int sz = dt_size(fields[f].datatype);
    APC_REQUIRE(ctx, sz > 0 && fields[f].offset >= 0 && (uint32_t)(fields[f].offset + sz) <= point_step,
                "output field outside the record");
    APC_REQUIRE(ctx, fields[f].source >= 0 && fields[f].source <= 5, "bad field source");
    APC_REQUIRE(ctx, fields[f].source != 5 || (fields[f].attr_dev && dt_size(fields[f].attr_datatype) > 0),
                "attribute source needs attr_dev and attr_datatype");
    prm.f[f] = RepackField{fields[f].offset, fields[f].datatype, fields[f].source, fields[f].attr_datatype, fields[f].attr_dev};
  }
  prm.n_fields = n_fields;
  prm.step = point_step;
  const uint32_t smem = APC_TILE_POINTS * point_step + 16;
  static bool configured[64] = {};   // the attribute is per device (as in frontend.cu set_smem)
  const int dev = ctx->device >= 0 && ctx->device < 64 ? ctx->device : 0;
  if (smem > 32 * 1024 && !configured[dev]) {
    APC_CUDA(ctx, cudaFuncSetAttribute(k_repack, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured[dev] = true;
  }
  APC_PROF(ctx, "k_repack", (cudaStream_t)stream);
  k_repack<<<apc_div_up(n_max, APC_TILE_POINTS), APC_TILE_THREADS, smem, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(xyzi), n_max, n_dev, prm, out_bytes);
  APC_LAUNCH_CHECK(ctx, "k_repack");
  return APC_OK;
}
