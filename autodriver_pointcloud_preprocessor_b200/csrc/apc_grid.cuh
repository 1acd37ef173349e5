// Neighbour-grid table shared by the kernels that build it (neighbors.cu, and the voxel finalize
// kernel when the pipeline inserts the centroids as it writes them) and the kernels that query it.
#pragma once
#include "apc_common.cuh"

#define GRID_EMPTY 0xffffffffffffffffull
#define GRID_CTR_CELLS 24     // ctrl->counters slot: occupied cells appended to GridDev::cells (zeroed by k_begin)
#define GRID_NOSLOT 0xffffffffu
// One cell of the open-addressing table: key, population and the start of its run in the sorted
// array share one 16-byte slot, so a query resolves a cell with ONE 128-bit load (the key probe and
// the {start, fill} read were two dependent L2 round trips when they lived in separate arrays:
// 22 % + 14 % of k_radius_query's stall samples, profiles/r1d_ncu_full.csv / hot_sass.py).
struct __align__(16) GridSlot {
  unsigned long long key;    // packed (level, ix, iy, iz); all ones = empty
  uint32_t fill;             // points in the cell
  uint32_t start;            // first sorted position of the cell
};
struct GridDev {
  GridSlot* slots;           // [cap]
  uint32_t cap_mask;
  uint32_t levels;
  uint32_t* slot;            // [levels][n_max] slot of point i at level l
  uint32_t* rank;            // [levels][n_max] arrival rank of point i inside its cell
  float4* sorted;            // [levels][n_max] xyz + original index (bits in w)
  float* cell;               // [levels] cell sizes (device)
  float cell0;               // > 0: single-level grid whose cell size the host knows (no k_grid_cells launch)
  float inv0;                // > 0: cell index = floor(x * inv0) instead of floor(x / cell0) (see grid_coord_g)
  uint32_t* cells;           // != NULL: grid_insert_items appends the slot of every newly occupied cell (the point
                             // that takes arrival rank 0) to this list, counted in ctrl->counters[GRID_CTR_CELLS]:
                             // the run assignment then walks ~11 k cells instead of 205 k points (k_grid_assign_cells)
  uint32_t cursor_base;      // first ctrl->counters slot of this grid's per-level scatter cursors (the radius
                             // grid and the KNN grid can both be built inside one pipeline run)
};
__device__ __forceinline__ float grid_cell_size(const GridDev& g, uint32_t level) {
  return g.cell0 > 0.0f ? g.cell0 : g.cell[level];
}

__device__ __forceinline__ bool grid_coord(float x, float y, float z, float c, int32_t& ix, int32_t& iy, int32_t& iz) {
  const float qx = floorf(__fdiv_rn(x, c)), qy = floorf(__fdiv_rn(y, c)), qz = floorf(__fdiv_rn(z, c));
  const float h = 262143.0f;  // one cell of margin for the +-1 neighbour offsets
  if (!(qx >= -h && qx < h && qy >= -h && qy < h && qz >= -h && qz < h)) return false;
  ix = (int32_t)qx; iy = (int32_t)qy; iz = (int32_t)qz;
  return true;
}
// Cell of a point in grid g.  The index only has to be the SAME function at insert and at query time and
// to respect the walk's slack; for the radius grid (cells of 2r, box-pruned walk with an explicit slack of
// 8 float32 ulps of the coordinate, neighbors.cu) the three IEEE divisions by the cell size - a quarter of
// the instructions of the kernel that inserts the centroids - are one multiplication each by the
// host-rounded reciprocal.  Every other grid (27-cell walks whose guarantee is the 2^-10 margin on the
// cell size) keeps the division.
__device__ __forceinline__ bool grid_coord_g(const GridDev& g, float c, float x, float y, float z, int32_t& ix, int32_t& iy,
                                             int32_t& iz) {
  if (!(g.inv0 > 0.0f)) return grid_coord(x, y, z, c, ix, iy, iz);
  const float qx = floorf(__fmul_rn(x, g.inv0)), qy = floorf(__fmul_rn(y, g.inv0)), qz = floorf(__fmul_rn(z, g.inv0));
  const float h = 262143.0f;
  if (!(qx >= -h && qx < h && qy >= -h && qy < h && qz >= -h && qz < h)) return false;
  ix = (int32_t)qx; iy = (int32_t)qy; iz = (int32_t)qz;
  return true;
}
__device__ __forceinline__ uint64_t grid_key(uint32_t level, int32_t ix, int32_t iy, int32_t iz) {
  return ((uint64_t)level << 57) | ((uint64_t)(uint32_t)(ix + 262144) << 38) |
         ((uint64_t)(uint32_t)(iy + 262144) << 19) | (uint64_t)(uint32_t)(iz + 262144);
}
// Inserts one point into the table: claims / finds the slot of its cell (linear probing on the key
// word) and takes the next arrival rank in the cell.  slot = GRID_NOSLOT on error (raised on ctrl).
__device__ __forceinline__ void grid_insert_point(const GridDev& g, uint32_t level, float x, float y, float z, float c,
                                                  ApcCtrl* ctrl, uint32_t& slot, uint32_t& rank) {
  int32_t ix, iy, iz;
  slot = GRID_NOSLOT;
  rank = 0;
  if (!grid_coord_g(g, c, x, y, z, ix, iy, iz)) {
    atomicOr(&ctrl->err, APC_DEVERR_KEY_RANGE);
    return;
  }
  const uint64_t key = grid_key(level, ix, iy, iz);
  uint32_t s = (uint32_t)mix64(key) & g.cap_mask;
  for (uint32_t probe = 0; probe <= g.cap_mask; ++probe) {
    const unsigned long long old = atomicCAS(&g.slots[s].key, GRID_EMPTY, (unsigned long long)key);
    if (old == GRID_EMPTY || old == key) { slot = s; break; }
    s = (s + 1) & g.cap_mask;
  }
  if (slot == GRID_NOSLOT) atomicOr(&ctrl->err, APC_DEVERR_CAPACITY);
  else rank = atomicAdd(&g.slots[slot].fill, 1u);
}

// The same for ITEMS points of one thread in lockstep: all the key CAS are issued before any result
// is looked at, then all the rank atomics, so a thread pays two atomic round trips, not 2 * ITEMS.
template <int ITEMS>
__device__ __forceinline__ void grid_insert_items(const GridDev& g, uint32_t level, const bool (&act)[ITEMS],
                                                  const float4 (&p)[ITEMS], float c, ApcCtrl* ctrl,
                                                  uint32_t (&slot)[ITEMS], uint32_t (&rank)[ITEMS]) {
  uint64_t key[ITEMS];
  uint32_t s[ITEMS];
  bool go[ITEMS];
  unsigned long long old[ITEMS];
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    int32_t ix = 0, iy = 0, iz = 0;
    go[j] = act[j] && grid_coord_g(g, c, p[j].x, p[j].y, p[j].z, ix, iy, iz);
    if (act[j] && !go[j]) atomicOr(&ctrl->err, APC_DEVERR_KEY_RANGE);
    key[j] = grid_key(level, ix, iy, iz);
    s[j] = (uint32_t)mix64(key[j]) & g.cap_mask;
    slot[j] = GRID_NOSLOT;
    rank[j] = 0;
  }
#pragma unroll
  for (int j = 0; j < ITEMS; ++j)
    old[j] = go[j] ? atomicCAS(&g.slots[s[j]].key, GRID_EMPTY, (unsigned long long)key[j]) : 0ull;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    if (!go[j]) continue;
    for (uint32_t probe = 1; old[j] != GRID_EMPTY && old[j] != key[j] && probe <= g.cap_mask; ++probe) {   // rare
      s[j] = (s[j] + 1) & g.cap_mask;
      old[j] = atomicCAS(&g.slots[s[j]].key, GRID_EMPTY, (unsigned long long)key[j]);
    }
    if (old[j] == GRID_EMPTY || old[j] == key[j]) slot[j] = s[j];
    else atomicOr(&ctrl->err, APC_DEVERR_CAPACITY);
  }
#pragma unroll
  for (int j = 0; j < ITEMS; ++j)
    if (slot[j] != GRID_NOSLOT) rank[j] = atomicAdd(&g.slots[slot[j]].fill, 1u);
  if (g.cells) {
#pragma unroll
    for (int j = 0; j < ITEMS; ++j)
      if (slot[j] != GRID_NOSLOT && rank[j] == 0u) g.cells[atomicAdd(&ctrl->counters[GRID_CTR_CELLS], 1u)] = slot[j];
  }
}
