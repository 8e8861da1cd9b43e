// grid.cuh — exact nearest-neighbour search on a uniform cell grid, built and queried on the device.
//
// Replaces the reference's CPU kd-tree (pcl::search::KdTree -> FLANN KDTreeSingleIndex, exact, results
// sorted ascending; call sites fast_gicp_impl.hpp:268,401,616,737) and simple-knn's Morton boxes
// (simple_knn.cu:147-183).  A pointer-chasing tree is the wrong structure for a GPU; a dense cell grid
// with ring expansion gives the same EXACT answer with coalesced cell scans:
//   build : bbox -> cell size (on device, no host sync) -> cell histogram -> exclusive scan -> scatter
//           points into cell order as float4 {x, y, z, bits(original index)}.
//   query : visit the cube of cells of Chebyshev radius r = 0,1,2,... around the query; stop when the
//           k-th best squared distance is below the squared distance to the nearest unsearched face;
//           beyond kMaxRing rings fall back to a linear scan (queries far outside the cloud).
// Determinism: candidates are ordered by (squared distance, original index), so results do not depend on
// the scatter order; distances are computed as ((dx*dx) + dy*dy) + dz*dz in fp32 with no fma contraction,
// bit-identical to oracle/gicp_oracle.cpp.
#pragma once
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <float.h>
#include <mutex>
#include <stdint.h>
#include "host_common.h"

namespace gsicp {

struct GridMeta {  // lives in device memory; written by grid_setup_kernel
  float ox, oy, oz;  // origin (bbox min)
  float cell, inv_cell;
  int nx, ny, nz;
  int ncells;
};

struct GridView {  // passed by value to query kernels
  const GridMeta* meta;
  const uint32_t* cell_start;  // [max_cells + 1]
  const float4* pts;           // [n] in cell order: x, y, z, original index bits
  int n;
};

constexpr int kMaxRing = 6;

__device__ __forceinline__ float dist2_nofma(float ax, float ay, float az, float bx, float by, float bz) {
  const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// ---- build kernels ----------------------------------------------------------------------------
__device__ __forceinline__ unsigned int f2ord(float f) {  // order-preserving float -> uint
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned int u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

static __global__ void grid_bbox_init_kernel(unsigned int* bb) {
  if (threadIdx.x < 3) bb[threadIdx.x] = 0xffffffffu;
  else if (threadIdx.x < 6) bb[threadIdx.x] = 0u;
}

static __global__ void grid_bbox_kernel(int n, const float* __restrict__ xyz, unsigned int* bb) {
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#pragma unroll
    for (int d = 0; d < 3; d++) {
      const float v = xyz[3 * (size_t)i + d];
      if (v == v) {  // ignore NaN coordinates
        mn[d] = fminf(mn[d], v);
        mx[d] = fmaxf(mx[d], v);
      }
    }
  }
#pragma unroll
  for (int d = 0; d < 3; d++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
      mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int d = 0; d < 3; d++) {
      atomicMin(&bb[d], f2ord(mn[d]));
      atomicMax(&bb[3 + d], f2ord(mx[d]));
    }
  }
}

// one thread: choose the cell size so that the dense grid has at most max_cells cells
static __global__ void grid_setup_kernel(const unsigned int* bb, int n, int max_cells, GridMeta* meta) {
  float lo[3], hi[3], ext[3];
  for (int d = 0; d < 3; d++) {
    lo[d] = ord2f(bb[d]);
    hi[d] = ord2f(bb[3 + d]);
    if (!(hi[d] >= lo[d])) { lo[d] = 0.f; hi[d] = 0.f; }
    ext[d] = hi[d] - lo[d];
  }
  const float emax = fmaxf(ext[0], fmaxf(ext[1], ext[2]));
  const double vol = (double)fmaxf(ext[0], emax * 1e-3f) * fmaxf(ext[1], emax * 1e-3f) * fmaxf(ext[2], emax * 1e-3f);
  // aim at ~2 cells per point (surfaces fill few cells; ~5-10 points per occupied cell)
  const double target = fmin((double)max_cells, fmax(64.0, 2.0 * (double)n));
  float cell = (float)cbrt(vol / target);
  if (!(cell > 0.f)) cell = 1.f;
  int nx, ny, nz;
  for (int it = 0; it < 64; it++) {
    nx = (int)fminf(ext[0] / cell, 2.0e6f) + 1;
    ny = (int)fminf(ext[1] / cell, 2.0e6f) + 1;
    nz = (int)fminf(ext[2] / cell, 2.0e6f) + 1;
    if ((double)nx * ny * nz <= (double)max_cells) break;
    cell *= 1.25f;
  }
  meta->ox = lo[0]; meta->oy = lo[1]; meta->oz = lo[2];
  meta->cell = cell;
  meta->inv_cell = 1.0f / cell;
  meta->nx = nx; meta->ny = ny; meta->nz = nz;
  meta->ncells = nx * ny * nz;
}

__device__ __forceinline__ int3 grid_cell_of(const GridMeta& m, float x, float y, float z) {
  int cx = (int)floorf((x - m.ox) * m.inv_cell);
  int cy = (int)floorf((y - m.oy) * m.inv_cell);
  int cz = (int)floorf((z - m.oz) * m.inv_cell);
  // NaN/inf -> clamp (the int conversion of NaN is 0 on the device)
  cx = min(max(cx, 0), m.nx - 1);
  cy = min(max(cy, 0), m.ny - 1);
  cz = min(max(cz, 0), m.nz - 1);
  return make_int3(cx, cy, cz);
}

static __global__ void grid_count_kernel(int n, const float* __restrict__ xyz, const GridMeta* __restrict__ meta,
                                  uint32_t* __restrict__ counts, uint32_t* __restrict__ cell_of_pt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const GridMeta m = *meta;
  const int3 c = grid_cell_of(m, xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2]);
  const uint32_t id = (uint32_t)((c.z * m.ny + c.y) * m.nx + c.x);
  cell_of_pt[i] = id;
  atomicAdd(&counts[id], 1u);
}

static __global__ void grid_scatter_kernel(int n, const float* __restrict__ xyz, const uint32_t* __restrict__ cell_of_pt,
                                    uint32_t* __restrict__ cursor, float4* __restrict__ pts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t slot = atomicAdd(&cursor[cell_of_pt[i]], 1u);
  pts[slot] = make_float4(xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2], __uint_as_float((uint32_t)i));
}

// ---- device-side storage ----------------------------------------------------------------------
struct DeviceGrid {
  Scratch meta_buf, bbox_buf, cell_start, cursor, cell_of_pt, pts, cub_tmp;
  int n = 0, max_cells = 0;

  GridView view() const {
    GridView v;
    v.meta = meta_buf.as<GridMeta>();
    v.cell_start = cell_start.as<uint32_t>();
    v.pts = pts.as<float4>();
    v.n = n;
    return v;
  }

  // d_xyz: device (n,3) fp32.  All work is stream-ordered; no host synchronisation.
  int build(const float* d_xyz, int n_points, cudaStream_t stream) {
    n = n_points;
    if (n <= 0) return GSICP_OK;
    long long want = 2LL * n;
    if (want < 4096) want = 4096;
    if (want > (1LL << 24)) want = (1LL << 24);
    max_cells = (int)want;
    int e;
    if ((e = meta_buf.ensure(sizeof(GridMeta)))) return e;
    if ((e = bbox_buf.ensure(6 * sizeof(unsigned int)))) return e;
    if ((e = cell_start.ensure(((size_t)max_cells + 1) * 4))) return e;
    if ((e = cursor.ensure(((size_t)max_cells + 1) * 4))) return e;
    if ((e = cell_of_pt.ensure((size_t)n * 4))) return e;
    if ((e = pts.ensure((size_t)n * sizeof(float4)))) return e;
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, cursor.as<uint32_t>(), cell_start.as<uint32_t>(), max_cells + 1, stream);
    if ((e = cub_tmp.ensure(tmp))) return e;

    GSICP_LAUNCH(grid_bbox_init_kernel, 1, 32, 0, stream, bbox_buf.as<unsigned int>());
    int blocks = (n + 255) / 256;
    if (blocks > 592) blocks = 592;
    GSICP_LAUNCH(grid_bbox_kernel, blocks, 256, 0, stream, n, d_xyz, bbox_buf.as<unsigned int>());
    GSICP_LAUNCH(grid_setup_kernel, 1, 1, 0, stream, bbox_buf.as<unsigned int>(), n, max_cells, meta_buf.as<GridMeta>());
    GSICP_CUDA(cudaMemsetAsync(cursor.ptr, 0, ((size_t)max_cells + 1) * 4, stream));
    GSICP_LAUNCH(grid_count_kernel, (n + 255) / 256, 256, 0, stream, n, d_xyz, meta_buf.as<GridMeta>(),
                 cursor.as<uint32_t>(), cell_of_pt.as<uint32_t>());
    tmp = cub_tmp.cap;
    GSICP_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp.ptr, tmp, cursor.as<uint32_t>(), cell_start.as<uint32_t>(),
                                             max_cells + 1, stream));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    GSICP_CUDA(cudaMemcpyAsync(cursor.ptr, cell_start.ptr, ((size_t)max_cells + 1) * 4, cudaMemcpyDeviceToDevice, stream));
    GSICP_LAUNCH(grid_scatter_kernel, (n + 255) / 256, 256, 0, stream, n, d_xyz, cell_of_pt.as<uint32_t>(),
                 cursor.as<uint32_t>(), pts.as<float4>());
    GSICP_CUDA(cudaGetLastError());
    return GSICP_OK;
  }
};

// ---- query ------------------------------------------------------------------------------------
// Sorted top-K list ordered by (d2, idx).  K is a compile-time capacity kept in registers.
template <int K>
struct TopK {
  float d2[K];
  uint32_t id[K];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int i = 0; i < K; i++) {
      d2[i] = FLT_MAX;
      id[i] = 0xffffffffu;
    }
  }
  __device__ __forceinline__ void push(float d, uint32_t idx) {
    if (!(d < d2[K - 1] || (d == d2[K - 1] && idx < id[K - 1]))) return;
    float cd = d;
    uint32_t ci = idx;
#pragma unroll
    for (int i = 0; i < K; i++) {
      const bool before = (cd < d2[i]) || (cd == d2[i] && ci < id[i]);
      if (before) {
        const float td = d2[i];
        const uint32_t ti = id[i];
        d2[i] = cd;
        id[i] = ci;
        cd = td;
        ci = ti;
      }
    }
  }
};

// Exact k-nearest search of (qx,qy,qz) in the grid by a GROUP of G consecutive lanes (G = 1, 2, 4, 8 ...).
//   kth     = number of neighbours that must be final (<= K);  exclude = original index to skip (0xffffffff: none)
//   g       = lane's rank inside its group, gmask = shuffle mask of the group's lanes (all must call together)
// The cells of ring r form (2r+1)^2 x-rows; the points of consecutive x cells are contiguous in `pts`, so a shell row is
// ONE contiguous point range (two cell_start loads).  Rows are dealt round-robin to the lanes of the group, each lane keeps
// its own sorted top-K, and the ring loop stops when the group has seen `want` candidates closer than the nearest
// unsearched cell face — the same exact criterion as a single sorted list (a lane's list truncates only when it alone
// holds K >= want such candidates).  Call grid_knn_merge afterwards to obtain the group's sorted top-kth in every lane.
template <int K, int G>
__device__ __forceinline__ void grid_knn(const GridView& g_, float qx, float qy, float qz, int kth, uint32_t exclude,
                                         TopK<K>& best, int g = 0, unsigned gmask = 0xffffffffu) {
  best.init();
  if (g_.n <= 0) return;
  const GridMeta m = *g_.meta;
  const int3 c0 = grid_cell_of(m, qx, qy, qz);
  const int want = min(kth, g_.n - (exclude != 0xffffffffu ? 1 : 0));
  bool finished = (want <= 0);
  auto scan_range = [&](uint32_t b, uint32_t e) {
    for (uint32_t i = b; i < e; i++) {
      const float4 p = g_.pts[i];
      const uint32_t pid = __float_as_uint(p.w);
      if (pid == exclude) continue;
      best.push(dist2_nofma(p.x, p.y, p.z, qx, qy, qz), pid);
    }
  };
  for (int r = 0; r <= kMaxRing && !finished; r++) {
    const int x0 = max(c0.x - r, 0), x1 = min(c0.x + r, m.nx - 1);
    const int y0 = max(c0.y - r, 0), y1 = min(c0.y + r, m.ny - 1);
    const int z0 = max(c0.z - r, 0), z1 = min(c0.z + r, m.nz - 1);
    int row = 0;
    for (int z = z0; z <= z1; z++) {
      for (int y = y0; y <= y1; y++, row++) {
        if (G > 1 && (row % G) != g) continue;
        const int base = (z * m.ny + y) * m.nx;
        if (abs(z - c0.z) == r || abs(y - c0.y) == r) {  // row lies on the shell: cells x0..x1 are one point range
          scan_range(g_.cell_start[base + x0], g_.cell_start[base + x1 + 1]);
        } else {  // interior row: only the two end cells belong to ring r
          if (c0.x - r >= 0) scan_range(g_.cell_start[base + c0.x - r], g_.cell_start[base + c0.x - r + 1]);
          if (c0.x + r <= m.nx - 1) scan_range(g_.cell_start[base + c0.x + r], g_.cell_start[base + c0.x + r + 1]);
        }
      }
    }
    if (x0 == 0 && y0 == 0 && z0 == 0 && x1 == m.nx - 1 && y1 == m.ny - 1 && z1 == m.nz - 1) {
      finished = true;  // the whole grid has been searched
      break;
    }
    // distance below which no unsearched point can exist: nearest face of the searched cube that has cells beyond it
    float bound = FLT_MAX;
    if (c0.x - r > 0) bound = fminf(bound, qx - (m.ox + (c0.x - r) * m.cell));
    if (c0.x + r < m.nx - 1) bound = fminf(bound, (m.ox + (c0.x + r + 1) * m.cell) - qx);
    if (c0.y - r > 0) bound = fminf(bound, qy - (m.oy + (c0.y - r) * m.cell));
    if (c0.y + r < m.ny - 1) bound = fminf(bound, (m.oy + (c0.y + r + 1) * m.cell) - qy);
    if (c0.z - r > 0) bound = fminf(bound, qz - (m.oz + (c0.z - r) * m.cell));
    if (c0.z + r < m.nz - 1) bound = fminf(bound, (m.oz + (c0.z + r + 1) * m.cell) - qz);
    // conservative: shrink by the fp32 error of the face coordinates and of the distances
    const float safe = fmaxf(bound, 0.f) * (1.0f - 1e-5f) - 1e-6f * m.cell;
    if (safe > 0.f) {
      const float lim = safe * safe * (1.0f - 1e-5f);
      int cnt = 0;
#pragma unroll
      for (int i = 0; i < K; i++) cnt += (best.d2[i] < lim) ? 1 : 0;
#pragma unroll
      for (int o = 1; o < G; o <<= 1) cnt += __shfl_xor_sync(gmask, cnt, o);
      if (cnt >= want) finished = true;
    }
  }
  if (!finished) {  // far outside the occupied cells: exact linear scan, interleaved over the group
    best.init();
    for (int i = g; i < g_.n; i += G) {
      const float4 p = g_.pts[i];
      const uint32_t pid = __float_as_uint(p.w);
      if (pid == exclude) continue;
      best.push(dist2_nofma(p.x, p.y, p.z, qx, qy, qz), pid);
    }
  }
}

// Merge the G per-lane lists of a group: afterwards every lane of the group holds the group's sorted top-`kth`
// (entries beyond kth are unspecified).  G = 1: no-op.
template <int K, int G>
__device__ __forceinline__ void grid_knn_merge(TopK<K>& best, int kth, unsigned gmask) {
  if (G == 1) return;
  TopK<K> out;
  out.init();
#pragma unroll
  for (int t = 0; t < K; t++) {
    if (t < kth) {
      float wd = best.d2[0];
      uint32_t wi = best.id[0];
#pragma unroll
      for (int o = 1; o < G; o <<= 1) {
        const float od = __shfl_xor_sync(gmask, wd, o);
        const uint32_t oi = __shfl_xor_sync(gmask, wi, o);
        if (od < wd || (od == wd && oi < wi)) {
          wd = od;
          wi = oi;
        }
      }
      out.d2[t] = wd;
      out.id[t] = wi;
      if (best.id[0] == wi && best.d2[0] == wd && wi != 0xffffffffu) {  // this lane owned the winner: pop it
#pragma unroll
        for (int i = 0; i < K - 1; i++) {
          best.d2[i] = best.d2[i + 1];
          best.id[i] = best.id[i + 1];
        }
        best.d2[K - 1] = FLT_MAX;
        best.id[K - 1] = 0xffffffffu;
      }
    }
  }
  best = out;
}

}  // namespace gsicp
