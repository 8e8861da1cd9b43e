// knn_dist2.cu — distCUDA2: mean squared distance to the 3 nearest other points.
//
// Replaces SimpleKNN::knn (submodules/simple-knn/simple_knn.cu:185-221: bbox reduce with two D2H copies ->
// 30-bit Morton codes -> radix sort -> 1024-point boxes -> per-point scan of every box that cannot be
// rejected).  Here: one uniform-grid build (grid.cuh, no host sync) and one exact 3-NN query per point.
// The result is the mathematically defined quantity the reference computes (self excluded by index,
// coincident points count with distance 0), so parity is to fp32 rounding of the distances.
#include "grid.cuh"

namespace gsicp {

constexpr int kD2Group = 4;  // lanes cooperating on one query

__global__ void __launch_bounds__(128)
dist2_kernel(GridView g, const float* __restrict__ xyz, float* __restrict__ out) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = tid / kD2Group, gl = tid % kD2Group;
  if (i >= g.n) return;  // group-uniform
  const unsigned gmask = ((1u << kD2Group) - 1u) << ((threadIdx.x & 31) & ~(kD2Group - 1));
  TopK<3> best;
  grid_knn<3, kD2Group>(g, xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2], 3, (uint32_t)i, best, gl, gmask);
  grid_knn_merge<3, kD2Group>(best, 3, gmask);
  if (gl == 0) out[i] = (best.d2[0] + best.d2[1] + best.d2[2]) / 3.0f;
}

static DeviceGrid g_dist2_grid;
static std::mutex g_dist2_mu;

}  // namespace gsicp

using namespace gsicp;

extern "C" int gsicp_dist2(int P, const float* d_points, float* d_out, void* stream_v) {
  if (P < 0 || (P > 0 && (!d_points || !d_out))) {
    set_error("gsicp_dist2: bad arguments");
    return GSICP_EINVAL;
  }
  if (P == 0) return GSICP_OK;
  cudaStream_t stream = (cudaStream_t)stream_v;
  std::lock_guard<std::mutex> lock(g_dist2_mu);
  ProfScope ps(kProfDist2, stream);
  if (int e = g_dist2_grid.build(d_points, P, stream)) return e;
  GSICP_LAUNCH(dist2_kernel, (int)(((size_t)P * kD2Group + 127) / 128), 128, 0, stream, g_dist2_grid.view(), d_points, d_out);
  GSICP_CUDA(cudaGetLastError());
  return GSICP_OK;
}
