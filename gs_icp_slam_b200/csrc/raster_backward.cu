// raster_backward.cu — backward pass of the B200 Gaussian-splat rasterizer.
//
// Replaces CudaRasterizer::Rasterizer::backward (DGR/cuda_rasterizer/rasterizer_impl.cu:351-454):
//   BACKWARD::render (backward.cu:429-657)  +  computeCov2DCUDA (:144-294)  +  preprocessCUDA (:369-426)
//
// render backward: the reference issues up to 12 global float atomicAdd per contributing
// pixel-Gaussian pair.  Here each warp owns an 8x4 sub-tile, culls the staged batch exactly as the
// forward does, reduces the 12 partial gradients of a surviving instance across its 32 lanes with
// shuffles, adds them to a per-CTA shared-memory accumulator, and the CTA flushes ONE set of 12
// global atomics per (tile, instance) at the end of each batch: global atomics drop from
// 12 * pairs to 12 * R.
//
// per-Gaussian backward: computeCov2D backward and the preprocess backward are one kernel (the
// intermediate dL_dcov3D / dL_dmeans never round-trip through HBM between two launches).
#include <cub/cub.cuh>
#include <mutex>
#include "host_common.h"
#include "raster_common.cuh"

namespace gsicp {

constexpr int kG = 12;  // gradient floats per instance: rgb(3) depth(1) mean2D(2) conic(3) cov_zx cov_yz opacity

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// moments[P][12]: per Gaussian, summed over every pixel it was blended into (one fp32 RED per lane-owned sum):
//   [0..2] sum w dL/dpix_rgb        (= dL/dcolour)             w = alpha * T
//   [3]    sum v                    (= dL/ddepth)              v = w * dL/dpix_depth
//   [4]    sum t                    (= dL/dopacity)            t = G (dL/dalpha + dL/dalpha_d)
//   [5..9] sum u dx, u dy, u dx^2, u dx dy, u dy^2             u = opacity * t
//   [10,11] sum v dx, v dy
// The reference's 12 per-pair gradient expressions (backward.cu:575-654) are linear in these moments with
// per-Gaussian coefficients; gaussian_backward_kernel forms them once per Gaussian.
template <bool kCull>
__global__ void __launch_bounds__(kTilePixels, 4)
render_backward_kernel(const uint32_t* __restrict__ tile_order, const uint2* __restrict__ ranges,
                       const uint32_t* __restrict__ point_list, int W, int H,
                       int tiles_x, const float* __restrict__ bg, const Splat* __restrict__ splats,
                       const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
                       const float* __restrict__ dL_dpix_color, const float* __restrict__ dL_dpix_depth,
                       float* __restrict__ moments, int shard_count, int shard_index) {
  const int tile = (int)tile_order[blockIdx.x];
  if (shard_count > 1 && (tile % shard_count) != shard_index) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile_x = tile % tiles_x, tile_y = tile / tiles_x;
  const int wx0 = tile_x * kTile + (warp & 1) * 8, wy0 = tile_y * kTile + (warp >> 1) * 4;
  const int px = wx0 + (lane & 7), py = wy0 + (lane >> 3);
  const bool inside = px < W && py < H;
  const float pxf = (float)px, pyf = (float)py;
  const int pix = py * W + px;

  // double-buffered staging of 256 instances: one barrier per batch, loads of batch k+1 overlap the math of batch k
  __shared__ float4 sA[2][kTilePixels], sB[2][kTilePixels], sC[2][kTilePixels];
  __shared__ uint32_t sId[2][kTilePixels];

  const uint2 range = ranges[tile];
  const int total = (int)(range.y - range.x);

  const float T_final = inside ? final_T[pix] : 0.f;
  float T = T_final;
  const int last_contributor = inside ? (int)n_contrib[pix] : 0;
  // instances behind the deepest contributor of this warp's pixels are never touched
  int warp_last = last_contributor;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) warp_last = max(warp_last, __shfl_xor_sync(0xffffffffu, warp_last, o));

  float dpr = 0.f, dpg = 0.f, dpb = 0.f, dpd = 0.f;
  if (inside) {
    const size_t HW = (size_t)H * W;
    dpr = dL_dpix_color[0 * HW + pix];
    dpg = dL_dpix_color[1 * HW + pix];
    dpb = dL_dpix_color[2 * HW + pix];
    dpd = dL_dpix_depth[pix];
  }
  const float bg_dot_dpixel = bg[0] * dpr + bg[1] * dpg + bg[2] * dpb;
  const float bg_dot_ddepth = 15.f * dpd;
  const float bg_term = T_final * (bg_dot_dpixel + bg_dot_ddepth);  // background share of dL/dalpha + dL/dalpha_d

  float last_alpha = 0.f, last_r = 0.f, last_g = 0.f, last_b = 0.f, last_depth = 0.f;
  float acc_r = 0.f, acc_g = 0.f, acc_b = 0.f, acc_d = 0.f;
  // lane roles of the recursive-halving reduction (see below)
  const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4, b2 = lane & 2;
  const int red_var = (b16 ? 6 : 0) + (b8 ? 3 : 0) + (b4 ? 2 : (b2 ? 1 : 0));
  const bool red_valid = !(b4 && b2) && !(lane & 1);

  // Back to front: batch `base` covers list positions [total-base-n, total-base), staged reversed
  // (slot k = position total-base-1-k) like backward.cu:519-531.
  auto stage = [&](int base, int buf) {
    const int n = min(kTilePixels, total - base);
    if (tid < n) {
      const uint32_t g = point_list[range.y - 1 - base - tid];
      const Splat* sp = splats + g;
      sId[buf][tid] = g;
      sA[buf][tid] = __ldg(&sp->a);
      sB[buf][tid] = __ldg(&sp->b);
      sC[buf][tid] = __ldg(&sp->c);
    }
  };
  if (total > 0) stage(0, 0);

  for (int base = 0, buf = 0; base < total; base += kTilePixels, buf ^= 1) {
    const int n = min(kTilePixels, total - base);
    __syncthreads();  // batch `base` is staged; every warp has finished reading the other buffer
    if (base + kTilePixels < total) stage(base + kTilePixels, buf ^ 1);

    const int first_pos = total - base;  // 1-based contributor id of slot 0
    if (first_pos - (n - 1) > warp_last) continue;  // the whole batch lies behind this warp's last contributor
    for (int c0 = 0; c0 < n; c0 += 32) {
      if (first_pos - c0 - 31 > warp_last && c0 + 32 <= n) continue;  // whole chunk behind the last contributor
      uint32_t mask;
      {
        const int j = c0 + lane;
        bool hit = (j < n) && (first_pos - j <= warp_last);
        if (kCull) hit = hit && subtile_hit(sA[buf][j < n ? j : 0], sB[buf][j < n ? j : 0], (float)wx0, (float)wy0, 7.f, 3.f);
        mask = __ballot_sync(0xffffffffu, hit);
      }
      while (mask) {
        const int bit = __ffs(mask) - 1;
        mask &= mask - 1;
        const int j = c0 + bit;
        const int contributor = first_pos - j;  // 1-based; reference compares (contributor-1) >= last (backward.cu:540-542)
        const float4 a = sA[buf][j], b = sB[buf][j];
        const float dx = a.x - pxf, dy = a.y - pyf;
        const float power = -0.5f * (a.z * dx * dx + b.x * dy * dy) - a.w * dx * dy;
        const float G = expf(power);
        const float alpha = fminf(0.99f, b.y * G);
        const bool active = inside && (contributor <= last_contributor) && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);

        const unsigned act = __ballot_sync(0xffffffffu, active);
        if (!act) continue;  // no pixel of this sub-tile blended the instance: skip the gradient arithmetic

        float g[kG];
#pragma unroll
        for (int k = 0; k < kG; k++) g[k] = 0.f;
        if (active) {
          const float4 c = sC[buf][j];
          const float inv = __frcp_rn(1.f - alpha);  // correctly rounded reciprocal, shared by the three divisions
          T = T * inv;                               // transmittance in front of this Gaussian (backward.cu:555)
          const float w = alpha * T;                 // d(pixel channel)/d(colour), also d(pixel depth)/d(depth)

          // colour and depth blended behind this Gaussian (backward.cu:563-576, 617-620)
          acc_r = last_alpha * last_r + (1.f - last_alpha) * acc_r;
          acc_g = last_alpha * last_g + (1.f - last_alpha) * acc_g;
          acc_b = last_alpha * last_b + (1.f - last_alpha) * acc_b;
          acc_d = last_alpha * last_depth + (1.f - last_alpha) * acc_d;
          const float czx = b.z, cyz = b.w;
          const float depth = c.w - (czx * a.z + cyz * a.w) * dx - (czx * a.w + cyz * b.x) * dy;
          last_r = c.x; last_g = c.y; last_b = c.z; last_depth = depth;
          last_alpha = alpha;
          // dL/dalpha (colour) + dL/dalpha_d (depth; same alpha and transmittance, T_d == T bit for bit)
          const float dsum = ((c.x - acc_r) * dpr + (c.y - acc_g) * dpg + (c.z - acc_b) * dpb + (depth - acc_d) * dpd) * T -
                             bg_term * inv;
          const float t = G * dsum;
          const float u = b.y * t;
          const float v = w * dpd;
          const float udx = u * dx, udy = u * dy;
          g[0] = w * dpr; g[1] = w * dpg; g[2] = w * dpb;
          g[3] = v;
          g[4] = t;
          g[5] = udx;
          g[6] = udy;
          g[7] = udx * dx;
          g[8] = udx * dy;
          g[9] = udy * dy;
          g[10] = v * dx;
          g[11] = v * dy;
        }
        float* dst = moments + (size_t)sId[buf][j] * kG;
        if (__popc(act) <= 2) {
          // one or two pixels of the sub-tile see this Gaussian (ellipse edge): add them directly
          if (active) {
#pragma unroll
            for (int k = 0; k < kG; k++) atomicAdd(dst + k, g[k]);
          }
        } else {
          // Recursive-halving reduction: at each step a lane keeps half of its running sums and hands the other half to
          // its partner, so the 12 sums over 32 lanes cost 6+3+2+1+1 = 13 shuffles (a butterfly per value: 60).
          // Sum k ends up in the lane with red_var == k, which issues ONE fp32 RED to moments[gaussian][k]:
          // 12 consecutive addresses per instance, fire-and-forget — no shared-memory accumulator (whose float add is
          // a CAS loop on this architecture), no flush, no extra barrier.
          float h[6], q[3];
#pragma unroll
          for (int i = 0; i < 6; i++) {
            const float send = b16 ? g[i] : g[i + 6], keep = b16 ? g[i + 6] : g[i];
            h[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
          }
#pragma unroll
          for (int i = 0; i < 3; i++) {
            const float send = b8 ? h[i] : h[i + 3], keep = b8 ? h[i + 3] : h[i];
            q[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
          }
          const float r0 = (b4 ? q[2] : q[0]) + __shfl_xor_sync(0xffffffffu, b4 ? q[0] : q[2], 4);
          const float r1 = (b4 ? 0.f : q[1]) + __shfl_xor_sync(0xffffffffu, b4 ? q[1] : 0.f, 4);
          float sum = (b2 ? r1 : r0) + __shfl_xor_sync(0xffffffffu, b2 ? r0 : r1, 2);
          sum += __shfl_xor_sync(0xffffffffu, sum, 1);
          if (red_valid) atomicAdd(dst + red_var, sum);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// per-Gaussian backward (backward.cu:144-294 + 298-364 + 369-426 + 20-139 fused)
// ------------------------------------------------------------------------------------------
struct BwdArgs {
  int P, D, M;
  float tan_fovx, tan_fovy, focal_x, focal_y, scale_modifier;
  const float *means, *scales, *rots, *shs, *cov_pre, *view, *proj, *campos;
};

__device__ __forceinline__ F3 dnormv(F3 v, F3 dv) {  // gradient through v/|v| (auxiliary.h:113-124)
  const float sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
  const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
  F3 r;
  r.x = ((+sum2 - v.x * v.x) * dv.x - v.y * v.x * dv.y - v.z * v.x * dv.z) * invsum32;
  r.y = (-v.x * v.y * dv.x + (sum2 - v.y * v.y) * dv.y - v.z * v.y * dv.z) * invsum32;
  r.z = (-v.x * v.z * dv.x - v.y * v.z * dv.y + (sum2 - v.z * v.z) * dv.z) * invsum32;
  return r;
}

__device__ __forceinline__ F3 sh_backward(int idx, const BwdArgs& a, uint8_t clamp_mask, F3 dL_dRGB,
                                          float* __restrict__ dL_dsh_out) {
  const F3 pos = {a.means[3 * idx], a.means[3 * idx + 1], a.means[3 * idx + 2]};
  const F3 cam = {a.campos[0], a.campos[1], a.campos[2]};
  const F3 dir_orig = pos - cam;
  const float len = sqrtf(dot3(dir_orig, dir_orig));
  const F3 dir = {dir_orig.x / len, dir_orig.y / len, dir_orig.z / len};
  const F3* sh = reinterpret_cast<const F3*>(a.shs) + (size_t)idx * a.M;
  F3* dL_dsh = reinterpret_cast<F3*>(dL_dsh_out) + (size_t)idx * a.M;

  dL_dRGB.x *= (clamp_mask & 1) ? 0.f : 1.f;
  dL_dRGB.y *= (clamp_mask & 2) ? 0.f : 1.f;
  dL_dRGB.z *= (clamp_mask & 4) ? 0.f : 1.f;

  F3 dRGBdx = {0, 0, 0}, dRGBdy = {0, 0, 0}, dRGBdz = {0, 0, 0};
  const float x = dir.x, y = dir.y, z = dir.z;
  dL_dsh[0] = kShC0 * dL_dRGB;
  if (a.D > 0) {
    dL_dsh[1] = (-kShC1 * y) * dL_dRGB;
    dL_dsh[2] = (kShC1 * z) * dL_dRGB;
    dL_dsh[3] = (-kShC1 * x) * dL_dRGB;
    dRGBdx = -kShC1 * sh[3];
    dRGBdy = -kShC1 * sh[1];
    dRGBdz = kShC1 * sh[2];
    if (a.D > 1) {
      const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
      dL_dsh[4] = (kShC2[0] * xy) * dL_dRGB;
      dL_dsh[5] = (kShC2[1] * yz) * dL_dRGB;
      dL_dsh[6] = (kShC2[2] * (2.f * zz - xx - yy)) * dL_dRGB;
      dL_dsh[7] = (kShC2[3] * xz) * dL_dRGB;
      dL_dsh[8] = (kShC2[4] * (xx - yy)) * dL_dRGB;
      dRGBdx = dRGBdx + (kShC2[0] * y * sh[4] + kShC2[2] * 2.f * -x * sh[6] + kShC2[3] * z * sh[7] + kShC2[4] * 2.f * x * sh[8]);
      dRGBdy = dRGBdy + (kShC2[0] * x * sh[4] + kShC2[1] * z * sh[5] + kShC2[2] * 2.f * -y * sh[6] + kShC2[4] * 2.f * -y * sh[8]);
      dRGBdz = dRGBdz + (kShC2[1] * y * sh[5] + kShC2[2] * 2.f * 2.f * z * sh[6] + kShC2[3] * x * sh[7]);
      if (a.D > 2) {
        dL_dsh[9] = (kShC3[0] * y * (3.f * xx - yy)) * dL_dRGB;
        dL_dsh[10] = (kShC3[1] * xy * z) * dL_dRGB;
        dL_dsh[11] = (kShC3[2] * y * (4.f * zz - xx - yy)) * dL_dRGB;
        dL_dsh[12] = (kShC3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy)) * dL_dRGB;
        dL_dsh[13] = (kShC3[4] * x * (4.f * zz - xx - yy)) * dL_dRGB;
        dL_dsh[14] = (kShC3[5] * z * (xx - yy)) * dL_dRGB;
        dL_dsh[15] = (kShC3[6] * x * (xx - 3.f * yy)) * dL_dRGB;
        dRGBdx = dRGBdx + (kShC3[0] * sh[9] * 3.f * 2.f * xy + kShC3[1] * sh[10] * yz + kShC3[2] * sh[11] * -2.f * xy +
                           kShC3[3] * sh[12] * -3.f * 2.f * xz + kShC3[4] * sh[13] * (-3.f * xx + 4.f * zz - yy) +
                           kShC3[5] * sh[14] * 2.f * xz + kShC3[6] * sh[15] * 3.f * (xx - yy));
        dRGBdy = dRGBdy + (kShC3[0] * sh[9] * 3.f * (xx - yy) + kShC3[1] * sh[10] * xz +
                           kShC3[2] * sh[11] * (-3.f * yy + 4.f * zz - xx) + kShC3[3] * sh[12] * -3.f * 2.f * yz +
                           kShC3[4] * sh[13] * -2.f * xy + kShC3[5] * sh[14] * -2.f * yz + kShC3[6] * sh[15] * -3.f * 2.f * xy);
        dRGBdz = dRGBdz + (kShC3[1] * sh[10] * xy + kShC3[2] * sh[11] * 4.f * 2.f * yz +
                           kShC3[3] * sh[12] * 3.f * (2.f * zz - xx - yy) + kShC3[4] * sh[13] * 4.f * 2.f * xz +
                           kShC3[5] * sh[14] * (xx - yy));
      }
    }
  }
  const F3 dL_ddir = {dot3(dRGBdx, dL_dRGB), dot3(dRGBdy, dL_dRGB), dot3(dRGBdz, dL_dRGB)};
  return dnormv(dir_orig, dL_ddir);
}

__global__ void __launch_bounds__(256)
gaussian_backward_kernel(BwdArgs a, const int32_t* __restrict__ radii, const uint8_t* __restrict__ clamped,
                         const float* __restrict__ moments, const Splat* __restrict__ splats, int W, int H,
                         float* __restrict__ dL_dmean2D, float* __restrict__ dL_dcolors,
                         float* __restrict__ dL_dopacity, float* __restrict__ dL_dmeans3D,
                         float* __restrict__ dL_dcov3D, float* __restrict__ dL_dsh, float* __restrict__ dL_dscales,
                         float* __restrict__ dL_drots) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.P || !(radii[idx] > 0)) return;

  const float3 mean = make_float3(a.means[3 * idx], a.means[3 * idx + 1], a.means[3 * idx + 2]);
  float cov3[6];
  float3 scale = make_float3(0, 0, 0);
  float4 q = make_float4(0, 0, 0, 1);
  if (a.cov_pre) {
#pragma unroll
    for (int i = 0; i < 6; i++) cov3[i] = a.cov_pre[6 * (size_t)idx + i];
  } else {
    scale = make_float3(a.scales[3 * idx], a.scales[3 * idx + 1], a.scales[3 * idx + 2]);
    q = reinterpret_cast<const float4*>(a.rots)[idx];
    cov3d_from_scale_rot(scale, a.scale_modifier, q, cov3);
  }
  // ---- render gradients from the moments (reference expressions: backward.cu:604-654) ----
  const float4 m0 = reinterpret_cast<const float4*>(moments)[3 * (size_t)idx];
  const float4 m1 = reinterpret_cast<const float4*>(moments)[3 * (size_t)idx + 1];
  const float4 m2 = reinterpret_cast<const float4*>(moments)[3 * (size_t)idx + 2];
  const float v0 = m0.w, tt = m1.x, ux = m1.y, uy = m1.z, uxx = m1.w, uxy = m2.x, uyy = m2.y, vx = m2.z, vy = m2.w;
  const Splat sp = splats[idx];
  const float cA = sp.a.z, cB = sp.a.w, cC = sp.b.x, czx = sp.b.z, cyz = sp.b.w;
  dL_dcolors[3 * (size_t)idx + 0] = m0.x;
  dL_dcolors[3 * (size_t)idx + 1] = m0.y;
  dL_dcolors[3 * (size_t)idx + 2] = m0.z;
  dL_dopacity[idx] = tt;
  const float g2x = (-cA * ux - cB * uy - (czx * cA + cyz * cB) * v0) * (0.5f * W);
  const float g2y = (-cC * uy - cB * ux - (czx * cB + cyz * cC) * v0) * (0.5f * H);
  dL_dmean2D[3 * (size_t)idx + 0] = g2x;
  dL_dmean2D[3 * (size_t)idx + 1] = g2y;
  const float dc_x = -0.5f * uxx - czx * vx;             // dL/dconic_xx
  const float dc_y = -0.5f * uxy - cyz * vx - czx * vy;  // dL/dconic_xy
  const float dc_z = -0.5f * uyy - cyz * vy;             // dL/dconic_yy
  const float dL_dcovzx = -cA * vx - cB * vy;
  const float dL_dcovyz = -cB * vx;  // the reference omits the conic_yy * dy term (backward.cu:616)
  const float dL_ddepth = v0;

  // ---- 2D covariance backward (backward.cu:144-294) ----
  const Ewa e = ewa_project(mean, a.focal_x, a.focal_y, a.tan_fovx, a.tan_fovy, cov3, a.view);
  const float limx = 1.3f * a.tan_fovx, limy = 1.3f * a.tan_fovy;
  const float x_grad_mul = (e.txtz < -limx || e.txtz > limx) ? 0.f : 1.f;
  const float y_grad_mul = (e.tytz < -limy || e.tytz > limy) ? 0.f : 1.f;
  const M3& T = e.T;
  const M3& V = e.Vrk;
  const M3& Wm = e.W;
  const float ca = e.cov.m[0][0] + 0.3f, cb = e.cov.m[0][1], cc = e.cov.m[1][1] + 0.3f;
  const float denom = ca * cc - cb * cb;
  float dL_da = 0, dL_db = 0, dL_dc = 0;
  const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
  float dcov[6];
  if (denom2inv != 0) {
    dL_da = denom2inv * (-cc * cc * dc_x + 2 * cb * cc * dc_y + (denom - ca * cc) * dc_z);
    dL_dc = denom2inv * (-ca * ca * dc_z + 2 * ca * cb * dc_y + (denom - ca * cc) * dc_x);
    dL_db = denom2inv * 2 * (cb * cc * dc_x - (denom + 2 * cb * cb) * dc_y + ca * cb * dc_z);
    dcov[0] = (T.m[0][0] * T.m[0][0] * dL_da + T.m[0][0] * T.m[1][0] * dL_db + T.m[0][0] * T.m[2][0] * dL_dcovzx +
               T.m[1][0] * T.m[1][0] * dL_dc + T.m[1][0] * T.m[2][0] * dL_dcovyz);
    dcov[3] = (T.m[0][1] * T.m[0][1] * dL_da + T.m[0][1] * T.m[1][1] * dL_db + T.m[0][1] * T.m[2][1] * dL_dcovzx +
               T.m[1][1] * T.m[1][1] * dL_dc + T.m[1][1] * T.m[2][1] * dL_dcovyz);
    dcov[5] = (T.m[0][2] * T.m[0][2] * dL_da + T.m[0][2] * T.m[1][2] * dL_db + T.m[0][2] * T.m[2][2] * dL_dcovzx +
               T.m[1][2] * T.m[1][2] * dL_dc + T.m[1][2] * T.m[2][2] * dL_dcovyz);
    dcov[1] = 2 * T.m[0][0] * T.m[0][1] * dL_da + (T.m[0][0] * T.m[1][1] + T.m[0][1] * T.m[1][0]) * dL_db +
              (T.m[0][1] * T.m[2][0] + T.m[0][0] * T.m[2][1]) * dL_dcovzx +
              (T.m[1][1] * T.m[2][0] + T.m[1][0] * T.m[2][1]) * dL_dcovyz + 2 * T.m[1][0] * T.m[1][1] * dL_dc;
    dcov[2] = 2 * T.m[0][0] * T.m[0][2] * dL_da + (T.m[0][0] * T.m[1][2] + T.m[0][2] * T.m[1][0]) * dL_db +
              (T.m[0][2] * T.m[2][0] + T.m[0][0] * T.m[2][2]) * dL_dcovzx +
              (T.m[1][2] * T.m[2][0] + T.m[1][0] * T.m[2][2]) * dL_dcovyz + 2 * T.m[1][0] * T.m[1][2] * dL_dc;
    dcov[4] = 2 * T.m[0][2] * T.m[0][1] * dL_da + (T.m[0][1] * T.m[1][2] + T.m[0][2] * T.m[1][1]) * dL_db +
              (T.m[0][2] * T.m[2][1] + T.m[0][1] * T.m[2][2]) * dL_dcovzx +
              (T.m[1][2] * T.m[2][1] + T.m[1][1] * T.m[2][2]) * dL_dcovyz + 2 * T.m[1][1] * T.m[1][2] * dL_dc;
  } else {
#pragma unroll
    for (int i = 0; i < 6; i++) dcov[i] = 0.f;
  }
#pragma unroll
  for (int i = 0; i < 6; i++) dL_dcov3D[6 * (size_t)idx + i] = dcov[i];

  // gradient w.r.t. the upper 2x3 of T.  Reference quirk reproduced: the cov_yz terms of
  // dL_dT11 / dL_dT12 are dropped by stray ';' (backward.cu:263-268).
  const float dL_dT00 = 2 * (T.m[0][0] * V.m[0][0] + T.m[0][1] * V.m[0][1] + T.m[0][2] * V.m[0][2]) * dL_da +
                        (T.m[1][0] * V.m[0][0] + T.m[1][1] * V.m[0][1] + T.m[1][2] * V.m[0][2]) * dL_db +
                        (T.m[2][0] * V.m[0][0] + T.m[2][1] * V.m[0][1] + T.m[2][2] * V.m[0][2]) * dL_dcovzx;
  const float dL_dT01 = 2 * (T.m[0][0] * V.m[1][0] + T.m[0][1] * V.m[1][1] + T.m[0][2] * V.m[1][2]) * dL_da +
                        (T.m[1][0] * V.m[1][0] + T.m[1][1] * V.m[1][1] + T.m[1][2] * V.m[1][2]) * dL_db +
                        (T.m[2][0] * V.m[1][0] + T.m[2][1] * V.m[1][1] + T.m[2][2] * V.m[1][2]) * dL_dcovzx;
  const float dL_dT02 = 2 * (T.m[0][0] * V.m[2][0] + T.m[0][1] * V.m[2][1] + T.m[0][2] * V.m[2][2]) * dL_da +
                        (T.m[1][0] * V.m[2][0] + T.m[1][1] * V.m[2][1] + T.m[1][2] * V.m[2][2]) * dL_db +
                        (T.m[2][0] * V.m[2][0] + T.m[2][1] * V.m[2][1] + T.m[2][2] * V.m[2][2]) * dL_dcovzx;
  const float dL_dT10 = 2 * (T.m[1][0] * V.m[0][0] + T.m[1][1] * V.m[0][1] + T.m[1][2] * V.m[0][2]) * dL_dc +
                        (T.m[0][0] * V.m[0][0] + T.m[0][1] * V.m[0][1] + T.m[0][2] * V.m[0][2]) * dL_db +
                        (T.m[2][0] * V.m[0][0] + T.m[2][1] * V.m[0][1] + T.m[2][2] * V.m[0][2]) * dL_dcovyz;
  const float dL_dT11 = 2 * (T.m[1][0] * V.m[1][0] + T.m[1][1] * V.m[1][1] + T.m[1][2] * V.m[1][2]) * dL_dc +
                        (T.m[0][0] * V.m[1][0] + T.m[0][1] * V.m[1][1] + T.m[0][2] * V.m[1][2]) * dL_db;
  const float dL_dT12 = 2 * (T.m[1][0] * V.m[2][0] + T.m[1][1] * V.m[2][1] + T.m[1][2] * V.m[2][2]) * dL_dc +
                        (T.m[0][0] * V.m[2][0] + T.m[0][1] * V.m[2][1] + T.m[0][2] * V.m[2][2]) * dL_db;

  const float dL_dJ00 = Wm.m[0][0] * dL_dT00 + Wm.m[0][1] * dL_dT01 + Wm.m[0][2] * dL_dT02;
  const float dL_dJ02 = Wm.m[2][0] * dL_dT00 + Wm.m[2][1] * dL_dT01 + Wm.m[2][2] * dL_dT02;
  const float dL_dJ11 = Wm.m[1][0] * dL_dT10 + Wm.m[1][1] * dL_dT11 + Wm.m[1][2] * dL_dT12;
  const float dL_dJ12 = Wm.m[2][0] * dL_dT10 + Wm.m[2][1] * dL_dT11 + Wm.m[2][2] * dL_dT12;

  const float tz = 1.f / e.t.z, tz2 = tz * tz, tz3 = tz2 * tz;
  const float hx = a.focal_x, hy = a.focal_y;
  const float dL_dtx = x_grad_mul * -hx * tz2 * dL_dJ02;
  const float dL_dty = y_grad_mul * -hy * tz2 * dL_dJ12;
  const float dL_dtz = -hx * tz2 * dL_dJ00 - hy * tz2 * dL_dJ11 + (2 * hx * e.t.x) * tz3 * dL_dJ02 +
                       (2 * hy * e.t.y) * tz3 * dL_dJ12;
  const float* vm = a.view;
  F3 dmean = {vm[0] * dL_dtx + vm[1] * dL_dty + vm[2] * dL_dtz, vm[4] * dL_dtx + vm[5] * dL_dty + vm[6] * dL_dtz,
              vm[8] * dL_dtx + vm[9] * dL_dty + vm[10] * dL_dtz};

  // ---- depth and 2D-mean paths (backward.cu:393-417) ----
  const float* proj = a.proj;
  const float4 mh = xform_point_4x4(mean, proj);
  const float m_w = 1.0f / (mh.w + 0.0000001f);
  const float mul = vm[2] * mean.x + vm[6] * mean.y + vm[10] * mean.z + vm[14];
  dmean.x += dL_ddepth * (vm[2] - vm[3] * mul);
  dmean.y += dL_ddepth * (vm[6] - vm[7] * mul);
  dmean.z += dL_ddepth * (vm[10] - vm[11] * mul);
  const float mul1 = (proj[0] * mean.x + proj[4] * mean.y + proj[8] * mean.z + proj[12]) * m_w * m_w;
  const float mul2 = (proj[1] * mean.x + proj[5] * mean.y + proj[9] * mean.z + proj[13]) * m_w * m_w;
  dmean.x += (proj[0] * m_w - proj[3] * mul1) * g2x + (proj[1] * m_w - proj[3] * mul2) * g2y;
  dmean.y += (proj[4] * m_w - proj[7] * mul1) * g2x + (proj[5] * m_w - proj[7] * mul2) * g2y;
  dmean.z += (proj[8] * m_w - proj[11] * mul1) * g2x + (proj[9] * m_w - proj[11] * mul2) * g2y;

  // ---- SH path ----
  if (a.shs) {
    const F3 dRGB = {m0.x, m0.y, m0.z};
    const F3 dm = sh_backward(idx, a, clamped[idx], dRGB, dL_dsh);
    dmean = dmean + dm;
  }
  dL_dmeans3D[3 * (size_t)idx + 0] = dmean.x;
  dL_dmeans3D[3 * (size_t)idx + 1] = dmean.y;
  dL_dmeans3D[3 * (size_t)idx + 2] = dmean.z;

  // ---- scale / rotation (backward.cu:298-364; no quaternion-normalisation backward) ----
  if (a.scales) {
    const M3 R = quat_to_m3(q.x, q.y, q.z, q.w);
    const float3 s = make_float3(a.scale_modifier * scale.x, a.scale_modifier * scale.y, a.scale_modifier * scale.z);
    M3 S;
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
      for (int r = 0; r < 3; r++) S.m[c][r] = 0.f;
    S.m[0][0] = s.x; S.m[1][1] = s.y; S.m[2][2] = s.z;
    const M3 Mx = m3_mul(S, R);
    M3 dSig;
    dSig.m[0][0] = dcov[0];        dSig.m[0][1] = 0.5f * dcov[1]; dSig.m[0][2] = 0.5f * dcov[2];
    dSig.m[1][0] = 0.5f * dcov[1]; dSig.m[1][1] = dcov[3];        dSig.m[1][2] = 0.5f * dcov[4];
    dSig.m[2][0] = 0.5f * dcov[2]; dSig.m[2][1] = 0.5f * dcov[4]; dSig.m[2][2] = dcov[5];
    M3 M2;
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
      for (int r = 0; r < 3; r++) M2.m[c][r] = 2.0f * Mx.m[c][r];
    const M3 dL_dM = m3_mul(M2, dSig);
    const M3 Rt = m3_transpose(R);
    M3 dMt = m3_transpose(dL_dM);
    dL_dscales[3 * (size_t)idx + 0] = Rt.m[0][0] * dMt.m[0][0] + Rt.m[0][1] * dMt.m[0][1] + Rt.m[0][2] * dMt.m[0][2];
    dL_dscales[3 * (size_t)idx + 1] = Rt.m[1][0] * dMt.m[1][0] + Rt.m[1][1] * dMt.m[1][1] + Rt.m[1][2] * dMt.m[1][2];
    dL_dscales[3 * (size_t)idx + 2] = Rt.m[2][0] * dMt.m[2][0] + Rt.m[2][1] * dMt.m[2][1] + Rt.m[2][2] * dMt.m[2][2];
#pragma unroll
    for (int r = 0; r < 3; r++) {
      dMt.m[0][r] *= s.x;
      dMt.m[1][r] *= s.y;
      dMt.m[2][r] *= s.z;
    }
    const float x = q.x, y = q.y, z = q.z, w = q.w;
    float4 dq;
    dq.x = 2 * y * (dMt.m[1][0] + dMt.m[0][1]) + 2 * z * (dMt.m[2][0] + dMt.m[0][2]) + 2 * w * (dMt.m[1][2] - dMt.m[2][1]) -
           4 * x * (dMt.m[2][2] + dMt.m[1][1]);
    dq.y = 2 * x * (dMt.m[1][0] + dMt.m[0][1]) + 2 * w * (dMt.m[2][0] - dMt.m[0][2]) + 2 * z * (dMt.m[1][2] + dMt.m[2][1]) -
           4 * y * (dMt.m[2][2] + dMt.m[0][0]);
    dq.z = 2 * w * (dMt.m[0][1] - dMt.m[1][0]) + 2 * x * (dMt.m[2][0] + dMt.m[0][2]) + 2 * y * (dMt.m[1][2] + dMt.m[2][1]) -
           4 * z * (dMt.m[1][1] + dMt.m[0][0]);
    dq.w = 2 * z * (dMt.m[0][1] - dMt.m[1][0]) + 2 * y * (dMt.m[2][0] - dMt.m[0][2]) + 2 * x * (dMt.m[1][2] - dMt.m[2][1]);
    reinterpret_cast<float4*>(dL_drots)[idx] = dq;
  }
}

extern int g_render_cull;

// ---- multi-GPU: all-reduce of the render moments of the VISIBLE Gaussians ----
// Every rank preprocesses all Gaussians, so the set {radii > 0} and its index order are identical on all ranks: the
// moments of those V Gaussians are gathered into a dense [V][12] buffer, summed over the ranks by the caller's
// collective (NCCL over NVLink) and scattered back; the per-Gaussian backward then runs replicated.  12*V floats
// travel instead of the 14*P parameter gradients (33k vs 300k Gaussians in config C3).
struct VisibleFlag {
  const int32_t* radii;
  __host__ __device__ uint32_t operator()(int i) const { return radii[i] > 0 ? 1u : 0u; }
};

__global__ void publish_visible_kernel(int P, const uint32_t* __restrict__ excl, const int32_t* __restrict__ radii,
                                       volatile unsigned long long* host_map, unsigned long long seq) {
  host_map[0] = (unsigned long long)(excl[P - 1] + (radii[P - 1] > 0 ? 1u : 0u));
  __threadfence_system();
  host_map[1] = seq;
  __threadfence_system();
}

template <bool kGather>
__global__ void moments_compact_kernel(int P, const int32_t* __restrict__ radii, const uint32_t* __restrict__ excl,
                                       float* __restrict__ moments, float* __restrict__ dense) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P || !(radii[i] > 0)) return;
  float4* m = reinterpret_cast<float4*>(moments) + 3 * (size_t)i;
  float4* d = reinterpret_cast<float4*>(dense) + 3 * (size_t)excl[i];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    if (kGather) d[k] = m[k];
    else m[k] = d[k];
  }
}

struct BwdShared {
  std::mutex mu;
  gsicp_allreduce_f32_fn fn = nullptr;
  void* user = nullptr;
  Scratch excl, dense, cub_tmp;
  unsigned long long* h_map = nullptr;
  unsigned long long* d_map = nullptr;
  unsigned long long seq = 0;
};
static BwdShared g_bwd;

static int allreduce_visible_moments(int P, const int32_t* d_radii, float* moments, cudaStream_t stream) {
  std::lock_guard<std::mutex> lock(g_bwd.mu);
  if (!g_bwd.h_map) {
    GSICP_CUDA(cudaHostAlloc((void**)&g_bwd.h_map, 2 * sizeof(unsigned long long), cudaHostAllocMapped));
    g_bwd.h_map[0] = g_bwd.h_map[1] = 0;
    GSICP_CUDA(cudaHostGetDevicePointer((void**)&g_bwd.d_map, g_bwd.h_map, 0));
  }
  if (int e = g_bwd.excl.ensure((size_t)P * 4)) return e;
  cub::CountingInputIterator<int> counting(0);
  cub::TransformInputIterator<uint32_t, VisibleFlag, cub::CountingInputIterator<int>> flags(counting, VisibleFlag{d_radii});
  size_t tmp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp, flags, g_bwd.excl.as<uint32_t>(), P, stream);
  if (int e = g_bwd.cub_tmp.ensure(tmp)) return e;
  tmp = g_bwd.cub_tmp.cap;
  GSICP_CUDA(cub::DeviceScan::ExclusiveSum(g_bwd.cub_tmp.ptr, tmp, flags, g_bwd.excl.as<uint32_t>(), P, stream));
  const unsigned long long seq = ++g_bwd.seq;
  GSICP_LAUNCH(publish_visible_kernel, 1, 1, 0, stream, P, g_bwd.excl.as<uint32_t>(), d_radii,
               (volatile unsigned long long*)g_bwd.d_map, seq);
  GSICP_CUDA(cudaGetLastError());
  volatile unsigned long long* pm = g_bwd.h_map;
  long spins = 0;
  while (pm[1] != seq) {
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
    if ((++spins & 0xfffff) == 0) {
      const cudaError_t q = cudaStreamQuery(stream);
      if (q != cudaSuccess && q != cudaErrorNotReady) {
        set_error("rasterizer backward failed: %s", cudaGetErrorString(q));
        return GSICP_ECUDA;
      }
    }
  }
  const size_t V = (size_t)pm[0];
  if (V == 0) return GSICP_OK;
  if (int e = g_bwd.dense.ensure(V * kG * sizeof(float))) return e;
  GSICP_LAUNCH(moments_compact_kernel<true>, (P + 255) / 256, 256, 0, stream, P, d_radii, g_bwd.excl.as<uint32_t>(), moments,
               g_bwd.dense.as<float>());
  const int rc = g_bwd.fn(g_bwd.user, g_bwd.dense.as<float>(), V * kG, (void*)stream);
  if (rc != 0) {
    set_error("rasterizer all-reduce callback failed (%d)", rc);
    return GSICP_ECUDA;
  }
  GSICP_LAUNCH(moments_compact_kernel<false>, (P + 255) / 256, 256, 0, stream, P, d_radii, g_bwd.excl.as<uint32_t>(), moments,
               g_bwd.dense.as<float>());
  GSICP_CUDA(cudaGetLastError());
  return GSICP_OK;
}

}  // namespace gsicp

using namespace gsicp;

extern "C" int gsicp_raster_set_allreduce(gsicp_allreduce_f32_fn fn, void* user) {
  std::lock_guard<std::mutex> lock(g_bwd.mu);
  g_bwd.fn = fn;
  g_bwd.user = user;
  return GSICP_OK;
}

extern "C" size_t gsicp_raster_backward_work_bytes(int P) { return (size_t)(P > 0 ? P : 0) * kG * sizeof(float) + 16; }

extern "C" int gsicp_raster_backward(const gsicp_raster_args* args, int num_rendered, const int32_t* d_radii,
                                     const void* d_geom, const void* d_binning, const void* d_image,
                                     const float* d_dL_dout_color, const float* d_dL_dout_depth, float* d_dL_dmeans2D,
                                     float* d_dL_dcolors, float* d_dL_dopacity, float* d_dL_dmeans3D, float* d_dL_dcov3D,
                                     float* d_dL_dsh, float* d_dL_dscales, float* d_dL_drotations, void* d_work,
                                     void* stream_v) {
  if (!args) return GSICP_EINVAL;
  const int P = args->P, W = args->width, H = args->height;
  if (P == 0) return GSICP_OK;
  if (!d_geom || !d_binning || !d_image || !d_work || !d_radii) {
    set_error("gsicp_raster_backward: null state buffer");
    return GSICP_EINVAL;
  }
  if (((uintptr_t)d_work & 15) != 0) {
    set_error("gsicp_raster_backward: d_work must be 16-byte aligned");
    return GSICP_EINVAL;
  }
  cudaStream_t stream = (cudaStream_t)stream_v;
  const int shard_count = args->tile_shard_count > 1 ? args->tile_shard_count : 1;
  const int shard_index = shard_count > 1 ? args->tile_shard_index : 0;
  const int tiles_x = (W + kTile - 1) / kTile, tiles_y = (H + kTile - 1) / kTile, tiles = tiles_x * tiles_y;
  GeomState geom = GeomState::from((char*)d_geom, P);
  BinState bin = BinState::from((char*)d_binning, num_rendered);
  ImgState img = ImgState::from((char*)d_image, (size_t)W * H, tiles);
  float* work = (float*)d_work;

  if (num_rendered > 0) {
    ProfScope ps(kProfRenderBwd, stream);
    if (g_render_cull) {
      GSICP_LAUNCH(render_backward_kernel<true>, tiles, kTilePixels, 0, stream, img.tile_order, img.ranges, bin.point_list, W, H, tiles_x,
                   args->d_background, geom.splats, img.final_T, img.n_contrib, d_dL_dout_color, d_dL_dout_depth, work,
                   shard_count, shard_index);
    } else {
      GSICP_LAUNCH(render_backward_kernel<false>, tiles, kTilePixels, 0, stream, img.tile_order, img.ranges, bin.point_list, W, H, tiles_x,
                   args->d_background, geom.splats, img.final_T, img.n_contrib, d_dL_dout_color, d_dL_dout_depth, work,
                   shard_count, shard_index);
    }
    if (args->debug) GSICP_CUDA(cudaStreamSynchronize(stream));
  }

  if (shard_count > 1 && g_bwd.fn) {
    if (int e = allreduce_visible_moments(P, d_radii, work, stream)) return e;
  }

  BwdArgs ba;
  ba.P = P; ba.D = args->D; ba.M = args->M;
  ba.tan_fovx = args->tan_fovx; ba.tan_fovy = args->tan_fovy;
  ba.focal_y = H / (2.0f * args->tan_fovy);
  ba.focal_x = W / (2.0f * args->tan_fovx);
  ba.scale_modifier = args->scale_modifier;
  ba.means = args->d_means3D; ba.scales = args->d_scales; ba.rots = args->d_rotations;
  ba.shs = args->d_shs; ba.cov_pre = args->d_cov3D_precomp; ba.view = args->d_viewmatrix; ba.proj = args->d_projmatrix;
  ba.campos = args->d_campos;
  ProfScope ps_gb(kProfGaussBwd, stream);
  GSICP_LAUNCH(gaussian_backward_kernel, (P + 255) / 256, 256, 0, stream, ba, d_radii, geom.clamped, work, geom.splats, W, H,
               d_dL_dmeans2D, d_dL_dcolors, d_dL_dopacity, d_dL_dmeans3D, d_dL_dcov3D, d_dL_dsh, d_dL_dscales, d_dL_drotations);
  if (args->debug) GSICP_CUDA(cudaStreamSynchronize(stream));
  GSICP_CUDA(cudaGetLastError());
  return GSICP_OK;
}
