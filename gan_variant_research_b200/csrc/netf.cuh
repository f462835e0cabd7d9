// PatchSampleF(use_mlp=True) as a module of its own (the north-star signature):
// backward-side helper.  The incoming gradient is w.r.t. the NORMALISED head output, fp32 rows (B*P, N) in the
// caller's row order; this kernel applies the normalise backward (F.normalize, eps = 1e-6) and writes d loss / d Y as
// the bf16 hi(+lo) row blob, sorted-slot order, that the head's backward GEMMs (GM_DH, k_wgrad_tc) consume -- the
// same blob the fused loss kernel produces in head mode.  Padding rows (slot >= P) are written as zeros: the weight
// gradient sums over every row of a 128-row tile.
//   warp <-> (layer, image, sorted slot); lane <-> 8 output columns (N <= 256)
#pragma once
#include "common.cuh"
#include "gemm_tc.cuh"

namespace pnce {

struct NetfPackJob {
  const float* g;            // (B*P, N) incoming gradient
  const float* y;            // (B*P, N) normalised output of the forward
  const float* inv;          // (B*P) norm bookkeeping of the forward (pnce_sample_fwd convention)
  const int* perm;           // sorted slot -> original index
  __nv_bfloat16 *hi, *lo;    // dY blob [tile][N/8][16][8][8]
  int P, Ppad, N;
};
struct NetfPackLaunch {
  NetfPackJob job[PNCE_MAX_LAYERS];
  long long start[PNCE_MAX_LAYERS + 1];   // CTA prefix (8 warps per CTA)
  int n, B;
};

__global__ void __launch_bounds__(kThreads) k_netf_dy_pack(const __grid_constant__ NetfPackLaunch a) {
  pdl_enter();
  int l = 0;
  for (int i = 1; i < a.n; ++i)
    if ((long long)blockIdx.x >= a.start[i]) l = i;
  const NetfPackJob& J = a.job[l];
  const int lane = threadIdx.x & 31;
  const long long w = ((long long)blockIdx.x - a.start[l]) * 8 + (threadIdx.x >> 5);
  if (w >= (long long)a.B * J.Ppad) return;
  const int b = (int)(w / J.Ppad), j = (int)(w - (long long)b * J.Ppad);
  const int N = J.N, N8 = N >> 3;
  float d[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) d[k] = 0.f;
  if (j < J.P) {
    const size_t row = (size_t)b * J.P + __ldg(J.perm + j);
    float g[8], y[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) g[k] = y[k] = 0.f;
    if (lane < N8) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(J.g + row * N + lane * 8)), g1 = __ldg(reinterpret_cast<const float4*>(J.g + row * N + lane * 8) + 1);
      const float4 y0 = __ldg(reinterpret_cast<const float4*>(J.y + row * N + lane * 8)), y1 = __ldg(reinterpret_cast<const float4*>(J.y + row * N + lane * 8) + 1);
      g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w; g[4] = g1.x; g[5] = g1.y; g[6] = g1.z; g[7] = g1.w;
      y[0] = y0.x; y[1] = y0.y; y[2] = y0.z; y[3] = y0.w; y[4] = y1.x; y[5] = y1.y; y[6] = y1.z; y[7] = y1.w;
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s = fmaf(g[k], y[k], s);
    s = warp_sum(s);
    const float inv = __ldg(J.inv + row);
#pragma unroll
    for (int k = 0; k < 8; ++k) d[k] = (inv < 0.f) ? g[k] * (-inv) : (g[k] - y[k] * s) * inv;   // g / eps when ||y|| < eps
  }
  if (lane < N8) {
    uint4 hi, lo;
    split8(d, hi, lo);
    const int tile = b * (J.Ppad >> 7) + (j >> 7), i = j & 127;
    const size_t off = ((size_t)tile * N8 * 16 + (size_t)(i >> 3)) * 64 + (size_t)(i & 7) * 8 + (size_t)lane * 1024;
    *reinterpret_cast<uint4*>(J.hi + off) = hi;
    if (J.lo != nullptr) *reinterpret_cast<uint4*>(J.lo + off) = lo;
  }
}

}  // namespace pnce
