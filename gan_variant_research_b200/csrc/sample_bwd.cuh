// Backward of PatchSampleF(use_mlp=False) for the module-split API: incoming d rows (B*P, C) ->
// normalise-backward -> dxT[b][c][rank[p]] (then the shared dense kernel writes d feat).
#pragma once
#include "common.cuh"

namespace pnce {

// grid = B * ntiles ; smem = 32*(C+1) floats
__global__ void __launch_bounds__(kThreads) k_rows_normbwd(const __grid_constant__ Params p,
                                                           const float* __restrict__ drows,
                                                           const float* __restrict__ rows, int l) {
  pdl_enter();
  extern __shared__ __align__(16) float st[];
  const LayerDev& L = p.L[l];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x % L.ntiles, b = blockIdx.x / L.ntiles;
  const int P = L.P, C = L.C, ldt = C + 1;
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    const int r = warp * 4 + rr, i = tile * kRowTile + r;
    if (i >= P) continue;                                       // warp-uniform
    const float* g = drows + ((size_t)b * P + i) * C;
    if (p.raw) {                                                // rows were the raw patches: dx = g
      for (int c = lane; c < C; c += 32) st[r * ldt + c] = g[c];
      continue;
    }
    const float* x = rows + ((size_t)b * P + i) * C;
    const float inv = L.qinv[(size_t)b * P + i];
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = fmaf(g[c], x[c], s);
    s = warp_sum(s);
    for (int c = lane; c < C; c += 32) {
      const float dx = (inv < 0.f) ? g[c] * (-inv) : (g[c] - x[c] * s) * inv;
      st[r * ldt + c] = dx;
    }
  }
  __syncthreads();
  const int i = tile * kRowTile + lane;
  if (i < P) {
    const int slot = L.rank[i];
    for (int c = warp; c < C; c += 8) L.dxT[((size_t)b * C + c) * L.dxpitch + slot] = st[lane * ldt + c];
  }
}

}  // namespace pnce
