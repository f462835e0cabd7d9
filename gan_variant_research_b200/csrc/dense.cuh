// The backward launch: write every layer's dense d loss / d tgt_feat (B,C,H,W) exactly once.
// Zero everywhere except the sampled positions, where the (duplicate-summed) gradient row value
// times the upstream scalar is patched into the same 16-byte store.  No memset + scatter pair, no
// atomics, deterministic.  Replaces index_put_(accumulate) + select_backward + zeros + adds of the
// reference's autograd chain -- this is the HBM-bound stage.
//
//   position h sampled?          bitmap[h>>5] bit (h&31)           (built by the prep CTA)
//   sorted-unique slot of h      u = prefix[h>>5] + popc(word & below(h))
//   its run of duplicate rows    j in [ustart[u], ustart[u+1])  ->  sum_j dxT[row][j]
#pragma once
#include "common.cuh"

namespace pnce {

constexpr int kDenseIters = 4;                 // 16-byte stores per thread per work item
struct DenseMap {
  long long start[PNCE_MAX_LAYERS + 1];        // work-item prefix per layer
  int tiles[PNCE_MAX_LAYERS];                  // items per (b,c) row
  int vec_ok[PNCE_MAX_LAYERS];                 // 1: HW % VEC == 0 and base 16 B aligned
  long long total;
};

template <typename T> struct Vec16;
template <> struct Vec16<float> { static constexpr int N = 4; };
template <> struct Vec16<__half> { static constexpr int N = 8; };
template <> struct Vec16<__nv_bfloat16> { static constexpr int N = 8; };

__device__ __forceinline__ float run_sum(const LayerDev& L, const float* dx, int u) {
  const int s = __ldg(L.ustart + u), e = __ldg(L.ustart + u + 1);
  float acc = 0.f;
  for (int j = s; j < e; ++j) acc += dx[j];
  return acc;
}

template <typename T>
__device__ __forceinline__ void dense_item(const LayerDev& L, long long local, int tiles, float g) {
  constexpr int VEC = Vec16<T>::N;
  const int HW = L.HW, P = L.P;
  const long long row = local / tiles;
  const int tile = (int)(local % tiles);
  T* out = reinterpret_cast<T*>(L.dtgt) + (size_t)row * HW;
  const float* dx = L.dxT + (size_t)row * P;
#pragma unroll
  for (int it = 0; it < kDenseIters; ++it) {
    const int chunk = (tile * kDenseIters + it) * kThreads + threadIdx.x;
    const int hw0 = chunk * VEC;
    if (hw0 >= HW) break;
    const unsigned word = __ldg(L.bitmap + (hw0 >> 5));
    const unsigned bits = (word >> (hw0 & 31)) & ((1u << VEC) - 1u);
    float v[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) v[e] = 0.f;
    if (bits != 0u) {
      int u = (int)__ldg(L.prefix + (hw0 >> 5)) + __popc(word & ((1u << (hw0 & 31)) - 1u));
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        if ((bits >> e) & 1u) {
          v[e] = run_sum(L, dx, u) * g;
          ++u;
        }
      }
    }
    if constexpr (VEC == 4) {
      float4 o = make_float4(v[0], v[1], v[2], v[3]);
      __stcs(reinterpret_cast<float4*>(out + hw0), o);
    } else {
      union { uint4 u4; T h[8]; } pk;
#pragma unroll
      for (int e = 0; e < 8; ++e) pk.h[e] = from_f32<T>(v[e]);
      __stcs(reinterpret_cast<uint4*>(out + hw0), pk.u4);
    }
  }
}

// Scalar fallback for maps whose H*W is not a multiple of the vector width (tests only).
template <typename T>
__device__ __forceinline__ void dense_item_scalar(const LayerDev& L, long long local, int tiles, float g) {
  constexpr int VEC = Vec16<T>::N;
  const int HW = L.HW, P = L.P;
  const long long row = local / tiles;
  const int tile = (int)(local % tiles);
  T* out = reinterpret_cast<T*>(L.dtgt) + (size_t)row * HW;
  const float* dx = L.dxT + (size_t)row * P;
  const int base = tile * kDenseIters * kThreads * VEC;
  for (int k = threadIdx.x; k < kDenseIters * kThreads * VEC; k += kThreads) {
    const int h = base + k;
    if (h >= HW) break;
    const unsigned word = __ldg(L.bitmap + (h >> 5));
    float v = 0.f;
    if ((word >> (h & 31)) & 1u) {
      const int u = (int)__ldg(L.prefix + (h >> 5)) + __popc(word & ((1u << (h & 31)) - 1u));
      v = run_sum(L, dx, u) * g;
    }
    out[h] = from_f32<T>(v);
  }
}

// Persistent grid (a multiple of the SM count); items are equal-sized so a static stride balances.
__global__ void __launch_bounds__(kThreads) k_dense_bwd(const __grid_constant__ Params p,
                                                        const __grid_constant__ DenseMap m) {
  const float g = p.grad_out ? __ldg(p.grad_out) : 1.0f;
  for (long long item = blockIdx.x; item < m.total; item += gridDim.x) {
    int l = 0;
    for (int i = 1; i < p.n_layers; ++i)
      if (item >= m.start[i]) l = i;
    const LayerDev& L = p.L[l];
    const long long local = item - m.start[l];
    const int tiles = m.tiles[l];
    if (m.vec_ok[l]) {
      if (p.dtype == PNCE_F32) dense_item<float>(L, local, tiles, g);
      else if (p.dtype == PNCE_F16) dense_item<__half>(L, local, tiles, g);
      else dense_item<__nv_bfloat16>(L, local, tiles, g);
    } else {
      if (p.dtype == PNCE_F32) dense_item_scalar<float>(L, local, tiles, g);
      else if (p.dtype == PNCE_F16) dense_item_scalar<__half>(L, local, tiles, g);
      else dense_item_scalar<__nv_bfloat16>(L, local, tiles, g);
    }
  }
}

}  // namespace pnce
