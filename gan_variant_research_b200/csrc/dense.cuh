// The backward launch: write every layer's dense d loss / d tgt_feat (B,C,H,W) exactly once.
// Replaces index_put_(accumulate) + select_backward + zeros + adds of the reference's autograd
// chain -- this is the HBM-bound stage of the path.
//
// Work item = one warp x one segment of 1024 consecutive positions of one (b,c) row:
//   1. the segment's 32 bitmap words and their popcount prefixes arrive as two coalesced 128 B loads
//      (L1-resident: every row of a layer reuses them);
//   2. the whole segment is zero-filled with straight, fully coalesced 16-byte streaming stores;
//   3. __syncwarp() orders the fill before the patches; each lane then walks the set bits of ITS OWN
//      word and overwrites those positions with (sum of the run of duplicate rows) x upstream scalar.
//      The patched sector is still dirty in L2, so DRAM sees one write per sector.
// No atomics, no memset + scatter pair over HBM, deterministic (runs are summed in sorted order).
//
//   position h sampled?            bitmap[h>>5] bit (h&31)            (prep CTA of the forward)
//   sorted-unique slot of h        u = prefix[h>>5] + popc(word & below(h))
//   run of duplicate rows of u     j in [ustart[u], ustart[u+1])  ->  sum_j dxT[row][j]
#pragma once
#include "common.cuh"

namespace pnce {

constexpr int kSegPos = 1024;                  // positions per warp work item (32 bitmap words)
struct DenseMap {
  long long start[PNCE_MAX_LAYERS + 1];        // work-item prefix per layer
  int segs[PNCE_MAX_LAYERS];                   // segments per (b,c) row
  int vec_ok[PNCE_MAX_LAYERS];                 // 1: HW % VEC == 0 and base 16 B aligned
  long long total;
};

template <typename T> struct Vec16;
template <> struct Vec16<float> { static constexpr int N = 4; };
template <> struct Vec16<__half> { static constexpr int N = 8; };
template <> struct Vec16<__nv_bfloat16> { static constexpr int N = 8; };

template <typename T>
__device__ __forceinline__ void dense_segment(const LayerDev& L, long long row, int seg, float g,
                                              bool vec_ok) {
  constexpr int VEC = Vec16<T>::N;
  const int lane = threadIdx.x & 31;
  const int HW = L.HW;
  T* out = reinterpret_cast<T*>(L.dtgt) + (size_t)row * HW;
  const int h0 = seg * kSegPos;
  const int w = (h0 >> 5) + lane;
  unsigned word = 0u, pre = 0u;
  if (w < L.nwords) {
    word = __ldg(L.bitmap + w);
    pre = __ldg(L.prefix + w);
  }
  // ---- zero fill ------------------------------------------------------------------------------
  if (vec_ok) {
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    uint4* o = reinterpret_cast<uint4*>(out + h0);
    constexpr int kChunks = kSegPos / VEC;                  // 16-byte chunks per segment
    const int n = min(kChunks, (HW - h0) / VEC);
#pragma unroll
    for (int it = 0; it < kChunks / 32; ++it) {
      const int q = it * 32 + lane;
      if (q < n) __stcs(o + q, z);
    }
  } else {
    const int n = min(kSegPos, HW - h0);
    for (int k = lane; k < n; k += 32) out[h0 + k] = from_f32<T>(0.f);
  }
  if (__ballot_sync(0xffffffffu, word != 0u) == 0u) return;
  __syncwarp();
  // ---- patch the sampled positions of this lane's word -----------------------------------------
  const float* dx = L.dxT + (size_t)row * L.P;
  int u = (int)pre;
  while (word != 0u) {
    const int bit = __ffs(word) - 1;
    word &= word - 1u;
    const int s = __ldg(L.ustart + u), e = __ldg(L.ustart + u + 1);
    float acc = dx[s];
    for (int j = s + 1; j < e; ++j) acc += dx[j];
    out[(w << 5) + bit] = from_f32<T>(acc * g);
    ++u;
  }
}

// Persistent grid of independent warps; items are equal-sized so a static stride balances.
__global__ void __launch_bounds__(kThreads) k_dense_bwd(const __grid_constant__ Params p,
                                                        const __grid_constant__ DenseMap m) {
  const float g = p.grad_out ? __ldg(p.grad_out) : 1.0f;
  const long long warps = (long long)gridDim.x * (kThreads / 32);
  for (long long item = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); item < m.total;
       item += warps) {
    int l = 0;
    for (int i = 1; i < p.n_layers; ++i)
      if (item >= m.start[i]) l = i;
    const LayerDev& L = p.L[l];
    const long long local = item - m.start[l];
    const int segs = m.segs[l];
    const long long row = local / segs;
    const int seg = (int)(local - row * segs);
    const bool vec_ok = m.vec_ok[l] != 0;
    if (p.dtype == PNCE_F32) dense_segment<float>(L, row, seg, g, vec_ok);
    else if (p.dtype == PNCE_F16) dense_segment<__half>(L, row, seg, g, vec_ok);
    else dense_segment<__nv_bfloat16>(L, row, seg, g, vec_ok);
  }
}

}  // namespace pnce
