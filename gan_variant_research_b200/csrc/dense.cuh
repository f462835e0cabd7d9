// The backward launch: write every layer's dense d loss / d tgt_feat (B,C,H,W) exactly once.
// Replaces index_put_(accumulate) + select_backward + zeros + adds of the reference's autograd
// chain -- this is the HBM-bound stage of the path.
#pragma once
#include "common.cuh"

namespace pnce {

// -------------------------------------------------------------------------------------------------
// ONE SMALL CTA PER 8 KB TILE, tiles dispatched in address order.
// Measured on B200 (scratch/fillbench.cu): a 3.2 GB fill by persistent CTAs -- LSU or bulk-copy
// (cp.async.bulk from shared memory), any chunk size -- saturates at ~6.3 TB/s, while many small CTAs that each write one contiguous 8 KB tile
// reach 7.5 TB/s (the hardware's in-order CTA dispatch keeps the DRAM write front linear).  So:
//   CTA = 128 threads, tile = 8 KB of one (b,c) row, staged in shared memory:
//     1. the tile's sorted-slot range [j0,j1) comes from k_prep's table (2 loads); thread t issues the
//        gradient-row loads of slots j0+t, j0+t+128, ... that head a run of equal ids (coalesced);
//     2. zero the staging tile (4 x STS.128 per thread), barrier, drop the run sums in, barrier;
//     3. copy out: 4 x (LDS.128 + STG.128) per thread, fully coalesced.
// DRAM sees one full-line write per line; no atomics, no partial sectors, deterministic.
// -------------------------------------------------------------------------------------------------
constexpr int kFlatBytes = 8192;

struct DenseFlatMap {
  long long start[PNCE_MAX_LAYERS + 1];        // tile prefix per layer
  int tiles[PNCE_MAX_LAYERS];                  // tiles per (b,c) row
  int flags;                                   // experiment bits: 1 = skip the patch values
};

template <typename T, int THREADS, bool VEC>
__global__ void __launch_bounds__(THREADS) k_dense_flat(const __grid_constant__ Params p,
                                                        const __grid_constant__ DenseFlatMap m) {
  constexpr int TP = kFlatBytes / (int)sizeof(T);          // positions per tile
  constexpr int SUB = TP / kTilePos;                       // k_prep table entries per tile
  constexpr int KS = 256 / THREADS;                        // slots per thread on the prefetching path
  constexpr int NV = kFlatBytes / 16 / THREADS;            // 16-byte vectors per thread
  __shared__ __align__(16) T tile[TP];
  const int tid = threadIdx.x;
  const long long item = blockIdx.x;
  int l = 0;
  for (int i = 1; i < p.n_layers; ++i)
    if (item >= m.start[i]) l = i;
  const LayerDev& L = p.L[l];
  const long long local = item - m.start[l];
  const int tiles = m.tiles[l];
  const long long row = local / tiles;
  const int t = (int)(local - row * tiles);
  const int h0 = t * TP;
  const int P = L.P, HW = L.HW;
  const int ntab = (HW + kTilePos - 1) / kTilePos;
  const int j0 = __ldg(L.cslot + t * SUB);
  const int j1 = __ldg(L.cslot + min(t * SUB + SUB, ntab));
  const float* __restrict__ dx = L.dxT + (size_t)row * L.dxpitch;
  const int* __restrict__ sid = L.sid;
  // the slots this thread owns: the gradient value and the ids around it are loaded independently
  // of each other (one round trip), before the zero fill and the barriers
  float val[KS];
  int ps[KS], prev[KS], next[KS];
#pragma unroll
  for (int s = 0; s < KS; ++s) {
    const int j = j0 + tid + THREADS * s;
    const bool in = j < j1 && !(m.flags & 1);
    val[s] = (in && !(m.flags & 16)) ? __ldcs(dx + j) : 0.f;   // streaming: last use of the row (see st_dx in loss_tc.cuh)
    ps[s] = in ? __ldg(sid + j) : -1;
    prev[s] = (in && j > 0) ? __ldg(sid + j - 1) : -2;
    next[s] = (in && j + 1 < P) ? __ldg(sid + j + 1) : -3;
  }
  const float g = p.grad_out ? __ldg(p.grad_out) : 1.0f;
  uint4* t4 = reinterpret_cast<uint4*>(tile);
#pragma unroll
  for (int k = 0; k < NV; ++k) t4[k * THREADS + tid] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
#pragma unroll
  for (int s = 0; s < KS; ++s) {
    if (ps[s] >= 0 && prev[s] != ps[s]) {                    // head of a run of duplicate ids
      float acc = val[s];
      if (next[s] == ps[s]) {                                // rare: sum the run in sorted order
        const int j = j0 + tid + THREADS * s;
        for (int jj = j + 1; jj < P && __ldg(sid + jj) == ps[s]; ++jj) acc += dx[jj];
      }
      tile[ps[s] - h0] = from_f32<T>((m.flags & 8) ? acc * 0.f : acc * g);   // 8: experiment, loads without values
    }
  }
  // more than 256 sampled positions in one tile (num_patches >> 256 on a small map)
  for (int j = j0 + tid + THREADS * KS; j < j1; j += THREADS) {
    const int q = __ldg(sid + j);
    if (__ldg(sid + j - 1) != q) {
      float acc = dx[j];
      for (int jj = j + 1; jj < P && __ldg(sid + jj) == q; ++jj) acc += dx[jj];
      if (!(m.flags & 1)) tile[q - h0] = from_f32<T>(acc * g);
    }
  }
  __syncthreads();
  const int npos = min(TP, HW - h0);
  T* drow = reinterpret_cast<T*>(L.dtgt) + (size_t)row * HW + h0;
  if (VEC) {
    const int n16 = (npos * (int)sizeof(T)) >> 4;
    uint4* dst = reinterpret_cast<uint4*>(drow);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int i = k * THREADS + tid;
      if (i < n16) __stcs(dst + i, t4[i]);               // streaming: keep dxT / sid resident in L2 instead
    }
  } else {                                                   // odd map sizes: element-wise, still coalesced
    for (int i = tid; i < npos; i += THREADS) drow[i] = tile[i];
  }
}

// -------------------------------------------------------------------------------------------------
// STORE-FIRST variant (same grid, same tile order, no shared memory): every thread issues its share of the
// tile's 128-bit ZERO stores at once -- they depend on nothing -- and only then walks the dependent chain
// cslot -> sorted ids / gradient rows; after one CTA barrier (which orders the zero stores before what
// follows, at CTA scope) the run heads overwrite their positions with scattered element stores.  The
// lines are still dirty in L2 a microsecond later, so the element stores merge there and DRAM still sees
// one full-line write per line; what disappears is the staging tile, a barrier and the two memory
// round trips that used to sit in front of every tile's first store.
// -------------------------------------------------------------------------------------------------
template <typename T, int THREADS>
__global__ void __launch_bounds__(THREADS) k_dense_direct(const __grid_constant__ Params p,
                                                          const __grid_constant__ DenseFlatMap m) {
  constexpr int TP = kFlatBytes / (int)sizeof(T);          // positions per tile
  constexpr int SUB = TP / kTilePos;                       // k_prep table entries per tile
  constexpr int KS = 256 / THREADS;                        // slots per thread on the prefetching path
  constexpr int NV = kFlatBytes / 16 / THREADS;            // 16-byte vectors per thread
  const int tid = threadIdx.x;
  const long long item = blockIdx.x;
  int l = 0;
  for (int i = 1; i < p.n_layers; ++i)
    if (item >= m.start[i]) l = i;
  const LayerDev& L = p.L[l];
  const long long local = item - m.start[l];
  const int tiles = m.tiles[l];
  const long long row = local / tiles;
  const int t = (int)(local - row * tiles);
  const int h0 = t * TP;
  const int P = L.P, HW = L.HW;
  const int npos = min(TP, HW - h0);
  T* drow = reinterpret_cast<T*>(L.dtgt) + (size_t)row * HW + h0;
  {
    const int n16 = (npos * (int)sizeof(T)) >> 4;
    uint4* dst = reinterpret_cast<uint4*>(drow);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int i = k * THREADS + tid;
      if (i < n16) dst[i] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  const int ntab = (HW + kTilePos - 1) / kTilePos;
  const int j0 = __ldg(L.cslot + t * SUB);
  const int j1 = __ldg(L.cslot + min(t * SUB + SUB, ntab));
  const float* __restrict__ dx = L.dxT + (size_t)row * L.dxpitch;
  const int* __restrict__ sid = L.sid;
  float val[KS];
  int ps[KS], prev[KS], next[KS];
#pragma unroll
  for (int s = 0; s < KS; ++s) {
    const int j = j0 + tid + THREADS * s;
    const bool in = j < j1 && !(m.flags & 1);
    val[s] = in ? dx[j] : 0.f;
    ps[s] = in ? __ldg(sid + j) : -1;
    prev[s] = (in && j > 0) ? __ldg(sid + j - 1) : -2;
    next[s] = (in && j + 1 < P) ? __ldg(sid + j + 1) : -3;
  }
  const float g = p.grad_out ? __ldg(p.grad_out) : 1.0f;
  __syncthreads();                                           // zero stores (any thread) before element stores
#pragma unroll
  for (int s = 0; s < KS; ++s) {
    if (ps[s] >= 0 && prev[s] != ps[s]) {                    // head of a run of duplicate ids
      float acc = val[s];
      if (next[s] == ps[s]) {                                // rare: sum the run in sorted order
        const int j = j0 + tid + THREADS * s;
        for (int jj = j + 1; jj < P && __ldg(sid + jj) == ps[s]; ++jj) acc += dx[jj];
      }
      drow[ps[s] - h0] = from_f32<T>(acc * g);
    }
  }
  // more than 256 sampled positions in one tile (num_patches >> 256 on a small map)
  for (int j = j0 + tid + THREADS * KS; j < j1; j += THREADS) {
    const int q = __ldg(sid + j);
    if (__ldg(sid + j - 1) != q) {
      float acc = dx[j];
      for (int jj = j + 1; jj < P && __ldg(sid + jj) == q; ++jj) acc += dx[jj];
      if (!(m.flags & 1)) drow[q - h0] = from_f32<T>(acc * g);
    }
  }
}

}  // namespace pnce
