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
template <typename T>
__global__ void __launch_bounds__(kThreads, 8) k_dense_bwd(const __grid_constant__ Params p,
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
    dense_segment<T>(L, row, seg, g, m.vec_ok[l] != 0);
  }
}

// -------------------------------------------------------------------------------------------------
// Bulk-copy variant (the default when every layer's rows are 16-byte tileable).  The dense gradient
// is produced chunk by chunk IN SHARED MEMORY and leaves the SM through the bulk-copy (TMA) engine:
// DRAM sees nothing but full-line writes, the LSU only touches the few sampled values.
//
//   item   = one chunk of kChunkPos consecutive positions of one (b,c) row
//   column = all rows of one (layer, chunk): the sampled POSITIONS of a column are the same for
//            every row (ids are shared by the batch and by all channels, patchnce_cut.py:63), only
//            the values differ.  Items are ordered (layer, chunk, row) and each CTA takes one
//            contiguous range, so inside a column the staging buffers are zeroed once and every
//            item merely overwrites the same few positions.
//   per item: thread t owns sorted slots j0+t, j0+t+128, ... of the chunk; a slot that starts a run
//            of equal ids sums the run's gradient rows (coalesced reads of dxT[row][j]) and stores
//            sum * upstream into the staging buffer; fence.proxy.async; one barrier; thread 0 issues
//            cp.async.bulk.global.shared::cta for the 16 KB chunk (SASS UBLKCP) and, before the
//            next barrier, waits until the copy engine has finished READING the other buffer.
// Deterministic (runs are summed in sorted order), no atomics, no partial-sector writes.
// -------------------------------------------------------------------------------------------------
constexpr int kChunkPos = 4096;                // positions per item (16 KB fp32, 8 KB fp16/bf16)
constexpr int kDenseTmaThreads = 128;
constexpr int kDenseBufs = 2;

struct DenseTmaMap {
  long long start[PNCE_MAX_LAYERS + 1];        // item prefix per layer
  int chunks[PNCE_MAX_LAYERS];                 // chunks per (b,c) row
  long long total;
  long long per_cta;                           // items per CTA (contiguous range)
  int flags;                                   // experiment bits: 1 = skip the patch values
};

__device__ __forceinline__ void bulk_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"(static_cast<uint32_t>(__cvta_generic_to_shared(ssrc))), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

// first sorted slot whose id is >= pos (sid ascending, P entries)
__device__ __forceinline__ int lower_slot(const int* __restrict__ sid, int P, int pos) {
  int lo = 0, hi = P;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(sid + mid) < pos) lo = mid + 1; else hi = mid;
  }
  return lo;
}

constexpr int kDenseSlots = 4;                 // sorted slots per thread and chunk on the prefetching path

template <typename T>
__global__ void __launch_bounds__(kDenseTmaThreads, 7) k_dense_tma(const __grid_constant__ Params p,
                                                                const __grid_constant__ DenseTmaMap m) {
  __shared__ __align__(128) T buf[kDenseBufs][kChunkPos];
  const int tid = threadIdx.x;
  const float g = p.grad_out ? __ldg(p.grad_out) : 1.0f;
  const long long first = (long long)blockIdx.x * m.per_cta;
  const long long last = min(first + m.per_cta, m.total);
  const bool lsu_out = (m.flags & 2) != 0;             // experiment: copy out with LDS/STG instead of UBLKCP
  int cur_l = -1, cur_chunk = -1;
  int j0 = 0, j1 = 0, h0 = 0, nbytes = 0, P = 0, HW = 0;
  long long rows = 1, lstart = 0;
  const int* __restrict__ sid = nullptr;
  const float* __restrict__ dxT = nullptr;
  T* dtgt = nullptr;
  // per-thread view of the column: slot s is sorted slot j0 + tid + 128 s; slen = run length when the
  // slot heads a run of equal ids, 0 otherwise
  int spos[kDenseSlots], slen[kDenseSlots];
  float cv[kDenseSlots], nv[kDenseSlots];
  bool fast = false;
#pragma unroll
  for (int s = 0; s < kDenseSlots; ++s) { spos[s] = 0; slen[s] = 0; cv[s] = 0.f; nv[s] = 0.f; }
  int k = 0;
  for (long long item = first; item < last; ++item, ++k) {
    // ---- (layer, chunk, row) of this item; items of a layer are ordered chunk-major ----------
    int l = cur_l < 0 ? 0 : cur_l;
    while (l + 1 < p.n_layers && item >= m.start[l + 1]) ++l;
    if (l != cur_l) {
      const LayerDev& L = p.L[l];
      P = L.P; HW = L.HW;
      rows = (long long)p.B * L.C;
      lstart = m.start[l];
      sid = L.sid; dxT = L.dxT; dtgt = reinterpret_cast<T*>(L.dtgt);
    }
    const long long local = item - lstart;
    const int chunk = (int)(local / rows);
    const long long row = local - (long long)chunk * rows;
    const float* __restrict__ dx = dxT + (size_t)row * P;
    if (l != cur_l || chunk != cur_chunk) {
      // new column: drain the copy engine's reads, re-zero the staging buffers, find the slot range
      cur_l = l; cur_chunk = chunk;
      if (tid == 0) bulk_wait_read<0>();
      __syncthreads();
      for (int i = tid; i < (int)(sizeof(buf) / 16); i += kDenseTmaThreads)
        reinterpret_cast<uint4*>(&buf[0][0])[i] = make_uint4(0u, 0u, 0u, 0u);
      h0 = chunk * kChunkPos;
      nbytes = min(kChunkPos, HW - h0) * (int)sizeof(T);
      j0 = lower_slot(sid, P, h0);
      j1 = lower_slot(sid, P, h0 + kChunkPos);
      fast = (j1 - j0) <= kDenseTmaThreads * kDenseSlots;
#pragma unroll
      for (int s = 0; s < kDenseSlots; ++s) {
        const int j = j0 + tid + kDenseTmaThreads * s;
        spos[s] = 0; slen[s] = 0;
        if (fast && j < j1) {
          const int pos = __ldg(sid + j);
          if (j == 0 || __ldg(sid + j - 1) != pos) {          // head of a run of duplicate ids
            int e = 1;
            while (j + e < P && __ldg(sid + j + e) == pos) ++e;
            spos[s] = pos - h0; slen[s] = e;
          }
        }
      }
      __syncthreads();
      if (fast) {
#pragma unroll
        for (int s = 0; s < kDenseSlots; ++s)
          if (slen[s] > 0) cv[s] = dx[j0 + tid + kDenseTmaThreads * s];
      }
    }
    // prefetch the next row's values of the same column: in flight across the barrier and the copy
    if (fast && item + 1 < last && row + 1 < rows) {
#pragma unroll
      for (int s = 0; s < kDenseSlots; ++s)
        if (slen[s] > 0) nv[s] = dx[P + j0 + tid + kDenseTmaThreads * s];
    }
    T* b = buf[k & 1];
    if (!(m.flags & 1)) {
      if (fast) {
#pragma unroll
        for (int s = 0; s < kDenseSlots; ++s) {
          if (slen[s] > 0) {
            float acc = cv[s];
            const int j = j0 + tid + kDenseTmaThreads * s;
            for (int e = 1; e < slen[s]; ++e) acc += dx[j + e];
            b[spos[s]] = from_f32<T>(acc * g);
          }
        }
      } else {
        for (int j = j0 + tid; j < j1; j += kDenseTmaThreads) {
          const int pos = __ldg(sid + j);
          if (j == 0 || __ldg(sid + j - 1) != pos) {
            float acc = dx[j];
            for (int jj = j + 1; jj < P && __ldg(sid + jj) == pos; ++jj) acc += dx[jj];
            b[pos - h0] = from_f32<T>(acc * g);
          }
        }
      }
    }
    T* dst = dtgt + (size_t)row * HW + h0;
    if (!lsu_out) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      // every bulk copy issued so far has finished READING shared memory => the buffer the next
      // iteration writes is free once the barrier below is passed
      if (tid == 0) bulk_wait_read<kDenseBufs - 2>();
      __syncthreads();
      if (tid == 0) {
        bulk_s2g(dst, b, (uint32_t)nbytes);
        bulk_commit();
      }
    } else {
      __syncthreads();
      const int n16 = nbytes >> 4;
      for (int i = tid; i < n16; i += kDenseTmaThreads)
        reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(b)[i];
    }
#pragma unroll
    for (int s = 0; s < kDenseSlots; ++s) cv[s] = nv[s];
  }
  if (tid == 0) bulk_wait_read<0>();
}

// -------------------------------------------------------------------------------------------------
// Flat variant (default): ONE SMALL CTA PER 8 KB TILE, tiles dispatched in address order.
// Measured on B200 (scratch/fillbench.cu): a 3.2 GB fill by persistent CTAs -- LSU or bulk-copy, any
// chunk size -- saturates at ~6.3 TB/s, while many small CTAs that each write one contiguous 8 KB tile
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

template <typename T, int THREADS>
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
  const float* __restrict__ dx = L.dxT + (size_t)row * P;
  const int* __restrict__ sid = L.sid;
  // the slots this thread owns: the gradient value and the ids around it are loaded independently
  // of each other (one round trip), before the zero fill and the barriers
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
      tile[ps[s] - h0] = from_f32<T>(acc * g);
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
  const int n16 = (min(TP, HW - h0) * (int)sizeof(T)) >> 4;
  uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<T*>(L.dtgt) + (size_t)row * HW + h0);
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = k * THREADS + tid;
    if (i < n16) dst[i] = t4[i];
  }
}

}  // namespace pnce
