// Channels-last (NHWC) maps on the tensor-core path: gather and dense backward.
//
// The reference's loss takes contiguous NCHW maps (`feat.view(B, C, -1)`, patchnce_cut.py:56), where every channel of
// a sampled patch sits in a DRAM line of its own: the gather of DESIGN.md 4 pulls 29.6 MB of 128-byte lines per image
// to use 1.57 MB of them.  A generator run in torch.channels_last hands out (B, H, W, C) storage, where a patch IS a
// contiguous row of C values -- the same algorithm then reads exactly the bytes it uses.  This is an EXTENSION of the
// reference's interface (its .view() rejects such maps), selected by the `layout` argument of pnce_fwd_ex / pnce_bwd_ex;
// ids, loss and gradient values follow the same law as the NCHW path (the oracle on the permuted data).
//   k_gather_tc_nhwc : warp <-> (side, image, 8 sorted patch slots, 32-channel chunk); lane <-> (slot, 8 channels):
//                      4 lanes read one patch's 128 contiguous bytes (fp32), 8 lanes write one 128-byte core matrix.
//   k_dense_nhwc     : one 128-thread CTA per <= 8 KB tile of d tgt (a few whole positions), address order; a tile
//                      without sampled positions (most of them) is a pure zero fill from registers.
// The gradient rows arrive ROW-major ([image][sorted slot][C], Params::nhwc) from the loss kernels.
#pragma once
#include "common.cuh"
#include "gather_tc.cuh"

namespace pnce {

template <typename T> struct Vec8;
template <> struct Vec8<float> {
  // one 256-bit load (sm_100: LDG.256) = this lane's whole 32-byte sector; two 128-bit loads ask L1 for it twice
  static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
    if ((reinterpret_cast<uintptr_t>(p) & 31u) == 0) {
      asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                   : "l"(p));
    } else {
      const float4 a = __ldcg(reinterpret_cast<const float4*>(p)), b = __ldcg(reinterpret_cast<const float4*>(p) + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
  }
};
template <> struct Vec8<__half> {
  static __device__ __forceinline__ void load(const __half* p, float (&v)[8]) {
    const uint4 a = __ldcg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __half2 h = *reinterpret_cast<const __half2*>(&w[k]);
      v[2 * k] = __low2float(h); v[2 * k + 1] = __high2float(h);
    }
  }
};
template <> struct Vec8<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 a = __ldcg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[2 * k] = __uint_as_float(w[k] << 16); v[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
    }
  }
};

// One warp item of the NHWC gather = (side, image, 8 sorted slots, 32-channel chunk); it writes exactly what
// gather_tc_chunk writes for the same coordinates.  Split in two so that a warp can have the loads of several items
// in flight before it touches the first value (the kernel is latency-bound: one dependent sid -> row chain per item).
struct NhwcItem {
  int s, b, side, p, c8;
  bool valid;
};
template <typename T>
__device__ __forceinline__ void gather_nhwc_load(const LayerDev& L, int b0, int B, long long witem, int lane, int side0,
                                                 NhwcItem& it, float (&v)[8], const unsigned long long* keys_s = nullptr) {
  // 32-bit index arithmetic (the launch checks that the item count fits): 64-bit divisions by run-time values cost ~100
  // instructions each, and this kernel has few others
  const unsigned nchunk = (unsigned)L.nchunk, np8 = (unsigned)L.Ppad >> 3, wi = (unsigned)witem;
  const unsigned g1 = wi / nchunk;
  it.s = (int)(wi - g1 * nchunk);
  const unsigned rest = g1 / np8;
  const int p8 = (int)(g1 - rest * np8);
  const unsigned sd = rest / (unsigned)B;
  it.b = b0 + (int)(rest - sd * (unsigned)B);
  it.side = (int)sd + side0;                                 // 0 = src (k), 1 = tgt (q); side0 = 1: tgt only
  const int C = L.C, HW = L.HW;
  it.p = p8 * 8 + (lane & 7);                                // sorted slot
  it.c8 = it.s * 4 + (lane >> 3);                            // 8-channel group
  const int c0 = it.c8 * 8;
  it.valid = it.p < L.P;
  const int id = it.valid ? (keys_s != nullptr ? (int)(keys_s[it.p] >> 32) : __ldg(L.sid + it.p)) : 0;
  const T* base = reinterpret_cast<const T*>(it.side ? L.tgt : L.src);
  const T* row = base + ((size_t)it.b * HW + id) * C + c0;
  const bool vec = (((size_t)C * sizeof(T)) & 15u) == 0 && (reinterpret_cast<uintptr_t>(base) & 15u) == 0;
  if (it.valid && vec && c0 + 8 <= C) {
    Vec8<T>::load(row, v);
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = (it.valid && c0 + k < C) ? to_f32<T>(__ldcg(row + k)) : 0.f;
  }
}
__device__ __forceinline__ void gather_nhwc_store(const LayerDev& L, const NhwcItem& it, int lane, const float (&v)[8]) {
  const int nchunk = L.nchunk, Ppad = L.Ppad, Cp8 = L.Cp >> 3;
  const int s = it.s, b = it.b, side = it.side, p = it.p, c8 = it.c8;
  float ss = 0.f;
  int bad = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    ss = fmaf(v[k], v[k], ss);
    bad |= !isfinite(v[k]);
  }
  ss += __shfl_xor_sync(0xffffffffu, ss, 8);
  ss += __shfl_xor_sync(0xffffffffu, ss, 16);
  bad |= __shfl_xor_sync(0xffffffffu, bad, 8);
  bad |= __shfl_xor_sync(0xffffffffu, bad, 16);
  float* ssbase = side ? L.qss : L.kss;                      // NULL in head mode (the head's output is normalised)
  if (ssbase != nullptr && (lane >> 3) == 0) ssbase[((size_t)b * nchunk + s) * Ppad + p] = bad ? __int_as_float(0x7fc00000) : ss;
  uint32_t hw[4], lw[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float a = v[2 * k], c2 = v[2 * k + 1];
    const __nv_bfloat16 ah = __float2bfloat16_rn(a), ch = __float2bfloat16_rn(c2);
    const __nv_bfloat16 al = __float2bfloat16_rn(a - __bfloat162float(ah));
    const __nv_bfloat16 cl = __float2bfloat16_rn(c2 - __bfloat162float(ch));
    hw[k] = pack_bf16x2(ah, ch);
    lw[k] = pack_bf16x2(al, cl);
  }
  __nv_bfloat16* hi_base = side ? L.qhi : L.khi;
  __nv_bfloat16* lo_base = side ? L.qlo : L.klo;
  size_t cm;                                                  // core-matrix index, as in gather_tc_chunk
  if (side || L.head_src_rows) cm = (((size_t)b * (Ppad >> 7) + (p >> 7)) * Cp8 + c8) * 16 + ((p & 127) >> 3);
  else {
    const int pb = p >> 8;
    const int nb8 = min(256, Ppad - pb * 256) >> 3;
    cm = (((size_t)b * Ppad + (size_t)pb * 256) * Cp8) / 8 + (size_t)c8 * nb8 + ((p & 255) >> 3);
  }
  const size_t off = cm * 64 + (size_t)(p & 7) * 8;
  *reinterpret_cast<uint4*>(hi_base + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
  if (lo_base != nullptr) *reinterpret_cast<uint4*>(lo_base + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
  if (!side && !L.head_src_rows) {
    const size_t off2 = (((size_t)b * (Ppad >> 3) + (p >> 3)) * Cp8 + c8) * 64 + (size_t)(p & 7) * 8;
    *reinterpret_cast<uint4*>(L.k2hi + off2) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    if (L.k2lo != nullptr) *reinterpret_cast<uint4*>(L.k2lo + off2) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
  }
}

constexpr int kNhwcItemsPerWarp = 2;
template <typename T>
__device__ __forceinline__ void gather_nhwc_warp(const LayerDev& L, int b0, int B, long long w0, long long nitems, int lane,
                                                 int side0, const unsigned long long* keys_s = nullptr) {
  NhwcItem it[kNhwcItemsPerWarp];
  float v[kNhwcItemsPerWarp][8];
#pragma unroll
  for (int n = 0; n < kNhwcItemsPerWarp; ++n)
    if (w0 + n < nitems) gather_nhwc_load<T>(L, b0, B, w0 + n, lane, side0, it[n], v[n], keys_s);
#pragma unroll
  for (int n = 0; n < kNhwcItemsPerWarp; ++n)
    if (w0 + n < nitems) gather_nhwc_store(L, it[n], lane, v[n]);
}

// grid = sum_l ceil(items_l / 16) CTAs of 8 warps, 2 items per warp; items_l = sides * B * (Ppad_l / 8) * nchunk_l
__global__ void __launch_bounds__(kThreads) k_gather_tc_nhwc(const __grid_constant__ Params p,
                                                             const __grid_constant__ BlockMap m) {
  pdl_enter();
  const long long blk = blockIdx.x;
  if (blk == 0 && threadIdx.x == 0 && p.b0 == 0 && p.counter != nullptr) {
    p.counter[0] = 0u; p.counter[1] = 0u;                      // launch-sequence state, as in k_gather_tc
    if (p.nonfinite != nullptr) p.nonfinite[1] = 0;
  }
  const int slot = find_layer(m, blk, p.n_layers);
  const int l = m.layer[slot];
  const LayerDev& L = p.L[l];
  const long long w0 = ((blk - m.start[slot]) * 8 + (threadIdx.x >> 5)) * kNhwcItemsPerWarp;
  const long long nitems = (long long)(2 - p.side0) * p.bn * (L.Ppad >> 3) * L.nchunk;
  if (w0 >= nitems) return;
  const int lane = threadIdx.x & 31;
  if (p.dtype == PNCE_F32) gather_nhwc_warp<float>(L, p.b0, p.bn, w0, nitems, lane, p.side0);
  else if (p.dtype == PNCE_F16) gather_nhwc_warp<__half>(L, p.b0, p.bn, w0, nitems, lane, p.side0);
  else gather_nhwc_warp<__nv_bfloat16>(L, p.b0, p.bn, w0, nitems, lane, p.side0);
}

// The same gather with the id prep folded in (small problems, see k_gather_tc_fold in gather_tc.cuh): every CTA sorts
// its layer's ids in shared memory, the first CTA of each layer writes the tables.  dynamic smem = N2max * 8 + 64.
__global__ void __launch_bounds__(kThreads) k_gather_tc_nhwc_fold(const __grid_constant__ Params p,
                                                                  const __grid_constant__ BlockMap m) {
  pdl_enter();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const long long blk = blockIdx.x;
  if (blk == 0 && threadIdx.x == 0 && p.b0 == 0 && p.counter != nullptr) {
    p.counter[0] = 0u; p.counter[1] = 0u;
    if (p.nonfinite != nullptr) p.nonfinite[1] = 0;
  }
  const int slot = find_layer(m, blk, p.n_layers);
  const int l = m.layer[slot];
  const LayerDev& L = p.L[l];
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);
  prep_layer(L, keys, p.rng_draw, p.rng_seed, p.rng_offset + 4ull * (unsigned long long)l, blk == m.start[slot]);
  __syncthreads();
  const long long w0 = ((blk - m.start[slot]) * 8 + (threadIdx.x >> 5)) * kNhwcItemsPerWarp;
  const long long nitems = (long long)(2 - p.side0) * p.bn * (L.Ppad >> 3) * L.nchunk;
  if (w0 >= nitems) return;
  const int lane = threadIdx.x & 31;
  if (p.dtype == PNCE_F32) gather_nhwc_warp<float>(L, p.b0, p.bn, w0, nitems, lane, p.side0, keys);
  else if (p.dtype == PNCE_F16) gather_nhwc_warp<__half>(L, p.b0, p.bn, w0, nitems, lane, p.side0, keys);
  else gather_nhwc_warp<__nv_bfloat16>(L, p.b0, p.bn, w0, nitems, lane, p.side0, keys);
}

// -------------------------------------------------------------------------------------------------
// Dense d tgt_feat for channels-last maps: (B, HW, C) storage, written exactly once, tiles in address order.
// A tile = TPOS whole positions (TPOS * C * sizeof(T) <= 8 KB).  The sorted slots that fall into a tile are a
// contiguous range [ja, jb) of sid[]; it is found from k_prep's 2048-position bucket table (one or two buckets,
// their candidates compared in parallel).  No sampled position in the tile: 128-bit zero stores straight from
// registers.  Otherwise the tile is staged in shared memory, thread <-> channel adds up the row-major gradient
// rows of every run of equal ids in sorted order (the order of the NCHW kernel), and the tile is copied out.
// (Measured and rejected: a store-first variant without shared memory -- zero stores issued before the slot lookup,
// slot range by warp ballots, hit rows overwritten after one barrier -- 460 vs 454 us at B=64.)
// -------------------------------------------------------------------------------------------------
struct DenseNhwcMap {
  long long start[PNCE_MAX_LAYERS + 1];        // tile prefix per layer
  int tiles[PNCE_MAX_LAYERS];                  // tiles per image
  int tpos[PNCE_MAX_LAYERS];                   // positions per tile
};

template <typename T, bool VEC>
__global__ void __launch_bounds__(128) k_dense_nhwc(const __grid_constant__ Params p,
                                                    const __grid_constant__ DenseNhwcMap m) {
  __shared__ __align__(16) T tile[kFlatBytes / sizeof(T)];
  __shared__ int s_ja, s_jb;
  const int tid = threadIdx.x;
  const long long item = blockIdx.x;
  int l = 0;
  for (int i = 1; i < p.n_layers; ++i)
    if (item >= m.start[i]) l = i;
  const LayerDev& L = p.L[l];
  const long long local = item - m.start[l];
  const int tiles = m.tiles[l], TPOS = m.tpos[l];
  const int b = (int)(local / tiles);
  const int t = (int)(local - (long long)b * tiles);
  const int P = L.P, HW = L.HW, C = L.C;
  const int pos0 = t * TPOS, npos = min(TPOS, HW - pos0), pos1 = pos0 + npos;
  const int ntab = (HW + kTilePos - 1) / kTilePos;
  const int jlo = __ldg(L.cslot + pos0 / kTilePos);
  const int jhi = __ldg(L.cslot + min((pos1 - 1) / kTilePos + 1, ntab));
  if (tid == 0) { s_ja = P; s_jb = 0; }
  __syncthreads();
  for (int j = jlo + tid; j < jhi; j += 128) {
    const int q = __ldg(L.sid + j);
    if (q >= pos0 && q < pos1) { atomicMin(&s_ja, j); atomicMax(&s_jb, j + 1); }
  }
  __syncthreads();
  const int ja = s_ja, jb = s_jb;
  const int nelem = npos * C;
  T* dst = reinterpret_cast<T*>(L.dtgt) + ((size_t)b * HW + pos0) * C;
  if (ja >= jb) {                                              // nothing sampled here: pure fill
    if (VEC) {
      const int n16 = (nelem * (int)sizeof(T)) >> 4;
      uint4* d4 = reinterpret_cast<uint4*>(dst);
      for (int i = tid; i < n16; i += 128) __stcs(d4 + i, make_uint4(0u, 0u, 0u, 0u));
    } else {
      for (int i = tid; i < nelem; i += 128) dst[i] = from_f32<T>(0.f);
    }
    return;
  }
  uint4* t4 = reinterpret_cast<uint4*>(tile);
  const int n16s = (nelem * (int)sizeof(T) + 15) >> 4;
  for (int i = tid; i < n16s; i += 128) t4[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  const float g = p.grad_out ? __ldg(p.grad_out) : 1.0f;
  const float* __restrict__ dx = L.dxT + ((size_t)b * L.dxpitch) * C;      // row-major rows [slot][C]
  for (int c = tid; c < C; c += 128) {
    int qprev = __ldg(L.sid + ja);
    float acc = 0.f;
    for (int j = ja; j < jb; ++j) {
      const int q = __ldg(L.sid + j);
      if (q != qprev) {
        tile[(size_t)(qprev - pos0) * C + c] = from_f32<T>(acc * g);
        acc = 0.f;
        qprev = q;
      }
      acc += __ldcs(dx + (size_t)j * C + c);                   // last use of the row: streaming
    }
    tile[(size_t)(qprev - pos0) * C + c] = from_f32<T>(acc * g);
  }
  __syncthreads();
  if (VEC) {
    const int n16 = (nelem * (int)sizeof(T)) >> 4;
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    for (int i = tid; i < n16; i += 128) __stcs(d4 + i, t4[i]);
  } else {
    for (int i = tid; i < nelem; i += 128) dst[i] = tile[i];
  }
}

// -------------------------------------------------------------------------------------------------
// TWO-LAUNCH variant (round 2): k_fill_zero writes every byte of every map as zeros -- no table, no ids, no shared
// memory: the bare fill, 0.38 ms for 3.2 GB on compressible gradient memory and 0.43 ms on plain memory (scratch/
// compbench.cu) -- and k_scatter_nhwc then overwrites the sampled positions, each a contiguous row of C values = whole
// 128-byte lines (50 MB at B = 64), summing the gradient rows of a run of equal ids in sorted order like k_dense_nhwc.
// A sampled line is written twice; everything else about the result is identical.  k_dense_nhwc pays ~30 us per step
// for looking up, for EVERY tile, whether it holds a sample (99 % do not).
// -------------------------------------------------------------------------------------------------
struct FillMap {
  long long start[PNCE_MAX_LAYERS + 1];        // 8 KB tile prefix per layer
  unsigned long long bytes[PNCE_MAX_LAYERS];   // bytes of the layer's gradient (a multiple of 16)
  void* base[PNCE_MAX_LAYERS];
  int n;
};
__global__ void __launch_bounds__(128) k_fill_zero(const __grid_constant__ FillMap m) {
  const unsigned item = blockIdx.x;
  int l = 0;
  for (int i = 1; i < m.n; ++i)
    if ((long long)item >= m.start[i]) l = i;
  const unsigned long long off = (unsigned long long)(item - (unsigned)m.start[l]) * kFlatBytes;
  const unsigned long long left = m.bytes[l] - off;
  const int n16 = (int)((left < (unsigned long long)kFlatBytes ? left : (unsigned long long)kFlatBytes) >> 4);
  uint4* dst = reinterpret_cast<uint4*>(static_cast<char*>(m.base[l]) + off);
#pragma unroll
  for (int k = 0; k < kFlatBytes / 16 / 128; ++k) {
    const int i = k * 128 + threadIdx.x;
    if (i < n16) __stcs(dst + i, make_uint4(0u, 0u, 0u, 0u));
  }
}

// four consecutive values of a gradient row: one 16-byte (fp32) or 8-byte (half precision) store
__device__ __forceinline__ void scatter_store4(float* d, float a, float b, float c, float e) {
  *reinterpret_cast<float4*>(d) = make_float4(a, b, c, e);
}
__device__ __forceinline__ void scatter_store4(__half* d, float a, float b, float c, float e) {
  const __half2 lo = __floats2half2_rn(a, b), hi = __floats2half2_rn(c, e);
  *reinterpret_cast<uint2*>(d) = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
}
__device__ __forceinline__ void scatter_store4(__nv_bfloat16* d, float a, float b, float c, float e) {
  const __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, e);
  *reinterpret_cast<uint2*>(d) = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
}

struct ScatterMap {
  long long start[PNCE_MAX_LAYERS + 1];        // warp-item prefix per layer: B * P items each
};
// warp <-> (layer, image, sorted slot); the head of a run of equal ids writes the position's C values
template <typename T>
__global__ void __launch_bounds__(256) k_scatter_nhwc(const __grid_constant__ Params p, const __grid_constant__ ScatterMap m) {
  const int lane = threadIdx.x & 31;
  const long long item = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (item >= m.start[p.n_layers]) return;
  int l = 0;
  for (int i = 1; i < p.n_layers; ++i)
    if (item >= m.start[i]) l = i;
  const LayerDev& L = p.L[l];
  const unsigned local = (unsigned)(item - m.start[l]);
  const int P = L.P, C = L.C, HW = L.HW;
  const int b = (int)(local / (unsigned)P), j = (int)(local - (unsigned)b * (unsigned)P);
  // the ids around the slot and the slot's own gradient row are loaded independently of each other (one round trip)
  const int q = __ldg(L.sid + j);
  const int prev = j > 0 ? __ldg(L.sid + j - 1) : -1;
  const int next = j + 1 < P ? __ldg(L.sid + j + 1) : -2;
  const float* __restrict__ dx = L.dxT + ((size_t)b * L.dxpitch + j) * C;      // row-major rows [slot][C]
  const int c0 = lane * 4, c1 = c0 + 128;                    // C <= 256 on the tensor-core path: at most two chunks per lane
  float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
  if (c0 < C) a0 = __ldcs(reinterpret_cast<const float4*>(dx + c0));            // last use of the row: streaming
  if (c1 < C) a1 = __ldcs(reinterpret_cast<const float4*>(dx + c1));
  if (prev == q) return;                                     // not the head of its run
  if (next == q) {                                           // rare: sum the run in sorted order
    for (int r = 1; j + r < P && __ldg(L.sid + j + r) == q; ++r) {
      if (c0 < C) { const float4 v = __ldcs(reinterpret_cast<const float4*>(dx + (size_t)r * C + c0)); a0.x += v.x; a0.y += v.y; a0.z += v.z; a0.w += v.w; }
      if (c1 < C) { const float4 v = __ldcs(reinterpret_cast<const float4*>(dx + (size_t)r * C + c1)); a1.x += v.x; a1.y += v.y; a1.z += v.z; a1.w += v.w; }
    }
  }
  const float g = p.grad_out ? __ldg(p.grad_out) : 1.0f;
  T* dst = reinterpret_cast<T*>(L.dtgt) + ((size_t)b * HW + q) * C;
  if (c0 < C) scatter_store4(dst + c0, a0.x * g, a0.y * g, a0.z * g, a0.w * g);
  if (c1 < C) scatter_store4(dst + c1, a1.x * g, a1.y * g, a1.z * g, a1.w * g);
}

}  // namespace pnce
