// Module-split API on the tensor cores: PatchNCELoss(feat_q, feat_k) gets fp32 rows (B*P, D), grouped per
// image; this kernel re-tiles them into the operand blobs the tcgen05 loss kernel consumes (loss_tc.cuh:
// Q blob, K blob, key-major K2 blob, bf16 hi + lo), rows in the order given (no ids here, so no sorting).
// The rows are used AS GIVEN (patchnce_cut.py:83-94 on already normalised rows): the per-chunk "sums of
// squares" are written as 1 for chunk 0 and 0 elsewhere, which makes every norm the loss kernel derives
// exactly 1 (NaN where a row holds a non-finite value, as the gather does).
//   thread <-> (row p, 8-channel slab c8); a warp = 8 rows x 4 slabs: 128 B coalesced per row on the way
//   in, 128 B coalesced per slab on the way out.
#pragma once
#include "common.cuh"
#include "gather_tc.cuh"

namespace pnce {

// grid = (ceil(2 * B * Ppad * Cp8 / 256) of the largest layer, n_layers); side 0 = k, 1 = q
__global__ void __launch_bounds__(kThreads) k_rows_pack(const __grid_constant__ Params p) {
  pdl_enter();
  const LayerDev& L = p.L[blockIdx.y];
  const int Ppad = L.Ppad, P = L.P, C = L.C, Cp8 = L.Cp >> 3, nchunk = L.nchunk;
  const long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
  const int r8 = (int)(t & 7), s4 = (int)((t >> 3) & 3);
  long long g = t >> 5;
  const int sb = (int)(g % (Cp8 >> 2)); g /= (Cp8 >> 2);
  const int rg = (int)(g % (Ppad >> 3)); g /= (Ppad >> 3);
  const int b = (int)(g % p.B);
  const int side = (int)(g / p.B);
  if (side > 1) return;                                        // whole warps: the grid is rounded up
  const int pr = rg * 8 + r8, c8 = sb * 4 + s4, c0 = c8 * 8;
  const float* rows = side ? L.qn : L.kn;
  float v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = 0.f;
  if (pr < P) {
    const float* src = rows + ((size_t)b * P + pr) * C + c0;
    if (c0 + 8 <= C && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0)) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(src)), c = __ldg(reinterpret_cast<const float4*>(src) + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (c0 + k < C) v[k] = __ldg(src + k);
    }
  }
  bool bad = false;
#pragma unroll
  for (int k = 0; k < 8; ++k) bad |= !isfinite(v[k]);
  // the four slabs of a 32-channel chunk sit in lanes r8 + 8 * {0,1,2,3}
  bad = __any_sync(0xffffffffu, bad) && ((__ballot_sync(0xffffffffu, bad) & (0x01010101u << r8)) != 0u);
  if (s4 == 0) {
    float* ss = side ? L.qss : L.kss;
    ss[((size_t)b * nchunk + sb) * Ppad + pr] = bad ? __int_as_float(0x7fc00000) : ((sb == 0 && pr < P) ? 1.0f : 0.f);
  }
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
  const uint4 hi = make_uint4(hw[0], hw[1], hw[2], hw[3]), lo = make_uint4(lw[0], lw[1], lw[2], lw[3]);
  if (side) {
    const size_t off = ((((size_t)b * (Ppad >> 7) + (pr >> 7)) * Cp8 + c8) * 16 + ((pr & 127) >> 3)) * 64 + (size_t)(pr & 7) * 8;
    *reinterpret_cast<uint4*>(L.qhi + off) = hi;
    if (L.qlo != nullptr) *reinterpret_cast<uint4*>(L.qlo + off) = lo;
  } else {                                                     // one key block (Ppad <= 256)
    const size_t off = (((size_t)b * Ppad * Cp8) / 8 + (size_t)c8 * (Ppad >> 3) + (pr >> 3)) * 64 + (size_t)(pr & 7) * 8;
    const size_t off2 = (((size_t)b * (Ppad >> 3) + (pr >> 3)) * Cp8 + c8) * 64 + (size_t)(pr & 7) * 8;
    *reinterpret_cast<uint4*>(L.khi + off) = hi;
    *reinterpret_cast<uint4*>(L.k2hi + off2) = hi;
    if (L.klo != nullptr) {
      *reinterpret_cast<uint4*>(L.klo + off) = lo;
      *reinterpret_cast<uint4*>(L.k2lo + off2) = lo;
    }
  }
}

}  // namespace pnce
