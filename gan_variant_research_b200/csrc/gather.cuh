// Launch 1 of the forward: random-patch gather + L2 normalise for every (layer, side, image,
// 32-patch tile), plus one "prep" CTA per layer that sorts the ids and builds the position bitmap
// the dense backward needs.  Replaces patchnce_cut.py:56-78 (view/permute, index, stack, normalize).
#pragma once
#include <curand_kernel.h>

#include "common.cuh"

namespace pnce {

// --------------------------------------------------------------------------------------------
// prep CTA: ids -> (sid, perm, rank, cslot).  P <= PNCE_MAX_PATCHES.   smem: keys[N2] (u64).
// --------------------------------------------------------------------------------------------
// draw != 0: the ids are DRAWN here first, bit for bit what `torch.randint(0, HW, (P,), device='cuda')` returns
// (patchnce_cut.py:63) for a CUDA generator at (seed, offset): ATen's random_from_to kernel
// (native/cuda/DistributionTemplates.h: distribution_elementwise_grid_stride_kernel with a grid that covers all P
// elements in one sweep) gives element i the first 32-bit word of the Philox4x32-10 block of subsequence i at counter
// offset / 4 -- curand_init(seed, i, offset) + curand4().x -- reduced by `% range` (ATen/core/TransformationHelper.h
// uniform_int_from_to; range = HW < 2^32, base = 0).  One randint call advances the generator's offset by 4; the host
// does the same (patchnce.py).  The int64 ids are written to L.ids, the caller's buffer, so that PatchNCELoss /
// PatchSampleF can hand them out like the reference does.
// write = false: only the sorted keys in shared memory are produced (k_gather_tc_fold: every gather CTA of a small
// problem sorts its layer's ids itself instead of waiting for a k_prep launch; ONE CTA per layer writes the tables).
__device__ void prep_layer(const LayerDev& L, unsigned long long* keys, int draw = 0, unsigned long long seed = 0,
                           unsigned long long offset = 0, bool write = true) {
  const int tid = threadIdx.x;
  const int P = L.P;
  int N2 = 1;
  while (N2 < P) N2 <<= 1;
  for (int i = tid; i < N2; i += kThreads) {
    unsigned long long k = ~0ull;
    if (i < P) {
      long long id;
      if (draw) {
        curandStatePhilox4_32_10_t st;
        curand_init(seed, (unsigned long long)i, offset, &st);
        const uint4 r = curand4(&st);
        id = (long long)(r.x % (unsigned)L.HW);
        if (write) const_cast<long long*>(L.ids)[i] = id;
      } else {
        id = L.ids[i];
        id = id < 0 ? 0 : (id >= L.HW ? L.HW - 1 : id);        // memory safety only; randint is in range
      }
      k = ((unsigned long long)(unsigned)id << 32) | (unsigned)i;
    }
    keys[i] = k;
  }
  __syncthreads();
  // bitonic sort, ascending by (id, original index): deterministic order inside runs of equal ids
  for (int k = 2; k <= N2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < N2; i += kThreads) {
        int ixj = i ^ j;
        if (ixj > i) {
          unsigned long long a = keys[i], b = keys[ixj];
          bool up = (i & k) == 0;
          if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
  if (!write) return;
  // sid / perm / rank
  for (int i = tid; i < P; i += kThreads) {
    const int id = (int)(keys[i] >> 32);
    const int pi = (int)(keys[i] & 0xffffffffu);
    L.sid[i] = id;
    L.perm[i] = pi;
    L.rank[pi] = i;
  }
  // tile -> first sorted slot (the dense backward's per-tile slot range)
  if (L.cslot != nullptr) {
    const int ntile = (L.HW + kTilePos - 1) / kTilePos;
    for (int t = tid; t <= ntile; t += kThreads) {
      const int pos = t * kTilePos;
      int lo = 0, hi = P;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((int)(keys[mid] >> 32) < pos) lo = mid + 1; else hi = mid;
      }
      L.cslot[t] = lo;
    }
  }
}

// --------------------------------------------------------------------------------------------
// gather + normalise one tile of 32 patches of one image of one side.
//   lane <-> patch, warp w <-> channels w, w+8, ...  (each load instruction touches 32 sectors of
//   one channel plane; C/8 independent loads per lane are in flight)
//   tile staged in smem [32][C+1] so the normalised rows leave as coalesced 128 B row segments.
// --------------------------------------------------------------------------------------------
template <typename T>
__device__ void gather_tile(const LayerDev& L, int B, int side0, int raw, long long local, float* tile_s, float* red_s) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nt = L.ntiles;
  const int tile = (int)(local % nt);
  const long long rest = local / nt;
  const int b = (int)(rest % B);
  const int side = (int)(rest / B) + side0;                   // 0 = src (k), 1 = tgt (q)
  const int C = L.C, HW = L.HW, P = L.P;
  const T* feat = reinterpret_cast<const T*>(side ? L.tgt : L.src) + (size_t)b * C * HW;
  float* out = (side ? L.qn : L.kn) + (size_t)b * P * C;
  const int p = tile * kRowTile + lane;
  const bool ok = p < P;
  long long id = ok ? L.ids[p] : 0;
  id = id < 0 ? 0 : (id >= HW ? HW - 1 : id);
  const int ldt = C + 1;
  const T* col = feat + id;
  float ss = 0.f;
  int bad = 0;
  // batches of 16 independent sector loads per lane before the first use (8 left the memory system idle:
  // 590 us for the five CUT maps of both sides at B=64 against 310 us for the tensor-core gather)
  for (int c0 = warp; c0 < C; c0 += 128) {
    float v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int c = c0 + 8 * k;
      v[k] = (ok && c < C) ? to_f32<T>(__ldcg(col + (size_t)c * HW)) : 0.f;   // L2-only: no 128 B L1 line fill
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int c = c0 + 8 * k;
      if (c < C) {
        tile_s[lane * ldt + c] = v[k];
        ss = fmaf(v[k], v[k], ss);
        bad |= !isfinite(v[k]);
      }
    }
  }
  red_s[warp * 32 + lane] = ss;
  red_s[256 + warp * 32 + lane] = __int_as_float(bad);
  __syncthreads();
  if (warp == 0) {
    float tot = 0.f;
    int anybad = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      tot += red_s[w * 32 + lane];
      anybad |= __float_as_int(red_s[256 + w * 32 + lane]);
    }
    float nrm = sqrtf(tot);
    float den = fmaxf(nrm, kNormEps);                          // clamp_min(norm, eps)
    if (!(nrm == nrm)) den = nrm;                              // NaN norm stays NaN (fmaxf would hide it)
    red_s[512 + lane] = raw ? 1.0f : den;
    if (side == 1 && ok && L.qinv != nullptr) {
      float inv = (nrm >= kNormEps) ? 1.0f / nrm : -1.0f / kNormEps;
      if (anybad) inv = __int_as_float(0x7fc00000);
      L.qinv[(size_t)b * P + p] = inv;
    }
  }
  __syncthreads();
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    const int r = warp * 4 + rr;
    const int pr = tile * kRowTile + r;
    if (pr < P) {
      const float den = red_s[512 + r];
      float* orow = out + (size_t)pr * C;
      for (int c = lane; c < C; c += 32) orow[c] = tile_s[r * ldt + c] / den;
    }
  }
}

// grid = sum_l 2*B*ntiles_l gather CTAs, then n_layers prep CTAs.
// dynamic smem = max(32*(Cmax+1)*4 + 544*4, N2max*8 + 64)
__global__ void __launch_bounds__(kThreads) k_gather_prep(const __grid_constant__ Params p,
                                                          const __grid_constant__ BlockMap m) {
  pdl_enter();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const long long blk = blockIdx.x;
  const long long n_gather = m.start[p.n_layers];
  if (blk >= n_gather) {
    const int l = (int)(blk - n_gather);
    if (l == 0 && threadIdx.x == 0 && p.counter != nullptr) {
      p.counter[0] = 0u; p.counter[1] = 0u;
      if (p.nonfinite != nullptr) p.nonfinite[1] = 0;
    }
    if (p.L[l].sid == nullptr) return;
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);
    int N2 = 1;
    while (N2 < p.L[l].P) N2 <<= 1;
    prep_layer(p.L[l], keys);
    return;
  }
  const int l = find_layer(m, blk, p.n_layers);
  const LayerDev& L = p.L[l];
  float* tile_s = reinterpret_cast<float*>(smem_raw);
  float* red_s = tile_s + kRowTile * (L.C + 1);
  const long long local = blk - m.start[l];
  if (p.dtype == PNCE_F32) gather_tile<float>(L, p.B, p.side0, p.raw, local, tile_s, red_s);
  else if (p.dtype == PNCE_F16) gather_tile<__half>(L, p.B, p.side0, p.raw, local, tile_s, red_s);
  else gather_tile<__nv_bfloat16>(L, p.B, p.side0, p.raw, local, tile_s, red_s);
}

}  // namespace pnce
