// D-side elementwise work of the training step: DiffAugment and the hinge losses.
// The reference builds both out of dozens of small ATen launches per call (training/diffaugment.py:6-60 -- three
// (B,H,W) int64 index grids, a padded NHWC copy and an advanced-indexing gather for the translation alone;
// losses/adv_hinge.py:6-62).  The images are 3 x 256 x 256: everything here is launch-bound, so the job is to make
// each call ONE or TWO launches.  The random draws stay in the Python host (the same torch.rand / torch.randint calls
// in the same order: bit-identical parameters, RNG stream aligned); the kernels take them as device arrays.
#pragma once
#include "common.cuh"

namespace pnce {

constexpr int kAugMaxC = 8;
constexpr int kAugThreads = 256;

// Per-call description; every pointer may be NULL when its stage is not in the policy.
struct AugParams {
  const void* x; void* y;                 // forward: input / output images (B, C, H, W); backward: d y / d x
  int dtype, B, C, H, W;
  const void* rb; const void* rs; const void* rc;      // color: rand(B) in the image dtype (brightness, saturation, contrast)
  const long long* tx; const long long* ty;            // translation: randint(-shift, shift + 1, (B,)) along H and W
  const long long* ox; const long long* oy;            // cutout: box centre draws along H and W
  int cut_h, cut_w;                                    // cutout size (0 = no cutout)
  float* part;                                         // [B][nblk] partial sums of pass 1
  int nblk;
};

template <typename T> __device__ __forceinline__ float aug_ld(const void* p, long long i) {
  return to_f32<T>(reinterpret_cast<const T*>(p)[i]);
}

__device__ __forceinline__ bool aug_cut(const AugParams& a, int b, int i, int j) {
  if (a.cut_h <= 0) return false;
  // mask[grid_batch, clamp(k + ox - ch/2, 0, H-1), clamp(k' + oy - cw/2, 0, W-1)] = 0     diffaugment.py:45-57
  const int x0 = (int)a.ox[b] - a.cut_h / 2, y0 = (int)a.oy[b] - a.cut_w / 2;
  const int xl = max(x0, 0), xh = min(x0 + a.cut_h - 1, a.H - 1);
  const int yl = max(y0, 0), yh = min(y0 + a.cut_w - 1, a.W - 1);
  return i >= xl && i <= xh && j >= yl && j <= yh;
}

// Pass 1 (color only).  Forward: per-image sum of the image after brightness and saturation (contrast's mean,
// diffaugment.py:20).  Backward: per-image sum of the upstream gradient over the output pixels that read a source
// pixel (not cut out, translation source inside the image).  grid = (nblk, B); fixed-order partials: deterministic.
template <typename T, bool BWD>
__global__ void __launch_bounds__(kAugThreads) k_aug_reduce(const __grid_constant__ AugParams a) {
  pdl_enter();
  const int b = blockIdx.y, HW = a.H * a.W, C = a.C;
  const long long img = (long long)b * C * HW;
  float acc = 0.f;
  const float rb = a.rb ? aug_ld<T>(a.rb, b) - 0.5f : 0.f;
  const float ss = a.rs ? aug_ld<T>(a.rs, b) * 2.0f : 1.0f;
  const int tx = a.tx ? (int)a.tx[b] : 0, ty = a.ty ? (int)a.ty[b] : 0;
  for (int pix = blockIdx.x * kAugThreads + threadIdx.x; pix < HW; pix += a.nblk * kAugThreads) {
    if (BWD) {
      const int i = pix / a.W, j = pix - i * a.W;
      const int u = i + tx, v = j + ty;
      if (u < 0 || u >= a.H || v < 0 || v >= a.W || aug_cut(a, b, i, j)) continue;
      _Pragma("unroll") for (int c = 0; c < kAugMaxC; ++c) if (c < C) acc += aug_ld<T>(a.x, img + (long long)c * HW + pix);
    } else {
      float x1[kAugMaxC], m = 0.f;
      _Pragma("unroll") for (int c = 0; c < kAugMaxC; ++c) if (c < C) { x1[c] = aug_ld<T>(a.x, img + (long long)c * HW + pix) + rb; m += x1[c]; }
      m /= (float)C;
      _Pragma("unroll") for (int c = 0; c < kAugMaxC; ++c) if (c < C) acc += (x1[c] - m) * ss + m;
    }
  }
  __shared__ float red[kAugThreads / 32];
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < kAugThreads / 32; ++w) t += red[w];
    a.part[(size_t)b * a.nblk + blockIdx.x] = t;
  }
}

// Sum of image b's pass-1 partials, once per CTA (warp 0, fixed shuffle tree: the same bits in every CTA), broadcast
// through shared memory.  Called by all threads of the CTA before anything returns.
__device__ __forceinline__ float aug_image_sum(const AugParams& a, int b) {
  __shared__ float s_tot;
  if (threadIdx.x < 32) {
    float v = 0.f;
    for (int k = threadIdx.x; k < a.nblk; k += 32) v += a.part[(size_t)b * a.nblk + k];
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) s_tot = v;
  }
  __syncthreads();
  return s_tot;
}

// Pass 2, forward: one thread per OUTPUT pixel, all channels.  y[b,:,i,j] = cutout(i,j) ? 0 : color(x[b,:,i+tx,j+ty])
// (zero when the source lies outside the image: the reference gathers from a zero-padded copy, :33-34).
template <typename T>
__global__ void __launch_bounds__(kAugThreads) k_aug_fwd(const __grid_constant__ AugParams a) {
  pdl_enter();
  const int b = blockIdx.y, HW = a.H * a.W, C = a.C;
  const int pix = blockIdx.x * kAugThreads + threadIdx.x;
  const long long img = (long long)b * C * HW;
  const int i = pix / a.W, j = pix - i * a.W;
  const int u = i + (a.tx ? (int)a.tx[b] : 0), v = j + (a.ty ? (int)a.ty[b] : 0);
  const bool live = pix < HW && u >= 0 && u < a.H && v >= 0 && v < a.W && !aug_cut(a, b, i, j);
  const int sp = u * a.W + v;
  float val[kAugMaxC];
  _Pragma("unroll") for (int c = 0; c < kAugMaxC; ++c) if (c < C) val[c] = live ? aug_ld<T>(a.x, img + (long long)c * HW + sp) : 0.f;
  const float tot = a.rb ? aug_image_sum(a, b) : 0.f;             // the loads above are in flight across its barrier
  if (pix >= HW) return;
  T* y = reinterpret_cast<T*>(a.y);
  if (!live) {
    _Pragma("unroll") for (int c = 0; c < kAugMaxC; ++c) if (c < C) y[img + (long long)c * HW + pix] = from_f32<T>(0.f);
    return;
  }
  if (a.rb) {                                                         // 'color' = brightness, saturation, contrast :6-23
    const float rb = aug_ld<T>(a.rb, b) - 0.5f, ss = aug_ld<T>(a.rs, b) * 2.0f, sc = aug_ld<T>(a.rc, b) + 0.5f;
    const float mu = tot / (float)((long long)C * HW);
    float m = 0.f;
    _Pragma("unroll") for (int c = 0; c < kAugMaxC; ++c) if (c < C) { val[c] += rb; m += val[c]; }
    m /= (float)C;
    _Pragma("unroll") for (int c = 0; c < kAugMaxC; ++c) if (c < C) val[c] = (((val[c] - m) * ss + m) - mu) * sc + mu;
  }
  _Pragma("unroll") for (int c = 0; c < kAugMaxC; ++c) if (c < C) y[img + (long long)c * HW + pix] = from_f32<T>(val[c]);
}

// Pass 2, backward: one thread per INPUT pixel (u,v); the only output pixel that read it is (u - tx, v - ty).
//   g_t = upstream there (0 if outside / cut out);  contrast: s_c g_t + (1 - s_c) mean_chw(g_t);
//   saturation: s_s d + (1 - s_s) mean_c(d);  brightness: identity.
template <typename T>
__global__ void __launch_bounds__(kAugThreads) k_aug_bwd(const __grid_constant__ AugParams a) {
  pdl_enter();
  const int b = blockIdx.y, HW = a.H * a.W, C = a.C;
  const int pix = blockIdx.x * kAugThreads + threadIdx.x;
  const long long img = (long long)b * C * HW;
  const int u = pix / a.W, v = pix - u * a.W;
  const int i = u - (a.tx ? (int)a.tx[b] : 0), j = v - (a.ty ? (int)a.ty[b] : 0);
  const bool live = pix < HW && i >= 0 && i < a.H && j >= 0 && j < a.W && !aug_cut(a, b, i, j);
  float d[kAugMaxC];
  _Pragma("unroll") for (int c = 0; c < kAugMaxC; ++c) if (c < C) d[c] = live ? aug_ld<T>(a.x, img + (long long)c * HW + i * a.W + j) : 0.f;
  const float tot = a.rb ? aug_image_sum(a, b) : 0.f;             // the loads above are in flight across its barrier
  if (pix >= HW) return;
  if (a.rb) {
    const float ss = aug_ld<T>(a.rs, b) * 2.0f, sc = aug_ld<T>(a.rc, b) + 0.5f;
    const float kappa = (1.0f - sc) * tot / (float)((long long)C * HW);
    float m = 0.f;
    _Pragma("unroll") for (int c = 0; c < kAugMaxC; ++c) if (c < C) { d[c] = sc * d[c] + kappa; m += d[c]; }
    m /= (float)C;
    _Pragma("unroll") for (int c = 0; c < kAugMaxC; ++c) if (c < C) d[c] = ss * d[c] + (1.0f - ss) * m;
  }
  T* y = reinterpret_cast<T*>(a.y);
  _Pragma("unroll") for (int c = 0; c < kAugMaxC; ++c) if (c < C) y[img + (long long)c * HW + pix] = from_f32<T>(d[c]);
}

// ---------------------------------------------------------------------------------------------------------------
// Hinge losses (losses/adv_hinge.py:6-62) over the list of discriminator outputs, one launch per direction.
//   mode 0, discriminator: loss = (1/S) sum_s 0.5 * (mean relu(1 - real_s) + mean relu(1 + fake_s))      :20-29
//   mode 1, generator:     loss = (1/S) sum_s -mean(fake_s)                                              :47-53
// One CTA; the tensors are PatchGAN maps of a few thousand values.  Backward: elementwise, same table.
constexpr int kHingeMaxScales = 8;
struct HingeParams {
  const void* real[kHingeMaxScales]; const void* fake[kHingeMaxScales];
  void* dreal[kHingeMaxScales]; void* dfake[kHingeMaxScales];
  long long n[kHingeMaxScales];
  int scales, mode, dtype;
  float* loss;                    // forward: [1]
  const float* grad_out;          // backward: upstream scalar
};

template <typename T>
__global__ void __launch_bounds__(kAugThreads) k_hinge_fwd(const __grid_constant__ HingeParams h) {
  pdl_enter();
  __shared__ float red[kAugThreads / 32];
  float total = 0.f;                                                  // thread 0 only
  for (int s = 0; s < h.scales; ++s) {
    float acc = 0.f;
    const long long n = h.n[s];
    for (long long i0 = threadIdx.x; i0 < n; i0 += 8 * kAugThreads) {   // 8 (x2) independent loads in flight per thread
      float f[8], r[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const long long i = i0 + (long long)k * kAugThreads;
        f[k] = i < n ? aug_ld<T>(h.fake[s], i) : (h.mode == 0 ? -1.0f : 0.f);       // neutral elements of the two sums
        r[k] = (h.mode == 0 && i < n) ? aug_ld<T>(h.real[s], i) : 1.0f;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (h.mode == 0) {
          // torch.relu propagates NaN (fmaxf would return the other operand and hide a diverged discriminator
          // from train_step's non-finite check, train_cutpp.py:326-329)
          const float vr = 1.0f - r[k], vf = 1.0f + f[k];
          acc += ((vr > 0.f || vr != vr) ? vr : 0.f) + ((vf > 0.f || vf != vf) ? vf : 0.f);
        } else acc -= f[k];
      }
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < kAugThreads / 32; ++w) t += red[w];
      total += (h.mode == 0 ? 0.5f : 1.0f) * t / (float)h.n[s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *h.loss = total / (float)h.scales;
}

template <typename T>
__global__ void __launch_bounds__(kAugThreads) k_hinge_bwd(const __grid_constant__ HingeParams h) {
  pdl_enter();
  const int s = blockIdx.y;
  const float g = *h.grad_out / (float)h.scales / (float)h.n[s];
  for (long long i = (long long)blockIdx.x * kAugThreads + threadIdx.x; i < h.n[s]; i += (long long)gridDim.x * kAugThreads) {
    if (h.mode == 0) {
      if (h.dreal[s]) reinterpret_cast<T*>(h.dreal[s])[i] = from_f32<T>(!(aug_ld<T>(h.real[s], i) >= 1.0f) ? -0.5f * g : 0.f);   // relu backward passes the gradient at NaN
      if (h.dfake[s]) reinterpret_cast<T*>(h.dfake[s])[i] = from_f32<T>(!(aug_ld<T>(h.fake[s], i) <= -1.0f) ? 0.5f * g : 0.f);
    } else if (h.dfake[s]) {
      reinterpret_cast<T*>(h.dfake[s])[i] = from_f32<T>(-g);
    }
  }
}

}  // namespace pnce
