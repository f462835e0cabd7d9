// Multi-tensor elementwise kernels for the optimiser-side "next" row: the
// reference walks its parameters in a Python loop and launches ~3 kernels + 2 allocations per tensor
// (utils/io_ckpt.py:23-29, EMA.update); here ONE launch covers every tensor through a device-resident table
// built once (parameter storage does not move during training).  HBM-bound: 12 bytes per element.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pnce {

constexpr int kMtChunk = 8192;             // elements per CTA chunk (32 KB of fp32)
constexpr int kMtThreads = 256;

// table layout (all device memory, caller-owned):
//   dst[t], src[t]  : tensor base pointers      chunk_tensor[c], chunk_start[c] : chunk -> (tensor, first element)
//   numel[t]        : elements of tensor t
// mode 0: dst = a * src + b * dst, each product and the sum rounded to fp32 separately (bit-identical to the
//         reference's (1 - decay) * param + decay * shadow, which ATen evaluates as two multiplies and an add);
// mode 1: dst = src (copy).
__global__ void __launch_bounds__(kMtThreads) k_multi_axpby(float* const* __restrict__ dst,
                                                            const float* const* __restrict__ src,
                                                            const long long* __restrict__ numel,
                                                            const int* __restrict__ chunk_tensor,
                                                            const long long* __restrict__ chunk_start, float a,
                                                            float b, int mode) {
  pdl_enter();
  const int t = chunk_tensor[blockIdx.x];
  const long long e0 = chunk_start[blockIdx.x];
  const long long n = numel[t];
  float* __restrict__ d = dst[t];
  const float* __restrict__ s = src[t];
  const long long e1 = (e0 + kMtChunk < n) ? e0 + kMtChunk : n;
  const bool vec = ((reinterpret_cast<uintptr_t>(d) | reinterpret_cast<uintptr_t>(s)) & 15u) == 0 && (e0 & 3) == 0;
  if (vec) {
    const long long nv = (e1 - e0) >> 2;
    float4* d4 = reinterpret_cast<float4*>(d + e0);
    const float4* s4 = reinterpret_cast<const float4*>(s + e0);
    for (long long i = threadIdx.x; i < nv; i += kMtThreads) {
      const float4 x = __ldcs(s4 + i);
      float4 y;
      if (mode == 1) y = x;
      else {
        const float4 o = d4[i];
        y.x = __fadd_rn(__fmul_rn(a, x.x), __fmul_rn(b, o.x));
        y.y = __fadd_rn(__fmul_rn(a, x.y), __fmul_rn(b, o.y));
        y.z = __fadd_rn(__fmul_rn(a, x.z), __fmul_rn(b, o.z));
        y.w = __fadd_rn(__fmul_rn(a, x.w), __fmul_rn(b, o.w));
      }
      d4[i] = y;
    }
    for (long long i = e0 + (nv << 2) + threadIdx.x; i < e1; i += kMtThreads)
      d[i] = (mode == 1) ? s[i] : __fadd_rn(__fmul_rn(a, s[i]), __fmul_rn(b, d[i]));
  } else {
    for (long long i = e0 + threadIdx.x; i < e1; i += kMtThreads)
      d[i] = (mode == 1) ? s[i] : __fadd_rn(__fmul_rn(a, s[i]), __fmul_rn(b, d[i]));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// AMP optimiser step: GradScaler.unscale_ + clip_grad_norm_ + Adam.step + GradScaler.update -- the reference's
// AMPContext.step_optimizer (utils/amp_utils.py:29-41) on an optim.Adam (training/sched_optim.py:20-25) -- in three
// launches whatever the number of parameter tensors, and without the host sync GradScaler.step() needs to decide
// whether to skip (found_inf.item()): the decision, the per-tensor step counters and the loss-scale update stay on
// the device.  The arithmetic follows ATen's foreach Adam operation for operation (each comment names the ATen op),
// so that with an inactive clip (coefficient clamped to 1) every value is bit-identical to torch's.
//
// scratch (floats, zeroed once by the caller): [0] found_inf (written as 1.0f by any chunk that saw a non-finite
// gradient), [1] total norm (diagnostic, written by k_amp_finish), [2 .. 2 + n_chunks) per-chunk sums of squares.
struct AmpAdamTable {
  float* const* param; float* const* grad; float* const* m; float* const* v; float* const* step;   // per tensor
  const long long* numel; const int* chunk_tensor; const long long* chunk_start;                   // per tensor / chunk
  int n_chunks, n_tensors;
  float* scale; int* growth_tracker;       // GradScaler._scale / ._growth_tracker (device scalars) or NULL (no scaler)
  float growth_factor, backoff_factor; int growth_interval;
  float max_norm;                          // < 0: no clipping
  double lr, beta1, beta2, eps, weight_decay;
  float* scratch;
};

__device__ __forceinline__ float amp_inv_scale(const float* scale) {
  // GradScaler._unscale_grads_: inv_scale = scale.double().reciprocal().float()
  return scale ? static_cast<float>(1.0 / static_cast<double>(*scale)) : 1.0f;
}

// pass 1: non-finite check on the raw gradients (_amp_foreach_non_finite_check_and_unscale_) and the sum of squares of
// the unscaled ones, one partial per chunk (reduced in a fixed order by pass 2: deterministic).
__global__ void __launch_bounds__(kMtThreads) k_amp_gradnorm(AmpAdamTable a) {
  pdl_enter();
  const int t = a.chunk_tensor[blockIdx.x];
  const long long e0 = a.chunk_start[blockIdx.x], n = a.numel[t];
  const long long e1 = (e0 + kMtChunk < n) ? e0 + kMtChunk : n;
  const float* __restrict__ g = a.grad[t];
  const float inv = amp_inv_scale(a.scale);
  float ss = 0.f; bool bad = false;
  if ((reinterpret_cast<uintptr_t>(g) & 15u) == 0) {      // chunk starts are multiples of 4 elements
    const long long nv = (e1 - e0) >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g + e0);
    for (long long i = threadIdx.x; i < nv; i += kMtThreads) {
      const float4 x = g4[i];
      bad |= !(isfinite(x.x) && isfinite(x.y) && isfinite(x.z) && isfinite(x.w));
      const float u0 = x.x * inv, u1 = x.y * inv, u2 = x.z * inv, u3 = x.w * inv;
      ss = fmaf(u0, u0, ss); ss = fmaf(u1, u1, ss); ss = fmaf(u2, u2, ss); ss = fmaf(u3, u3, ss);
    }
    for (long long i = e0 + (nv << 2) + threadIdx.x; i < e1; i += kMtThreads) {
      const float x = g[i]; bad |= !isfinite(x); const float u = x * inv; ss = fmaf(u, u, ss);
    }
  } else {
    for (long long i = e0 + threadIdx.x; i < e1; i += kMtThreads) {
      const float x = g[i]; bad |= !isfinite(x); const float u = x * inv; ss = fmaf(u, u, ss);
    }
  }
  __shared__ float red[kMtThreads / 32];
  for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  const int any_bad = __syncthreads_or(bad);
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < kMtThreads / 32; ++w) tot += red[w];
    a.scratch[2 + blockIdx.x] = tot;
    if (any_bad) a.scratch[0] = 1.0f;
  }
}

// fixed-order block reduction of the per-chunk partials: every CTA gets the same bits
__device__ __forceinline__ float amp_total_sumsq(const float* part, int n, float* red) {
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += kMtThreads) acc += static_cast<double>(part[i]);
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  double* dred = reinterpret_cast<double*>(red);
  if ((threadIdx.x & 31) == 0) dred[threadIdx.x >> 5] = acc;
  __syncthreads();
  double tot = 0.0;
  for (int w = 0; w < kMtThreads / 32; ++w) tot += dred[w];
  return static_cast<float>(sqrt(tot));
}

__device__ __forceinline__ void amp_adam_elem(float& p, float& g, float& m, float& v, float inv, float coef, bool clip,
                                              bool skip, float wd, float w1, float b2, float w2, float bc2s,
                                              float eps, float neg_step) {
  g = g * inv;                                   // unscale_: grad.mul_(inv_scale)
  if (clip) g = g * coef;                        // clip_grad_norm_: _foreach_mul_(grads, clip_coef_clamped)
  if (skip) return;                              // GradScaler.step: optimizer.step() skipped on found_inf
  float gg = g;
  if (wd != 0.f) gg = fmaf(wd, p, gg);           // _foreach_add(grads, params, alpha=weight_decay)
  const float diff = gg - m;                     // _foreach_lerp_(exp_avgs, grads, 1 - beta1): ATen lerp()
  m = (fabsf(w1) < 0.5f) ? fmaf(w1, diff, m) : fmaf(-diff, 1.0f - w1, gg);
  const float vv = __fmul_rn(v, b2);             // _foreach_mul_(exp_avg_sqs, beta2)
  v = fmaf(w2, __fmul_rn(gg, gg), vv);           // _foreach_addcmul_(exp_avg_sqs, grads, grads, 1 - beta2)
  const float den = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), bc2s), eps);   // _foreach_sqrt, _foreach_div_, _foreach_add_
  p = fmaf(neg_step, __fdiv_rn(m, den), p);      // _foreach_addcdiv_(params, exp_avgs, denom, -lr / bias_correction1)
}

// pass 2: unscale, clip, Adam.  Gradients are written back unscaled and clipped, as the reference leaves them.
__global__ void __launch_bounds__(kMtThreads) k_amp_adam(AmpAdamTable a) {
  pdl_enter();
  __shared__ double redbuf[kMtThreads / 32];
  const int t = a.chunk_tensor[blockIdx.x];
  const long long e0 = a.chunk_start[blockIdx.x], n = a.numel[t];
  const long long e1 = (e0 + kMtChunk < n) ? e0 + kMtChunk : n;
  const float inv = amp_inv_scale(a.scale);
  const bool clip = a.max_norm >= 0.f;
  float coef = 1.0f;
  if (clip) {
    const float total = amp_total_sumsq(a.scratch + 2, a.n_chunks, reinterpret_cast<float*>(redbuf));
    const float raw = a.max_norm / (total + 1e-6f);           // clip_coef = max_norm / (total_norm + 1e-6)
    coef = (raw != raw) ? raw : fminf(raw, 1.0f);             // clamp(max=1.0); a NaN norm stays NaN, as torch.clamp keeps it
  }
  const bool skip = a.scale != nullptr && a.scratch[0] != 0.f;
  // bias corrections in double from the step counter, as the Python of torch.optim.adam computes them
  const double step = static_cast<double>(*a.step[t]) + 1.0;
  const double bc1 = 1.0 - pow(a.beta1, step), bc2 = 1.0 - pow(a.beta2, step);
  const float neg_step = static_cast<float>((a.lr / bc1) * -1.0);
  const float bc2s = static_cast<float>(sqrt(bc2));
  const float w1 = static_cast<float>(1.0 - a.beta1), b2 = static_cast<float>(a.beta2),
              w2 = static_cast<float>(1.0 - a.beta2), eps = static_cast<float>(a.eps),
              wd = static_cast<float>(a.weight_decay);
  float* __restrict__ P = a.param[t]; float* __restrict__ G = a.grad[t];
  float* __restrict__ M = a.m[t]; float* __restrict__ V = a.v[t];
  const bool vec = ((reinterpret_cast<uintptr_t>(P) | reinterpret_cast<uintptr_t>(G) | reinterpret_cast<uintptr_t>(M) |
                     reinterpret_cast<uintptr_t>(V)) & 15u) == 0;
  long long done = e0;
  if (vec) {
    const long long nv = (e1 - e0) >> 2;
    float4 *p4 = reinterpret_cast<float4*>(P + e0), *g4 = reinterpret_cast<float4*>(G + e0),
           *m4 = reinterpret_cast<float4*>(M + e0), *v4 = reinterpret_cast<float4*>(V + e0);
    for (long long i = threadIdx.x; i < nv; i += kMtThreads) {
      float4 p = p4[i], g = g4[i], m = m4[i], v = v4[i];
      amp_adam_elem(p.x, g.x, m.x, v.x, inv, coef, clip, skip, wd, w1, b2, w2, bc2s, eps, neg_step);
      amp_adam_elem(p.y, g.y, m.y, v.y, inv, coef, clip, skip, wd, w1, b2, w2, bc2s, eps, neg_step);
      amp_adam_elem(p.z, g.z, m.z, v.z, inv, coef, clip, skip, wd, w1, b2, w2, bc2s, eps, neg_step);
      amp_adam_elem(p.w, g.w, m.w, v.w, inv, coef, clip, skip, wd, w1, b2, w2, bc2s, eps, neg_step);
      g4[i] = g;
      if (!skip) { p4[i] = p; m4[i] = m; v4[i] = v; }
    }
    done = e0 + (nv << 2);
  }
  for (long long i = done + threadIdx.x; i < e1; i += kMtThreads) {
    float p = P[i], g = G[i], m = M[i], v = V[i];
    amp_adam_elem(p, g, m, v, inv, coef, clip, skip, wd, w1, b2, w2, bc2s, eps, neg_step);
    G[i] = g;
    if (!skip) { P[i] = p; M[i] = m; V[i] = v; }
  }
}

// pass 3 (one CTA): step counters, GradScaler.update() (_amp_update_scale_), scratch reset for the next call.
__global__ void __launch_bounds__(kMtThreads) k_amp_finish(AmpAdamTable a) {
  pdl_enter();
  __shared__ double redbuf[kMtThreads / 32];
  const float total = amp_total_sumsq(a.scratch + 2, a.n_chunks, reinterpret_cast<float*>(redbuf));
  const bool found = a.scratch[0] != 0.f;
  const bool skip = a.scale != nullptr && found;
  if (!skip)
    for (int t = threadIdx.x; t < a.n_tensors; t += kMtThreads) *a.step[t] += 1.0f;
  __syncthreads();
  if (threadIdx.x == 0) {
    a.scratch[1] = total;
    a.scratch[0] = 0.f;
    if (a.scale && a.growth_tracker) {
      if (found) { *a.scale = *a.scale * a.backoff_factor; *a.growth_tracker = 0; }
      else {
        const int ok = *a.growth_tracker + 1;
        if (ok == a.growth_interval) {
          const float ns = *a.scale * a.growth_factor;
          if (isfinite(ns)) *a.scale = ns;
          *a.growth_tracker = 0;
        } else *a.growth_tracker = ok;
      }
    }
  }
}

}  // namespace pnce
