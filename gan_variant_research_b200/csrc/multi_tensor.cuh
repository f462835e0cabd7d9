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

}  // namespace pnce
