// Micro-benchmark: how fast can 3.2 GB be zero-filled on B200, by store path?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void lsu_persist(uint4* p, size_t n16) {
  const uint4 z = make_uint4(0, 0, 0, 0);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) p[i] = z;
}
__global__ void lsu_persist_cs(uint4* p, size_t n16) {
  const uint4 z = make_uint4(0, 0, 0, 0);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) __stcs(p + i, z);
}
// torch-like: each CTA writes a contiguous tile, 4 stores per thread
__global__ void lsu_flat(uint4* p, size_t n16) {
  const uint4 z = make_uint4(0, 0, 0, 0);
  size_t base = (size_t)blockIdx.x * blockDim.x * 4 + threadIdx.x;
#pragma unroll
  for (int k = 0; k < 4; ++k) { size_t i = base + (size_t)k * blockDim.x; if (i < n16) p[i] = z; }
}
// each CTA owns a contiguous range, written 16 KB at a time (like k_dense_tma's order), LSU stores
__global__ void lsu_range(uint4* p, size_t n16, size_t per_cta16) {
  const uint4 z = make_uint4(0, 0, 0, 0);
  size_t b = (size_t)blockIdx.x * per_cta16, e = b + per_cta16; if (e > n16) e = n16;
  for (size_t i = b + threadIdx.x; i < e; i += blockDim.x) p[i] = z;
}
template <int HINT>
__global__ void tma_range(unsigned char* p, size_t nbytes, size_t per_cta, int chunk, int depth_unused) {
  extern __shared__ __align__(128) unsigned char zeros[];
  for (int i = threadIdx.x; i < chunk / 16; i += blockDim.x) reinterpret_cast<uint4*>(zeros)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x != 0) return;
  size_t b = (size_t)blockIdx.x * per_cta, e = b + per_cta; if (e > nbytes) e = nbytes;
  uint64_t pol = 0;
  if (HINT == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  if (HINT == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  uint32_t s = (uint32_t)__cvta_generic_to_shared(zeros);
  for (size_t o = b; o < e; o += chunk) {
    uint32_t n = (uint32_t)((e - o) < (size_t)chunk ? (e - o) : chunk);
    if (HINT == 0) asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p + o), "r"(s), "r"(n) : "memory");
    else asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(p + o), "r"(s), "r"(n), "l"(pol) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 8;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
// strided across CTAs (item k of CTA c = chunk c + k*grid)
__global__ void tma_strided(unsigned char* p, size_t nbytes, int chunk) {
  extern __shared__ __align__(128) unsigned char zeros[];
  for (int i = threadIdx.x; i < chunk / 16; i += blockDim.x) reinterpret_cast<uint4*>(zeros)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x != 0) return;
  uint32_t s = (uint32_t)__cvta_generic_to_shared(zeros);
  for (size_t o = (size_t)blockIdx.x * chunk; o < nbytes; o += (size_t)gridDim.x * chunk) {
    uint32_t n = (uint32_t)((nbytes - o) < (size_t)chunk ? (nbytes - o) : chunk);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p + o), "r"(s), "r"(n) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 8;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <typename F> float timeit(F f, int n = 10) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) f();
  cudaEventRecord(a);
  for (int i = 0; i < n; ++i) f();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms / n * 1e3f;
}

int main() {
  const size_t nbytes = 3221225472ull;   // 64 images x 50.3 MB
  unsigned char* p; CK(cudaMalloc(&p, nbytes));
  const size_t n16 = nbytes / 16;
  auto rep = [&](const char* name, float us) { printf("%-40s %8.1f us  %7.0f GB/s\n", name, us, nbytes / us / 1e3); };
  rep("cudaMemsetAsync", timeit([&] { cudaMemsetAsync(p, 0, nbytes); }));
  for (int per_sm : {4, 8, 16}) {
    char nm[64];
    snprintf(nm, 64, "lsu_persist 256thr x %d/SM", per_sm);
    rep(nm, timeit([&] { lsu_persist<<<148 * per_sm, 256>>>((uint4*)p, n16); }));
    snprintf(nm, 64, "lsu_persist_cs 256thr x %d/SM", per_sm);
    rep(nm, timeit([&] { lsu_persist_cs<<<148 * per_sm, 256>>>((uint4*)p, n16); }));
  }
  rep("lsu_flat 128thr x4 (torch-like)", timeit([&] { lsu_flat<<<(unsigned)((n16 + 511) / 512), 128>>>((uint4*)p, n16); }));
  rep("lsu_flat 256thr x4", timeit([&] { lsu_flat<<<(unsigned)((n16 + 1023) / 1024), 256>>>((uint4*)p, n16); }));
  for (int per_sm : {4, 8}) {
    size_t grid = 148 * per_sm, per = ((n16 + grid - 1) / grid + 1023) / 1024 * 1024;
    char nm[64]; snprintf(nm, 64, "lsu_range 128thr x %d/SM", per_sm);
    rep(nm, timeit([&] { lsu_range<<<(unsigned)grid, 128>>>((uint4*)p, n16, per); }));
  }
  CK(cudaFuncSetAttribute(tma_range<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CK(cudaFuncSetAttribute(tma_range<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CK(cudaFuncSetAttribute(tma_range<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CK(cudaFuncSetAttribute(tma_strided, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  for (int chunk : {4096, 16384, 65536}) for (int per_sm : {1, 2, 4}) {
    size_t grid = 148 * per_sm, per = ((nbytes + grid - 1) / grid + chunk - 1) / chunk * chunk;
    char nm[64];
    snprintf(nm, 64, "tma_range chunk=%d x %d/SM", chunk, per_sm);
    rep(nm, timeit([&] { tma_range<0><<<(unsigned)grid, 32, chunk>>>(p, nbytes, per, chunk, 0); }));
    if (chunk == 16384) {
      snprintf(nm, 64, "tma_range evict_first chunk=%d x %d/SM", chunk, per_sm);
      rep(nm, timeit([&] { tma_range<1><<<(unsigned)grid, 32, chunk>>>(p, nbytes, per, chunk, 0); }));
      snprintf(nm, 64, "tma_range evict_last chunk=%d x %d/SM", chunk, per_sm);
      rep(nm, timeit([&] { tma_range<2><<<(unsigned)grid, 32, chunk>>>(p, nbytes, per, chunk, 0); }));
    }
    snprintf(nm, 64, "tma_strided chunk=%d x %d/SM", chunk, per_sm);
    rep(nm, timeit([&] { tma_strided<<<(unsigned)grid, 32, chunk>>>(p, nbytes, chunk); }));
  }
  CK(cudaDeviceSynchronize());
  CK(cudaGetLastError());
  return 0;
}
