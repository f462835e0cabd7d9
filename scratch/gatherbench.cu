// Micro-benchmark: DRAM cost of random 4-byte reads (one per 32 B sector) by load flavour and by
// the cudaLimitMaxL2FetchGranularity hint.  Mimics k_gather_tc: thread <-> random position, 32 channel
// planes 256 KB apart.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int MODE>
__device__ __forceinline__ float ld(const float* p) {
  float v;
  if (MODE == 0) v = __ldcg(p);
  else if (MODE == 1) v = __ldg(p);
  else if (MODE == 2) v = __ldcv(p);
  else if (MODE == 3) asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  else if (MODE == 4) asm volatile("ld.global.cg.L2::64B.f32 %0, [%1];" : "=f"(v) : "l"(p));
  else if (MODE == 5) asm volatile("ld.global.cs.f32 %0, [%1];" : "=f"(v) : "l"(p));
  else { uint64_t pol; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
         asm volatile("ld.global.cg.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol)); }
  return v;
}

// grid: nplanes/32 * nimg CTAs of 256 threads; positions random in [0, HW)
template <int MODE>
__global__ void gather(const float* __restrict__ base, const int* __restrict__ ids, int HW, int C, float* out) {
  const int p = threadIdx.x;
  const int chunk = blockIdx.x % (C / 32), img = blockIdx.x / (C / 32);
  const float* col = base + ((size_t)img * C + chunk * 32) * HW + ids[p];
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < 32; ++k) acc += ld<MODE>(col + (size_t)k * HW);
  if (acc == 12345.678f) out[0] = acc;
}

// MODE 7: every thread fetches its 32 elements with 16-byte cp.async.bulk copies (TMA engine) into
// shared memory: does the L2 still fill whole 128 B lines for a TMA-side miss?
__global__ void gather_bulk(const float* __restrict__ base, const int* __restrict__ ids, int HW, int C, float* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ uint64_t bar;
  const int p = threadIdx.x;
  const int chunk = blockIdx.x % (C / 32), img = blockIdx.x / (C / 32);
  const float* col = base + ((size_t)img * C + chunk * 32) * HW + (ids[p] & ~3);
  const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
  if (p == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  if (p == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(256 * 32 * 16));
  __syncthreads();
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sm) + p * 512;
#pragma unroll
  for (int k = 0; k < 32; ++k)
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];"
                 ::"r"(dst + k * 16), "l"(col + (size_t)k * HW), "r"(bar_a) : "memory");
  uint32_t ok = 0;
  while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar_a) : "memory");
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < 32; ++k) acc += *reinterpret_cast<float*>(sm + p * 512 + k * 16);
  if (acc == 12345.678f) out[0] = acc;
}

// Occupies one SM per CTA (227 KB of dynamic shared memory) and spins for `cycles`: emulates a persistent
// kernel that owns S SMs while the gather runs on the others.
__global__ void blocker(volatile int* stop, int* resident, float* out) {
  extern __shared__ unsigned char bs[];
  if (threadIdx.x == 0) atomicAdd(resident, 1);
  const long long t0 = clock64();
  while (clock64() - t0 < 40000000ll) {                     // released by the host (bounded: ~20 ms)
    if (threadIdx.x == 0 && *stop != 0) break;              // one poll per ~10 us, by one thread: no memory traffic to speak of
    __nanosleep(10000);
  }
  if (bs[threadIdx.x] == 77 && *stop == 12345) out[0] = 1.f;
}

template <typename F> float timeit(F f, int n = 5) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f();
  cudaEventRecord(a);
  for (int i = 0; i < n; ++i) f();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms / n * 1e3f;
}

int main(int argc, char** argv) {
  if (argc > 1) {
    CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(argv[1])));
  }
  size_t gran = 0; CK(cudaDeviceGetLimit(&gran, cudaLimitMaxL2FetchGranularity));
  for (int cfg = 0; cfg < 2; ++cfg) {
    const int HW = cfg ? 4096 : 65536, C = cfg ? 256 : 64, nimg = cfg ? 512 : 256;   // 2 GB / 4 GB of maps
    float* base; CK(cudaMalloc(&base, (size_t)nimg * C * HW * 4));
    CK(cudaMemset(base, 0, (size_t)nimg * C * HW * 4));
    int h[256]; srand(1); for (int i = 0; i < 256; ++i) h[i] = rand() % HW;
    int* ids; CK(cudaMalloc(&ids, 1024)); CK(cudaMemcpy(ids, h, 1024, cudaMemcpyHostToDevice));
    float* out; CK(cudaMalloc(&out, 4));
    const unsigned grid = nimg * (C / 32);
    const double sectors = (double)grid * 256 * 32;
    auto rep = [&](const char* nm, float us) { printf("gran=%zu HW=%d %-28s %8.1f us  %6.1f Gsector/s  (%.0f GB/s at 32 B/sector)\n", gran, HW, nm, us, sectors / us / 1e3, sectors * 32 / us / 1e3); };
    rep("ld.cg", timeit([&] { gather<0><<<grid, 256>>>(base, ids, HW, C, out); }));
    rep("ld.nc (ldg)", timeit([&] { gather<1><<<grid, 256>>>(base, ids, HW, C, out); }));
    rep("ld.cv", timeit([&] { gather<2><<<grid, 256>>>(base, ids, HW, C, out); }));
    rep("ld L1::no_allocate", timeit([&] { gather<3><<<grid, 256>>>(base, ids, HW, C, out); }));
    rep("ld.cg L2::64B", timeit([&] { gather<4><<<grid, 256>>>(base, ids, HW, C, out); }));
    rep("ld.cs", timeit([&] { gather<5><<<grid, 256>>>(base, ids, HW, C, out); }));
    rep("ld.cg evict_first hint", timeit([&] { gather<6><<<grid, 256>>>(base, ids, HW, C, out); }));
    // gather speed when S SMs are owned by another (persistent) kernel
    {
      CK(cudaFuncSetAttribute(blocker, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      cudaStream_t sa, sb; CK(cudaStreamCreateWithFlags(&sa, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&sb, cudaStreamNonBlocking));
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      int* stop; CK(cudaHostAlloc(&stop, 8, cudaHostAllocMapped));
      volatile int* res = stop + 1;
      for (int S : {0, 16, 32, 48, 64, 80}) {
        CK(cudaDeviceSynchronize());
        stop[0] = 0; stop[1] = 0;
        if (S) blocker<<<S, 1, 227 * 1024, sa>>>(stop, stop + 1, out);
        while (*res < S) { }                                             // all blocker CTAs are resident
        gather<0><<<grid, 256, 0, sb>>>(base, ids, HW, C, out);          // warm
        CK(cudaEventRecord(e0, sb));
        for (int r = 0; r < 3; ++r) gather<0><<<grid, 256, 0, sb>>>(base, ids, HW, C, out);
        CK(cudaEventRecord(e1, sb));
        CK(cudaEventSynchronize(e1));
        stop[0] = 1;
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        char nm[64]; snprintf(nm, sizeof(nm), "ld.cg, %d SMs blocked", S);
        rep(nm, ms / 3 * 1e3f);
      }
      CK(cudaDeviceSynchronize());
    }
    CK(cudaFuncSetAttribute(gather_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
    rep("cp.async.bulk 16 B (TMA)", timeit([&] { gather_bulk<<<grid, 256, 131072>>>(base, ids, HW, C, out); }));
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    cudaFree(base);
  }
  return 0;
}
