// Micro-benchmark: does L2/DRAM compute-data compression (cuMemCreate with CU_MEM_ALLOCATION_COMP_GENERIC) speed up the
// dense backward's write stream?  d tgt_feat is a zero fill with a few per cent of sampled values: fills of zeros,
// of "mostly zeros" (a fraction of the 32-byte sectors non-zero) and of random data, plus a read-back sum, on a plain
// cudaMalloc buffer and on a compressible VMM allocation.  Built ON the GPU box (scratch/run_comp.sh), never committed
// as a binary:  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -cudart shared -o /tmp/compbench scratch/compbench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
#define CU(x) do { CUresult e = (x); if (e != CUDA_SUCCESS) { const char* s; cuGetErrorString(e, &s); printf("CU error %s at %d\n", s, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t hash(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

// one CTA of 128 threads per 8 KB tile (the shape of k_dense_flat); mode 0: zeros, 1: a fraction `pm` per mille of the
// 32-byte sectors holds random values, 2: all random
template <int CS>
__global__ void __launch_bounds__(128) k_fill(float4* out, int mode, uint32_t pm, uint32_t seed) {
  const size_t base = (size_t)blockIdx.x * 512;          // float4 per tile
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const size_t i = base + k * 128 + threadIdx.x;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (mode == 3) {                                     // a fraction of the 128-byte LINES holds ONE non-zero float
      const uint32_t line = (uint32_t)(i >> 3);
      const uint32_t h = hash(line * 2654435761u + seed);
      if ((h % 1000u) < pm && (i & 7) == ((h >> 12) & 7)) v.x = __uint_as_float((hash(h) & 0x007fffffu) | 0x3f800000u);
    } else if (mode != 0) {
      const uint32_t sector = (uint32_t)(i >> 1);
      const uint32_t h = hash(sector * 2654435761u + seed);
      if (mode == 2 || (h % 1000u) < pm) {
        const uint32_t g = hash((uint32_t)i + seed);
        v = make_float4(__uint_as_float((g & 0x007fffffu) | 0x3f800000u), __uint_as_float((hash(g) & 0x007fffffu) | 0x3f800000u),
                        __uint_as_float((hash(g + 1) & 0x007fffffu) | 0x3f800000u), __uint_as_float((hash(g + 2) & 0x007fffffu) | 0x3f800000u));
      }
    }
    if (CS) __stcs(out + i, v); else out[i] = v;
  }
}
// the same tile shape, staged in shared memory and written with ONE bulk copy (cp.async.bulk shared -> global)
__global__ void __launch_bounds__(128) k_fill_bulk(float4* out, int mode, uint32_t pm, uint32_t seed) {
  __shared__ __align__(128) float4 tile[512];
  const size_t base = (size_t)blockIdx.x * 512;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const size_t i = base + k * 128 + threadIdx.x;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (mode == 3) {
      const uint32_t line = (uint32_t)(i >> 3);
      const uint32_t h = hash(line * 2654435761u + seed);
      if ((h % 1000u) < pm && (i & 7) == ((h >> 12) & 7)) v.x = __uint_as_float((hash(h) & 0x007fffffu) | 0x3f800000u);
    }
    tile[k * 128 + threadIdx.x] = v;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(tile);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + base), "r"(s), "r"(8192u) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}
static float time_fill_bulk(float4* buf, size_t bytes, int mode, uint32_t pm) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const unsigned grid = (unsigned)(bytes / 8192);
  for (int w = 0; w < 3; ++w) k_fill_bulk<<<grid, 128>>>(buf, mode, pm, 17u + w);
  CK(cudaEventRecord(e0));
  const int reps = 10;
  for (int r = 0; r < reps; ++r) k_fill_bulk<<<grid, 128>>>(buf, mode, pm, 100u + r);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / reps;
}
// one CTA per CHUNK of several contiguous 8 KB tiles (a whole (b, c) row of the gradient): does the fill keep its rate when
// a CTA walks many tiles?  (on plain memory persistent / multi-tile CTAs lose the linear DRAM write front: DESIGN.md 4.2)
template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_fill_chunk(float4* out, int tiles_per_cta, uint32_t pm, uint32_t seed) {
  const size_t base = (size_t)blockIdx.x * tiles_per_cta * 512;
  for (int t = 0; t < tiles_per_cta; ++t) {
#pragma unroll
    for (int k = 0; k < 512 / THREADS; ++k) {
      const size_t i = base + (size_t)t * 512 + k * THREADS + threadIdx.x;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      const uint32_t line = (uint32_t)(i >> 3);
      const uint32_t h = hash(line * 2654435761u + seed);
      if ((h % 1000u) < pm && (i & 7) == ((h >> 12) & 7)) v.x = __uint_as_float((hash(h) & 0x007fffffu) | 0x3f800000u);
      __stcs(out + i, v);
    }
  }
}
template <int THREADS>
static float time_fill_chunk(float4* buf, size_t bytes, int tiles_per_cta, uint32_t pm) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const unsigned grid = (unsigned)(bytes / 8192 / tiles_per_cta);
  for (int w = 0; w < 3; ++w) k_fill_chunk<THREADS><<<grid, THREADS>>>(buf, tiles_per_cta, pm, 17u + w);
  CK(cudaEventRecord(e0));
  const int reps = 10;
  for (int r = 0; r < reps; ++r) k_fill_chunk<THREADS><<<grid, THREADS>>>(buf, tiles_per_cta, pm, 100u + r);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / reps;
}
__global__ void __launch_bounds__(256) k_sum(const float4* in, size_t n4, float* out) {
  float acc = 0.f;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256) {
    const float4 v = in[i];
    acc += v.x + v.y + v.z + v.w;
  }
  if (acc == 12345.678f) out[0] = acc;
}

template <int CS>
static float time_fill(float4* buf, size_t bytes, int mode, uint32_t pm) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const unsigned grid = (unsigned)(bytes / 8192);
  for (int w = 0; w < 3; ++w) k_fill<CS><<<grid, 128>>>(buf, mode, pm, 17u + w);
  CK(cudaEventRecord(e0));
  const int reps = 10;
  for (int r = 0; r < reps; ++r) k_fill<CS><<<grid, 128>>>(buf, mode, pm, 100u + r);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / reps;
}
static float time_sum(const float4* buf, size_t bytes, float* out) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int w = 0; w < 2; ++w) k_sum<<<148 * 8, 256>>>(buf, bytes / 16, out);
  CK(cudaEventRecord(e0));
  const int reps = 10;
  for (int r = 0; r < reps; ++r) k_sum<<<148 * 8, 256>>>(buf, bytes / 16, out);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / reps;
}

int main() {
  CK(cudaSetDevice(0));
  CK(cudaFree(0));
  CUdevice dev; CU(cuDeviceGet(&dev, 0));
  int comp = 0;
  CU(cuDeviceGetAttribute(&comp, CU_DEVICE_ATTRIBUTE_GENERIC_COMPRESSION_SUPPORTED, dev));
  printf("generic compression supported: %d\n", comp);
  const size_t want = (size_t)3221225472ull;               // the B=64 dense gradient: 3.2 GB
  float4* plain = nullptr;
  CK(cudaMalloc(&plain, want));
  float* out; CK(cudaMalloc(&out, 256));
  CUmemAllocationProp prop = {};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = 0;
  prop.allocFlags.compressionType = CU_MEM_ALLOCATION_COMP_GENERIC;
  size_t gran = 0;
  CU(cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  const size_t size = (want + gran - 1) / gran * gran;
  CUmemGenericAllocationHandle h;
  CUresult rc = cuMemCreate(&h, size, &prop, 0);
  if (rc != CUDA_SUCCESS) { const char* s; cuGetErrorString(rc, &s); printf("cuMemCreate(compressible) failed: %s\n", s); return 0; }
  CUmemAllocationProp got = {};
  CU(cuMemGetAllocationPropertiesFromHandle(&got, h));
  printf("granularity %zu, compressionType granted: %d\n", gran, (int)got.allocFlags.compressionType);
  CUdeviceptr va;
  CU(cuMemAddressReserve(&va, size, 0, 0, 0));
  CU(cuMemMap(va, size, 0, h, 0));
  CUmemAccessDesc acc = {};
  acc.location = prop.location;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  CU(cuMemSetAccess(va, size, &acc, 1));
  float4* compb = reinterpret_cast<float4*>(va);
  const char* names[2] = {"plain cudaMalloc", "compressible VMM"};
  float4* bufs[2] = {plain, compb};
  for (int b = 0; b < 2; ++b) {
    const float z = time_fill<0>(bufs[b], want, 0, 0);
    const float rz = time_sum(bufs[b], want, out);
    const float s = time_fill<0>(bufs[b], want, 1, 120);     // 12 % of the sectors non-zero (a 256^2 C=64 layer's lines with samples)
    const float rs = time_sum(bufs[b], want, out);
    const float s5 = time_fill<0>(bufs[b], want, 1, 500);
    const float l1 = time_fill<0>(bufs[b], want, 3, 117), l2 = time_fill<0>(bufs[b], want, 3, 390), l3 = time_fill<0>(bufs[b], want, 3, 860);
    const float c1 = time_fill<1>(bufs[b], want, 3, 117), c0 = time_fill<1>(bufs[b], want, 0, 0);
    printf("%-18s CTA per chunk (11.7 %% lines): 128 thr x 2 / 8 / 32 tiles: %.3f / %.3f / %.3f ms | 256 thr x 2 / 8 / 32 tiles: %.3f / %.3f / %.3f ms\n", names[b],
           time_fill_chunk<128>(bufs[b], want, 2, 117), time_fill_chunk<128>(bufs[b], want, 8, 117), time_fill_chunk<128>(bufs[b], want, 32, 117),
           time_fill_chunk<256>(bufs[b], want, 2, 117), time_fill_chunk<256>(bufs[b], want, 8, 117), time_fill_chunk<256>(bufs[b], want, 32, 117));
    printf("%-18s staged + one bulk store per tile: zeros %.3f ms | one float in 11.7 %% of the lines %.3f ms | 86 %%: %.3f ms\n", names[b],
           time_fill_bulk(bufs[b], want, 0, 0), time_fill_bulk(bufs[b], want, 3, 117), time_fill_bulk(bufs[b], want, 3, 860));
    printf("%-18s st.global.cs: zeros %.3f ms | one float in 11.7 %% of the lines %.3f ms\n", names[b], c0, c1);
    printf("%-18s one float in 11.7 %% of the lines: %.3f ms | 39 %%: %.3f ms | 86 %%: %.3f ms\n", names[b], l1, l2, l3);
    const float r = time_fill<0>(bufs[b], want, 2, 0);
    const float rr = time_sum(bufs[b], want, out);
    printf("%-18s fill zeros %.3f ms (%.0f GB/s) | 12%% sectors %.3f ms (%.0f GB/s) | 50%% %.3f ms | random %.3f ms (%.0f GB/s) || read zeros %.3f ms (%.0f GB/s) read 12%% %.3f ms read random %.3f ms (%.0f GB/s)\n",
           names[b], z, want / z / 1e6, s, want / s / 1e6, s5, r, want / r / 1e6, rz, want / rz / 1e6, rs, rr, want / rr / 1e6);
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
