// Compressible device memory for the dense gradients (B200: L2 / HBM compute-data compression).
//
// d tgt_feat is a zero fill with one sampled float in a few per cent of its 128-byte lines, written once per step by
// k_dense_flat / k_dense_nhwc -- the dominant kernel of the path, bound by HBM writes.  Memory created
// with CU_MEM_ALLOCATION_COMP_GENERIC is compressed on its way from L2 to HBM: the same kernel then fills 3.2 GB in
// 0.38 ms instead of 0.43 ms (scratch/compbench.cu: 8.5 vs 7.5 TB/s, independent of how many lines carry a sample),
// and whoever reads the gradient next (the generator's backward) reads the zero lines at 9.6 instead of 6.8 TB/s.
// Transparent to every kernel: only the ALLOCATION differs.  The two functions below have the signature of
// torch.cuda.memory.CUDAPluggableAllocator and back a torch.cuda.MemPool that only the gradient tensors come from
// (patchnce.py: _grad_pool); torch's caching allocator keeps the blocks, so the driver calls here happen once per size.
// The driver API is reached through cudaGetDriverEntryPoint: libpnce.so keeps no link-time dependency on libcuda.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <atomic>
#include <cstddef>

namespace pnce {

struct CompDriver {
  CUresult (*getGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*release)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*reserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*addressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*unmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*setAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*getProps)(CUmemAllocationProp*, CUmemGenericAllocationHandle) = nullptr;
  CUresult (*devGetAttr)(int*, CUdevice_attribute, CUdevice) = nullptr;
  bool ok = false;
};

struct CompBlock {
  CUmemGenericAllocationHandle handle;
  size_t size;
  int device;
  bool compressed;
};

// a fixed table and a spin lock: the library stays free of libstdc++ (blocks are few: one per gradient tensor size)
constexpr int kCompMaxBlocks = 4096;
struct CompEntry { void* ptr; CompBlock blk; };
static CompEntry g_comp_blocks[kCompMaxBlocks];
static std::atomic_flag g_comp_mu = ATOMIC_FLAG_INIT;
struct CompLock {
  CompLock() { while (g_comp_mu.test_and_set(std::memory_order_acquire)) { } }
  ~CompLock() { g_comp_mu.clear(std::memory_order_release); }
};
static CompEntry* comp_find(void* ptr) {
  for (int i = 0; i < kCompMaxBlocks; ++i)
    if (g_comp_blocks[i].ptr == ptr) return &g_comp_blocks[i];
  return nullptr;
}
static CompDriver g_comp_drv;
static bool g_comp_drv_tried = false;

template <typename F>
static bool comp_entry(const char* name, F* fn) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult st;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess || !p) {
    (void)cudaGetLastError();
    return false;
  }
  *fn = reinterpret_cast<F>(p);
  return true;
}

// caller holds g_comp_mu
static const CompDriver& comp_driver() {
  if (!g_comp_drv_tried) {
    g_comp_drv_tried = true;
    CompDriver d;
    d.ok = comp_entry("cuMemGetAllocationGranularity", &d.getGranularity) && comp_entry("cuMemCreate", &d.create) &&
           comp_entry("cuMemRelease", &d.release) && comp_entry("cuMemAddressReserve", &d.reserve) &&
           comp_entry("cuMemAddressFree", &d.addressFree) && comp_entry("cuMemMap", &d.map) &&
           comp_entry("cuMemUnmap", &d.unmap) && comp_entry("cuMemSetAccess", &d.setAccess) &&
           comp_entry("cuMemGetAllocationPropertiesFromHandle", &d.getProps) &&
           comp_entry("cuDeviceGetAttribute", &d.devGetAttr);
    g_comp_drv = d;
  }
  return g_comp_drv;
}

}  // namespace pnce

// 1: device `device` creates compressible allocations (and the driver entry points resolve), 0: it does not
extern "C" int pnce_comp_supported(int device) {
  using namespace pnce;
  CompLock lk;
  const CompDriver& d = comp_driver();
  if (!d.ok) return 0;
  int v = 0;
  if (d.devGetAttr(&v, CU_DEVICE_ATTRIBUTE_GENERIC_COMPRESSION_SUPPORTED, (CUdevice)device) != CUDA_SUCCESS) return 0;
  return v ? 1 : 0;
}

// CUDAPluggableAllocator alloc_fn: `size` bytes of device memory on `device`, compressible when the device grants it
// (plain VMM memory otherwise); NULL on failure (torch then reports out-of-memory).
extern "C" void* pnce_comp_alloc(ptrdiff_t size, int device, void* /*stream*/) {
  using namespace pnce;
  if (size <= 0) return nullptr;
  CompLock lk;
  const CompDriver& d = comp_driver();
  if (!d.ok) return nullptr;
  int prev = -1;
  if (cudaGetDevice(&prev) != cudaSuccess) return nullptr;
  if (prev != device && cudaSetDevice(device) != cudaSuccess) return nullptr;
  (void)cudaFree(nullptr);                                   // make sure the primary context exists and is current
  void* result = nullptr;
  CompEntry* slot = comp_find(nullptr);                      // a free table entry
  if (slot == nullptr) return nullptr;
  CUmemAllocationProp prop = {};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device;
  for (int attempt = 0; attempt < 2 && result == nullptr; ++attempt) {
    prop.allocFlags.compressionType = attempt == 0 ? CU_MEM_ALLOCATION_COMP_GENERIC : CU_MEM_ALLOCATION_COMP_NONE;
    size_t gran = 0;
    if (d.getGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || gran == 0) continue;
    const size_t bytes = ((size_t)size + gran - 1) / gran * gran;
    CUmemGenericAllocationHandle h;
    if (d.create(&h, bytes, &prop, 0) != CUDA_SUCCESS) continue;
    CUdeviceptr va = 0;
    if (d.reserve(&va, bytes, 0, 0, 0) != CUDA_SUCCESS) { d.release(h); continue; }
    if (d.map(va, bytes, 0, h, 0) != CUDA_SUCCESS) { d.addressFree(va, bytes); d.release(h); continue; }
    CUmemAccessDesc acc = {};
    acc.location = prop.location;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    if (d.setAccess(va, bytes, &acc, 1) != CUDA_SUCCESS) { d.unmap(va, bytes); d.addressFree(va, bytes); d.release(h); continue; }
    CUmemAllocationProp got = {};
    const bool comp = d.getProps(&got, h) == CUDA_SUCCESS && got.allocFlags.compressionType == CU_MEM_ALLOCATION_COMP_GENERIC;
    result = reinterpret_cast<void*>(va);
    slot->ptr = result;
    slot->blk = CompBlock{h, bytes, device, comp};
  }
  if (prev != device) (void)cudaSetDevice(prev);
  return result;
}

// CUDAPluggableAllocator free_fn
extern "C" void pnce_comp_free(void* ptr, ptrdiff_t /*size*/, int /*device*/, void* /*stream*/) {
  using namespace pnce;
  if (ptr == nullptr) return;
  CompLock lk;
  CompEntry* e = comp_find(ptr);
  if (e == nullptr) return;
  const CompBlock b = e->blk;
  e->ptr = nullptr;
  const CompDriver& d = comp_driver();
  int prev = -1;
  (void)cudaGetDevice(&prev);
  if (prev != b.device) (void)cudaSetDevice(b.device);
  (void)cudaDeviceSynchronize();                             // unmapping does not wait for work in flight (cudaFree does)
  d.unmap(reinterpret_cast<CUdeviceptr>(ptr), b.size);
  d.release(b.handle);
  d.addressFree(reinterpret_cast<CUdeviceptr>(ptr), b.size);
  if (prev != b.device && prev >= 0) (void)cudaSetDevice(prev);
}

// 1 when `ptr` lies inside a live block of this allocator that the driver made compressible, 0 otherwise (tests, bench)
extern "C" int pnce_comp_is_compressed(const void* ptr) {
  using namespace pnce;
  CompLock lk;
  if (ptr == nullptr) return 0;
  const char* q = static_cast<const char*>(ptr);
  for (int i = 0; i < kCompMaxBlocks; ++i) {
    const CompEntry& e = g_comp_blocks[i];
    if (e.ptr != nullptr && q >= static_cast<const char*>(e.ptr) && q < static_cast<const char*>(e.ptr) + e.blk.size)
      return e.blk.compressed ? 1 : 0;
  }
  return 0;
}
