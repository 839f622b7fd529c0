// vmm.cu -- compressible allocations for the CSR value arrays (default; FECB200_COMPRESS=0 opts out).
//
// Why: 27.6 GB of the fused kernel's 45.6 GB of DRAM traffic at 192^3 are zeros -- the in-kernel clear of the idle value
// array writes 13.8 GB of them and the first RED to every line reads them back.  Blackwell's L2 can keep such lines
// compressed in HBM ("compute data compression") when the allocation is created compressible, which needs the driver's
// virtual-memory API (cuMemCreate with CU_MEM_ALLOCATION_COMP_GENERIC) instead of cudaMalloc.  libcuda is resolved with
// dlopen at run time so that the library still loads on machines without a driver (CPU-side symbol checks).
// Measured at 192^3 (ncu, fused k_mat2): dram read 17.30 -> 8.68 GB, write 28.33 -> 19.90 GB per launch, i.e. exactly the
// algorithmic 28.6 GB; step 16.39 -> 16.23 ms (the kernel is bound on chip, not by HBM).
#include "common.cuh"
#include <cuda.h>
#include <dlfcn.h>

namespace fec {

namespace {
struct Driver {
  bool ok = false;
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice) = nullptr;
  Driver() {
    void* lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return;
    auto sym = [&](const char* n) { return dlsym(lib, n); };
    MemCreate = reinterpret_cast<decltype(MemCreate)>(sym("cuMemCreate"));
    MemAddressReserve = reinterpret_cast<decltype(MemAddressReserve)>(sym("cuMemAddressReserve"));
    MemMap = reinterpret_cast<decltype(MemMap)>(sym("cuMemMap"));
    MemSetAccess = reinterpret_cast<decltype(MemSetAccess)>(sym("cuMemSetAccess"));
    MemUnmap = reinterpret_cast<decltype(MemUnmap)>(sym("cuMemUnmap"));
    MemRelease = reinterpret_cast<decltype(MemRelease)>(sym("cuMemRelease"));
    MemAddressFree = reinterpret_cast<decltype(MemAddressFree)>(sym("cuMemAddressFree"));
    MemGetAllocationGranularity = reinterpret_cast<decltype(MemGetAllocationGranularity)>(sym("cuMemGetAllocationGranularity"));
    DeviceGetAttribute = reinterpret_cast<decltype(DeviceGetAttribute)>(sym("cuDeviceGetAttribute"));
    ok = MemCreate && MemAddressReserve && MemMap && MemSetAccess && MemUnmap && MemRelease && MemAddressFree &&
         MemGetAllocationGranularity && DeviceGetAttribute;
  }
};
Driver& drv() { static Driver d; return d; }
}  // namespace

// returns nullptr when compression is not available (the caller falls back to cudaMalloc)
void* vmm_alloc_compressible(int device, size_t bytes, size_t* mapped, unsigned long long* handle) {
  Driver& d = drv();
  if (!d.ok || !bytes) return nullptr;
  int supported = 0;
  if (d.DeviceGetAttribute(&supported, CU_DEVICE_ATTRIBUTE_GENERIC_COMPRESSION_SUPPORTED, device) != CUDA_SUCCESS || !supported)
    return nullptr;
  CUmemAllocationProp prop{};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device;
  prop.allocFlags.compressionType = CU_MEM_ALLOCATION_COMP_GENERIC;
  size_t gran = 0;
  if (d.MemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || !gran) return nullptr;
  const size_t size = ((bytes + gran - 1) / gran) * gran;
  CUmemGenericAllocationHandle hnd;
  if (d.MemCreate(&hnd, size, &prop, 0) != CUDA_SUCCESS) return nullptr;
  CUdeviceptr ptr = 0;
  if (d.MemAddressReserve(&ptr, size, gran, 0, 0) != CUDA_SUCCESS) { d.MemRelease(hnd); return nullptr; }
  if (d.MemMap(ptr, size, 0, hnd, 0) != CUDA_SUCCESS) { d.MemAddressFree(ptr, size); d.MemRelease(hnd); return nullptr; }
  CUmemAccessDesc acc{};
  acc.location = prop.location;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  if (d.MemSetAccess(ptr, size, &acc, 1) != CUDA_SUCCESS) {
    d.MemUnmap(ptr, size); d.MemAddressFree(ptr, size); d.MemRelease(hnd);
    return nullptr;
  }
  *mapped = size;
  *handle = (unsigned long long)hnd;
  return reinterpret_cast<void*>(ptr);
}

void vmm_free(void* p, size_t mapped, unsigned long long handle) {
  Driver& d = drv();
  if (!d.ok || !p) return;
  d.MemUnmap((CUdeviceptr)p, mapped);
  d.MemAddressFree((CUdeviceptr)p, mapped);
  d.MemRelease((CUmemGenericAllocationHandle)handle);
}

}  // namespace fec
