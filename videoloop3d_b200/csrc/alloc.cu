// Optional allocation helper: device memory with the GENERIC compression attribute (CUDA virtual memory management).
//
// The gradient buffer of the dynamic atlas spends two of its four HBM crossings per step as ZEROS (written back by the
// Adam items, fetched again by the first RED of the next step: 2 x 22.6 GB at 720p).  Lines of a compressible allocation
// that hold all-zero data move through the memory system compressed, so those crossings cost a fraction of their size;
// lines with real gradients are stored uncompressed as before.  Results never depend on where a buffer lives: this is
// purely a placement hint, used when the device supports it (cudaDevAttr / CU_DEVICE_ATTRIBUTE_GENERIC_COMPRESSION_SUPPORTED).
// The library keeps no other state; the handles of these allocations are kept in a small table so that
// vl3d_free_compressible can unmap them.
#include <cuda.h>

#include <mutex>
#include <unordered_map>

#include "vl3d_common.cuh"

namespace vl3d {

struct VmmApi {
    CUresult (*getAttr)(int*, CUdevice_attribute, CUdevice) = nullptr;
    CUresult (*granularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
    CUresult (*reserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*setAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    CUresult (*unmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*release)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*addressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*getProps)(CUmemAllocationProp*, CUmemGenericAllocationHandle) = nullptr;
    bool ok = false;
};

static const VmmApi& vmm() {
    static VmmApi api = [] {
        VmmApi a;
        auto get = [](const char* name) -> void* {
            void* p = nullptr;
            cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
            (void)cudaGetLastError();
            return p;
        };
        a.getAttr = reinterpret_cast<decltype(a.getAttr)>(get("cuDeviceGetAttribute"));
        a.granularity = reinterpret_cast<decltype(a.granularity)>(get("cuMemGetAllocationGranularity"));
        a.create = reinterpret_cast<decltype(a.create)>(get("cuMemCreate"));
        a.reserve = reinterpret_cast<decltype(a.reserve)>(get("cuMemAddressReserve"));
        a.map = reinterpret_cast<decltype(a.map)>(get("cuMemMap"));
        a.setAccess = reinterpret_cast<decltype(a.setAccess)>(get("cuMemSetAccess"));
        a.unmap = reinterpret_cast<decltype(a.unmap)>(get("cuMemUnmap"));
        a.release = reinterpret_cast<decltype(a.release)>(get("cuMemRelease"));
        a.addressFree = reinterpret_cast<decltype(a.addressFree)>(get("cuMemAddressFree"));
        a.getProps = reinterpret_cast<decltype(a.getProps)>(get("cuMemGetAllocationPropertiesFromHandle"));
        a.ok = a.getAttr && a.granularity && a.create && a.reserve && a.map && a.setAccess && a.unmap && a.release && a.addressFree;
        return a;
    }();
    return api;
}

struct VmmEntry { CUmemGenericAllocationHandle handle; size_t bytes; };
static std::mutex g_vmm_mutex;
static std::unordered_map<void*, VmmEntry> g_vmm;

}  // namespace vl3d

using namespace vl3d;

extern "C" int vl3d_alloc_compressible(int64_t bytes, void** ptr_out, int64_t* bytes_out, int32_t* compressed_out) {
    VL3D_REQUIRE(ptr_out && bytes_out && bytes > 0, VL3D_EINVAL, "alloc_compressible: bad arguments");
    *ptr_out = nullptr; *bytes_out = 0;
    if (compressed_out) *compressed_out = 0;
    const VmmApi& A = vmm();
    VL3D_REQUIRE(A.ok, VL3D_EINVAL, "alloc_compressible: the CUDA virtual memory management entry points are not available");
    int dev = 0;
    cudaError_t ce = cudaGetDevice(&dev);
    if (ce != cudaSuccess) return set_err((int)ce, "alloc_compressible: %s", cudaGetErrorString(ce));
    cudaFree(nullptr);                                              // make sure the primary context exists
    int supported = 0;
    A.getAttr(&supported, CU_DEVICE_ATTRIBUTE_GENERIC_COMPRESSION_SUPPORTED, (CUdevice)dev);
    VL3D_REQUIRE(supported != 0, VL3D_EINVAL, "alloc_compressible: the device does not support generic compression");
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = dev;
    prop.allocFlags.compressionType = CU_MEM_ALLOCATION_COMP_GENERIC;
    size_t gran = 0;
    CUresult r = A.granularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
    if (r != CUDA_SUCCESS || gran == 0) return set_err(VL3D_EINVAL, "alloc_compressible: cuMemGetAllocationGranularity failed (%d)", (int)r);
    const size_t size = ((size_t)bytes + gran - 1) / gran * gran;
    CUmemGenericAllocationHandle h;
    r = A.create(&h, size, &prop, 0);
    if (r != CUDA_SUCCESS) return set_err(VL3D_EINVAL, "alloc_compressible: cuMemCreate(%zu bytes) failed (%d)", size, (int)r);
    CUdeviceptr va = 0;
    r = A.reserve(&va, size, gran, 0, 0);
    if (r != CUDA_SUCCESS) { A.release(h); return set_err(VL3D_EINVAL, "alloc_compressible: cuMemAddressReserve failed (%d)", (int)r); }
    r = A.map(va, size, 0, h, 0);
    if (r != CUDA_SUCCESS) { A.addressFree(va, size); A.release(h); return set_err(VL3D_EINVAL, "alloc_compressible: cuMemMap failed (%d)", (int)r); }
    CUmemAccessDesc acc = {};
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = dev;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    r = A.setAccess(va, size, &acc, 1);
    if (r != CUDA_SUCCESS) { A.unmap(va, size); A.addressFree(va, size); A.release(h); return set_err(VL3D_EINVAL, "alloc_compressible: cuMemSetAccess failed (%d)", (int)r); }
    if (compressed_out && A.getProps) {                             // did the driver grant the compression attribute?
        CUmemAllocationProp got = {};
        if (A.getProps(&got, h) == CUDA_SUCCESS) *compressed_out = got.allocFlags.compressionType == CU_MEM_ALLOCATION_COMP_GENERIC ? 1 : 0;
    }
    {
        std::lock_guard<std::mutex> lock(g_vmm_mutex);
        g_vmm[(void*)va] = VmmEntry{h, size};
    }
    *ptr_out = (void*)va; *bytes_out = (int64_t)size;
    return 0;
}

extern "C" int vl3d_free_compressible(void* ptr) {
    if (ptr == nullptr) return 0;
    VmmEntry e;
    {
        std::lock_guard<std::mutex> lock(g_vmm_mutex);
        auto it = g_vmm.find(ptr);
        VL3D_REQUIRE(it != g_vmm.end(), VL3D_EINVAL, "free_compressible: unknown pointer");
        e = it->second;
        g_vmm.erase(it);
    }
    const VmmApi& A = vmm();
    A.unmap((CUdeviceptr)ptr, e.bytes);
    A.addressFree((CUdeviceptr)ptr, e.bytes);
    A.release(e.handle);
    return 0;
}
