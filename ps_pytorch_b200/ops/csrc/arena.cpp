// Symmetric peer-mapped arena: CUDA virtual-memory-management allocations that every rank of the box maps into its own
// address space (P2P loads/stores over NVLink 5 through plain pointers) plus an NVLS multicast mapping (one store ->
// every GPU, replicated in the NVSwitch).  This is the replacement for the reference's numpy receive buffers and MPI
// messages.  Handles are exported as POSIX file descriptors; the python side passes them between
// ranks over AF_UNIX sockets (SCM_RIGHTS).  Driver entry points are resolved at run time (cudaGetDriverEntryPoint) so
// the extension imports on a machine without libcuda (the CPU authoring pod).
#include <torch/extension.h>
#include <cuda.h>
#include <cuda_runtime.h>
#include <unistd.h>

namespace psb {

#define DRV(name) \
    static decltype(&name) fn_##name = nullptr; \
    if (!fn_##name) { \
        cudaDriverEntryPointQueryResult q; \
        void* p = nullptr; \
        cudaError_t e = cudaGetDriverEntryPoint(#name, &p, cudaEnableDefault, &q); \
        TORCH_CHECK(e == cudaSuccess && p != nullptr, "cannot resolve driver entry point " #name); \
        fn_##name = reinterpret_cast<decltype(&name)>(p); \
    }
#define CU_OK(call) do { CUresult r_ = (call); if (r_ != CUDA_SUCCESS) { const char* s_ = nullptr; \
    DRV(cuGetErrorString); fn_cuGetErrorString(r_, &s_); TORCH_CHECK(false, #call " failed: ", s_ ? s_ : "?", " (", (int)r_, ")"); } } while (0)

static CUmemAllocationProp alloc_prop(int dev) {
    CUmemAllocationProp p{};
    p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    p.location.id = dev;
    p.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    return p;
}

static void set_access(CUdeviceptr va, size_t size, int dev) {
    DRV(cuMemSetAccess);
    CUmemAccessDesc d{};
    d.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    d.location.id = dev;
    d.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    CU_OK(fn_cuMemSetAccess(va, size, &d, 1));
}

// -> (va, handle, size, fd)
std::vector<int64_t> vmm_alloc(int64_t nbytes, int64_t dev, int64_t min_granularity) {
    DRV(cuMemGetAllocationGranularity); DRV(cuMemCreate); DRV(cuMemAddressReserve); DRV(cuMemMap); DRV(cuMemExportToShareableHandle);
    cudaFree(0);   // make sure the primary context exists
    CUmemAllocationProp prop = alloc_prop((int)dev);
    size_t gran = 0;
    CU_OK(fn_cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
    gran = std::max<size_t>(gran, (size_t)min_granularity);
    const size_t size = ((size_t)nbytes + gran - 1) / gran * gran;
    CUmemGenericAllocationHandle h;
    CU_OK(fn_cuMemCreate(&h, size, &prop, 0));
    CUdeviceptr va = 0;
    CU_OK(fn_cuMemAddressReserve(&va, size, gran, 0, 0));
    CU_OK(fn_cuMemMap(va, size, 0, h, 0));
    set_access(va, size, (int)dev);
    int fd = -1;
    CU_OK(fn_cuMemExportToShareableHandle(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    return {(int64_t)va, (int64_t)h, (int64_t)size, (int64_t)fd};
}

// -> (va, handle)
std::vector<int64_t> vmm_import(int64_t fd, int64_t size, int64_t dev) {
    DRV(cuMemImportFromShareableHandle); DRV(cuMemAddressReserve); DRV(cuMemMap);
    cudaFree(0);
    CUmemGenericAllocationHandle h;
    CU_OK(fn_cuMemImportFromShareableHandle(&h, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
    CUdeviceptr va = 0;
    CU_OK(fn_cuMemAddressReserve(&va, (size_t)size, 0, 0, 0));
    CU_OK(fn_cuMemMap(va, (size_t)size, 0, h, 0));
    set_access(va, (size_t)size, (int)dev);
    return {(int64_t)va, (int64_t)h};
}

void vmm_unmap(int64_t va, int64_t handle, int64_t size) {
    DRV(cuMemUnmap); DRV(cuMemRelease); DRV(cuMemAddressFree);
    fn_cuMemUnmap((CUdeviceptr)va, (size_t)size);
    fn_cuMemRelease((CUmemGenericAllocationHandle)handle);
    fn_cuMemAddressFree((CUdeviceptr)va, (size_t)size);
}

bool multicast_supported(int64_t dev) {
    DRV(cuDeviceGetAttribute);
    int v = 0;
    if (fn_cuDeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, (CUdevice)dev) != CUDA_SUCCESS) return false;
    return v != 0;
}

static CUmulticastObjectProp mc_prop(int ndev, size_t size) {
    CUmulticastObjectProp p{};
    p.numDevices = (unsigned)ndev;
    p.size = size;
    p.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    return p;
}

int64_t mc_granularity(int64_t ndev, int64_t size) {
    DRV(cuMulticastGetGranularity);
    cudaFree(0);
    CUmulticastObjectProp p = mc_prop((int)ndev, (size_t)size);
    size_t g = 0;
    CU_OK(fn_cuMulticastGetGranularity(&g, &p, CU_MULTICAST_GRANULARITY_RECOMMENDED));
    return (int64_t)g;
}

// rank 0: -> (mc_handle, fd)
std::vector<int64_t> mc_create(int64_t ndev, int64_t size) {
    DRV(cuMulticastCreate); DRV(cuMemExportToShareableHandle);
    CUmulticastObjectProp p = mc_prop((int)ndev, (size_t)size);
    CUmemGenericAllocationHandle h;
    CU_OK(fn_cuMulticastCreate(&h, &p));
    int fd = -1;
    CU_OK(fn_cuMemExportToShareableHandle(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    return {(int64_t)h, (int64_t)fd};
}

int64_t mc_import(int64_t fd) {
    DRV(cuMemImportFromShareableHandle);
    cudaFree(0);
    CUmemGenericAllocationHandle h;
    CU_OK(fn_cuMemImportFromShareableHandle(&h, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
    return (int64_t)h;
}

void mc_add_device(int64_t mc, int64_t dev) {
    DRV(cuMulticastAddDevice);
    CU_OK(fn_cuMulticastAddDevice((CUmemGenericAllocationHandle)mc, (CUdevice)dev));
}

void mc_bind(int64_t mc, int64_t mem_handle, int64_t size) {
    DRV(cuMulticastBindMem);
    CU_OK(fn_cuMulticastBindMem((CUmemGenericAllocationHandle)mc, 0, (CUmemGenericAllocationHandle)mem_handle, 0, (size_t)size, 0));
}

int64_t mc_map(int64_t mc, int64_t size, int64_t dev, int64_t gran) {
    DRV(cuMemAddressReserve); DRV(cuMemMap);
    CUdeviceptr va = 0;
    CU_OK(fn_cuMemAddressReserve(&va, (size_t)size, (size_t)gran, 0, 0));
    CU_OK(fn_cuMemMap(va, (size_t)size, 0, (CUmemGenericAllocationHandle)mc, 0));
    set_access(va, (size_t)size, (int)dev);
    return (int64_t)va;
}

// Tear-down of the multicast side (the order the driver requires): unmap + free the multicast VA, unbind this device's memory from the object,
// release this process's reference to the object.  After this the local allocation can be unmapped/released like any other.
void mc_unmap(int64_t mc_va, int64_t mc, int64_t size, int64_t dev) {
    DRV(cuMemUnmap); DRV(cuMemAddressFree); DRV(cuMulticastUnbind); DRV(cuMemRelease);
    if (mc_va) {
        CU_OK(fn_cuMemUnmap((CUdeviceptr)mc_va, (size_t)size));
        CU_OK(fn_cuMemAddressFree((CUdeviceptr)mc_va, (size_t)size));
    }
    if (mc) {
        CU_OK(fn_cuMulticastUnbind((CUmemGenericAllocationHandle)mc, (CUdevice)dev, 0, (size_t)size));
        CU_OK(fn_cuMemRelease((CUmemGenericAllocationHandle)mc));
    }
}

void close_fd(int64_t fd) { if (fd >= 0) ::close((int)fd); }

// A uint8 view over raw device memory owned by the arena (never freed by torch).
at::Tensor tensor_from_ptr(int64_t ptr, int64_t nbytes, int64_t dev) {
    return at::from_blob(reinterpret_cast<void*>(ptr), {nbytes}, [](void*) {},
                         at::TensorOptions().dtype(at::kByte).device(at::kCUDA, (c10::DeviceIndex)dev));
}

}  // namespace psb
