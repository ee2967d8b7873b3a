// Shareable device memory + NVSwitch multicast objects for the NVLS all-reduce (sm_100a, NVLink 5 / NVSwitch).
//
// The in-switch reduction (multimem.ld_reduce / multimem.st in peer_comm.cu) needs every rank's arena bound to ONE
// multicast object.  That is driver VMM API territory (cuMemCreate / cuMulticast*), reached here through
// cudaGetDriverEntryPoint so that libdmlb.so keeps no link-time dependency on libcuda.so.1 (the library must load on a
// box without a driver: tests/test_abi.py).  Handles cross process boundaries as POSIX file descriptors; the Python host
// passes them over unix sockets (gradsync.PeerComm).
//
// No reference counterpart: the reference reaches NVLS only through NCCL inside torch DDP (pipeline.py:74).
#include <cuda.h>
#include <unistd.h>

#include "dmlb_common.cuh"

namespace dmlb {

struct Driver {
    bool ok = false;
    CUresult (*MemCreate)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long) = nullptr;
    CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*MemAddressReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
    CUresult (*MemExport)(void *, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
    CUresult (*MemImport)(CUmemGenericAllocationHandle *, void *, CUmemAllocationHandleType) = nullptr;
    CUresult (*MemGetGranularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*McCreate)(CUmemGenericAllocationHandle *, const CUmulticastObjectProp *) = nullptr;
    CUresult (*McAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
    CUresult (*McBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long) = nullptr;
    CUresult (*McGetGranularity)(size_t *, const CUmulticastObjectProp *, CUmulticastGranularity_flags) = nullptr;
    CUresult (*DeviceGet)(CUdevice *, int) = nullptr;
    CUresult (*DeviceGetAttribute)(int *, CUdevice_attribute, CUdevice) = nullptr;
};

template <class F>
static bool entry(const char *name, F &fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p)
        return false;
    fn = reinterpret_cast<F>(p);
    return true;
}

static const Driver &driver() {
    static Driver d = [] {
        Driver x;
        x.ok = entry("cuMemCreate", x.MemCreate) && entry("cuMemRelease", x.MemRelease) &&
               entry("cuMemAddressReserve", x.MemAddressReserve) && entry("cuMemAddressFree", x.MemAddressFree) &&
               entry("cuMemMap", x.MemMap) && entry("cuMemUnmap", x.MemUnmap) && entry("cuMemSetAccess", x.MemSetAccess) &&
               entry("cuMemExportToShareableHandle", x.MemExport) && entry("cuMemImportFromShareableHandle", x.MemImport) &&
               entry("cuMemGetAllocationGranularity", x.MemGetGranularity) && entry("cuMulticastCreate", x.McCreate) &&
               entry("cuMulticastAddDevice", x.McAddDevice) && entry("cuMulticastBindMem", x.McBindMem) &&
               entry("cuMulticastGetGranularity", x.McGetGranularity) && entry("cuDeviceGet", x.DeviceGet) &&
               entry("cuDeviceGetAttribute", x.DeviceGetAttribute);
        return x;
    }();
    return d;
}

#define DMLB_CU(x)                                   \
    do {                                             \
        CUresult _r = (x);                           \
        if (_r != CUDA_SUCCESS) return -20000 - (int)_r; \
    } while (0)

static CUmemAccessDesc rw(int device) {
    CUmemAccessDesc a = {};
    a.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    a.location.id = device;
    a.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    return a;
}

static int map_handle(const Driver &d, int device, CUmemGenericAllocationHandle h, size_t bytes, size_t align, void **ptr) {
    CUdeviceptr va = 0;
    DMLB_CU(d.MemAddressReserve(&va, bytes, align, 0, 0));
    CUresult r = d.MemMap(va, bytes, 0, h, 0);
    if (r == CUDA_SUCCESS) {
        CUmemAccessDesc a = rw(device);
        r = d.MemSetAccess(va, bytes, &a, 1);
        if (r != CUDA_SUCCESS) d.MemUnmap(va, bytes);
    }
    if (r != CUDA_SUCCESS) {
        d.MemAddressFree(va, bytes);
        return -20000 - (int)r;
    }
    *ptr = reinterpret_cast<void *>(va);
    return DMLB_OK;
}

}  // namespace dmlb

using namespace dmlb;

extern "C" {

size_t dmlb_vmm_granularity(int device, int world) {
    const Driver &d = driver();
    if (!d.ok || world < 1) return 0;
    if (cudaSetDevice(device) != cudaSuccess || cudaFree(0) != cudaSuccess) return 0;
    CUdevice dev;
    int supported = 0;
    if (d.DeviceGet(&dev, device) != CUDA_SUCCESS) return 0;
    if (d.DeviceGetAttribute(&supported, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) != CUDA_SUCCESS || !supported) return 0;
    CUmulticastObjectProp mp = {};
    mp.numDevices = (unsigned)world;
    mp.size = 2u << 20;
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t g = 0;
    if (d.McGetGranularity(&g, &mp, CU_MULTICAST_GRANULARITY_MINIMUM) != CUDA_SUCCESS) return 0;
    return g;
}

int dmlb_vmm_alloc(int device, size_t bytes, void **ptr, int *fd, uint64_t *handle) {
    const Driver &d = driver();
    if (!d.ok) return DMLB_ESTATE;
    if (!ptr || !fd || !handle || bytes == 0) return DMLB_EINVAL;
    DMLB_CUDA(cudaSetDevice(device));
    DMLB_CUDA(cudaFree(0));
    CUmemAllocationProp ap = {};
    ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    ap.location.id = device;
    ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t g = 0;
    DMLB_CU(d.MemGetGranularity(&g, &ap, CU_MEM_ALLOC_GRANULARITY_MINIMUM));
    if (g == 0 || bytes % g) return DMLB_EALIGN;
    CUmemGenericAllocationHandle h;
    DMLB_CU(d.MemCreate(&h, bytes, &ap, 0));
    int rc = map_handle(d, device, h, bytes, g, ptr);
    if (rc != DMLB_OK) {
        d.MemRelease(h);
        return rc;
    }
    int out = -1;
    CUresult r = d.MemExport(&out, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
    if (r != CUDA_SUCCESS) {
        d.MemUnmap((CUdeviceptr)*ptr, bytes);
        d.MemAddressFree((CUdeviceptr)*ptr, bytes);
        d.MemRelease(h);
        return -20000 - (int)r;
    }
    DMLB_CUDA(cudaMemset(*ptr, 0, bytes));
    DMLB_CUDA(cudaDeviceSynchronize());
    *fd = out;
    *handle = (uint64_t)h;
    return DMLB_OK;
}

int dmlb_vmm_import(int device, int fd, size_t bytes, void **ptr, uint64_t *handle) {
    const Driver &d = driver();
    if (!d.ok) return DMLB_ESTATE;
    if (!ptr || !handle || fd < 0 || bytes == 0) return DMLB_EINVAL;
    DMLB_CUDA(cudaSetDevice(device));
    CUmemGenericAllocationHandle h;
    DMLB_CU(d.MemImport(&h, (void *)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
    int rc = map_handle(d, device, h, bytes, 2u << 20, ptr);
    if (rc != DMLB_OK) {
        d.MemRelease(h);
        return rc;
    }
    *handle = (uint64_t)h;
    return DMLB_OK;
}

int dmlb_vmm_free(void *ptr, size_t bytes, uint64_t handle) {
    const Driver &d = driver();
    if (!d.ok) return DMLB_ESTATE;
    if (ptr) {
        d.MemUnmap((CUdeviceptr)ptr, bytes);
        d.MemAddressFree((CUdeviceptr)ptr, bytes);
    }
    if (handle) d.MemRelease((CUmemGenericAllocationHandle)handle);
    return DMLB_OK;
}

int dmlb_mc_create(int world, size_t bytes, int *fd, uint64_t *mc_handle) {
    const Driver &d = driver();
    if (!d.ok) return DMLB_ESTATE;
    if (!fd || !mc_handle || world < 2 || bytes == 0) return DMLB_EINVAL;
    CUmulticastObjectProp mp = {};
    mp.numDevices = (unsigned)world;
    mp.size = bytes;
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    CUmemGenericAllocationHandle h;
    DMLB_CU(d.McCreate(&h, &mp));
    int out = -1;
    CUresult r = d.MemExport(&out, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
    if (r != CUDA_SUCCESS) {
        d.MemRelease(h);
        return -20000 - (int)r;
    }
    *fd = out;
    *mc_handle = (uint64_t)h;
    return DMLB_OK;
}

int dmlb_mc_import(int fd, uint64_t *mc_handle) {
    const Driver &d = driver();
    if (!d.ok) return DMLB_ESTATE;
    if (!mc_handle || fd < 0) return DMLB_EINVAL;
    CUmemGenericAllocationHandle h;
    DMLB_CU(d.MemImport(&h, (void *)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
    *mc_handle = (uint64_t)h;
    return DMLB_OK;
}

int dmlb_mc_add_device(uint64_t mc_handle, int device) {
    const Driver &d = driver();
    if (!d.ok) return DMLB_ESTATE;
    CUdevice dev;
    DMLB_CU(d.DeviceGet(&dev, device));
    DMLB_CU(d.McAddDevice((CUmemGenericAllocationHandle)mc_handle, dev));
    return DMLB_OK;
}

int dmlb_mc_bind(uint64_t mc_handle, int device, uint64_t mem_handle, size_t bytes, void **mc_ptr) {
    const Driver &d = driver();
    if (!d.ok) return DMLB_ESTATE;
    if (!mc_ptr) return DMLB_EINVAL;
    DMLB_CUDA(cudaSetDevice(device));
    DMLB_CU(d.McBindMem((CUmemGenericAllocationHandle)mc_handle, 0, (CUmemGenericAllocationHandle)mem_handle, 0, bytes, 0));
    return map_handle(d, device, (CUmemGenericAllocationHandle)mc_handle, bytes, 2u << 20, mc_ptr);
}

int dmlb_mc_release(uint64_t mc_handle, void *mc_ptr, size_t bytes) {
    const Driver &d = driver();
    if (!d.ok) return DMLB_ESTATE;
    if (mc_ptr) {
        d.MemUnmap((CUdeviceptr)mc_ptr, bytes);
        d.MemAddressFree((CUdeviceptr)mc_ptr, bytes);
    }
    if (mc_handle) d.MemRelease((CUmemGenericAllocationHandle)mc_handle);
    return DMLB_OK;
}

}  // extern "C"
