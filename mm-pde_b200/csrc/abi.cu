// Library-level entry points of the C ABI.
#include "common.cuh"

extern "C" int mmpde_abi_version(void) { return 1; }

extern "C" int mmpde_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) return (int)e;
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return MMPDE_OK;
}

extern "C" int mmpde_set_persistent_ctas(int n) {
    if (n < 0) return MMPDE_EINVAL;
    mmpde::persistent_cap() = n;
    return MMPDE_OK;
}
