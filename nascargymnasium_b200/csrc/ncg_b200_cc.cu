// ncg_b200_cc.cu -- the shared-world variant of the step kernel (NcgConfig.car_contacts: the cars of an env collide with each
// other; optional, default off) in a translation unit of its own, see ncg_step.cuh.
// (the out-of-line __host__ __device__ functions of the shared headers get their own names in this unit)
#define ncg ncg_cc
#include "ncg_step.cuh"

extern "C" __attribute__((visibility("hidden"))) const char* ncg_cc_launch(const void* kparams, size_t kparams_bytes, int n_ctas, int smem_bytes, cudaStream_t stream) {
    if (kparams_bytes != sizeof(KParams)) return "KParams differs between the translation units";
    KParams p = *static_cast<const KParams*>(kparams);
    void (*k)(KParams) = ncg_step_kernel<4, 2, 1, true>;       // one shape: an env's cars sit in one physics warp
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    k<<<n_ctas, 32 * (1 + 16 / 4), (size_t)smem_bytes, stream>>>(p);
    e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}
