// ncg_b200_res.cu -- the resident variant of the step kernel (RES = true, see ncg_step.cuh) in a translation unit of its own,
// so that the default kernels carry none of its code.  ncg_step_mapped uses it for a host-driven step loop: the kernel is
// launched once, keeps the records and the track table in shared memory and takes one command per step from a mailbox in
// page-locked host memory, instead of paying a grid launch, the table staging and the record load / store on every step.
#define ncg ncg_res
#include "ncg_step.cuh"

// shape: the (RPL, MINB) pair launch_step picked for the batch; NULL = launched, else why not (the caller falls back to launches)
extern "C" __attribute__((visibility("hidden"))) const char* ncg_res_launch(const void* kparams, size_t kparams_bytes, int rpl, int minb, int n_ctas,
                                                                            int num_sms, int smem_bytes, cudaStream_t stream) {
    if (kparams_bytes != sizeof(KParams)) return "KParams differs between the translation units";
    KParams p = *static_cast<const KParams*>(kparams);
    void (*k)(KParams) = nullptr;
    if (rpl == 2 && minb == 1) k = ncg_step_kernel<2, 1, 1, false, true>;
    else if (rpl == 4 && minb == 2) k = ncg_step_kernel<4, 2, 1, false, true>;
    else return "no resident kernel of this shape";
    const int threads = 32 * (1 + 16 / rpl);
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    // every CTA must be on an SM at the same time: the grid's step completes when its last CTA has counted itself
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, threads, (size_t)smem_bytes);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    if (per_sm * num_sms < n_ctas) return "the grid does not fit on the SMs at once";
    k<<<n_ctas, threads, (size_t)smem_bytes, stream>>>(p);
    e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}
