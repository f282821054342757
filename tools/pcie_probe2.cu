#include <cuda_runtime.h>
#include <stdio.h>
#include <chrono>
// 147 CTAs, each writes its own contiguous region of `rowbytes` with float2 stores; base offset c*stride
__global__ void wr_rows(float2* dst, int n2, int stride2) {
    float2* d = dst + (size_t)blockIdx.x * stride2;
    for (int i = threadIdx.x; i < n2; i += blockDim.x) d[i] = make_float2(i, blockIdx.x);
}
__global__ void wr_bytes(unsigned char* dst, int n) { if (threadIdx.x < n) dst[blockIdx.x * n + threadIdx.x] = 1; }
__global__ void rd_rel(const float2* src, int n, float* out) {
    float2 v = make_float2(0, 0);
    if (threadIdx.x < n) asm volatile("ld.relaxed.sys.global.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(src + blockIdx.x * n + threadIdx.x) : "memory");
    if (v.x == 123.f) *out = v.y;
}
__global__ void rd_plain(const float2* src, int n, float* out) {
    float2 v = make_float2(0, 0);
    if (threadIdx.x < n) v = src[blockIdx.x * n + threadIdx.x];
    if (v.x == 123.f) *out = v.y;
}
__global__ void fence_only(float* dst) { if (threadIdx.x == 0) { dst[blockIdx.x * 32] = 1.f; __threadfence_system(); dst[blockIdx.x * 32 + 1] = 2.f; } }
int main() {
    const size_t bytes = 647168; void *h, *d; float* o;
    cudaHostAlloc(&h, 4 << 20, cudaHostAllocMapped | cudaHostAllocPortable); cudaMalloc(&d, 4 << 20); cudaMalloc(&o, 4);
    cudaStream_t s; cudaStreamCreate(&s);
    auto time = [&](const char* name, auto f) {
        for (int i = 0; i < 20; ++i) { f(); cudaStreamSynchronize(s); }
        auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < 200; ++i) { f(); cudaStreamSynchronize(s); }
        double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / 200;
        printf("%-60s %7.1f us\n", name, us);
    };
    time("empty", [&] { wr_bytes<<<1, 32, 0, s>>>((unsigned char*)d, 0); });
    time("rows: 147 x 4256 B contiguous (8-B aligned regions), float2", [&] { wr_rows<<<147, 256, 0, s>>>((float2*)h, 532, 532); });
    time("rows: 147 x 4256 B in 4352-B (128-aligned) regions", [&] { wr_rows<<<147, 256, 0, s>>>((float2*)h, 532, 544); });
    time("rows: 147 x 4352 B (whole lines)", [&] { wr_rows<<<147, 256, 0, s>>>((float2*)h, 544, 544); });
    time("rows -> device", [&] { wr_rows<<<147, 256, 0, s>>>((float2*)d, 532, 532); });
    time("flags: 147 x 28 single bytes", [&] { wr_bytes<<<147, 32, 0, s>>>((unsigned char*)h, 28); });
    time("actions: 147 x 28 float2 ld.relaxed.sys", [&] { rd_rel<<<147, 32, 0, s>>>((const float2*)h, 28, o); });
    time("actions: 147 x 28 float2 plain ld", [&] { rd_plain<<<147, 32, 0, s>>>((const float2*)h, 28, o); });
    time("147 x (store, fence.sys, store)", [&] { fence_only<<<147, 32, 0, s>>>((float*)h); });
    return 0;
}
