// Probe: how fast can a kernel write N bytes into mapped host memory, by store width, vs a DMA copy.
#include <cuda_runtime.h>
#include <stdio.h>
#include <chrono>
template <typename V> __global__ void wr(V* dst, size_t n, float v) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    V x; float* f = (float*)&x; for (int k = 0; k < (int)(sizeof(V) / 4); ++k) f[k] = v + k;
    for (; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = x;
}
__global__ void rd(const float4* src, size_t n, float* out) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; float a = 0;
    for (; i < n; i += (size_t)gridDim.x * blockDim.x) { float4 v = src[i]; a += v.x + v.w; }
    if (a == 123.f) *out = a;
}
int main() {
    const size_t bytes = 647168; void *h, *d; float* o;
    cudaHostAlloc(&h, bytes + 4096, cudaHostAllocMapped | cudaHostAllocPortable); cudaMalloc(&d, bytes + 4096); cudaMalloc(&o, 4);
    cudaStream_t s; cudaStreamCreate(&s);
    auto time = [&](const char* name, auto f) {
        for (int i = 0; i < 20; ++i) { f(); cudaStreamSynchronize(s); }
        auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < 200; ++i) { f(); cudaStreamSynchronize(s); }
        double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / 200;
        printf("%-44s %7.1f us  %6.1f GB/s\n", name, us, bytes / us / 1e3);
    };
    time("empty kernel + sync", [&] { wr<float><<<1, 32, 0, s>>>((float*)d, 0, 1.f); });
    time("kernel float stores -> mapped host (148x256)", [&] { wr<float><<<148, 256, 0, s>>>((float*)h, bytes / 4, 1.f); });
    time("kernel float2 stores -> mapped host", [&] { wr<float2><<<148, 256, 0, s>>>((float2*)h, bytes / 8, 1.f); });
    time("kernel float4 stores -> mapped host", [&] { wr<float4><<<148, 256, 0, s>>>((float4*)h, bytes / 16, 1.f); });
    time("kernel float4 stores -> mapped host (16x256)", [&] { wr<float4><<<16, 256, 0, s>>>((float4*)h, bytes / 16, 1.f); });
    time("kernel float4 stores -> device", [&] { wr<float4><<<148, 256, 0, s>>>((float4*)d, bytes / 16, 1.f); });
    time("cudaMemcpyAsync D2H", [&] { cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, s); });
    time("kernel -> device + cudaMemcpyAsync D2H", [&] { wr<float4><<<148, 256, 0, s>>>((float4*)d, bytes / 16, 1.f); cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, s); });
    time("cudaMemcpyAsync H2D 32 KB", [&] { cudaMemcpyAsync(d, h, 32768, cudaMemcpyHostToDevice, s); });
    time("kernel float4 loads <- mapped host 32 KB", [&] { rd<<<16, 128, 0, s>>>((const float4*)h, 32768 / 16, o); });
    return 0;
}
