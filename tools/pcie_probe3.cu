// Probe: latency (clock64, one warp per CTA, 147 CTAs at once) of fetching 28 float2 = 224 B per CTA from mapped host memory
// that the CPU has just written, by access pattern.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
__device__ __forceinline__ long long clk() { long long c; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c) :: "memory"); return c; }
template <int MODE> __global__ void rd(const float2* src, long long* out, float* sink, const unsigned long long* dev) {
    const int lane = threadIdx.x; const float2* p = src + blockIdx.x * 28;
    float2 v = make_float2(0, 0);
    __syncwarp();
    const long long t0 = clk();
    if (MODE == 0) { if (lane < 28) v = __ldcv(p + lane); }                                  // one warp-wide uncached load
    if (MODE == 1) { if (lane < 28) asm volatile("ld.relaxed.sys.global.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p + lane) : "memory"); }
    if (MODE == 2) { if (lane == 0) v = __ldcv(p); }                                          // one lane, 8 bytes
    if (MODE == 3) {                                                                          // one instruction per 32-byte sector
#pragma unroll
        for (int k = 0; k < 7; ++k) if ((lane >> 2) == k) v = __ldcv(p + lane);
    }
    if (MODE == 4) { if (lane < 14) { float4 w = __ldcv(reinterpret_cast<const float4*>(p) + lane); v.x = w.x + w.z; v.y = w.y + w.w; } }   // 16 bytes per lane
    if (MODE == 5) { if (lane < 28) v = p[lane]; }                                            // plain (weak) load
    if (MODE == 6) { if (lane < 28) v = __ldcg(p + lane); }
    if (MODE == 7 || MODE == 8) {                                                             // the resident kernel's sequence
        unsigned long long c = 0;
        if (MODE == 8) { if (lane == 0) asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(c) : "l"(dev + 8) : "memory"); c = __shfl_sync(0xffffffffu, c, 0); }
        if (lane < 28) v = __ldcv(p + lane);
        const unsigned long long q0 = __ldcg(dev), q1 = __ldcg(dev + 1), q2 = __ldcg(dev + 2), q3 = __ldcg(dev + 3);
        v.x += (float)(q0 + q1 + q2 + q3 + c);
    }
    if (MODE == 9) {                                                                          // device loads only
        const unsigned long long q0 = __ldcg(dev), q1 = __ldcg(dev + 1), q2 = __ldcg(dev + 2), q3 = __ldcg(dev + 3);
        v.x += (float)(q0 + q1 + q2 + q3);
    }
    float s = v.x + v.y;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const long long t1 = clk();
    if (lane == 0) { out[blockIdx.x] = t1 - t0; if (s == 123456.f) *sink = s; }
}
int main() {
    float2* h; long long *d_out, h_out[147]; float* sink;
    cudaHostAlloc(&h, 1 << 20, cudaHostAllocMapped | cudaHostAllocPortable); cudaMalloc(&d_out, sizeof(h_out)); cudaMalloc(&sink, 4);
    unsigned long long* dev; cudaMalloc(&dev, 256); cudaMemset(dev, 0, 256);
    const char* names[10] = {"warp-wide ld.global.cv, 28 x 8 B", "warp-wide ld.relaxed.sys", "one lane, 8 B", "7 instructions, one per 32-B sector",
                            "14 lanes x 16 B", "plain ld.global", "ld.global.cg", "warp ld.cv + 4 ld.cg of device words", "lane-0 ld.acquire + shfl, then the same", "4 ld.cg of device words only"};
    for (int mode = 0; mode < 10; ++mode) {
        double sum = 0, mx = 0; int n = 0;
        for (int it = 0; it < 60; ++it) {
            for (int i = 0; i < 147 * 28; ++i) h[i] = make_float2((float)it, (float)i);      // the CPU has just written the lines
            switch (mode) {
                case 0: rd<0><<<147, 32>>>(h, d_out, sink, dev); break; case 1: rd<1><<<147, 32>>>(h, d_out, sink, dev); break;
                case 2: rd<2><<<147, 32>>>(h, d_out, sink, dev); break; case 3: rd<3><<<147, 32>>>(h, d_out, sink, dev); break;
                case 4: rd<4><<<147, 32>>>(h, d_out, sink, dev); break; case 5: rd<5><<<147, 32>>>(h, d_out, sink, dev); break;
                case 6: rd<6><<<147, 32>>>(h, d_out, sink, dev); break; case 7: rd<7><<<147, 32>>>(h, d_out, sink, dev); break; case 8: rd<8><<<147, 32>>>(h, d_out, sink, dev); break;
                default: rd<9><<<147, 32>>>(h, d_out, sink, dev); break;
            }
            cudaMemcpy(h_out, d_out, sizeof(h_out), cudaMemcpyDeviceToHost);
            if (it < 10) continue;
            for (int c = 0; c < 147; ++c) { sum += h_out[c]; if (h_out[c] > mx) mx = h_out[c]; ++n; }
        }
        printf("%-40s mean %6.2f us  max %6.2f us per CTA (1.965 GHz)\n", names[mode], sum / n / 1965.0, mx / 1965.0);
    }
    return 0;
}
