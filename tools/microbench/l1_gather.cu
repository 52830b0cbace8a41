// Microbenchmark: cost of a warp-wide 8-byte (and 16-byte) gather from L1-resident data as a function of
// the lane -> address pattern (rows of a 32-texel brick row = 256 B). Informs the march kernel's warp
// tile shape and brick layout. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l1_gather l1_gather.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_gather64(const uint2* __restrict__ buf, const int* __restrict__ laneOff, int iters, int span, unsigned long long* out) {
    const int lane = threadIdx.x & 31;
    const uint2* p = buf + laneOff[lane] + (threadIdx.x >> 5) * 64;
    unsigned acc = 0;
    int step = 0;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            uint2 v = __ldg(p + step);
            acc ^= v.x + v.y;
            step = (step + 32 * 33) & (span - 1);  // walk rows inside an L1-resident window
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}
__global__ void k_gather128(const uint4* __restrict__ buf, const int* __restrict__ laneOff, int iters, int span, unsigned long long* out) {
    const int lane = threadIdx.x & 31;
    const uint4* p = buf + laneOff[lane] + (threadIdx.x >> 5) * 32;
    unsigned acc = 0;
    int step = 0;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            uint4 v = __ldg(p + step);
            acc ^= v.x + v.y + v.z + v.w;
            step = (step + 16 * 33) & (span - 1);
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}

int main() {
    const int texels = 1 << 13;  // 64 KB window of 8-byte texels (+ slack), stays in L1
    uint2* buf; cudaMalloc(&buf, (texels + 4096 + 65536) * sizeof(uint2)); cudaMemset(buf, 1, (texels + 4096 + 65536) * sizeof(uint2));
    int* dOff; cudaMalloc(&dOff, 32 * sizeof(int));
    unsigned long long* dOut; cudaMalloc(&dOut, 8);
    struct Pat { const char* name; int off[32]; bool wide; } pats[16];
    int np = 0;
    auto add = [&](const char* name, auto f, bool wide = false) { pats[np].name = name; pats[np].wide = wide; for (int i = 0; i < 32; i++) pats[np].off[i] = f(i); np++; };
    const int ROW = 32, PLANE = 1024;
    add("32x1 contiguous (256 B, 2 lines)", [&](int i) { return i; });
    add("32x1 contiguous, misaligned by 5 texels", [&](int i) { return i + 5; });
    add("16x2 rows", [&](int i) { return (i / 16) * ROW + (i % 16) + 3; });
    add("8x4 rows", [&](int i) { return (i / 8) * ROW + (i % 8) + 3; });
    add("8x4 rows, x ^= (row&1)<<3 swizzle", [&](int i) { int r = i / 8, x = (i % 8) + 3; return r * ROW + (x ^ ((r & 1) << 3)); });
    add("8x4 rows, x ^= (row&3)<<2 swizzle", [&](int i) { int r = i / 8, x = (i % 8) + 3; return r * ROW + (x ^ ((r & 3) << 2)); });
    add("8x4, rows in 2 planes (2 rows x 2 z)", [&](int i) { int r = i / 8; return (r & 1) * ROW + (r >> 1) * PLANE + (i % 8) + 3; });
    add("4x8 rows", [&](int i) { return (i / 4) * ROW + (i % 4) + 3; });
    add("1x32 rows (same x)", [&](int i) { return i * ROW + 3; });
    add("8x4 rows, stride 9/8 texel (9 texels wide)", [&](int i) { return (i / 8) * ROW + ((i % 8) * 9) / 8 + 3; });
    add("all lanes same texel (broadcast)", [&](int i) { return 7; });
    add("8x4 rows, row stride 33 texels (padded rows)", [&](int i) { return (i / 8) * 33 + (i % 8) + 3; });
    add("8x4 rows, row stride 40 texels (pad 64 B)", [&](int i) { return (i / 8) * 40 + (i % 8) + 3; });
    add("LDG.128: 8x4 rows of 16-B pairs (pair idx)", [&](int i) { return (i / 8) * 16 + (i % 8) + 1; }, true);
    add("LDG.128: 16x2 rows of 16-B pairs", [&](int i) { return (i / 16) * 16 + (i % 16); }, true);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int dev = 0; cudaDeviceProp prop; cudaGetDeviceProperties(&prop, dev);
    const int blocks = prop.multiProcessorCount * 4, threads = 256, iters = 4000;
    for (int pi = 0; pi < np; pi++) {
        cudaMemcpy(dOff, pats[pi].off, sizeof(int) * 32, cudaMemcpyHostToDevice);
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0);
            if (pats[pi].wide) k_gather128<<<blocks, threads>>>((const uint4*)buf, dOff, iters, texels / 2, dOut);
            else k_gather64<<<blocks, threads>>>(buf, dOff, iters, texels, dOut);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        double warpLoads = (double)blocks * (threads / 32) * iters * 8;
        double perSm = warpLoads / prop.multiProcessorCount;
        double clk = 1.965e9;  // max SM clock; cycles are approximate
        printf("%-48s %8.3f ms  %6.2f cycles per warp-load per SM\n", pats[pi].name, best, best * 1e-3 * clk / perSm);
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
