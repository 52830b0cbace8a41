// tma_roundtrip.cu - validates the cp.async.bulk.tensor wrappers of csrc/vpe_sweep_tma.cuh on a small 4D tensor of 8-byte
// texels (x, y, z, brick) with padded rows: load a [8][rows][32] box, add 1 to every word, store it back (optionally one
// slice lower, i.e. with an out-of-bounds slice). build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_roundtrip tma_roundtrip.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
namespace vpe { constexpr int FILLC_THREADS = 256; }
#define VPE_TMA_STANDALONE
#include "../../volumetric-particles-for-unity_b200/csrc/vpe_tma.cuh"
using namespace vpe;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void k(const __grid_constant__ CUtensorMap tmap, int zLoad, int zStore, int brick, int y) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ unsigned long long bar;
    uint2* buf = reinterpret_cast<uint2*>(smem);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if (threadIdx.x == 0) { mbar_expect_tx(&bar, 8 * 256 * 8); tma_load_4d(buf, &tmap, &bar, 0, y, zLoad, brick); }
    mbar_wait(&bar, 0);
    for (int j = 0; j < 8; j++) { uint2 t = buf[j * 256 + threadIdx.x]; t.x += 1; t.y += 1; buf[j * 256 + threadIdx.x] = t; }
    fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0) { tma_store_4d(&tmap, buf, 0, y, zStore, brick); tma_commit(); tma_wait_all(); }
}

int main() {
    const int N = 32, RS = 40, B = 3;
    const size_t texels = (size_t)B * N * N * RS;
    std::vector<uint2> h(texels);
    for (size_t i = 0; i < texels; i++) h[i] = make_uint2((unsigned)i, (unsigned)(i * 7));
    uint2* d; CK(cudaMalloc(&d, texels * 8)); CK(cudaMemcpy(d, h.data(), texels * 8, cudaMemcpyHostToDevice));
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                           CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    CUtensorMap map;
    const cuuint64_t dims[4] = {RS, N, N, B}, strides[3] = {RS * 8, (cuuint64_t)N * RS * 8, (cuuint64_t)N * N * RS * 8};
    const cuuint32_t box[4] = {N, 256 / N, 8, 1}, es[4] = {1, 1, 1, 1};
    CUresult r = ((Fn)p)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)r);
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    struct { int zl, zs, b, y; } cases[] = {{8, 8, 1, 8}, {1, 0, 2, 16}, {25, 24, 0, 0}};
    for (auto c : cases) {
        k<<<1, 256, 16384>>>(map, c.zl, c.zs, c.b, c.y);
        cudaError_t e = cudaDeviceSynchronize();
        printf("load z=%d store z=%d brick %d y %d: %s\n", c.zl, c.zs, c.b, c.y, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        std::vector<uint2> g(texels); CK(cudaMemcpy(g.data(), d, texels * 8, cudaMemcpyDeviceToHost));
        size_t bad = 0, changed = 0;
        for (int z = 0; z < N; z++) for (int yy = 0; yy < N; yy++) for (int x = 0; x < RS; x++) {
            size_t i = (((size_t)c.b * N + z) * N + yy) * RS + x;
            bool inStore = x < N && yy >= c.y && yy < c.y + 8 && z >= c.zs && z < c.zs + 8;
            if (inStore) {
                int zl = z - c.zs + c.zl;  // the loaded slice this slot came from
                size_t src = (((size_t)c.b * N + zl) * N + yy) * RS + x;
                uint2 want = zl < N ? make_uint2(h[src].x + 1, h[src].y + 1) : make_uint2(1, 1);
                if (g[i].x != want.x || g[i].y != want.y) bad++;
                changed++;
            } else if (g[i].x != h[i].x || g[i].y != h[i].y) bad++;
        }
        printf("   %zu texels stored, %zu wrong\n", changed, bad);
        h = g;
    }
    return 0;
}
