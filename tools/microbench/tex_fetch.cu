// tex_fetch.cu - is the texture unit a cheaper way to fetch + convert the march's 2x2x2 fp16 footprint than LDG.64 + HADD2.F32?
// Measures, for an 8x4 warp tile of rays 0.7 texel apart stepping 0.81 texel per sample through a 3D volume of z-paired
// (r,d)[z],(r,d)[z+1] half4 texels (the engine's grey brick texel):
//   tex : 4 point-sampled tex3D<float4> fetches per sample (cudaArray, block-linear; the unit converts half -> float)
//   ldg : 4 LDG.64 from linear memory + 16 HADD2.F32 (what k_march_flat does)
// both followed by the same 7 packed lerps. Prints ns per warp-sample and the implied cycles per SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tex_fetch tex_fetch.cu ; run: ./tex_fetch [edge=384]
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ float2 lerp2(float2 a, float2 b, float w) { return make_float2(fmaf(w, b.x - a.x, a.x), fmaf(w, b.y - a.y, a.y)); }

struct RayState { float x, y, z, dx, dy, dz; };
__device__ RayState make_ray(int E, int steps) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    // every warp starts somewhere else; lanes form an 8x4 tile 0.7 texel apart, perpendicular to the (oblique) direction
    unsigned h = warp * 2654435761u;
    RayState r;
    const float span = (float)E - 0.81f * steps * 0.8f - 16.0f;
    r.x = 4.0f + (h & 1023) * (span / 1024.0f) + 0.7f * (lane & 7);
    r.y = 4.0f + ((h >> 10) & 1023) * (span / 1024.0f) + 0.7f * (lane >> 3);
    r.z = 4.0f + ((h >> 20) & 1023) * (span / 1024.0f) * 0.5f;
    r.dx = 0.81f * 0.35f; r.dy = 0.81f * 0.25f; r.dz = 0.81f * 0.9f;
    return r;
}

__global__ void k_tex(cudaTextureObject_t tex, int E, int steps, float* out) {
    RayState r = make_ray(E, steps);
    float2 acc = make_float2(0.f, 1.f);
    for (int i = 0; i < steps; i++) {
        const float fx = floorf(r.x), fy = floorf(r.y), fz = floorf(r.z);
        const float wx = r.x - fx, wy = r.y - fy, wz = r.z - fz;
        const float4 t00 = tex3D<float4>(tex, fx + 0.5f, fy + 0.5f, fz + 0.5f), t10 = tex3D<float4>(tex, fx + 1.5f, fy + 0.5f, fz + 0.5f);
        const float4 t01 = tex3D<float4>(tex, fx + 0.5f, fy + 1.5f, fz + 0.5f), t11 = tex3D<float4>(tex, fx + 1.5f, fy + 1.5f, fz + 0.5f);
        float2 a0 = lerp2(make_float2(t00.x, t00.y), make_float2(t10.x, t10.y), wx), a1 = lerp2(make_float2(t01.x, t01.y), make_float2(t11.x, t11.y), wx);
        float2 b0 = lerp2(make_float2(t00.z, t00.w), make_float2(t10.z, t10.w), wx), b1 = lerp2(make_float2(t01.z, t01.w), make_float2(t11.z, t11.w), wx);
        float2 v = lerp2(lerp2(a0, a1, wy), lerp2(b0, b1, wy), wz);
        const float bl = __frcp_rn(1.0f + v.y);
        acc.x = fmaf(bl, acc.x - v.x, v.x); acc.y *= bl;
        r.x += r.dx; r.y += r.dy; r.z += r.dz;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y;
}

__global__ void k_ldg(const uint2* __restrict__ vol, int E, int steps, float* out) {
    RayState r = make_ray(E, steps);
    float2 acc = make_float2(0.f, 1.f);
    const size_t RS = E, SS = (size_t)E * E;
    for (int i = 0; i < steps; i++) {
        const float fx = floorf(r.x), fy = floorf(r.y), fz = floorf(r.z);
        const float wx = r.x - fx, wy = r.y - fy, wz = r.z - fz;
        const uint2* p = vol + (size_t)fz * SS + (size_t)fy * RS + (size_t)fx;
        const uint2 u00 = __ldg(p), u10 = __ldg(p + 1), u01 = __ldg(p + RS), u11 = __ldg(p + RS + 1);
        auto lo = [](uint2 t) { return __half22float2(*reinterpret_cast<const __half2*>(&t.x)); };
        auto hi = [](uint2 t) { return __half22float2(*reinterpret_cast<const __half2*>(&t.y)); };
        float2 a0 = lerp2(lo(u00), lo(u10), wx), a1 = lerp2(lo(u01), lo(u11), wx);
        float2 b0 = lerp2(hi(u00), hi(u10), wx), b1 = lerp2(hi(u01), hi(u11), wx);
        float2 v = lerp2(lerp2(a0, a1, wy), lerp2(b0, b1, wy), wz);
        const float bl = __frcp_rn(1.0f + v.y);
        acc.x = fmaf(bl, acc.x - v.x, v.x); acc.y *= bl;
        r.x += r.dx; r.y += r.dy; r.z += r.dz;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y;
}

int main(int argc, char** argv) {
    const int E = argc > 1 ? atoi(argv[1]) : 384, steps = 256;
    const size_t n = (size_t)E * E * E;
    std::vector<uint2> h(n);
    for (size_t i = 0; i < n; i++) {
        __half2 a = __floats2half2_rn((float)(i % 97) / 97.0f, (float)(i % 31) / 310.0f), b = __floats2half2_rn((float)((i + E * E) % 97) / 97.0f, (float)((i + E * E) % 31) / 310.0f);
        h[i].x = *reinterpret_cast<unsigned*>(&a); h[i].y = *reinterpret_cast<unsigned*>(&b);
    }
    uint2* dlin; CK(cudaMalloc(&dlin, n * 8)); CK(cudaMemcpy(dlin, h.data(), n * 8, cudaMemcpyHostToDevice));
    cudaChannelFormatDesc cd = cudaCreateChannelDescHalf4();
    cudaArray_t arr; CK(cudaMalloc3DArray(&arr, &cd, make_cudaExtent(E, E, E)));
    cudaMemcpy3DParms cp = {}; cp.srcPtr = make_cudaPitchedPtr(h.data(), (size_t)E * 8, E, E); cp.dstArray = arr; cp.extent = make_cudaExtent(E, E, E); cp.kind = cudaMemcpyHostToDevice;
    CK(cudaMemcpy3D(&cp));
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
    cudaTextureDesc td = {}; td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp; td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
    cudaTextureObject_t tex; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    const int blocks = 148 * 40, threads = 128;
    float* out; CK(cudaMalloc(&out, (size_t)blocks * threads * 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double warpSamples = (double)blocks * threads / 32 * steps;
    for (int which = 0; which < 2; which++) {
        float best = 1e9f;
        for (int it = 0; it < 5; it++) {
            cudaEventRecord(e0);
            if (which == 0) k_tex<<<blocks, threads>>>(tex, E, steps, out); else k_ldg<<<blocks, threads>>>(dlin, E, steps, out);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("%s  edge %d: %.3f ms, %.3f ns per warp-sample, %.1f SM-cycles per warp-sample (148 SMs, 1.965 GHz)\n", which == 0 ? "tex" : "ldg", E, best,
               best * 1e6 / warpSamples, best * 1e-3 * 1.965e9 * 148 / warpSamples);
    }
    std::vector<float> ho((size_t)blocks * threads); CK(cudaMemcpy(ho.data(), out, ho.size() * 4, cudaMemcpyDeviceToHost));
    double s = 0; for (float v : ho) s += v; printf("checksum %.3f\n", s);
    return 0;
}
