// vpe_sweep_tma.cuh — the light sweep (Fill.shader:211-269) as a TMA + mbarrier pipeline.
//
// The sweep is the one HBM-bound kernel of the path: 8 B read + 8 B written per voxel, a few flops in between. The
// register-staged k_sweep_columns keeps 16 slices per thread in flight (126 registers, one CTA per SM). Here the tiles move
// with the Tensor Memory Accelerator instead: one elected thread issues cp.async.bulk.tensor loads of [S slices][rows][N]
// boxes of a brick into a ring of shared-memory stages (completion on an mbarrier), all 256 threads turn their own voxel
// column of the stage into final texels in place, and the stage goes back with one cp.async.bulk.tensor store. Bytes in
// flight are bounded by shared memory, not registers: ~40 registers per thread, so the kernel can share an SM with the
// density pass (k_sweep_tma<.., OVERLAP = true> is the persistent variant that runs concurrently with it).
// Arithmetic and its order are those of k_sweep_columns / k_fill_columns: the volume stays bit-identical.
#pragma once
#include <cuda.h>

#include "vpe_tma.cuh"

namespace vpe {

constexpr int TMA_S = 8;  // slices per stage

// Chunks of one block of voxel columns, in sweep order: the covered metavoxels of the column, nearest the light first, S slices at a time
struct ChunkIter {
    int zz, k0, entry;
    __device__ __forceinline__ void seek(const GridParams& g, const int* __restrict__ brickOf, int cellBase) {  // first covered metavoxel at or after zz
        entry = -1;
        while (zz < g.z1) {
            entry = __ldg(brickOf + zz * g.NX * g.NY + cellBase);
            if (entry >= 0) break;
            zz++;
        }
    }
    __device__ __forceinline__ bool valid(const GridParams& g) const { return zz < g.z1; }
    __device__ __forceinline__ void next(const GridParams& g, const int* __restrict__ brickOf, int cellBase, int N) {
        k0 += TMA_S;
        if (k0 >= N) { k0 = 0; zz++; seek(g, brickOf, cellBase); }
    }
};

// NT = voxels per metavoxel edge (32 or 64): a CTA of 256 threads owns NT x (256 / NT) voxel columns, the same block of
// columns as k_fill_columns / k_sweep_columns (so the density flags and the sheet link use the same block numbering);
// thread t owns column (t % NT, t / NT) of the block: a warp reads 32 consecutive texels of a stage row, conflict-free.
// Requires N % TMA_S == 0 (true for 32, 64).
template <int NT, bool GRAY, bool LINKED, bool OVERLAP, int STAGES>
__global__ void __launch_bounds__(FILLC_THREADS, OVERLAP ? 4 : 2)
k_sweep_tma(const __grid_constant__ CUtensorMap tmap, GridParams g, FillArgs a, const int* __restrict__ brickOf, SheetLink link, int gridX, int numBlocks) {
    constexpr int ROWS = FILLC_THREADS / NT;
    constexpr int SLICE_TEXELS = ROWS * NT;                 // 256
    constexpr int STAGE_TEXELS = TMA_S * SLICE_TEXELS;
    constexpr unsigned STAGE_BYTES = STAGE_TEXELS * 8u;
    constexpr int LOAD_SHIFT = GRAY ? 1 : 0;  // GRAY boxes are loaded one slice ahead of the slices they are stored to (see the slice loop)
    extern __shared__ __align__(128) unsigned char tmaSmem[];
    uint2* const buf = reinterpret_cast<uint2*>(tmaSmem);   // [STAGES][TMA_S][ROWS][NT]
    __shared__ unsigned long long full[STAGES];
    const int tid = threadIdx.x;
    const int px = tid % NT, row = tid / NT;
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int N = NT;
    const float Nf = g.Nf;
    const int borderVoxelIndex = N - g.border;
    const size_t NN = (size_t)N * g.rowStride;
    const int rw = a.x1 - a.x0;
    unsigned G = 0;  // chunks consumed by this CTA so far: stage = G % STAGES, barrier parity = (G / STAGES) & 1
    for (int block = blockIdx.x; block < numBlocks; block += gridDim.x) {
        const int bx = block % gridX, by = block / gridX;
        const int xx = a.x0 + bx % rw, yy = a.y0 + bx / rw;
        const int py = by * ROWS + row;
        const int cellBase = yy * g.NX + xx;
        // the column's constants, as column_thread computes them (Fill.shader:100-105, 214-218)
        const float posx = (float)px + 0.5f, posy = (float)py + 0.5f;
        const F3 nrm = f3((posx - Nf / 2.0f) / Nf, (posy - Nf / 2.0f) / Nf, (0.0f - Nf / 2.0f) / Nf);
        const float lx = (g.Ab.m[0][0] * nrm.x + g.Ab.m[0][1] * nrm.y) + g.Ab.m[0][2] * nrm.z;
        const float ly = (g.Ab.m[1][0] * nrm.x + g.Ab.m[1][1] * nrm.y) + g.Ab.m[1][2] * nrm.z;
        const float lz = (g.Ab.m[2][0] * nrm.x + g.Ab.m[2][1] * nrm.y) + g.Ab.m[2][2] * nrm.z;
        const size_t sheetIdx = (size_t)(py + yy * N) * (size_t)(g.NX * N) + (size_t)(px + xx * N);
        float dmap = 1.0f;
        if (a.depth) {
            const float u = (posx + (float)xx * Nf) / ((float)g.NX * Nf), v = (posy + (float)yy * Nf) / ((float)g.NY * Nf);
            dmap = sample_depth(a.depth, g.NX * N, g.NY * N, u, v);
        }
        const float lsSceneDepth = (dmap - g.depthB) * g.depthRcpA;
        // dependencies of this block: its own densities (overlapped with the density pass), the upstream slab's sheet values
        if (OVERLAP) {
            if (tid == 0) {
                const long long t0 = clock64();
                while ((int)(ld_acquire_gpu(a.densityDone + block) - a.densityEpoch) < 0) {
                    if (clock64() - t0 > link.spinLimit) { atomicAdd(link.timeouts, 1u); atomicAdd(link.timeouts + 1, 1u); break; }
                    __nanosleep(256);
                }
                fence_proxy_async();  // the densities were written through the generic proxy; the TMA reads them through the async proxy
            }
            __syncthreads();
        }
        float incoming = 1.0f;  // the cleared sheet (VPR.cs:498-499)
        if (LINKED && link.hasUp) {
            link_wait(link.flagIn + block, link.epoch, link);
            incoming = ld_relaxed_sys(link.inbox + sheetIdx);
            __syncthreads();  // every thread has its value: the upstream rank may reuse the inbox block
            if (tid == 0) st_release_sys(link.upAck + block, link.epoch);
        }
        // ---- the pipeline over this block's chunks ----
        ChunkIter itc;
        itc.zz = g.z0; itc.k0 = 0;
        itc.seek(g, brickOf, cellBase);
        ChunkIter itp = itc;  // producer side (thread 0): the next chunk to load
        if (tid == 0) {
            tma_wait_read<0>();  // stores of the previous block have left their stages
            for (int i = 0; i < STAGES - 1 && itp.valid(g); i++) {
                const unsigned s = (G + i) % STAGES;
                mbar_expect_tx(&full[s], STAGE_BYTES);
                tma_load_4d(buf + (size_t)s * STAGE_TEXELS, &tmap, &full[s], 0, by * ROWS, itp.k0 + LOAD_SHIFT, itp.entry);
                itp.next(g, brickOf, cellBase, N);
            }
        }
        float carried = 0.0f, transmitted = 0.0f, propagated = 0.0f;
        bool haveCarried = false;
        int shadowIndex = 0;
        unsigned prevWord = 0;
        while (itc.valid(g)) {
            const unsigned s = G % STAGES;
            if (itc.k0 == 0) {  // a new metavoxel: Fill.shader:211-229
                const F3 c = mv_center(g, xx, yy, itc.zz);
                const F3 voxel0 = f3(lx + c.x, ly + c.y, lz + c.z);
                const float lsZ = ((g.w2lcRow2[0] * voxel0.x + g.w2lcRow2[1] * voxel0.y) + g.w2lcRow2[2] * voxel0.z) + g.w2lcRow2[3];
                shadowIndex = ftoi_sat((lsSceneDepth - lsZ) / g.oneVoxelSize);
                transmitted = (itc.zz == 0) ? 1.0f : (haveCarried ? carried : (LINKED ? incoming : a.sheet[sheetIdx]));
                propagated = transmitted;
                if (GRAY) {  // slice 0 is not in the (shifted) boxes: every thread sweeps its own texel of it from global memory
                    const uint2 t = __ldcg(a.bricks + (size_t)itc.entry * NN * N + (size_t)py * g.rowStride + px);
                    prevWord = sweep_voxel_gray(g, 0, shadowIndex, borderVoxelIndex, __uint_as_float(t.x), __uint_as_float(t.y), transmitted, propagated);
                }
            }
            mbar_wait(&full[s], (G / STAGES) & 1u);
            uint2* const sb = buf + (size_t)s * STAGE_TEXELS + row * NT + px;
            const int entry = itc.entry, k0 = itc.k0;
            const bool lastOfBrick = k0 + TMA_S >= N;
#pragma unroll
            for (int j = 0; j < TMA_S; j++) {
                // GRAY: z-paired (r, density) texels. The box was loaded one slice ahead (slot j = the intermediate of slice
                // k0 + j + 1), so that slot j can become the final texel of slice k0 + j = {word[k0 + j], word[k0 + j + 1]};
                // slice N of the last box is out of bounds (the TMA fills zeros): the pair's upper half there is never sampled.
                const int slice = k0 + j + LOAD_SHIFT;
                const uint2 t = sb[j * SLICE_TEXELS];
                if (GRAY) {
                    unsigned word = 0u;
                    if (slice < N)
                        word = sweep_voxel_gray(g, slice, shadowIndex, borderVoxelIndex, __uint_as_float(t.x), __uint_as_float(t.y), transmitted, propagated);
                    sb[j * SLICE_TEXELS] = make_uint2(prevWord, word);
                    prevWord = word;
                } else {
                    sb[j * SLICE_TEXELS] = sweep_voxel(g, slice, shadowIndex, borderVoxelIndex, __uint_as_float(t.x), __uint_as_float(t.y), transmitted, propagated);
                }
            }
            if (lastOfBrick) {
                carried = propagated;  // Fill.shader:250
                haveCarried = true;
            }
            fence_proxy_async();  // the stage was written through the generic proxy, the store reads it through the async proxy
            __syncthreads();
            if (tid == 0) {
                tma_store_4d(&tmap, buf + (size_t)s * STAGE_TEXELS, 0, by * ROWS, k0, entry);
                tma_commit();
                if (itp.valid(g)) {
                    tma_wait_read<1>();  // the store of the previous chunk has left the stage that is loaded next
                    const unsigned sn = (G + STAGES - 1) % STAGES;
                    mbar_expect_tx(&full[sn], STAGE_BYTES);
                    tma_load_4d(buf + (size_t)sn * STAGE_TEXELS, &tmap, &full[sn], 0, by * ROWS, itp.k0 + LOAD_SHIFT, itp.entry);
                    itp.next(g, brickOf, cellBase, N);
                }
            }
            G++;
            itc.next(g, brickOf, cellBase, N);
        }
        if (!LINKED) {
            if (haveCarried) a.sheet[sheetIdx] = carried;
        } else {
            // the sheet as this slab leaves it: untouched columns pass the incoming value on
            const float outgoing = haveCarried ? carried : incoming;
            a.sheet[sheetIdx] = outgoing;
            if (link.hasDown) {
                link_wait(link.ackIn + block, link.epoch - 1u, link, 3);  // the previous fill's values have been read
                link.downInbox[sheetIdx] = outgoing;
                __threadfence_system();
                __syncthreads();
                if (tid == 0) st_release_sys(link.downFlag + block, link.epoch);
            }
        }
        __syncthreads();
    }
    if (tid == 0) tma_wait_all();
}

}  // namespace vpe
