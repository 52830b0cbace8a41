// vpe_cuda.cu — libvpe_cuda.so: the C-ABI of include/vpe.h over hand-written sm_100a kernels.
//
// Host orchestration mirrors the per-frame loop of VolumetricParticleRenderer (VPR.cs:181-220):
//   vpe_fill   ≙ BinParticlesToMetavoxels + FillMetavoxels   (VPR.cs:397-520)
//   vpe_march  ≙ RenderMetavoxels                            (VPR.cs:637-713)
// but instead of one draw call per metavoxel it issues a handful of launches: particle set-up and
// counting-sort binning, then ONE fill launch in which every voxel column walks all light-axis slices
// (the light dependency is carried in a register), then one march launch for the whole image that walks
// the metavoxels per ray in the reference's submission order (the host-buffer path launches the image in
// bands so that the copy home overlaps).  Multi-GPU entry points split the fill at the light sheet and
// hand it over through peer memory (vpe_sheet_link_*).  There is NO CPU fallback: every entry point that
// computes runs on the GPU.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/vpe.h"
#include "vpe_kernels.cuh"

using namespace vpe;

#define CUDA_TRY(ctx, expr)                                                                          \
    do {                                                                                             \
        cudaError_t e_ = (expr);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            char buf_[512];                                                                          \
            snprintf(buf_, sizeof(buf_), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
            (ctx)->err = buf_;                                                                       \
            return e_ == cudaErrorMemoryAllocation ? VPE_E_OUT_OF_MEMORY : VPE_E_CUDA;               \
        }                                                                                            \
    } while (0)

// Entry points run on the context's device and leave the caller's current device as they found it.
struct DeviceScope {
    int prev = -1;
    bool ok = true;
    explicit DeviceScope(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != device) ok = cudaSetDevice(device) == cudaSuccess;
        else prev = -1;
    }
    ~DeviceScope() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;  // elements
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, n * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct VpeContext {
    VpeConfig cfg;
    std::string err;
    int device = 0;
    cudaStream_t stream = nullptr;
    bool ownStream = false;
    // state
    VpeTransform light;
    float center[3] = {0, 0, 0};
    bool lightSet = false, cubeSet = false, prepared = false, filledOnce = false;
    GridParams g;
    M3 lightRot;
    F3 lightFwdRaw;
    // device buffers
    DevBuf<float> dParticles;        // n*7 staging for host particle input
    DevBuf<ParticleFill> dPfill;
    DevBuf<ParticleBin> dPbin;
    DevBuf<int> dCellCount, dCellStart, dBrickOf, dCovered, dSliceStart, dPairs, dTotals;
    DevBuf<int2> dBlockSums;
    DevBuf<float> dCube, dDepth, dSheet;
    DevBuf<float4> dCubeFp;          // bilinear footprints of the cubemap, [6][E+1][E+1]
    DevBuf<uint2> dBricks;
    DevBuf<unsigned> dNz, dOcc;      // dNz: per brick [warp tile][z] words, non-zero density bits of the fill's 8x4 tiles; dOcc: per brick [z][y] rows of
                                     // occRowWords words, bit x: occupied sample base (k_occ_build)
    int occRowWords = 0;             // ceil(N / 32)
    VpeDebugOptions dbg;             // vpe_set_debug_options (all zero = production behaviour)
    DevBuf<float4> dMvCam;
    DevBuf<int> dRank, dPixels, dSamples;
    DevBuf<float4> dImage, dImage2;
    DevBuf<unsigned long long> dTotalSamples;
    DevBuf<unsigned long long> dSliceSamples;   // [NZ] ray samples per slice of the last march (VpeDebugOptions.profileSlices)
    std::vector<int> slicePairs;                // [NZ + 1] prefix of (particle, metavoxel) pairs per slice of the last fill
    DevBuf<const float4*> dParts;
    // sheet link (multi-GPU sweep over peer memory): own buffer [inbox | flagIn | ackIn | timeouts] and the
    // neighbours' buffers mapped into this process
    void* linkOwn = nullptr;
    void* linkUp = nullptr;
    void* linkDown = nullptr;
    bool linkUpIpc = false, linkDownIpc = false;
    int linkBlocks = 0;
    unsigned linkEpoch = 0;
    // linked sweep overlapped with the density pass: per-block "densities in place" flags, a high-priority stream for the
    // persistent sweep kernel, and the events that order it against the context's stream
    DevBuf<unsigned> dDensityDone;
    unsigned densityEpoch = 0;
    bool densitySignalled = false;   // the density pass of the current fill raises the flags
    cudaStream_t sweepStream = nullptr;
    cudaEvent_t evDensityBegin = nullptr, evSweepDone = nullptr;
    int numSMs = 148;
    // brick pool as a TMA tensor (x, y, z, brick) of 8-byte texels, for k_sweep_tma; re-encoded when the pool moves
    CUtensorMap brickMap;
    const void* brickMapBase = nullptr;
    size_t brickMapBricks = 0;
    int brickMapRowStride = 0;
    bool brickMapOk = false;
    // image link (multi-GPU march): receive buffer [2 parities][over|under][slab][row][col] float4 + flags, and
    // every rank's buffer mapped into this process
    void* imgOwn = nullptr;
    void* imgPeers[64] = {};
    bool imgPeerIpc[64] = {};
    DevBuf<float4*> dImgPeers;
    int imgWorld = 0, imgRank = 0, imgW = 0, imgH = 0, imgPer = 0;
    size_t imgFlagsOff = 0, imgTimeoutOff = 0, imgBytes = 0;
    unsigned imgEpoch = 0;
    unsigned imgKinds = 3;           // partials of the last linked march: bit 0 OVER, bit 1 UNDER
    bool imgCompositePending = false;  // vpe_march_linked has run, its vpe_composite_linked has not
    bool imgConnected = false;
    // host-path march: bands on two auxiliary streams, each copied home as soon as it is done
    cudaStream_t aux[3] = {nullptr, nullptr, nullptr};   // two compute streams + one copy stream
    cudaEvent_t evFork = nullptr, evJoin[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t evBand[32] = {};
    // around the path: light camera, march options
    Affine w2lc;                     // lightCamera.transform.worldToLocalMatrix (VPR.cs:365-366)
    int targetFormat = 0, debugMode = 0;
    DevBuf<float> dSceneDepth;
    int sceneW = 0, sceneH = 0;
    bool sceneDepthSet = false;
    DevBuf<int> dOrderOf;
    DevBuf<float> dTris;
    DevBuf<float4> dScene;
    bool depthSet = false;
    bool bricksGray = false;         // layout of the bricks of the last fill (GridParams::gray at that time)
    int cubeEdge = 0;
    // pinned host staging
    int* hCounts = nullptr;  // [NZ + 3]: sliceStart[NZ+1], totals[2]
    unsigned long long* hTotalSamples = nullptr;
    std::vector<int> sliceStart;
    int nParticles = 0, nCovered = 0, nPairs = 0;
    int numCells = 0;
    // timing
    cudaEvent_t evFill0 = nullptr, evFill1 = nullptr, evMarch0 = nullptr, evMarch1 = nullptr;
    cudaEvent_t evFillK0 = nullptr, evMarchK0 = nullptr, evMarchK1 = nullptr;
    bool fillTimed = false, marchTimed = false;
    VpeStats stats;
};

namespace {

int fail(VpeContext* c, int code, const char* msg) {
    if (c) c->err = msg;
    return code;
}

int validate_config(const VpeConfig& c, std::string& why) {
    if (c.numMetavoxelsX < 1 || c.numMetavoxelsY < 1 || c.numMetavoxelsZ < 1) { why = "grid dims must be >= 1"; return -1; }
    if ((long long)c.numMetavoxelsX * c.numMetavoxelsY * c.numMetavoxelsZ > (1ll << 26)) { why = "grid too large"; return -1; }
    if (c.numVoxelsInMetavoxel < 2 || c.numVoxelsInMetavoxel > 256) { why = "numVoxelsInMetavoxel must be in [2,256]"; return -1; }
    // fill clamps the border to [0,N-2] (VPR.cs:528), the march does not (VPR.cs:726)
    if (c.numBorderVoxels < 0 || c.numBorderVoxels > (c.numVoxelsInMetavoxel - 2) / 2) { why = "numBorderVoxels out of range"; return -1; }
    if (!(c.mvScale > 0.0f)) { why = "mvScale must be > 0"; return -1; }
    if (c.rayMarchSteps < 1 || c.rayMarchSteps > 32768) { why = "rayMarchSteps must be in [1, 32768]"; return -1; }
    // values that would divide by zero or poison every voxel with NaN later on
    if (!(c.lightFar > c.lightNear) || !std::isfinite(c.lightFar) || !std::isfinite(c.lightNear)) { why = "lightFar must be > lightNear (finite)"; return -1; }
    if (c.softParticleStepDistance < 0) { why = "softParticleStepDistance must be >= 0"; return -1; }
    if (!std::isfinite(c.mvScale) || !std::isfinite(c.opacityFactor) || !std::isfinite(c.displacementScale) || !std::isfinite(c.lightCameraDistance) ||
        !std::isfinite(c.ambientColor[0]) || !std::isfinite(c.ambientColor[1]) || !std::isfinite(c.ambientColor[2]) ||
        !std::isfinite(c.marchEarlyOutTransmittance)) { why = "non-finite float in the configuration"; return -1; }
    if (c.slabZBegin < 0 || c.slabZEnd > c.numMetavoxelsZ || c.slabZBegin >= c.slabZEnd) { why = "bad slab range"; return -1; }
    return 0;
}

void normalise_slab(VpeConfig& c) {
    if (c.slabZBegin == 0 && c.slabZEnd == 0) c.slabZEnd = c.numMetavoxelsZ;
}

F3 forward_of(const M3& r) { return f3(r.m[0][2], r.m[1][2], r.m[2][2]); }

// Everything the kernels need that depends only on config + light (VPR.cs:139, 361-394, 523-554).
void rebuild_grid_params(VpeContext* c) {
    GridParams& g = c->g;
    const VpeConfig& k = c->cfg;
    memset(&g, 0, sizeof(g));
    g.NX = k.numMetavoxelsX; g.NY = k.numMetavoxelsY; g.NZ = k.numMetavoxelsZ;
    g.N = k.numVoxelsInMetavoxel;
    g.border = std::min(std::max(k.numBorderVoxels, 0), g.N - 2);  // VPR.cs:528
    g.z0 = k.slabZBegin; g.z1 = k.slabZEnd;
    g.binMode = k.binMode;
    g.s = k.mvScale;
    g.sb = k.mvScale * (float)g.N / (float)(g.N - 2 * k.numBorderVoxels);  // VPR.cs:139
    g.Nf = (float)g.N;
    g.center = f3(c->center[0], c->center[1], c->center[2]);
    c->lightRot = quat_to_m3(c->light.rotation);
    g.L2W = trs(f3(c->light.position[0], c->light.position[1], c->light.position[2]), c->lightRot, 1.0f);
    g.W2L = affine_inverse(g.L2W);
    g.lsCenter = xform_point(g.W2L, g.center);  // VPR.cs:380
    c->lightFwdRaw = forward_of(c->lightRot);
    {
        // Vector3.normalized (VPR.cs:535)
        float mag = sqrtf(dot3(c->lightFwdRaw, c->lightFwdRaw));
        g.lightFwd = mag > 1e-5f ? f3(c->lightFwdRaw.x / mag, c->lightFwdRaw.y / mag, c->lightFwdRaw.z / mag) : f3(0, 0, 0);
    }
    g.Ab = trs(f3(0, 0, 0), c->lightRot, g.sb);
    g.Lb = affine_inverse_linear(g.Ab);
    g.As = trs(f3(0, 0, 0), c->lightRot, g.s);
    g.Ls = affine_inverse_linear(g.As);
    {
        // lightCamera.transform: position = centre - forward * 200, the light's rotation (VPR.cs:365-366)
        F3 lc = f3(g.center.x - c->lightFwdRaw.x * k.lightCameraDistance, g.center.y - c->lightFwdRaw.y * k.lightCameraDistance,
                   g.center.z - c->lightFwdRaw.z * k.lightCameraDistance);
        Affine w2lc = affine_inverse(trs(lc, c->lightRot, 1.0f));
        c->w2lc = w2lc;
        for (int j = 0; j < 4; j++) g.w2lcRow2[j] = w2lc.m[2][j];
    }
    g.oneVoxelSize = g.sb / g.Nf;  // Fill.shader:160
    g.lightStep = f3(g.lightFwd.x * g.oneVoxelSize, g.lightFwd.y * g.oneVoxelSize, g.lightFwd.z * g.oneVoxelSize);
    for (int i = 0; i < 3; i++) g.ambient[i] = k.ambientColor[i];
    g.ds = k.displacementScale;
    g.opacityFactor = k.opacityFactor;
    g.fade = k.fadeOutParticles;
    {
        float a = 1.0f / (k.lightFar - k.lightNear);  // Fill.shader:217
        g.depthB = -k.lightNear * a;
        g.depthRcpA = 1.0f / a;
    }
    g.cubeEdge = c->cubeEdge;
    {
        const float maxG = (float)std::max(g.NX, std::max(g.NY, g.NZ));
        const float cmax = std::max(fabsf(g.center.x), std::max(fabsf(g.center.y), fabsf(g.center.z)));
        g.worldReach = cmax + 0.5f * 1.7320508f * (maxG + 2.0f) * g.sb * 1.01f;
    }
    g.rowStride = g.N + ((g.N % 16 == 0 && !c->dbg.noRowPad) ? ROW_PAD : 0);
    // r == g == b in every texel iff the three ambient components are the same bits (Fill.shader:244)
    g.gray = (memcmp(&k.ambientColor[0], &k.ambientColor[1], sizeof(float)) == 0 &&
              memcmp(&k.ambientColor[1], &k.ambientColor[2], sizeof(float)) == 0 && !c->dbg.noGray) ? 1 : 0;
}

int sync_stream(VpeContext* c) {
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return VPE_OK;
}

inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }
#ifndef VPE_TMA_STAGES_OVERLAP
#define VPE_TMA_STAGES_OVERLAP 3
#endif
#ifndef VPE_TMA_STAGES_ALONE
#define VPE_TMA_STAGES_ALONE 4
#endif
constexpr int TMA_STAGES_OVERLAP = VPE_TMA_STAGES_OVERLAP, TMA_STAGES_ALONE = VPE_TMA_STAGES_ALONE;  // shared-memory stages of k_sweep_tma (16 KB each)

// CUDA loads kernels lazily, and loading one may wait until the device is idle. This library runs kernels that spin on
// flags raised by other kernels (sheet link, image link, the overlapped sweep): if the producer is a kernel that still has
// to be loaded while the consumer already spins, the load waits for the consumer and the consumer for the producer. Every
// kernel is therefore loaded when the first context of a device is created, before anything can spin.
template <class K>
void preload(K kernel) {
    cudaFuncAttributes at;
    if (cudaFuncGetAttributes(&at, kernel) != cudaSuccess) cudaGetLastError();
}
// Kernels that wait on flags raised by other kernels of the same GPU must be able to share an SM with the producer. The
// shared-memory / L1 split of an SM can only change while the SM is idle: a resident spinning CTA of a kernel that asked for
// "no shared memory" would keep every CTA of a producer that needs shared memory (k_fill_columns: 36 KB) off its SM - on all
// SMs at once for a persistent kernel, i.e. for as long as it spins. All of them therefore ask for the same carve-out.
template <class K>
void prefer_max_shared(K kernel) {
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared) != cudaSuccess) cudaGetLastError();
}
template <int NT>
void preload_march() {
#define VPE_PRELOAD_MARCH(S, G, P) preload(k_march_flat<NT, S, G, P>); preload(k_march<NT, false, S, G, P>);
    VPE_PRELOAD_MARCH(false, false, false) VPE_PRELOAD_MARCH(false, false, true) VPE_PRELOAD_MARCH(false, true, false) VPE_PRELOAD_MARCH(false, true, true)
    VPE_PRELOAD_MARCH(true, false, false) VPE_PRELOAD_MARCH(true, false, true) VPE_PRELOAD_MARCH(true, true, false) VPE_PRELOAD_MARCH(true, true, true)
#undef VPE_PRELOAD_MARCH
}
void preload_kernels(int device) {
    static bool done[64] = {};
    if (device < 0 || device >= 64 || done[device]) return;
    done[device] = true;
    preload(k_particle_setup); preload(k_scatter_pairs); preload(k_sort_lists); preload(k_scan_reduce); preload(k_scan_blocks);
    preload(k_scan_final); preload(k_fill_value); preload(k_occ_build); preload(k_mv_camera); preload(k_popcount);
    preload(k_fill_columns<false, false>); preload(k_fill_columns<false, true>); preload(k_fill_columns<true, false>);
    preload(k_fill_columns<false, false, true>); preload(k_fill_columns<false, true, true>);
    preload(k_sweep_columns<false, false>); preload(k_sweep_columns<true, false>); preload(k_sweep_columns<false, true>);
    preload(k_sweep_columns<true, true>); preload(k_sweep_overlapped<false>); preload(k_sweep_overlapped<true>);
    preload_march<0>(); preload_march<32>(); preload_march<64>();
    preload(k_march<-1, true, false, false, false>); preload(k_march<-1, false, false, false, false>);
    preload(k_image_signal); preload(k_composite_linked); preload(k_composite); preload(k_raster_depth);
    preload(k_composite_scene); preload(k_order_index);
    prefer_max_shared(k_fill_columns<false, false>); prefer_max_shared(k_fill_columns<false, true>); prefer_max_shared(k_fill_columns<true, false>);
    prefer_max_shared(k_fill_columns<false, false, true>); prefer_max_shared(k_fill_columns<false, true, true>);
    prefer_max_shared(k_sweep_columns<false, true>); prefer_max_shared(k_sweep_columns<true, true>);
    prefer_max_shared(k_sweep_overlapped<false>); prefer_max_shared(k_sweep_overlapped<true>);
    prefer_max_shared(k_composite_linked); prefer_max_shared(k_occ_build);
#define VPE_TMA_ATTR(NT, GRAY, LINKED, OVERLAP)                                                                                          \
    {                                                                                                                                   \
        constexpr int ST = OVERLAP ? TMA_STAGES_OVERLAP : TMA_STAGES_ALONE;                                                             \
        auto k = k_sweep_tma<NT, GRAY, LINKED, OVERLAP, ST>;                                                                            \
        preload(k);                                                                                                                     \
        if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, ST * TMA_S * FILLC_THREADS * 8) != cudaSuccess) cudaGetLastError(); \
        prefer_max_shared(k);                                                                                                           \
    }
#define VPE_TMA_ATTR4(NT, GRAY) VPE_TMA_ATTR(NT, GRAY, false, false) VPE_TMA_ATTR(NT, GRAY, true, false) VPE_TMA_ATTR(NT, GRAY, true, true)
    VPE_TMA_ATTR4(32, false) VPE_TMA_ATTR4(32, true) VPE_TMA_ATTR4(64, false) VPE_TMA_ATTR4(64, true)
#undef VPE_TMA_ATTR4
#undef VPE_TMA_ATTR
}

void update_brick_map(VpeContext* c);

// ---- fill -------------------------------------------------------------------------------------
int fill_prepare_impl(VpeContext* c, const float* particlesDev, int n, const VpeTransform* em) {
    GridParams& g = c->g;
    const int cells = c->numCells;
    CUDA_TRY(c, cudaEventRecord(c->evFill0, c->stream));
    c->fillTimed = false;
    c->stats.fillLaunches = 0;
    CUDA_TRY(c, c->dPfill.ensure(std::max(n, 1)));
    CUDA_TRY(c, c->dPbin.ensure(std::max(n, 1)));
    CUDA_TRY(c, cudaMemsetAsync(c->dCellCount.p, 0, sizeof(int) * cells, c->stream));
    EmitterParams ep;
    M3 er = quat_to_m3(em->rotation);
    ep.E2W = trs(f3(em->position[0], em->position[1], em->position[2]), er, 1.0f);
    ep.forward = forward_of(er);
    if (n > 0) {
        k_particle_setup<<<div_up(n, 128), 128, 0, c->stream>>>(g, ep, particlesDev, n, c->dPfill.p, c->dPbin.p, c->dCellCount.p);
        c->stats.fillLaunches++;
    }
    const int nb = div_up(cells, SCAN_BLOCK);
    k_scan_reduce<<<nb, SCAN_BLOCK, 0, c->stream>>>(c->dCellCount.p, cells, c->dBlockSums.p);
    k_scan_blocks<<<1, SCAN_BLOCK, 0, c->stream>>>(c->dBlockSums.p, nb, c->dTotals.p);
    k_scan_final<<<nb, SCAN_BLOCK, 0, c->stream>>>(c->dCellCount.p, cells, g.NX * g.NY, g.NZ, c->dBlockSums.p, c->dTotals.p,
                                                  c->dCellStart.p, c->dBrickOf.p, c->dCovered.p, c->dSliceStart.p);
    c->stats.fillLaunches += 3;
    // the host needs the covered count (brick pool size) and the per-slice counts (launch grids)
    CUDA_TRY(c, cudaMemcpyAsync(c->hCounts, c->dSliceStart.p, sizeof(int) * (g.NZ + 1), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(c->hCounts + g.NZ + 1, c->dTotals.p, sizeof(int) * 2, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    c->sliceStart.assign(c->hCounts, c->hCounts + g.NZ + 1);
    if (c->dbg.profileSlices) {  // pairs before each slice = cellStart at the slice's first cell (load balancing of the slabs)
        c->slicePairs.assign((size_t)g.NZ + 1, 0);
        CUDA_TRY(c, cudaMemcpy2D(c->slicePairs.data(), sizeof(int), c->dCellStart.p, sizeof(int) * (size_t)g.NX * g.NY, sizeof(int), (size_t)g.NZ + 1,
                                 cudaMemcpyDeviceToHost));
    }
    c->nPairs = c->hCounts[g.NZ + 1];
    c->nCovered = c->hCounts[g.NZ + 2];
    c->nParticles = n;
    CUDA_TRY(c, c->dPairs.ensure(std::max(c->nPairs, 1)));
    if (n > 0 && c->nPairs > 0) {
        k_scatter_pairs<<<div_up(n, 128), 128, 0, c->stream>>>(g, c->dPbin.p, n, c->dCellCount.p, c->dCellStart.p, c->dPairs.p);
        k_sort_lists<<<div_up(c->nCovered, 128), 128, 0, c->stream>>>(c->dCovered.p, c->dTotals.p + 1, c->dCellStart.p, c->dPairs.p);
        c->stats.fillLaunches += 2;
    }
    // brick pool: one N^3 half4 brick per covered metavoxel (≙ lazily created mvFillTextures, VPR.cs:565-566)
    const size_t brickTexels = (size_t)g.N * g.N * g.rowStride;  // stored size (rows padded), DESIGN.md §4
    size_t want = (size_t)c->nCovered;
    if (want * brickTexels > c->dBricks.cap) {
        size_t padded = std::min((size_t)cells, want + want / 16 + 8);
        cudaError_t e = c->dBricks.ensure(padded * brickTexels);
        if (e != cudaSuccess) {
            cudaGetLastError();
            e = c->dBricks.ensure(want * brickTexels);
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(c, VPE_E_OUT_OF_MEMORY, "brick pool does not fit in device memory");
        }
    }
    update_brick_map(c);
    {
        // empty-space bitmaps: every word of a covered brick is rewritten by each fill (k_fill_columns / k_occ_build)
        const size_t words = std::max<size_t>(1, (size_t)c->nCovered * g.N * g.N * c->occRowWords);
        const size_t nzWords = std::max<size_t>(1, (size_t)c->nCovered * ((g.N + 7) / 8) * ((g.N + 3) / 4) * g.N);  // [brick][warp tile][z]
        if (words > c->dOcc.cap) CUDA_TRY(c, c->dOcc.ensure(words + words / 16));
        if (nzWords > c->dNz.cap) CUDA_TRY(c, c->dNz.ensure(nzWords + nzWords / 16));
    }
    // VPR.cs:498-499: clear the light propagation texture to 1
    const size_t sheetN = (size_t)g.NX * g.N * g.NY * g.N;
    k_fill_value<<<std::min(div_up(sheetN, 256), 148 * 8), 256, 0, c->stream>>>(c->dSheet.p, sheetN, 1.0f);
    c->stats.fillLaunches++;
    CUDA_TRY(c, cudaGetLastError());
    c->prepared = true;
    c->stats.numParticles = n;
    c->stats.numMetavoxelsCovered = c->nCovered;
    c->stats.numParticlePairs = c->nPairs;
    c->stats.voxelsFilled = (int64_t)c->nCovered * (int64_t)g.N * g.N * g.N;
    return VPE_OK;
}

enum FillPhase { FILL_FUSED, FILL_DENSITY, FILL_SWEEP, FILL_SWEEP_LINKED, FILL_FUSED_HEAD };

// cuTensorMapEncodeTiled through the runtime (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
        else cudaGetLastError();
    }
    return fn;
}

// The brick pool as the tensor k_sweep_tma moves: box = one block of voxel columns (N x 256/N) x TMA_S slices of one brick.
void update_brick_map(VpeContext* c) {
    const GridParams& g = c->g;
    const size_t brickTexels = (size_t)g.N * g.N * g.rowStride;
    const size_t bricks = brickTexels ? c->dBricks.cap / brickTexels : 0;
    if (c->brickMapOk && c->brickMapBase == c->dBricks.p && c->brickMapBricks == bricks && c->brickMapRowStride == g.rowStride) return;
    c->brickMapOk = false;
    c->brickMapBase = c->dBricks.p; c->brickMapBricks = bricks; c->brickMapRowStride = g.rowStride;
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc || bricks == 0 || (g.N != 32 && g.N != 64) || (g.rowStride * 8) % 16 != 0) return;
    const cuuint64_t dims[4] = {(cuuint64_t)g.rowStride, (cuuint64_t)g.N, (cuuint64_t)g.N, (cuuint64_t)bricks};
    const cuuint64_t strides[3] = {(cuuint64_t)g.rowStride * 8, (cuuint64_t)g.N * g.rowStride * 8, (cuuint64_t)brickTexels * 8};
    const cuuint32_t box[4] = {(cuuint32_t)g.N, (cuuint32_t)(FILLC_THREADS / g.N), (cuuint32_t)TMA_S, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    c->brickMapOk = enc(&c->brickMap, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, c->dBricks.p, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <bool LINKED, bool OVERLAP>
void launch_sweep_tma(VpeContext* c, const FillArgs& a, const SheetLink& link, dim3 grid, cudaStream_t st) {
    const GridParams& g = c->g;
    constexpr int STAGES = OVERLAP ? TMA_STAGES_OVERLAP : TMA_STAGES_ALONE;
    const size_t smem = (size_t)STAGES * TMA_S * FILLC_THREADS * 8;
    const int numBlocks = (int)(grid.x * grid.y);
    const int ctas = OVERLAP ? std::min(numBlocks, c->numSMs) : numBlocks;
#define VPE_SWEEP_TMA(NT, GRAY) k_sweep_tma<NT, GRAY, LINKED, OVERLAP, STAGES><<<ctas, FILLC_THREADS, smem, st>>>(c->brickMap, g, a, c->dBrickOf.p, link, (int)grid.x, numBlocks)
    if (g.N == 32) { if (g.gray) VPE_SWEEP_TMA(32, true); else VPE_SWEEP_TMA(32, false); }
    else { if (g.gray) VPE_SWEEP_TMA(64, true); else VPE_SWEEP_TMA(64, false); }
#undef VPE_SWEEP_TMA
}

// Layout of a sheet link buffer; identical on every rank (same grid and voxel count).
struct LinkLayout {
    size_t sheetN, flagOff, ackOff, timeoutOff, bytes;
    int blocks;
};
LinkLayout link_layout(const GridParams& g) {
    LinkLayout l;
    const int warpTiles = ((g.N + 7) / 8) * ((g.N + 3) / 4);
    l.blocks = g.NX * g.NY * ((warpTiles + FILLC_THREADS / 32 - 1) / (FILLC_THREADS / 32));
    l.sheetN = (size_t)g.NX * g.N * g.NY * g.N;
    l.flagOff = (l.sheetN * sizeof(float) + 255) / 256 * 256;
    l.ackOff = l.flagOff + (size_t)l.blocks * sizeof(unsigned);
    l.timeoutOff = l.ackOff + (size_t)l.blocks * sizeof(unsigned);
    l.bytes = l.timeoutOff + 256;
    return l;
}

int fill_region_impl(VpeContext* c, int x0, int x1, int y0, int y1, FillPhase phase = FILL_FUSED) {
    GridParams& g = c->g;
    FillArgs a;
    a.covered = c->dCovered.p; a.sliceStart = c->dSliceStart.p; a.cellStart = c->dCellStart.p; a.pairs = c->dPairs.p;
    a.pfill = c->dPfill.p; a.cube = c->dCube.p; a.depth = c->depthSet ? c->dDepth.p : nullptr;
    a.sheet = c->dSheet.p; a.bricks = c->dBricks.p;
    a.nz = c->dNz.p;
    a.x0 = x0; a.x1 = x1; a.y0 = y0; a.y1 = y1;
    a.densityDone = nullptr; a.densityEpoch = 0;
    const bool wholeGrid = x0 == 0 && y0 == 0 && x1 == g.NX && y1 == g.NY;
    if (phase == FILL_DENSITY) {
        // with a sheet link in place the sweep that follows runs concurrently (k_sweep_overlapped): raise a flag per block
        c->densitySignalled = c->linkOwn && c->sweepStream && wholeGrid && c->dbg.sweepOverlap != 0;
        if (c->densitySignalled) {
            a.densityDone = c->dDensityDone.p;
            a.densityEpoch = ++c->densityEpoch;
            CUDA_TRY(c, cudaEventRecord(c->evDensityBegin, c->stream));
        }
    }
    CUDA_TRY(c, cudaEventRecord(c->evFillK0, c->stream));
    if (phase != FILL_DENSITY) c->bricksGray = g.gray != 0;
    if (c->nCovered > 0 || phase == FILL_SWEEP_LINKED || phase == FILL_FUSED_HEAD) {  // a linked fill always runs: its flags must flow
        // one launch: every voxel column of the region walks all slices of the slab
        const int warpTiles = ((g.N + 7) / 8) * ((g.N + 3) / 4);
        const dim3 grid((x1 - x0) * (y1 - y0), div_up(warpTiles, FILLC_THREADS / 32));
        SheetLink link;
        memset(&link, 0, sizeof(link));
        if (phase == FILL_SWEEP_LINKED || phase == FILL_FUSED_HEAD) {
            const LinkLayout l = link_layout(g);
            char* own = static_cast<char*>(c->linkOwn);
            link.inbox = reinterpret_cast<const float*>(own);
            link.flagIn = reinterpret_cast<const unsigned*>(own + l.flagOff);
            link.ackIn = reinterpret_cast<const unsigned*>(own + l.ackOff);
            link.timeouts = reinterpret_cast<unsigned*>(own + l.timeoutOff);
            link.hasUp = c->linkUp != nullptr;
            link.hasDown = c->linkDown != nullptr;
            if (link.hasUp) link.upAck = reinterpret_cast<unsigned*>(static_cast<char*>(c->linkUp) + l.ackOff);
            if (link.hasDown) {
                link.downInbox = reinterpret_cast<float*>(c->linkDown);
                link.downFlag = reinterpret_cast<unsigned*>(static_cast<char*>(c->linkDown) + l.flagOff);
            }
            link.epoch = ++c->linkEpoch;
            link.spinLimit = (long long)(c->dbg.linkSpinMs > 0 ? c->dbg.linkSpinMs : 2000) * 2000000ll;  // ~2 GHz ticks
        }
        if (phase == FILL_FUSED && g.gray) k_fill_columns<false, true><<<grid, FILLC_THREADS, 0, c->stream>>>(g, a, c->dBrickOf.p, c->dCubeFp.p, link);
        else if (phase == FILL_FUSED) k_fill_columns<false, false><<<grid, FILLC_THREADS, 0, c->stream>>>(g, a, c->dBrickOf.p, c->dCubeFp.p, link);
        else if (phase == FILL_FUSED_HEAD && g.gray) k_fill_columns<false, true, true><<<grid, FILLC_THREADS, 0, c->stream>>>(g, a, c->dBrickOf.p, c->dCubeFp.p, link);
        else if (phase == FILL_FUSED_HEAD) k_fill_columns<false, false, true><<<grid, FILLC_THREADS, 0, c->stream>>>(g, a, c->dBrickOf.p, c->dCubeFp.p, link);
        else if (phase == FILL_DENSITY) k_fill_columns<true, false><<<grid, FILLC_THREADS, 0, c->stream>>>(g, a, c->dBrickOf.p, c->dCubeFp.p, link);
        else {
            if (phase == FILL_SWEEP_LINKED) {
                const bool tma = c->brickMapOk && !c->dbg.noTmaSweep;
                if (c->densitySignalled) {
                    // persistent kernel on its own stream, concurrent with the density pass launched just before
                    a.densityDone = c->dDensityDone.p;
                    a.densityEpoch = c->densityEpoch;
                    const int numBlocks = (int)(grid.x * grid.y);
                    const int ctas = std::min(numBlocks, c->numSMs);
                    CUDA_TRY(c, cudaStreamWaitEvent(c->sweepStream, c->evDensityBegin, 0));
                    if (tma) launch_sweep_tma<true, true>(c, a, link, grid, c->sweepStream);
                    else if (g.gray) k_sweep_overlapped<true><<<ctas, FILLC_THREADS, 0, c->sweepStream>>>(g, a, c->dBrickOf.p, link, (int)grid.x, numBlocks);
                    else k_sweep_overlapped<false><<<ctas, FILLC_THREADS, 0, c->sweepStream>>>(g, a, c->dBrickOf.p, link, (int)grid.x, numBlocks);
                    CUDA_TRY(c, cudaEventRecord(c->evSweepDone, c->sweepStream));
                    CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->evSweepDone, 0));
                    c->densitySignalled = false;
                } else if (tma) launch_sweep_tma<true, false>(c, a, link, grid, c->stream);
                else if (g.gray) k_sweep_columns<true, true><<<grid, FILLC_THREADS, 0, c->stream>>>(g, a, c->dBrickOf.p, link);
                else k_sweep_columns<false, true><<<grid, FILLC_THREADS, 0, c->stream>>>(g, a, c->dBrickOf.p, link);
            } else if (c->brickMapOk && !c->dbg.noTmaSweep) launch_sweep_tma<false, false>(c, a, link, grid, c->stream);
            else if (g.gray) k_sweep_columns<true, false><<<grid, FILLC_THREADS, 0, c->stream>>>(g, a, c->dBrickOf.p, link);
            else k_sweep_columns<false, false><<<grid, FILLC_THREADS, 0, c->stream>>>(g, a, c->dBrickOf.p, link);
        }
        c->stats.fillLaunches++;
        if ((phase == FILL_FUSED || phase == FILL_FUSED_HEAD || phase == FILL_DENSITY) && c->nCovered > 0) {  // the densities are final: derive the march's bitmap
            k_occ_build<<<c->nCovered, 256, 0, c->stream>>>(g, c->dCovered.p, c->dTotals.p + 1, c->dNz.p, c->dOcc.p, c->occRowWords, x0, x1, y0, y1);
            c->stats.fillLaunches++;
        }
    }
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaEventRecord(c->evFill1, c->stream));
    c->fillTimed = true;
    c->filledOnce = true;
    return VPE_OK;
}

// ---- march ------------------------------------------------------------------------------------
struct SortData {  // ≙ MetavoxelSortData, VPR.cs:43-62
    int x, y;
    float distance;
};

// Host destination of a full-image march: when given (and pinned), the image is marched in bands of CTA rows on
// two auxiliary streams and every band is copied to the host as soon as it is done, so that the device-to-host
// copy of the image overlaps the march instead of following it.
struct HostImage {
    float* rgba = nullptr;
    int32_t* samples = nullptr;
    bool copied = false;   // out: the bands have been copied (the caller must not copy again)
};

int march_impl(VpeContext* c, const VpeCamera* cam, const int* pixelsDev, int nPixels, float4* rgbaDev, float4* underDev,
               int* samplesDev, bool partial, unsigned* footprint = nullptr, HostImage* host = nullptr, bool linked = false) {
    GridParams& g = c->g;
    const VpeConfig& k = c->cfg;
    if (cam->width < 1 || cam->height < 1) return fail(c, VPE_E_INVALID_ARG, "bad image size");
    MarchParams m;
    memset(&m, 0, sizeof(m));
    m.W = cam->width; m.H = cam->height;
    m.Wf = (float)cam->width; m.Hf = (float)cam->height;
    m.aspect = m.Wf / m.Hf;                                                   // March.shader:190
    {
        float fovRad = 0.0174532924f * cam->fovYDegrees;                      // VPR.cs:734
        float tanHalf = (float)tan((double)(fovRad / 2.0f));                  // March.shader:193
        m.negRcpTan = -(1.0f / tanHalf);
    }
    F3 camPos = f3(cam->transform.position[0], cam->transform.position[1], cam->transform.position[2]);
    M3 camRot = quat_to_m3(cam->transform.rotation);
    Affine camL2W = trs(camPos, camRot, 1.0f);
    Affine C2W = camL2W;  // cameraToWorldMatrix: OpenGL convention, -Z forward
    for (int i = 0; i < 3; i++) C2W.m[i][2] = -C2W.m[i][2];
    Affine W2C = affine_inverse(camL2W);  // worldToCameraMatrix
    for (int j = 0; j < 4; j++) W2C.m[2][j] = -W2C.m[2][j];
    const float maxGridDim = (float)std::max(g.NX, std::max(g.NY, g.NZ));     // March.shader:207-208
    {
        float csVolOriginZ = ((W2C.m[2][0] * g.center.x + W2C.m[2][1] * g.center.y) + W2C.m[2][2] * g.center.z) + W2C.m[2][3];  // :206
        float csVolHalfZ = 1.73205f * 0.5f * maxGridDim * g.s;                // :210
        m.csZVolMin = csVolOriginZ + csVolHalfZ;                              // :211
        float csRayLength = 2.0f * csVolHalfZ;                                // :214
        float total = maxGridDim * (float)k.rayMarchSteps;                    // :221
        float oneOver = 1.0f / total;                                         // :222
        float mvRayLength = csRayLength * (1.0f / g.s);                       // :223
        m.stepSize = mvRayLength * oneOver;                                   // :224
    }
    m.borderVoxelOffset = (1.0f / g.Nf) * (float)k.numBorderVoxels;           // :245
    m.sampleScale = 1.0f - 2.0f * m.borderVoxelOffset;                        // :258
    m.softDistance = k.softParticleStepDistance;
    m.softRcp = 1.0f / (float)k.softParticleStepDistance;                     // :269
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++)
            m.C2Mlin[i][j] = (g.Ls.b[i][0] * C2W.m[0][j] + g.Ls.b[i][1] * C2W.m[1][j]) + g.Ls.b[i][2] * C2W.m[2][j];  // VPR.cs:778
        m.c2wT[i] = C2W.m[i][3];
    }
    // VPR.cs:613-632 SortMetavoxelSlicesFarToNearFromEye: keys from slice 0; ties by (y,x)
    std::vector<SortData> asc;
    asc.reserve((size_t)g.NX * g.NY);
    for (int yy = 0; yy < g.NY; yy++)
        for (int xx = 0; xx < g.NX; xx++) {
            F3 d = sub(mv_center(g, xx, yy, 0), camPos);
            asc.push_back(SortData{xx, yy, dot3(d, d)});
        }
    std::stable_sort(asc.begin(), asc.end(), [](const SortData& a, const SortData& b) { return a.distance < b.distance; });
    std::vector<int> rank((size_t)g.NX * g.NY);
    for (size_t i = 0; i < asc.size(); i++) rank[(size_t)asc[i].y * g.NX + asc[i].x] = (int)i;
    // VPR.cs:642-648
    {
        F3 lsCam = xform_point(g.W2L, camPos);
        float lsFirst = xform_point(g.W2L, mv_center(g, 0, 0, 0)).z;
        float blendOverIndex = (lsCam.z - lsFirst) / g.s;
        int zB = (int)nearbyint((double)blendOverIndex);  // Mathf.RoundToInt
        m.zBoundary = std::min(std::max(zB, -1), g.NZ - 1);
    }
    m.zOverBegin = std::max(0, g.z0);
    m.zOverEnd = std::max(m.zOverBegin, std::min(m.zBoundary + 1, g.z1));
    m.zUnderBegin = std::max(m.zBoundary + 1, g.z0);
    m.zUnderEnd = std::max(m.zUnderBegin, g.z1);
    m.earlyOut = k.marchEarlyOutTransmittance;
    m.numPixels = pixelsDev ? nPixels : cam->width * cam->height;
    m.maxSamplesPerMv = (int)(1.7320508f / m.stepSize) + 2;
    m.wrap = k.numBorderVoxels == 0 ? 1 : 0;
    m.rowStride = g.rowStride;
    m.gray = c->bricksGray ? 1 : 0;
    // march options (vpe_set_march_options): rendered by the general kernel
    const bool optionsActive = c->targetFormat != 0 || c->debugMode != 0 || c->sceneDepthSet;
    if (optionsActive && (partial || footprint)) return fail(c, VPE_E_UNSUPPORTED, "march options are not available for slab partial images");
    if (c->sceneDepthSet && (c->sceneW != cam->width || c->sceneH != cam->height || pixelsDev))
        return fail(c, VPE_E_INVALID_ARG, "scene depth buffer does not match the camera's image (full-image march only)");
    m.targetFormat = c->targetFormat;
    m.debugMode = c->debugMode;
    m.numCovered = c->nCovered;

    CUDA_TRY(c, cudaEventRecord(c->evMarch0, c->stream));
    c->marchTimed = false;
    CUDA_TRY(c, cudaMemcpyAsync(c->dRank.p, rank.data(), sizeof(int) * rank.size(), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaMemsetAsync(c->dTotalSamples.p, 0, 2 * sizeof(unsigned long long), c->stream));
    k_mv_camera<<<div_up(c->numCells, 256), 256, 0, c->stream>>>(g, m, c->dBrickOf.p, c->dMvCam.p);
    MarchArgs a;
    a.mvCam = c->dMvCam.p; a.rankAsc = c->dRank.p; a.bricks = c->dBricks.p; a.pixels = pixelsDev;
    a.rgba = rgbaDev; a.under = underDev; a.samples = samplesDev; a.totalSamples = c->dTotalSamples.p;
    a.footprint = footprint;
    a.totalSkipped = footprint ? c->dTotalSamples.p + 1 : nullptr;
    a.sliceSamples = nullptr;
    if (c->dbg.profileSlices && !footprint) {
        CUDA_TRY(c, c->dSliceSamples.ensure((size_t)g.NZ));
        CUDA_TRY(c, cudaMemsetAsync(c->dSliceSamples.p, 0, sizeof(unsigned long long) * (size_t)g.NZ, c->stream));
        a.sliceSamples = c->dSliceSamples.p;
    }
    a.peerRecv = nullptr;
    a.linkWorld = a.linkRank = a.linkPer = a.linkParity = a.linkW = 0;
    a.linkKinds = 3;
    if (linked) {
        // a slab entirely on one side of zBoundary has one partial that cannot be anything but zero: it stays at home
        a.linkKinds = (m.zOverEnd > m.zOverBegin ? 1 : 0) | (m.zUnderEnd > m.zUnderBegin ? 2 : 0);
        c->imgKinds = (unsigned)a.linkKinds;
        a.peerRecv = c->dImgPeers.p;
        a.linkWorld = c->imgWorld; a.linkRank = c->imgRank; a.linkPer = c->imgPer; a.linkW = c->imgW;
        a.linkParity = (int)(c->imgEpoch & 1u);
    }
    a.sceneDepth = c->sceneDepthSet ? c->dSceneDepth.p : nullptr;
    a.orderOf = nullptr;
    if (c->debugMode == 1) {
        CUDA_TRY(c, c->dOrderOf.ensure(c->numCells));
        k_order_index<<<div_up(c->numCells, 128), 128, 0, c->stream>>>(g, m, c->dBrickOf.p, c->dRank.p, c->dSliceStart.p, c->dOrderOf.p);
        a.orderOf = c->dOrderOf.p;
    }
    const bool skip = !c->dbg.noSkip;
    a.occ = (skip || footprint) ? c->dOcc.p : nullptr; a.occRowWords = c->occRowWords;
    // Warp pixel tile. Measured on cfg3 (profiles/): the compact 8x4 tile wins over strips that follow
    // the bricks' 128-byte rows, because lanes of a strip sit in different bricks and diverge.
    {
        m.tileLog2W = (c->dbg.marchTileLog2W >= 1 && c->dbg.marchTileLog2W <= 6) ? c->dbg.marchTileLog2W - 1 : 3;
    }
    dim3 grid, block(128);
    int ctaRows = 1;  // image rows per row of CTAs
    if (pixelsDev) grid = dim3(div_up(nPixels, 128));
    else {
        const int lw = m.tileLog2W, tw = 1 << lw, th = 32 >> lw;
        const int cw = lw >= 4 ? tw : (lw <= 1 ? 4 * tw : 2 * tw), ch = lw >= 4 ? 4 * th : (lw <= 1 ? th : 2 * th);
        grid = dim3(div_up(cam->width, cw), div_up(cam->height, ch));
        ctaRows = ch;
    }
    // bands (host path): only worth it for a large image going to pinned memory
    int numBands = 1;
    if (host && host->rgba && !pixelsDev && !partial && !footprint && (int)grid.y >= 32 && m.numPixels >= (1 << 19) && c->dbg.marchBands != 1) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, host->rgba) == cudaSuccess && at.type == cudaMemoryTypeHost)
            numBands = c->dbg.marchBands > 1 ? std::min(c->dbg.marchBands, 32) : 6;  // measured on cfg3: 4-8 bands 7.40-7.46 ms, 16 bands 8.7, one launch + one copy 7.86
        else cudaGetLastError();
        if (numBands > 1 && host->samples) {
            if (!(cudaPointerGetAttributes(&at, host->samples) == cudaSuccess && at.type == cudaMemoryTypeHost)) { cudaGetLastError(); numBands = 1; }
        }
        if (numBands > 1 && !c->aux[0]) {
            bool ok = cudaEventCreateWithFlags(&c->evFork, cudaEventDisableTiming) == cudaSuccess;
            for (int i = 0; i < 3; i++)
                ok = ok && cudaStreamCreateWithFlags(&c->aux[i], cudaStreamNonBlocking) == cudaSuccess &&
                     cudaEventCreateWithFlags(&c->evJoin[i], cudaEventDisableTiming) == cudaSuccess;
            for (int i = 0; i < 32; i++) ok = ok && cudaEventCreateWithFlags(&c->evBand[i], cudaEventDisableTiming) == cudaSuccess;
            if (!ok) { cudaGetLastError(); numBands = 1; }
        }
    }
    CUDA_TRY(c, cudaEventRecord(c->evMarchK0, c->stream));
    if (numBands > 1) {
        CUDA_TRY(c, cudaEventRecord(c->evFork, c->stream));
        for (int i = 0; i < 3; i++) CUDA_TRY(c, cudaStreamWaitEvent(c->aux[i], c->evFork, 0));
    }
    const dim3 fullGrid = grid;
    for (int band = 0; band < numBands && m.numPixels > 0; band++) {
        cudaStream_t st = numBands > 1 ? c->aux[band & 1] : c->stream;
        const int by0 = (int)((long long)band * fullGrid.y / numBands), by1 = (int)((long long)(band + 1) * fullGrid.y / numBands);
        if (numBands > 1) { grid = dim3(fullGrid.x, by1 - by0); m.blockYBase = by0; }
        const bool legacy = m.wrap || optionsActive || c->dbg.marchKernel == 1;
        const bool gray = c->bricksGray;  // the layout the fill wrote (z-paired grey texels or half4)
        // k_march_flat (one sample loop per ray and slice) is the production kernel; round 1's per-fragment loop
        // remains selectable for comparison (VpeDebugOptions.marchKernel = 2).
        const bool flat = c->dbg.marchKernel != 2;
#define VPE_LAUNCH_MARCH3(KERNEL, ...)                                                                  \
    do {                                                                                                \
        if (skip && gray) KERNEL<__VA_ARGS__, true, true, PAD_><<<grid, block, 0, st>>>(g, m, a); \
        else if (skip) KERNEL<__VA_ARGS__, true, false, PAD_><<<grid, block, 0, st>>>(g, m, a);   \
        else if (gray) KERNEL<__VA_ARGS__, false, true, PAD_><<<grid, block, 0, st>>>(g, m, a);   \
        else KERNEL<__VA_ARGS__, false, false, PAD_><<<grid, block, 0, st>>>(g, m, a);            \
    } while (0)
#define VPE_LAUNCH_MARCH2(NT, PAD)                           \
    do {                                                     \
        constexpr bool PAD_ = PAD;                           \
        if (flat) VPE_LAUNCH_MARCH3(k_march_flat, NT);       \
        else VPE_LAUNCH_MARCH3(k_march, NT, false);          \
    } while (0)
#define VPE_LAUNCH_MARCH(NT)                     \
    do {                                         \
        if (m.rowStride != g.N) VPE_LAUNCH_MARCH2(NT, true);  \
        else VPE_LAUNCH_MARCH2(NT, false);       \
    } while (0)
        if (footprint) k_march<-1, true, false, false, false><<<grid, block, 0, st>>>(g, m, a);
        else if (legacy) k_march<-1, false, false, false, false><<<grid, block, 0, st>>>(g, m, a);
        else if (g.N == 32) VPE_LAUNCH_MARCH(32);
        else if (g.N == 64) VPE_LAUNCH_MARCH(64);
        else VPE_LAUNCH_MARCH(0);
#undef VPE_LAUNCH_MARCH2
#undef VPE_LAUNCH_MARCH3
#undef VPE_LAUNCH_MARCH
        if (numBands > 1) {  // this band's rows go home (copy stream) while the next bands are marched
            const int r0 = by0 * ctaRows, r1 = std::min(cam->height, by1 * ctaRows);
            const size_t off = (size_t)r0 * cam->width, cnt = (size_t)(r1 - r0) * cam->width;
            CUDA_TRY(c, cudaEventRecord(c->evBand[band], st));
            CUDA_TRY(c, cudaStreamWaitEvent(c->aux[2], c->evBand[band], 0));
            CUDA_TRY(c, cudaMemcpyAsync(host->rgba + off * 4, rgbaDev + off, cnt * sizeof(float4), cudaMemcpyDeviceToHost, c->aux[2]));
            if (host->samples && samplesDev) CUDA_TRY(c, cudaMemcpyAsync(host->samples + off, samplesDev + off, cnt * sizeof(int), cudaMemcpyDeviceToHost, c->aux[2]));
        }
    }
    if (numBands > 1) {
        for (int i = 0; i < 3; i++) {
            CUDA_TRY(c, cudaEventRecord(c->evJoin[i], c->aux[i]));
            CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->evJoin[i], 0));
        }
        host->copied = true;
    }
    CUDA_TRY(c, cudaEventRecord(c->evMarchK1, c->stream));
    c->stats.marchLaunches = 1 + numBands;
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaMemcpyAsync(c->hTotalSamples, c->dTotalSamples.p, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaEventRecord(c->evMarch1, c->stream));
    c->marchTimed = true;
    c->stats.zBoundary = m.zBoundary;
    return VPE_OK;
}

int check_ready_for_march(VpeContext* c) {
    if (!c->lightSet) return fail(c, VPE_E_NOT_READY, "vpe_set_light has not been called");
    if (!c->filledOnce) return fail(c, VPE_E_NOT_READY, "vpe_fill has not been called");
    return VPE_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
extern "C" {

void vpe_default_config(VpeConfig* cfg) {
    memset(cfg, 0, sizeof(*cfg));
    cfg->numMetavoxelsX = cfg->numMetavoxelsY = cfg->numMetavoxelsZ = 10;
    cfg->mvScale = 3.0f;
    cfg->numVoxelsInMetavoxel = 32;
    cfg->numBorderVoxels = 1;
    cfg->rayMarchSteps = 64;
    cfg->ambientColor[0] = cfg->ambientColor[1] = cfg->ambientColor[2] = 0.2f;
    cfg->displacementScale = 0.7f;
    cfg->fadeOutParticles = 0;
    cfg->opacityFactor = 0.04f;
    cfg->softParticleStepDistance = 20;
    cfg->lightNear = 0.3f;
    cfg->lightFar = 1000.0f;
    cfg->lightCameraDistance = 200.0f;
    cfg->binMode = VPE_BIN_REFERENCE;
    cfg->marchEarlyOutTransmittance = 0.0f;
    cfg->slabZBegin = cfg->slabZEnd = 0;
}

int vpe_create(const VpeConfig* cfg, int device, VpeContext** out) {
    if (!cfg || !out) return VPE_E_INVALID_ARG;
    VpeConfig c2 = *cfg;
    normalise_slab(c2);
    std::string why;
    if (validate_config(c2, why)) return VPE_E_INVALID_ARG;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return VPE_E_CUDA;  // no fallback
    DeviceScope deviceScope(device);
    if (!deviceScope.ok) { cudaGetLastError(); return VPE_E_CUDA; }
    preload_kernels(device);
    VpeContext* c = new VpeContext();
    c->cfg = c2;
    c->device = device;
    memset(&c->stats, 0, sizeof(c->stats));
    memset(&c->light, 0, sizeof(c->light));
    c->light.rotation[3] = 1.0f;
    c->numCells = c2.numMetavoxelsX * c2.numMetavoxelsY * c2.numMetavoxelsZ;
    c->occRowWords = (c2.numVoxelsInMetavoxel + 31) / 32;
    memset(&c->dbg, 0, sizeof(c->dbg));
    rebuild_grid_params(c);
    bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess;
    c->ownStream = ok;
    ok = ok && cudaEventCreate(&c->evFill0) == cudaSuccess && cudaEventCreate(&c->evFill1) == cudaSuccess;
    ok = ok && cudaEventCreate(&c->evMarch0) == cudaSuccess && cudaEventCreate(&c->evMarch1) == cudaSuccess;
    ok = ok && cudaEventCreate(&c->evFillK0) == cudaSuccess && cudaEventCreate(&c->evMarchK0) == cudaSuccess && cudaEventCreate(&c->evMarchK1) == cudaSuccess;
    const int cells = c->numCells;
    const size_t sheetN = (size_t)c2.numMetavoxelsX * c2.numMetavoxelsY * c2.numVoxelsInMetavoxel * c2.numVoxelsInMetavoxel;
    ok = ok && c->dCellCount.ensure(cells) == cudaSuccess && c->dCellStart.ensure(cells + 1) == cudaSuccess;
    ok = ok && c->dBrickOf.ensure(cells) == cudaSuccess && c->dCovered.ensure(cells) == cudaSuccess;
    ok = ok && c->dSliceStart.ensure(c2.numMetavoxelsZ + 1) == cudaSuccess && c->dTotals.ensure(2) == cudaSuccess;
    ok = ok && c->dBlockSums.ensure(div_up(cells, SCAN_BLOCK)) == cudaSuccess;
    ok = ok && c->dSheet.ensure(sheetN) == cudaSuccess && c->dMvCam.ensure(cells) == cudaSuccess;
    ok = ok && c->dRank.ensure((size_t)c2.numMetavoxelsX * c2.numMetavoxelsY) == cudaSuccess;
    ok = ok && c->dTotalSamples.ensure(2) == cudaSuccess;
    ok = ok && cudaMallocHost(&c->hCounts, sizeof(int) * (c2.numMetavoxelsZ + 3)) == cudaSuccess;
    ok = ok && cudaMallocHost(&c->hTotalSamples, 2 * sizeof(unsigned long long)) == cudaSuccess;
    if (ok) {
        ok = cudaMemset(c->dBrickOf.p, 0xff, sizeof(int) * cells) == cudaSuccess;
        k_fill_value<<<div_up(sheetN, 256), 256>>>(c->dSheet.p, sheetN, 1.0f);
        ok = ok && cudaDeviceSynchronize() == cudaSuccess;
    }
    if (!ok) {
        cudaGetLastError();
        vpe_destroy(c);
        return VPE_E_CUDA;
    }
    c->hTotalSamples[0] = c->hTotalSamples[1] = 0;
    *out = c;
    return VPE_OK;
}

int vpe_destroy(VpeContext* c) {
    if (!c) return VPE_OK;
    DeviceScope deviceScope(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    c->dParticles.release(); c->dPfill.release(); c->dPbin.release();
    c->dCellCount.release(); c->dCellStart.release(); c->dBrickOf.release(); c->dCovered.release();
    c->dSliceStart.release(); c->dPairs.release(); c->dTotals.release(); c->dBlockSums.release();
    c->dCube.release(); c->dCubeFp.release(); c->dDepth.release(); c->dSheet.release(); c->dBricks.release(); c->dNz.release(); c->dOcc.release(); c->dMvCam.release();
    c->dRank.release(); c->dPixels.release(); c->dSamples.release(); c->dImage.release(); c->dImage2.release();
    c->dTotalSamples.release(); c->dSliceSamples.release(); c->dParts.release();
    c->dSceneDepth.release(); c->dOrderOf.release(); c->dTris.release(); c->dScene.release();
    for (int i = 0; i < 3; i++) {
        if (c->aux[i]) cudaStreamDestroy(c->aux[i]);
        if (c->evJoin[i]) cudaEventDestroy(c->evJoin[i]);
    }
    for (int i = 0; i < 32; i++)
        if (c->evBand[i]) cudaEventDestroy(c->evBand[i]);
    if (c->evFork) cudaEventDestroy(c->evFork);
    for (int q = 0; q < 64; q++)
        if (c->imgPeers[q] && c->imgPeerIpc[q]) cudaIpcCloseMemHandle(c->imgPeers[q]);
    if (c->imgOwn) cudaFree(c->imgOwn);
    c->dImgPeers.release();
    if (c->linkUp && c->linkUpIpc) cudaIpcCloseMemHandle(c->linkUp);
    if (c->linkDown && c->linkDownIpc) cudaIpcCloseMemHandle(c->linkDown);
    if (c->linkOwn) cudaFree(c->linkOwn);
    c->dDensityDone.release();
    if (c->sweepStream) cudaStreamDestroy(c->sweepStream);
    if (c->evDensityBegin) cudaEventDestroy(c->evDensityBegin);
    if (c->evSweepDone) cudaEventDestroy(c->evSweepDone);
    if (c->hCounts) cudaFreeHost(c->hCounts);
    if (c->hTotalSamples) cudaFreeHost(c->hTotalSamples);
    if (c->evFill0) cudaEventDestroy(c->evFill0);
    if (c->evFill1) cudaEventDestroy(c->evFill1);
    if (c->evMarch0) cudaEventDestroy(c->evMarch0);
    if (c->evMarch1) cudaEventDestroy(c->evMarch1);
    if (c->evFillK0) cudaEventDestroy(c->evFillK0);
    if (c->evMarchK0) cudaEventDestroy(c->evMarchK0);
    if (c->evMarchK1) cudaEventDestroy(c->evMarchK1);
    if (c->ownStream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return VPE_OK;
}

int vpe_set_config(VpeContext* c, const VpeConfig* cfg) {
    if (!c || !cfg) return VPE_E_INVALID_ARG;
    VpeConfig c2 = *cfg;
    normalise_slab(c2);
    std::string why;
    if (validate_config(c2, why)) return fail(c, VPE_E_INVALID_ARG, why.c_str());
    if (c2.numMetavoxelsX != c->cfg.numMetavoxelsX || c2.numMetavoxelsY != c->cfg.numMetavoxelsY ||
        c2.numMetavoxelsZ != c->cfg.numMetavoxelsZ || c2.numVoxelsInMetavoxel != c->cfg.numVoxelsInMetavoxel)
        return fail(c, VPE_E_INVALID_ARG, "grid dims and voxel count are fixed at create");
    if (c2.slabZBegin != c->cfg.slabZBegin || c2.slabZEnd != c->cfg.slabZEnd) {
        // the slab may move between fills (load balancing, slabs.py): the volume of the old slab is gone
        c->prepared = false;
        c->filledOnce = false;
    }
    c->cfg = c2;
    rebuild_grid_params(c);  // SetGridScale re-places the metavoxels (VPR.cs:1059-1063)
    return VPE_OK;
}

int vpe_set_light(VpeContext* c, const VpeTransform* light, const float gridCenter[3]) {
    if (!c || !light || !gridCenter) return fail(c, VPE_E_INVALID_ARG, "null argument");
    c->light = *light;
    for (int i = 0; i < 3; i++) c->center[i] = gridCenter[i];
    c->lightSet = true;
    rebuild_grid_params(c);
    return VPE_OK;
}

int vpe_set_displacement_cubemap(VpeContext* c, const uint8_t* r8, int edge) {
    if (!c || !r8 || edge < 1) return fail(c, VPE_E_INVALID_ARG, "bad cubemap");
    DeviceScope deviceScope(c->device);
    size_t n = (size_t)6 * edge * edge;
    std::vector<float> f(n);
    for (size_t i = 0; i < n; i++) f[i] = (float)r8[i] / 255.0f;  // UNORM8 -> float, as the sampler would
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    c->dCube.release();
    CUDA_TRY(c, c->dCube.ensure(n));
    CUDA_TRY(c, cudaMemcpy(c->dCube.p, f.data(), n * sizeof(float), cudaMemcpyHostToDevice));
    {
        // bilinear footprints with clamp addressing baked in: entry (i,j) = texels (i-1..i, j-1..j)
        const int E1 = edge + 1;
        std::vector<float4> fp((size_t)6 * E1 * E1);
        auto cl = [&](int v) { return std::min(std::max(v, 0), edge - 1); };
        for (int face = 0; face < 6; face++)
            for (int j = 0; j < E1; j++)
                for (int i = 0; i < E1; i++) {
                    const float* t = f.data() + (size_t)face * edge * edge;
                    const int x0 = cl(i - 1), x1 = cl(i), y0 = cl(j - 1), y1 = cl(j);
                    fp[((size_t)face * E1 + j) * E1 + i] = make_float4(t[y0 * edge + x0], t[y0 * edge + x1], t[y1 * edge + x0], t[y1 * edge + x1]);
                }
        c->dCubeFp.release();
        CUDA_TRY(c, c->dCubeFp.ensure(fp.size()));
        CUDA_TRY(c, cudaMemcpy(c->dCubeFp.p, fp.data(), fp.size() * sizeof(float4), cudaMemcpyHostToDevice));
    }
    c->cubeEdge = edge;
    c->g.cubeEdge = edge;
    c->cubeSet = true;
    return VPE_OK;
}

int vpe_set_light_depth_map(VpeContext* c, const float* depth01) {
    if (!c) return VPE_E_INVALID_ARG;
    DeviceScope deviceScope(c->device);
    if (!depth01) { c->depthSet = false; return VPE_OK; }
    size_t n = (size_t)c->g.NX * c->g.N * c->g.NY * c->g.N;
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    CUDA_TRY(c, c->dDepth.ensure(n));
    CUDA_TRY(c, cudaMemcpy(c->dDepth.p, depth01, n * sizeof(float), cudaMemcpyHostToDevice));
    c->depthSet = true;
    return VPE_OK;
}

int vpe_render_light_depth_map(VpeContext* c, const float* tris, int numTriangles) {
    if (!c || (!tris && numTriangles > 0) || numTriangles < 0) return fail(c, VPE_E_INVALID_ARG, "bad triangle list");
    if (!c->lightSet) return fail(c, VPE_E_NOT_READY, "vpe_set_light has not been called");
    DeviceScope deviceScope(c->device);
    const GridParams& g = c->g;
    DepthRasterParams p;
    p.W2LC = c->w2lc;
    p.r = (float)g.NX * g.s * 0.5f; p.t = (float)g.NY * g.s * 0.5f;  // VPR.cs:340
    p.zn = c->cfg.lightNear; p.zf = c->cfg.lightFar;
    p.W = g.NX * g.N; p.H = g.NY * g.N;
    const size_t n = (size_t)p.W * p.H;
    CUDA_TRY(c, c->dDepth.ensure(n));
    CUDA_TRY(c, c->dTris.ensure((size_t)std::max(numTriangles, 1) * 9));
    if (numTriangles > 0)
        CUDA_TRY(c, cudaMemcpyAsync(c->dTris.p, tris, sizeof(float) * 9 * (size_t)numTriangles, cudaMemcpyHostToDevice, c->stream));
    k_fill_value<<<std::min(div_up(n, 256), 148 * 8), 256, 0, c->stream>>>(c->dDepth.p, n, 1.0f);  // CameraClearFlags.Depth
    if (numTriangles > 0)
        k_raster_depth<<<std::min(numTriangles, 148 * 16), 256, 0, c->stream>>>(p, c->dTris.p, numTriangles, reinterpret_cast<unsigned*>(c->dDepth.p));
    CUDA_TRY(c, cudaGetLastError());
    c->depthSet = true;
    return sync_stream(c);
}

int vpe_read_light_depth_map(VpeContext* c, float* depth01) {
    if (!c || !depth01) return VPE_E_INVALID_ARG;
    DeviceScope deviceScope(c->device);
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    const size_t n = (size_t)c->g.NX * c->g.N * c->g.NY * c->g.N;
    if (!c->depthSet) {
        for (size_t i = 0; i < n; i++) depth01[i] = 1.0f;
        return VPE_OK;
    }
    CUDA_TRY(c, cudaMemcpy(depth01, c->dDepth.p, n * sizeof(float), cudaMemcpyDeviceToHost));
    return VPE_OK;
}

int vpe_set_march_options(VpeContext* c, const VpeMarchOptions* o) {
    if (!c || !o) return VPE_E_INVALID_ARG;
    if (o->targetFormat < 0 || o->targetFormat > 1 || o->debugMode < 0 || o->debugMode > 3) return fail(c, VPE_E_INVALID_ARG, "bad march option");
    if (o->sceneDepth && (o->sceneWidth < 1 || o->sceneHeight < 1)) return fail(c, VPE_E_INVALID_ARG, "bad scene depth size");
    DeviceScope deviceScope(c->device);
    c->targetFormat = o->targetFormat;
    c->debugMode = o->debugMode;
    c->sceneDepthSet = false;
    if (o->sceneDepth) {
        const size_t n = (size_t)o->sceneWidth * o->sceneHeight;
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        CUDA_TRY(c, c->dSceneDepth.ensure(n));
        CUDA_TRY(c, cudaMemcpy(c->dSceneDepth.p, o->sceneDepth, n * sizeof(float), cudaMemcpyHostToDevice));
        c->sceneW = o->sceneWidth; c->sceneH = o->sceneHeight;
        c->sceneDepthSet = true;
    }
    return VPE_OK;
}

int vpe_composite_scene(VpeContext* c, const float* particles, float* scene, int numPixels, int targetFormat) {
    if (!c || !particles || !scene || numPixels < 0 || targetFormat < 0 || targetFormat > 1) return fail(c, VPE_E_INVALID_ARG, "bad argument");
    if (numPixels == 0) return VPE_OK;
    DeviceScope deviceScope(c->device);
    CUDA_TRY(c, c->dImage.ensure((size_t)numPixels));
    CUDA_TRY(c, c->dScene.ensure((size_t)numPixels));
    CUDA_TRY(c, cudaMemcpyAsync(c->dImage.p, particles, sizeof(float4) * (size_t)numPixels, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(c->dScene.p, scene, sizeof(float4) * (size_t)numPixels, cudaMemcpyHostToDevice, c->stream));
    k_composite_scene<<<div_up(numPixels, 256), 256, 0, c->stream>>>(c->dImage.p, c->dScene.p, numPixels, targetFormat);
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaMemcpyAsync(scene, c->dScene.p, sizeof(float4) * (size_t)numPixels, cudaMemcpyDeviceToHost, c->stream));
    return sync_stream(c);
}

int vpe_set_debug_options(VpeContext* c, const VpeDebugOptions* o) {
    if (!c || !o) return VPE_E_INVALID_ARG;
    if (o->marchKernel < 0 || o->marchKernel > 2 || o->marchBands < 0 || o->marchTileLog2W < 0 || o->marchTileLog2W > 6 || o->linkSpinMs < 0)
        return fail(c, VPE_E_INVALID_ARG, "bad debug option");
    const bool layoutChanged = (o->noGray != 0) != (c->dbg.noGray != 0) || (o->noRowPad != 0) != (c->dbg.noRowPad != 0);
    c->dbg = *o;
    if (layoutChanged) {  // the brick layout of the next fill changes: the volume of the last fill is gone
        c->prepared = false;
        c->filledOnce = false;
    }
    rebuild_grid_params(c);
    return VPE_OK;
}

int vpe_set_stream(VpeContext* c, void* cudaStream) {
    if (!c) return VPE_E_INVALID_ARG;
    DeviceScope deviceScope(c->device);
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (c->ownStream) cudaStreamDestroy(c->stream);
    c->stream = (cudaStream_t)cudaStream;
    c->ownStream = false;
    return VPE_OK;
}

int vpe_fill_prepare(VpeContext* c, const VpeParticle* particles, int n, const VpeTransform* emitter, int onDevice) {
    if (!c || (!particles && n > 0) || n < 0 || !emitter) return fail(c, VPE_E_INVALID_ARG, "bad particle input");
    if (!c->lightSet) return fail(c, VPE_E_NOT_READY, "vpe_set_light has not been called");
    if (!c->cubeSet) return fail(c, VPE_E_NOT_READY, "vpe_set_displacement_cubemap has not been called");
    DeviceScope deviceScope(c->device);
    const float* dev = reinterpret_cast<const float*>(particles);
    if (!onDevice) {
        CUDA_TRY(c, c->dParticles.ensure((size_t)std::max(n, 1) * 7));
        if (n > 0) CUDA_TRY(c, cudaMemcpyAsync(c->dParticles.p, particles, sizeof(VpeParticle) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
        dev = c->dParticles.p;
    }
    return fill_prepare_impl(c, dev, n, emitter);
}

int vpe_fill_region(VpeContext* c, int x0, int x1, int y0, int y1) {
    if (!c) return VPE_E_INVALID_ARG;
    if (!c->prepared) return fail(c, VPE_E_NOT_READY, "vpe_fill_prepare has not been called");
    if (x0 < 0 || y0 < 0 || x1 > c->g.NX || y1 > c->g.NY || x0 >= x1 || y0 >= y1) return fail(c, VPE_E_INVALID_ARG, "bad region");
    DeviceScope deviceScope(c->device);
    return fill_region_impl(c, x0, x1, y0, y1);
}

int vpe_fill(VpeContext* c, const VpeParticle* particles, int n, const VpeTransform* emitter) {
    int rc = vpe_fill_prepare(c, particles, n, emitter, 0);
    if (rc) return rc;
    rc = fill_region_impl(c, 0, c->g.NX, 0, c->g.NY);
    if (rc) return rc;
    return sync_stream(c);
}

int vpe_fill_device(VpeContext* c, const VpeParticle* particles_dev, int n, const VpeTransform* emitter) {
    int rc = vpe_fill_prepare(c, particles_dev, n, emitter, 1);
    if (rc) return rc;
    return fill_region_impl(c, 0, c->g.NX, 0, c->g.NY);
}

int vpe_fill_density(VpeContext* c) {
    if (!c) return VPE_E_INVALID_ARG;
    if (!c->prepared) return fail(c, VPE_E_NOT_READY, "vpe_fill_prepare has not been called");
    DeviceScope deviceScope(c->device);
    return fill_region_impl(c, 0, c->g.NX, 0, c->g.NY, FILL_DENSITY);
}

int vpe_fill_sweep_region(VpeContext* c, int x0, int x1, int y0, int y1) {
    if (!c) return VPE_E_INVALID_ARG;
    if (!c->prepared) return fail(c, VPE_E_NOT_READY, "vpe_fill_prepare has not been called");
    if (x0 < 0 || y0 < 0 || x1 > c->g.NX || y1 > c->g.NY || x0 >= x1 || y0 >= y1) return fail(c, VPE_E_INVALID_ARG, "bad region");
    DeviceScope deviceScope(c->device);
    return fill_region_impl(c, x0, x1, y0, y1, FILL_SWEEP);
}

float* vpe_light_sheet_device(VpeContext* c) { return c ? c->dSheet.p : nullptr; }

int vpe_sheet_link_create(VpeContext* c, void* ipcHandle64, void** devPtr) {
    if (!c) return VPE_E_INVALID_ARG;
    DeviceScope deviceScope(c->device);
    const LinkLayout l = link_layout(c->g);
    if (!c->linkOwn) {
        CUDA_TRY(c, cudaMalloc(&c->linkOwn, l.bytes));
        CUDA_TRY(c, cudaMemset(c->linkOwn, 0, l.bytes));
        c->linkBlocks = l.blocks;
        c->linkEpoch = 0;
        CUDA_TRY(c, c->dDensityDone.ensure((size_t)l.blocks));
        CUDA_TRY(c, cudaMemset(c->dDensityDone.p, 0, sizeof(unsigned) * (size_t)l.blocks));
        c->densityEpoch = 0;
        if (!c->sweepStream) {
            int lo = 0, hi = 0;
            CUDA_TRY(c, cudaDeviceGetStreamPriorityRange(&lo, &hi));
            CUDA_TRY(c, cudaStreamCreateWithPriority(&c->sweepStream, cudaStreamNonBlocking, hi));
            CUDA_TRY(c, cudaEventCreateWithFlags(&c->evDensityBegin, cudaEventDisableTiming));
            CUDA_TRY(c, cudaEventCreateWithFlags(&c->evSweepDone, cudaEventDisableTiming));
            CUDA_TRY(c, cudaDeviceGetAttribute(&c->numSMs, cudaDevAttrMultiProcessorCount, c->device));
        }
    }
    if (ipcHandle64) {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size is part of the ABI");
        cudaIpcMemHandle_t h;
        CUDA_TRY(c, cudaIpcGetMemHandle(&h, c->linkOwn));
        memcpy(ipcHandle64, &h, sizeof(h));
    }
    if (devPtr) *devPtr = c->linkOwn;
    return VPE_OK;
}

namespace {
int link_unmap(VpeContext* c, void*& p, bool& ipc) {
    if (p && ipc) cudaIpcCloseMemHandle(p);
    p = nullptr;
    ipc = false;
    (void)c;
    return VPE_OK;
}
int link_map(VpeContext* c, const void* src, int isIpc, void*& out, bool& outIpc) {
    if (!src) return VPE_OK;
    if (isIpc) {
        cudaIpcMemHandle_t h;
        memcpy(&h, src, sizeof(h));
        CUDA_TRY(c, cudaIpcOpenMemHandle(&out, h, cudaIpcMemLazyEnablePeerAccess));
        outIpc = true;
    } else {
        void* p = *static_cast<void* const*>(src);
        cudaPointerAttributes at;
        CUDA_TRY(c, cudaPointerGetAttributes(&at, p));
        if (at.type != cudaMemoryTypeDevice) return fail(c, VPE_E_INVALID_ARG, "sheet link: not a device pointer");
        if (at.device != c->device) {
            cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else CUDA_TRY(c, e);
        }
        out = p;
        outIpc = false;
    }
    return VPE_OK;
}
}  // namespace

int vpe_sheet_link_connect(VpeContext* c, const void* upstream, const void* downstream, int handlesAreIpc) {
    if (!c) return VPE_E_INVALID_ARG;
    if (!c->linkOwn) return fail(c, VPE_E_NOT_READY, "vpe_sheet_link_create has not been called");
    DeviceScope deviceScope(c->device);
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    link_unmap(c, c->linkUp, c->linkUpIpc);
    link_unmap(c, c->linkDown, c->linkDownIpc);
    int rc = link_map(c, upstream, handlesAreIpc, c->linkUp, c->linkUpIpc);
    if (rc) return rc;
    return link_map(c, downstream, handlesAreIpc, c->linkDown, c->linkDownIpc);
}

int vpe_fill_sweep_linked(VpeContext* c) {
    if (!c) return VPE_E_INVALID_ARG;
    if (!c->prepared) return fail(c, VPE_E_NOT_READY, "vpe_fill_prepare has not been called");
    if (!c->linkOwn) return fail(c, VPE_E_NOT_READY, "vpe_sheet_link_create has not been called");
    DeviceScope deviceScope(c->device);
    return fill_region_impl(c, 0, c->g.NX, 0, c->g.NY, FILL_SWEEP_LINKED);
}

int vpe_fill_linked(VpeContext* c) {
    if (!c) return VPE_E_INVALID_ARG;
    if (!c->prepared) return fail(c, VPE_E_NOT_READY, "vpe_fill_prepare has not been called");
    if (!c->linkOwn) return fail(c, VPE_E_NOT_READY, "vpe_sheet_link_create has not been called");
    DeviceScope deviceScope(c->device);
    if (!c->linkUp && c->linkDown && !c->dbg.noHeadFused)  // head of the chain: nothing to wait for, no reason to split
        return fill_region_impl(c, 0, c->g.NX, 0, c->g.NY, FILL_FUSED_HEAD);
    const int rc = fill_region_impl(c, 0, c->g.NX, 0, c->g.NY, FILL_DENSITY);
    if (rc) return rc;
    return fill_region_impl(c, 0, c->g.NX, 0, c->g.NY, FILL_SWEEP_LINKED);
}

int vpe_sheet_link_status(VpeContext* c, int* timeouts) {
    if (!c || !timeouts) return VPE_E_INVALID_ARG;
    *timeouts = 0;
    if (!c->linkOwn) return VPE_OK;
    DeviceScope deviceScope(c->device);
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    unsigned t[4] = {0, 0, 0, 0};  // total, density-flag waits, upstream waits, acknowledge waits
    CUDA_TRY(c, cudaMemcpy(t, static_cast<char*>(c->linkOwn) + link_layout(c->g).timeoutOff, sizeof(t), cudaMemcpyDeviceToHost));
    *timeouts = (int)t[0];
    if (t[0]) {
        char buf[160];
        snprintf(buf, sizeof(buf), "sheet link waits that gave up: %u (own densities %u, upstream sheet %u, downstream acknowledge %u)", t[0], t[1], t[2], t[3]);
        c->err = buf;
    }
    return VPE_OK;
}

int vpe_march_device(VpeContext* c, const VpeCamera* cam, float* rgba_dev, int32_t* samples_dev) {
    if (!c || !cam || !rgba_dev) return fail(c, VPE_E_INVALID_ARG, "null argument");
    int rc = check_ready_for_march(c);
    if (rc) return rc;
    DeviceScope deviceScope(c->device);
    return march_impl(c, cam, nullptr, 0, reinterpret_cast<float4*>(rgba_dev), nullptr, samples_dev, false);
}

int vpe_march_partial_device(VpeContext* c, const VpeCamera* cam, float* over_dev, float* under_dev, int32_t* samples_dev) {
    if (!c || !cam || !over_dev || !under_dev) return fail(c, VPE_E_INVALID_ARG, "null argument");
    int rc = check_ready_for_march(c);
    if (rc) return rc;
    DeviceScope deviceScope(c->device);
    return march_impl(c, cam, nullptr, 0, reinterpret_cast<float4*>(over_dev), reinterpret_cast<float4*>(under_dev), samples_dev, true);
}

// ---- image link: the slab partial images go straight into the compositing ranks' memory ----
int vpe_image_link_create(VpeContext* c, int world, int rank, int width, int height, void* ipcHandle64, void** devPtr) {
    if (!c || world < 1 || world > 64 || rank < 0 || rank >= world || width < 1 || height < 1) return fail(c, VPE_E_INVALID_ARG, "bad image link geometry");
    DeviceScope deviceScope(c->device);
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (c->imgOwn && (c->imgWorld != world || c->imgRank != rank || c->imgW != width || c->imgH != height)) {
        for (int q = 0; q < 64; q++) {
            if (c->imgPeers[q] && c->imgPeerIpc[q]) cudaIpcCloseMemHandle(c->imgPeers[q]);
            c->imgPeers[q] = nullptr; c->imgPeerIpc[q] = false;
        }
        cudaFree(c->imgOwn);
        c->imgOwn = nullptr;
        c->imgConnected = false;
    }
    if (!c->imgOwn) {
        c->imgWorld = world; c->imgRank = rank; c->imgW = width; c->imgH = height;
        c->imgPer = (height + world - 1) / world;
        const size_t recv = (size_t)2 * 2 * world * c->imgPer * width * sizeof(float4);
        c->imgFlagsOff = (recv + 255) / 256 * 256;
        c->imgTimeoutOff = c->imgFlagsOff + (size_t)4 * world * sizeof(unsigned);  // flags[2][world], kinds[2][world]
        c->imgBytes = c->imgTimeoutOff + 256;
        CUDA_TRY(c, cudaMalloc(&c->imgOwn, c->imgBytes));
        CUDA_TRY(c, cudaMemset(c->imgOwn, 0, c->imgBytes));
        c->imgEpoch = 0;
    }
    if (ipcHandle64) {
        cudaIpcMemHandle_t h;
        CUDA_TRY(c, cudaIpcGetMemHandle(&h, c->imgOwn));
        memcpy(ipcHandle64, &h, sizeof(h));
    }
    if (devPtr) *devPtr = c->imgOwn;
    return VPE_OK;
}

int vpe_image_link_connect(VpeContext* c, const void* const* peers, int handlesAreIpc) {
    if (!c || !peers) return VPE_E_INVALID_ARG;
    if (!c->imgOwn) return fail(c, VPE_E_NOT_READY, "vpe_image_link_create has not been called");
    DeviceScope deviceScope(c->device);
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    std::vector<float4*> ptrs(c->imgWorld);
    for (int q = 0; q < c->imgWorld; q++) {
        if (c->imgPeers[q] && c->imgPeerIpc[q]) cudaIpcCloseMemHandle(c->imgPeers[q]);
        c->imgPeers[q] = nullptr; c->imgPeerIpc[q] = false;
        if (q == c->imgRank) { ptrs[q] = static_cast<float4*>(c->imgOwn); continue; }
        if (!peers[q]) return fail(c, VPE_E_INVALID_ARG, "image link: a peer is missing");
        int rc = link_map(c, peers[q], handlesAreIpc, c->imgPeers[q], c->imgPeerIpc[q]);
        if (rc) return rc;
        ptrs[q] = static_cast<float4*>(c->imgPeers[q]);
    }
    CUDA_TRY(c, c->dImgPeers.ensure(c->imgWorld));
    CUDA_TRY(c, cudaMemcpy(c->dImgPeers.p, ptrs.data(), sizeof(float4*) * c->imgWorld, cudaMemcpyHostToDevice));
    c->imgConnected = true;
    return VPE_OK;
}

int vpe_march_linked(VpeContext* c, const VpeCamera* cam, int32_t* samples_dev) {
    if (!c || !cam) return fail(c, VPE_E_INVALID_ARG, "null argument");
    int rc = check_ready_for_march(c);
    if (rc) return rc;
    if (!c->imgConnected) return fail(c, VPE_E_NOT_READY, "vpe_image_link_connect has not been called");
    if (cam->width != c->imgW || cam->height != c->imgH) return fail(c, VPE_E_INVALID_ARG, "camera image size differs from the image link's");
    // two buffer parities make acknowledgements unnecessary only if every rank composites frame f before it marches frame
    // f + 1: marching twice in a row could overwrite rows a peer is still compositing
    if (c->imgCompositePending) return fail(c, VPE_E_NOT_READY, "vpe_march_linked twice without vpe_composite_linked in between");
    DeviceScope deviceScope(c->device);
    c->imgEpoch++;
    // the kernel stores into the receive buffers; rgba/under are placeholders that are never written
    rc = march_impl(c, cam, nullptr, 0, static_cast<float4*>(c->imgOwn), static_cast<float4*>(c->imgOwn), samples_dev, true, nullptr, nullptr, true);
    if (rc) return rc;
    k_image_signal<<<1, 64, 0, c->stream>>>(c->dImgPeers.p, c->imgFlagsOff, c->imgWorld, c->imgRank, (int)(c->imgEpoch & 1u), c->imgEpoch, c->imgKinds);
    CUDA_TRY(c, cudaGetLastError());
    c->stats.marchLaunches++;
    c->imgCompositePending = true;
    return VPE_OK;
}

int vpe_composite_linked(VpeContext* c, float* band_rgba_dev) {
    if (!c || !band_rgba_dev) return fail(c, VPE_E_INVALID_ARG, "null argument");
    if (!c->imgConnected || c->imgEpoch == 0) return fail(c, VPE_E_NOT_READY, "vpe_march_linked has not been called");
    DeviceScope deviceScope(c->device);
    const int np = c->imgPer * c->imgW;
    const long long spin = (long long)(c->dbg.linkSpinMs > 0 ? c->dbg.linkSpinMs : 2000) * 2000000ll;
    char* own = static_cast<char*>(c->imgOwn);
    k_composite_linked<<<div_up(np, 256), 256, 0, c->stream>>>(reinterpret_cast<const float4*>(own), reinterpret_cast<const unsigned*>(own + c->imgFlagsOff),
                                                             c->imgWorld, c->imgPer, c->imgW, (int)(c->imgEpoch & 1u), c->imgEpoch, spin,
                                                             reinterpret_cast<unsigned*>(own + c->imgTimeoutOff), reinterpret_cast<float4*>(band_rgba_dev));
    CUDA_TRY(c, cudaGetLastError());
    c->imgCompositePending = false;
    return VPE_OK;
}

int vpe_image_link_status(VpeContext* c, int* timeouts) {
    if (!c || !timeouts) return VPE_E_INVALID_ARG;
    *timeouts = 0;
    if (!c->imgOwn) return VPE_OK;
    DeviceScope deviceScope(c->device);
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    unsigned t = 0;
    CUDA_TRY(c, cudaMemcpy(&t, static_cast<char*>(c->imgOwn) + c->imgTimeoutOff, sizeof(t), cudaMemcpyDeviceToHost));
    *timeouts = (int)t;
    return VPE_OK;
}

int vpe_composite_device(VpeContext* c, const float* const* parts_dev, int numSlabs, int numPixels, float* rgba_dev) {
    if (!c || !parts_dev || numSlabs < 1 || numPixels < 0 || !rgba_dev) return fail(c, VPE_E_INVALID_ARG, "bad argument");
    DeviceScope deviceScope(c->device);
    CUDA_TRY(c, c->dParts.ensure((size_t)2 * numSlabs));
    CUDA_TRY(c, cudaMemcpyAsync(c->dParts.p, parts_dev, sizeof(float4*) * 2 * numSlabs, cudaMemcpyHostToDevice, c->stream));
    if (numPixels > 0)
        k_composite<<<div_up(numPixels, 256), 256, 0, c->stream>>>(c->dParts.p, numSlabs, numPixels, reinterpret_cast<float4*>(rgba_dev));
    CUDA_TRY(c, cudaGetLastError());
    return VPE_OK;
}

int vpe_march(VpeContext* c, const VpeCamera* cam, float* rgba, int32_t* samples) {
    if (!c || !cam || !rgba) return fail(c, VPE_E_INVALID_ARG, "null argument");
    int rc = check_ready_for_march(c);
    if (rc) return rc;
    DeviceScope deviceScope(c->device);
    const size_t np = (size_t)cam->width * cam->height;
    CUDA_TRY(c, c->dImage.ensure(np));
    if (samples) CUDA_TRY(c, c->dSamples.ensure(np));
    HostImage host;
    host.rgba = rgba;
    host.samples = samples;
    rc = march_impl(c, cam, nullptr, 0, c->dImage.p, nullptr, samples ? c->dSamples.p : nullptr, false, nullptr, &host);
    if (rc) return rc;
    if (!host.copied) {
        CUDA_TRY(c, cudaMemcpyAsync(rgba, c->dImage.p, np * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
        if (samples) CUDA_TRY(c, cudaMemcpyAsync(samples, c->dSamples.p, np * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    }
    return sync_stream(c);
}

int vpe_march_pixels(VpeContext* c, const VpeCamera* cam, const int32_t* pixels, int n, float* rgba, int32_t* samples) {
    if (!c || !cam || !rgba || !pixels || n < 0) return fail(c, VPE_E_INVALID_ARG, "null argument");
    int rc = check_ready_for_march(c);
    if (rc) return rc;
    const int total = cam->width * cam->height;
    for (int i = 0; i < n; i++)
        if (pixels[i] < 0 || pixels[i] >= total) return fail(c, VPE_E_INVALID_ARG, "pixel index out of range");
    DeviceScope deviceScope(c->device);
    CUDA_TRY(c, c->dPixels.ensure(std::max(n, 1)));
    CUDA_TRY(c, c->dImage.ensure(std::max(n, 1)));
    CUDA_TRY(c, c->dSamples.ensure(std::max(n, 1)));
    CUDA_TRY(c, cudaMemcpyAsync(c->dPixels.p, pixels, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    rc = march_impl(c, cam, c->dPixels.p, n, c->dImage.p, nullptr, c->dSamples.p, false);
    if (rc) return rc;
    CUDA_TRY(c, cudaMemcpyAsync(rgba, c->dImage.p, (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
    if (samples) CUDA_TRY(c, cudaMemcpyAsync(samples, c->dSamples.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    return sync_stream(c);
}

int vpe_march_footprint(VpeContext* c, const VpeCamera* cam, int64_t* uniqueTexels) {
    if (!c || !cam || !uniqueTexels) return fail(c, VPE_E_INVALID_ARG, "null argument");
    int rc = check_ready_for_march(c);
    if (rc) return rc;
    DeviceScope deviceScope(c->device);
    const size_t np = (size_t)cam->width * cam->height;
    const size_t texels = (size_t)c->nCovered * c->g.N * c->g.N * c->g.N;
    const size_t words = texels / 32 + 1;
    DevBuf<unsigned> bitmap;
    DevBuf<unsigned long long> count;
    CUDA_TRY(c, c->dImage.ensure(np));
    CUDA_TRY(c, bitmap.ensure(words));
    CUDA_TRY(c, count.ensure(1));
    CUDA_TRY(c, cudaMemsetAsync(bitmap.p, 0, words * sizeof(unsigned), c->stream));
    CUDA_TRY(c, cudaMemsetAsync(count.p, 0, sizeof(unsigned long long), c->stream));
    rc = march_impl(c, cam, nullptr, 0, c->dImage.p, nullptr, nullptr, false, bitmap.p);
    if (rc == VPE_OK) {
        k_popcount<<<148 * 8, 256, 0, c->stream>>>(bitmap.p, words, count.p);
        unsigned long long h = 0;
        cudaError_t e = cudaMemcpyAsync(&h, count.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) { c->err = cudaGetErrorString(e); rc = VPE_E_CUDA; }
        *uniqueTexels = (int64_t)h;
    }
    cudaStreamSynchronize(c->stream);
    if (rc == VPE_OK) {  // the instrumented launch also counted the samples the production kernels skip
        c->stats.raySamples = (int64_t)c->hTotalSamples[0];
        c->stats.raySamplesSkipped = (int64_t)c->hTotalSamples[1];
    }
    bitmap.release();
    count.release();
    c->marchTimed = false;  // an instrumented march is never reported as a timing
    return rc;
}

// ---- test hooks -------------------------------------------------------------------------------
int vpe_read_brick(VpeContext* c, int x, int y, int z, uint16_t* half4, int* covered) {
    if (!c || !covered) return VPE_E_INVALID_ARG;
    if (x < 0 || y < 0 || z < 0 || x >= c->g.NX || y >= c->g.NY || z >= c->g.NZ) return fail(c, VPE_E_INVALID_ARG, "metavoxel index out of range");
    DeviceScope deviceScope(c->device);
    *covered = 0;
    if (!c->filledOnce) return VPE_OK;
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    int flat = (z * c->g.NY + y) * c->g.NX + x, brick = -1;
    CUDA_TRY(c, cudaMemcpy(&brick, c->dBrickOf.p + flat, sizeof(int), cudaMemcpyDeviceToHost));
    if (brick < 0) return VPE_OK;
    *covered = 1;
    size_t texels = (size_t)c->g.N * c->g.N * c->g.N;
    if (half4) {
        // the hook returns [slice][row][col] half4: drop the row padding
        const int N = c->g.N, RS = c->g.rowStride;
        CUDA_TRY(c, cudaMemcpy2D(half4, (size_t)N * sizeof(uint2), c->dBricks.p + (size_t)brick * N * N * RS, (size_t)RS * sizeof(uint2),
                                 (size_t)N * sizeof(uint2), (size_t)N * N, cudaMemcpyDeviceToHost));
        if (c->bricksGray) {  // z-paired grey texel -> the half4 (r, r, r, density) it stands for
            for (size_t i = 0; i < texels; i++) {
                const uint16_t r = half4[4 * i], d = half4[4 * i + 1];
                half4[4 * i + 1] = r; half4[4 * i + 2] = r; half4[4 * i + 3] = d;
            }
        }
    }
    return VPE_OK;
}

int vpe_read_sample_bitmap(VpeContext* c, int x, int y, int z, uint32_t* words, int* covered) {
    if (!c || !covered) return VPE_E_INVALID_ARG;
    if (x < 0 || y < 0 || z < 0 || x >= c->g.NX || y >= c->g.NY || z >= c->g.NZ) return fail(c, VPE_E_INVALID_ARG, "metavoxel index out of range");
    DeviceScope deviceScope(c->device);
    *covered = 0;
    if (!c->filledOnce || !c->dOcc.p) return VPE_OK;
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    int flat = (z * c->g.NY + y) * c->g.NX + x, brick = -1;
    CUDA_TRY(c, cudaMemcpy(&brick, c->dBrickOf.p + flat, sizeof(int), cudaMemcpyDeviceToHost));
    if (brick < 0) return VPE_OK;
    *covered = 1;
    const size_t n = (size_t)c->g.N * c->g.N * c->occRowWords;
    if (words) CUDA_TRY(c, cudaMemcpy(words, c->dOcc.p + (size_t)brick * n, n * sizeof(unsigned), cudaMemcpyDeviceToHost));
    return VPE_OK;
}

int vpe_read_light_sheet(VpeContext* c, float* sheet) {
    if (!c || !sheet) return VPE_E_INVALID_ARG;
    DeviceScope deviceScope(c->device);
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    size_t n = (size_t)c->g.NX * c->g.N * c->g.NY * c->g.N;
    CUDA_TRY(c, cudaMemcpy(sheet, c->dSheet.p, n * sizeof(float), cudaMemcpyDeviceToHost));
    return VPE_OK;
}

int vpe_read_particle_list(VpeContext* c, int x, int y, int z, int32_t* idx, int cap, int* n) {
    if (!c || !n) return VPE_E_INVALID_ARG;
    if (x < 0 || y < 0 || z < 0 || x >= c->g.NX || y >= c->g.NY || z >= c->g.NZ) return fail(c, VPE_E_INVALID_ARG, "metavoxel index out of range");
    *n = 0;
    if (!c->prepared) return VPE_OK;
    DeviceScope deviceScope(c->device);
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    int flat = (z * c->g.NY + y) * c->g.NX + x, se[2];
    CUDA_TRY(c, cudaMemcpy(se, c->dCellStart.p + flat, sizeof(int) * 2, cudaMemcpyDeviceToHost));
    *n = se[1] - se[0];
    int take = std::min(*n, cap);
    if (idx && take > 0) CUDA_TRY(c, cudaMemcpy(idx, c->dPairs.p + se[0], sizeof(int) * take, cudaMemcpyDeviceToHost));
    return VPE_OK;
}

int vpe_read_metavoxel_position(VpeContext* c, int x, int y, int z, float pos[3]) {
    if (!c || !pos) return VPE_E_INVALID_ARG;
    if (!c->lightSet) return fail(c, VPE_E_NOT_READY, "vpe_set_light has not been called");
    if (x < 0 || y < 0 || z < 0 || x >= c->g.NX || y >= c->g.NY || z >= c->g.NZ) return fail(c, VPE_E_INVALID_ARG, "metavoxel index out of range");
    F3 p = mv_center(c->g, x, y, z);  // same routine the kernels call
    pos[0] = p.x; pos[1] = p.y; pos[2] = p.z;
    return VPE_OK;
}

int vpe_debug_div_rn(VpeContext* c, const float* a, const float* b, float* q, int n) {
    if (!c || !a || !b || !q || n < 0) return VPE_E_INVALID_ARG;
    if (n == 0) return VPE_OK;
    DeviceScope deviceScope(c->device);
    DevBuf<float> da, db, dq;
    CUDA_TRY(c, da.ensure(n)); CUDA_TRY(c, db.ensure(n)); CUDA_TRY(c, dq.ensure(n));
    CUDA_TRY(c, cudaMemcpyAsync(da.p, a, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(db.p, b, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    k_debug_div<<<div_up(n, 256), 256, 0, c->stream>>>(da.p, db.p, dq.p, n);
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaMemcpyAsync(q, dq.p, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    da.release(); db.release(); dq.release();
    return VPE_OK;
}

int vpe_read_slice_profile(VpeContext* c, int64_t* pairs, int64_t* covered, int64_t* samples) {
    if (!c) return VPE_E_INVALID_ARG;
    if (!c->dbg.profileSlices) return fail(c, VPE_E_NOT_READY, "VpeDebugOptions.profileSlices is off");
    DeviceScope deviceScope(c->device);
    const int NZ = c->g.NZ;
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    for (int z = 0; z < NZ; z++) {
        if (pairs) pairs[z] = (c->prepared && (int)c->slicePairs.size() == NZ + 1) ? (int64_t)c->slicePairs[z + 1] - c->slicePairs[z] : 0;
        if (covered) covered[z] = (c->prepared && (int)c->sliceStart.size() == NZ + 1) ? (int64_t)c->sliceStart[z + 1] - c->sliceStart[z] : 0;
    }
    if (samples) {
        std::vector<unsigned long long> h((size_t)NZ, 0ull);
        if (c->dSliceSamples.p) CUDA_TRY(c, cudaMemcpy(h.data(), c->dSliceSamples.p, sizeof(unsigned long long) * (size_t)NZ, cudaMemcpyDeviceToHost));
        for (int z = 0; z < NZ; z++) samples[z] = (int64_t)h[z];
    }
    return VPE_OK;
}

int vpe_get_stats(VpeContext* c, VpeStats* s) {
    if (!c || !s) return VPE_E_INVALID_ARG;
    DeviceScope deviceScope(c->device);
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (c->fillTimed) {
        cudaEventElapsedTime(&c->stats.fillMs, c->evFill0, c->evFill1);
        cudaEventElapsedTime(&c->stats.fillKernelMs, c->evFillK0, c->evFill1);
    }
    if (c->marchTimed) {
        cudaEventElapsedTime(&c->stats.marchMs, c->evMarch0, c->evMarch1);
        cudaEventElapsedTime(&c->stats.marchKernelMs, c->evMarchK0, c->evMarchK1);
        c->stats.raySamples = (int64_t)*c->hTotalSamples;
    }
    c->stats.brickPoolBytes = (int64_t)(c->dBricks.cap * sizeof(uint2));
    *s = c->stats;
    return VPE_OK;
}

const char* vpe_last_error(VpeContext* c) { return c ? c->err.c_str() : "null context"; }
int vpe_abi_version(void) { return VPE_ABI_VERSION; }
const char* vpe_backend(void) { return "cuda"; }

}  // extern "C"
