// vpe_common.cuh — device-visible parameter blocks shared by the host side (vpe_cuda.cu) and the
// kernels (vpe_kernels.cuh).  All pass-wide constants are computed once on the host with the
// normative arithmetic of vpe_math.cuh and handed to kernels by value.
#pragma once
#include "vpe_math.cuh"

namespace vpe {

// ≙ VolumeConstants / LightConstants / ParticleConstants of Fill.shader:62-93 plus the grid state of
// VPR.cs:106-109,370-394.
struct GridParams {
    int NX, NY, NZ, N, border;  // border already clamped to [0,N-2] (VPR.cs:528)
    int z0, z1;                 // light-axis slab owned by this context
    int binMode;
    float s, sb, Nf;            // mvScale.x, mvScaleWithBorder.x (VPR.cs:139), (float)N
    Affine L2W, W2L;            // dirLight.transform.localToWorldMatrix / worldToLocalMatrix
    F3 center, lsCenter;        // wsGridCenter, W2L * wsGridCenter (VPR.cs:380,420)
    F3 lightFwd;                // dirLight.transform.forward.normalized (VPR.cs:535)
    Affine Ab;                  // linear part of TRS(., lightRot, sb): binning + fill
    AffineInvLin Lb;
    Affine As;                  // linear part of TRS(., lightRot, s): march
    AffineInvLin Ls;
    float w2lcRow2[4];          // row 2 of lightCamera.transform.worldToLocalMatrix (Fill.shader:211)
    float oneVoxelSize;         // _MetavoxelScaleZ / _NumVoxels (Fill.shader:160)
    F3 lightStep;               // _LightForward * oneVoxelSize (Fill.shader:183)
    float ambient[3];
    float ds, opacityFactor;    // _DisplacementScale, _OpacityFactor
    int fade;                   // _FadeOutParticles
    float depthB, depthRcpA;    // Fill.shader:217-218: b, rcp(a)
    int cubeEdge;
    float worldReach;           // bound on |coordinate| of any voxel centre of the grid (pre-test error band)
    int rowStride;              // texels between stored brick rows: N + 8 when N % 16 == 0 (bank spread), else N
    int gray;                   // ambient r == g == b (same bits): bricks hold z-paired (r,density) texels
};

// Texel (x,y,z) of a brick. Texels are 8 bytes, x fastest; rows are rowStride texels apart: padded by
// half a 128-byte line when N % 16 == 0 so that vertically adjacent rows fall into different L1 banks
// (the march's warp tile reads 4 rows at about the same x in one load instruction).
VPE_HD size_t brick_texel_index(int N, int rowStride, int x, int y, int z) {
    return ((size_t)z * N + y) * rowStride + (size_t)x;
}

// world-space centre of metavoxel (x,y,z), VPR.cs:388-390
VPE_HD F3 mv_center(const GridParams& g, int x, int y, int z) {
    F3 off = f3((float)(g.NX / 2 - x) * g.s, (float)(g.NY / 2 - y) * g.s, (float)(g.NZ / 2 - z) * g.s);
    return xform_point(g.L2W, sub(g.lsCenter, off));
}

// per-particle fill input (≙ DisplacedParticle, VPR.cs:34-40: only mWorldToLocal and mOpacity are
// read by the shader, Fill.shader:169,174). 80 bytes so a warp-uniform read is five LDS.128.
struct __align__(16) ParticleFill {
    float m[3][4];
    float opacity;
    float rejectAbove;  // 0.25 + error band of the fused evaluations (k_particle_setup); see k_fill_columns
    float pad[2];
    float dq[3];        // mWorldToLocal's 3x3 block times the light step: particle-space motion per slice
    float dqLen2;       // dot(dq, dq)
};
constexpr int PARTICLE_FILL_VEC4 = 5;  // sizeof(ParticleFill) / 16

// per-particle binning record
struct ParticleBin {
    F3 ws;          // world-space centre (VPR.cs:418)
    float radius;   // size / 2
    int lo[3], hi[3];  // inclusive candidate cell range (already clipped to grid and slab)
};

// ≙ CameraConstants / VolumeConstants of March.shader:46-76 and the per-frame state of
// VPR.cs:637-648,716-763.
struct MarchParams {
    int W, H;
    float Wf, Hf, aspect;        // _ScreenRes, W/H (March.shader:190)
    float negRcpTan;             // -rcp(tan(_Fov/2)) (March.shader:193)
    float csZVolMin;             // March.shader:206-211
    float stepSize;              // March.shader:221-224
    float borderVoxelOffset;     // March.shader:245
    float sampleScale;           // 1 - 2*borderVoxelOffset (March.shader:258)
    int softDistance;
    float softRcp;               // rcp(_SoftDistance) (March.shader:269)
    float C2Mlin[3][3];          // 3x3 block of _CameraToMetavoxel (identical for every metavoxel)
    float c2wT[3];               // translation column of cameraToWorldMatrix
    int zBoundary;               // VPR.cs:648
    int zOverBegin, zOverEnd;    // phase 1 slices [begin,end) in this slab (ascending, OVER)
    int zUnderBegin, zUnderEnd;  // phase 2 slices [begin,end) in this slab (ascending, UNDER)
    float earlyOut;              // 0 = exact
    int numPixels;               // W*H or the length of the pixel list
    int maxSamplesPerMv;         // hang guard: (int)(sqrt(3)/stepSize) + 2
    int wrap;                    // border == 0: repeat addressing can trigger (VPR.cs:770)
    int tileLog2W;               // warp pixel tile = 2^tileLog2W x (32 >> tileLog2W)
    int blockYBase;              // first row of CTAs of this launch (the host path launches the image in bands)
    int rowStride;               // = GridParams::rowStride
    int gray;                    // layout of the bricks being marched (GridParams::gray at fill time)
    int targetFormat, debugMode; // VpeMarchOptions (legacy kernel only)
    int numCovered;              // _NumMetavoxelsCovered (VPR.cs:755)
};

}  // namespace vpe
