/*
 * vpe.h — C-ABI of the B200-native sparse volumetric particle engine.
 *
 * Drop-in boundary for ONE hot path of rajabala/Volumetric-Particles-For-Unity:
 *   Fill Volume  (BinParticlesToMetavoxels + FillMetavoxels + FillVolume.shader)
 *   Ray March    (RenderMetavoxels + RayMarchVoxel.shader)
 *
 * The reference has no FFI boundary of its own: the path sits behind Unity's managed
 * graphics API (Graphics.Blit / DrawMeshNow / Material.Set*).  Each entry point below cites
 * the reference interface it replaces.  Citations are relative to /root/reference:
 *   VPR.cs       = Assets/Main Scene/VolumetricParticleRenderer.cs
 *   Fill.shader  = Assets/Shaders/Metavoxel/FillVolume.shader
 *   March.shader = Assets/Shaders/Metavoxel/RayMarchVoxel.shader
 *
 * Two libraries export exactly these symbols:
 *   libvpe_cuda.so  — the product: hand-written sm_100a CUDA kernels (no CPU fallback).
 *   libvpe_ref.so   — the CPU oracle under oracle/ (test infrastructure only).
 *
 * Conventions
 *   - plain C, no exceptions cross the boundary; every call returns VPE_OK (0) or a negative
 *     VPE_E_* code; the message is available from vpe_last_error().
 *   - the caller owns every pointer it passes; inputs are consumed before the call returns;
 *     outputs are caller-allocated.  "_device" variants take CUDA device pointers.
 *   - a context is not thread-safe (the reference runs on Unity's main thread only).
 *   - Unity conventions: left-handed world, +Z forward, quaternions (x,y,z,w), column vectors.
 */
#ifndef VPE_H_
#define VPE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VPE_ABI_VERSION 2   /* 2: VpeStats.raySamplesSkipped, VpeDebugOptions, vpe_fill_linked, slice profile / bitmap / division hooks */

enum {
    VPE_OK = 0,
    VPE_E_INVALID_ARG = -1,   /* NULL pointer, bad dimension, border out of range ...      */
    VPE_E_NOT_READY = -2,     /* march before fill, missing cubemap ...                    */
    VPE_E_CUDA = -3,          /* a CUDA runtime call or kernel failed                      */
    VPE_E_OUT_OF_MEMORY = -4, /* brick pool does not fit                                   */
    VPE_E_UNSUPPORTED = -5    /* entry point not available in this library                 */
};

enum { VPE_BIN_REFERENCE = 0, /* bug-compatible candidate range, VPR.cs:425-438           */
       VPE_BIN_EXACT = 1      /* test every cell the enlarged box can reach               */ };

/* Unity Transform.{position,rotation} (VPR.cs:136,188-192,380,418,616). Scale is 1. */
typedef struct VpeTransform {
    float position[3];
    float rotation[4]; /* quaternion x,y,z,w */
} VpeTransform;

/* Public inspector fields of VolumetricParticleRenderer, VPR.cs:82-101, plus the constants the
 * reference hard-codes for its light camera (VPR.cs:342,365). */
typedef struct VpeConfig {
    int32_t numMetavoxelsX, numMetavoxelsY, numMetavoxelsZ; /* VPR.cs:82                     */
    float mvScale;                    /* VPR.cs:83 — cubic metavoxels only (VPR.cs:422,425)  */
    int32_t numVoxelsInMetavoxel;     /* VPR.cs:84  (any N >= 2; the shader's 32 cap lifted) */
    int32_t numBorderVoxels;          /* VPR.cs:85  0 <= b <= (N-2)/2                        */
    int32_t rayMarchSteps;            /* VPR.cs:89  steps per metavoxel                      */
    float ambientColor[3];            /* VPR.cs:90                                           */
    float displacementScale;          /* VPR.cs:94                                           */
    int32_t fadeOutParticles;         /* VPR.cs:95                                           */
    float opacityFactor;              /* VPR.cs:100                                          */
    int32_t softParticleStepDistance; /* VPR.cs:101                                          */
    float lightNear, lightFar;        /* VPR.cs:342 (0.3, 1000)                              */
    float lightCameraDistance;        /* VPR.cs:365 (200)                                    */
    int32_t binMode;                  /* VPE_BIN_*                                           */
    float marchEarlyOutTransmittance; /* 0 = exact reference semantics (default); > 0 lets the
                                         CUDA march stop a ray once 1-alpha falls below it   */
    int32_t slabZBegin, slabZEnd;     /* light-axis slab [begin,end) owned by this context;
                                         0,0 = whole grid (single GPU)                       */
} VpeConfig;

/* The ParticleSystem.Particle fields the reference reads (VPR.cs:418,425,583-586). 28 bytes. */
typedef struct VpeParticle {
    float position[3]; /* emitter-local                           */
    float size;        /* diameter, world units                   */
    float rotationDeg; /* about the emitter's forward axis        */
    float lifetime;    /* remaining                               */
    float startLifetime;
} VpeParticle;

/* Camera.main.{transform, fieldOfView}, Screen.{width,height} (VPR.cs:616,642,733-737,778). */
typedef struct VpeCamera {
    VpeTransform transform;
    float fovYDegrees;
    int32_t width, height;
} VpeCamera;

typedef struct VpeStats {
    int32_t numParticles;        /* VPR.cs:124                                              */
    int32_t numMetavoxelsCovered;/* VPR.cs:125 (in this context's slab)                     */
    int64_t numParticlePairs;    /* sum of list lengths after binning                       */
    int64_t voxelsFilled;        /* numMetavoxelsCovered * N^3                              */
    int64_t raySamples;          /* iterations of March.shader:254-279 in the last march    */
    int32_t zBoundary;           /* VPR.cs:648 of the last march                            */
    int32_t fillLaunches;        /* kernels launched by the last fill (0 in the oracle)     */
    int32_t marchLaunches;       /* kernels launched by the last march                      */
    float fillMs, marchMs;       /* device time of the last fill / march (CUDA events);
                                    wall time in the oracle                                 */
    int64_t brickPoolBytes;
    float fillKernelMs;          /* device time of the per-slice fill kernels alone (last fill)  */
    float marchKernelMs;         /* device time of the march kernel alone (last march)           */
    int64_t raySamplesSkipped;   /* of raySamples: samples in empty space (all 8 texels of the footprint have density 0)
                                    that the production march kernels do not fetch; counted by vpe_march_footprint   */
} VpeStats;

typedef struct VpeContext VpeContext;

/* Defaults = the demo scene's inspector values (Assets/Volumetric_Particle_System.unity:9013-9026). */
void vpe_default_config(VpeConfig* cfg);

/* ≙ Start()/CreateResources()/CreateMetavoxelGrid(), VPR.cs:132-149,224-317.
 * device: CUDA ordinal (ignored by the oracle). */
int vpe_create(const VpeConfig* cfg, int device, VpeContext** out);
int vpe_destroy(VpeContext* ctx);

/* ≙ the GUI setters VPR.cs:1040-1119. Grid dims / N / slab may not change after create. */
int vpe_set_config(VpeContext* ctx, const VpeConfig* cfg);

/* ≙ UpdateMetavoxelPositions + UpdatePositionOfCameraAtLight, VPR.cs:361-394;
 * light = dirLight.transform, gridCenter = gridCenter.transform.position. */
int vpe_set_light(VpeContext* ctx, const VpeTransform* light, const float gridCenter[3]);

/* ≙ Material "_DisplacementTexture" (Fill.shader:57,116): 6 faces x edge x edge, R channel,
 * faces in Unity order +X,-X,+Y,-Y,+Z,-Z. */
int vpe_set_displacement_cubemap(VpeContext* ctx, const uint8_t* r8, int edge);

/* ≙ lightDepthMap (VPR.cs:274, Fill.shader:59,216): (NY*N) rows x (NX*N) floats in [0,1];
 * NULL = no occluders (all 1.0). */
int vpe_set_light_depth_map(VpeContext* ctx, const float* depth01);

/* ≙ lightCamera.RenderWithShader(generateLightDepthMapShader), VPR.cs:184, with the light camera of
 * InitCameraAtLight (VPR.cs:320-367) and GenerateLightDepthMap.shader:6 (Cull Front, ZWrite On, ZTest
 * Less): rasterises the scene's occluders, numTriangles x 3 world-space vertices (9 floats each, Unity
 * winding: clockwise = front), into the context's light depth map; later fills read it (Fill.shader:216).
 * vpe_read_light_depth_map is the test hook ((NY*N) rows x (NX*N) floats). */
int vpe_render_light_depth_map(VpeContext* ctx, const float* trianglesWorld, int numTriangles);
int vpe_read_light_depth_map(VpeContext* ctx, float* depth01);

/* ≙ BinParticlesToMetavoxels + FillMetavoxels, VPR.cs:397-520 (+ FillVolume.shader).
 * emitter = particleSys.transform. */
int vpe_fill(VpeContext* ctx, const VpeParticle* particles, int n, const VpeTransform* emitter);

/* ≙ RenderMetavoxels, VPR.cs:637-713 (+ RayMarchVoxel.shader).
 * rgba: height*width*4 floats, premultiplied RGB + coverage alpha (≙ particlesRT, VPR.cs:228,
 * as float4, no 8-bit quantisation); row 0 is screen-space y = 0 of March.shader:189.
 * samples: optional height*width int32 = loop iterations per pixel (≙ _ShowNumSamples). */
int vpe_march(VpeContext* ctx, const VpeCamera* cam, float* rgba, int32_t* samples);

/* What surrounds the march in the reference's frame; defaults = all zero.
 * Non-default options are rendered by the general march kernel, not the specialised fast one. */
typedef struct VpeMarchOptions {
    int32_t targetFormat;    /* 0 = float4; 1 = UNORM8 like the reference's ARGB32 particlesRT (VPR.cs:228):
                                the target is quantised after every metavoxel's ROP blend                 */
    int32_t debugMode;       /* 0 = off; 1 = draw order (March.shader:123-138,170); 2 = blend function
                                (:174-181); 3 = sample-count bands (:283-299)                              */
    const float* sceneDepth; /* host, sceneHeight*sceneWidth eye-space depths of the opaque scene, or NULL:
                                ≙ `ZTest Less` against mainSceneRT.depthBuffer (March.shader:14, VPR.cs:204):
                                a metavoxel's fragment exists only where its back face is nearer            */
    int32_t sceneWidth, sceneHeight;
} VpeMarchOptions;
int vpe_set_march_options(VpeContext* ctx, const VpeMarchOptions* options);

/* ≙ Graphics.Blit(particlesRT, mainSceneRT, matBlendParticles), VPR.cs:210, with
 * CompositeParticles.shader:10 `Blend One OneMinusSrcAlpha, One One`: scene.rgb = p.rgb + (1-p.a)*scene.rgb,
 * scene.a += p.a; host buffers, numPixels*4 floats each; targetFormat 1 quantises the result to UNORM8. */
int vpe_composite_scene(VpeContext* ctx, const float* particlesRgba, float* sceneRgba, int numPixels,
                        int targetFormat);

/* Same, for a list of pixel indices (y*width + x); rgba is n*4, samples n. Used for parity at
 * sizes where the scalar oracle cannot render the full image. */
int vpe_march_pixels(VpeContext* ctx, const VpeCamera* cam, const int32_t* pixels, int n,
                     float* rgba, int32_t* samples);

/* ---- device-resident variants (CUDA library only; the oracle returns VPE_E_UNSUPPORTED) ---- */
int vpe_set_stream(VpeContext* ctx, void* cudaStream);
int vpe_fill_device(VpeContext* ctx, const VpeParticle* particles_dev, int n,
                    const VpeTransform* emitter);
int vpe_march_device(VpeContext* ctx, const VpeCamera* cam, float* rgba_dev, int32_t* samples_dev);

/* ---- multi-GPU light-axis slabs ----
 * vpe_fill_prepare bins the particles into this context's slab and sizes the brick pool;
 * vpe_fill_region sweeps the metavoxel columns [x0,x1) x [y0,y1) through the slab's z-slices,
 * reading and updating the context's light sheet; the caller moves sheet tiles between slabs
 * (NCCL send/recv on the pointer returned by vpe_light_sheet_device). vpe_fill == prepare +
 * region(whole grid). */
int vpe_fill_prepare(VpeContext* ctx, const VpeParticle* particles, int n,
                     const VpeTransform* emitter, int particlesOnDevice);
int vpe_fill_region(VpeContext* ctx, int x0, int x1, int y0, int y1);
/* The same fill split at its only cross-slab dependency, so that slabs scale: vpe_fill_density runs
 * the particle loop of the whole slab (Fill.shader:155-208: density and ambient-occlusion term of
 * every voxel; no light involved, every GPU runs it at once), vpe_fill_sweep_region then runs the
 * light sweep (Fill.shader:211-269) of a metavoxel-column region through the slab's slices, reading
 * and updating the light sheet like vpe_fill_region. density + sweep(region) == fill_region(region),
 * bit for bit. Between the two calls the bricks hold an intermediate and must not be marched. */
int vpe_fill_density(VpeContext* ctx);
int vpe_fill_sweep_region(VpeContext* ctx, int x0, int x1, int y0, int y1);
float* vpe_light_sheet_device(VpeContext* ctx);
/* The same sweep with the hand-over of the sheet between neighbouring slabs inside the kernel, over
 * NVLink peer memory (CUDA library only): every rank owns a link buffer (inbox for the sheet values of
 * the slab nearer the light + per-block flags). vpe_sheet_link_create allocates it and returns a
 * 64-byte CUDA IPC handle for other processes and/or the device pointer for contexts of this process;
 * vpe_sheet_link_connect maps the neighbours' buffers (NULL = no neighbour on that side; handlesAreIpc:
 * the arguments point to IPC handles, else to device pointers as returned in *devPtr);
 * vpe_fill_sweep_linked sweeps all metavoxel columns: each block of voxel columns waits for its values
 * from upstream, sweeps the slab, stores its exit values into the downstream inbox and raises the flag.
 * All ranks must call it once per fill, after vpe_fill_density. Result: identical to the other paths,
 * bit for bit. vpe_sheet_link_status reports waits that gave up (a peer that never arrived). */
int vpe_sheet_link_create(VpeContext* ctx, void* ipcHandle64, void** devPtr);
int vpe_sheet_link_connect(VpeContext* ctx, const void* upstream, const void* downstream, int handlesAreIpc);
int vpe_fill_sweep_linked(VpeContext* ctx);
/* The whole linked fill of this rank after vpe_fill_prepare, in the cheapest form its place in the chain allows: the rank
 * with no upstream neighbour (the slab nearest the light, VPR.cs:505 starts there) has nothing to wait for and runs the
 * FUSED kernel of vpe_fill, which stores its exit values into the downstream inbox itself; every other rank runs
 * vpe_fill_density + vpe_fill_sweep_linked. All ranks call it once per fill. Same bits as every other path. */
int vpe_fill_linked(VpeContext* ctx);
int vpe_sheet_link_status(VpeContext* ctx, int* timeouts);
/* Slab-local march: two premultiplied RGBA partial images (device, height*width*4 floats each):
 * over = the slab's slices <= zBoundary composited back-to-front (phase 1, VPR.cs:652-681),
 * under = its slices > zBoundary composited front-to-back (phase 2, VPR.cs:688-711). */
int vpe_march_partial_device(VpeContext* ctx, const VpeCamera* cam, float* over_dev,
                             float* under_dev, int32_t* samples_dev);
/* The same with the exchange of the partial images inside the march kernel (CUDA library only): every rank owns
 * a receive buffer for the image rows it composites (rows [q*per, (q+1)*per), per = ceil(height / world));
 * vpe_image_link_create allocates it (64-byte CUDA IPC handle and/or device pointer out), vpe_image_link_connect
 * maps all ranks' buffers (peers[q] = rank q's handle or pointer to its device pointer; peers[rank] is ignored).
 * vpe_march_linked marches the slab and stores every pixel's two partials straight into the receive buffer of the
 * rank that owns its row (peer stores over NVLink), then raises this rank's flag on every rank;
 * vpe_composite_linked waits for all ranks' flags and composites this rank's band (per*width*4 floats, device) in
 * the reference's order. No collective is involved. All ranks call both once per frame. */
int vpe_image_link_create(VpeContext* ctx, int world, int rank, int width, int height, void* ipcHandle64, void** devPtr);
int vpe_image_link_connect(VpeContext* ctx, const void* const* peers, int handlesAreIpc);
int vpe_march_linked(VpeContext* ctx, const VpeCamera* cam, int32_t* samples_dev);
int vpe_composite_linked(VpeContext* ctx, float* band_rgba_dev);
int vpe_image_link_status(VpeContext* ctx, int* timeouts);
/* Composite R slabs' partial images for a pixel range in reference order (phase 1 slabs in
 * ascending z with OVER, then phase 2 slabs in ascending z with UNDER). parts_dev[2*r+0] = over,
 * parts_dev[2*r+1] = under of slab r (device pointers, numPixels*4 floats each). */
int vpe_composite_device(VpeContext* ctx, const float* const* parts_dev, int numSlabs,
                         int numPixels, float* rgba_dev);

/* ---- debug / experiment switches (CUDA library; the oracle accepts and ignores them) ----
 * All zero = production behaviour. Read by the calls that follow; a change of noGray / noRowPad changes the
 * brick layout, so the volume must be filled again before the next march. */
typedef struct VpeDebugOptions {
    int32_t marchKernel;    /* 0 = production (k_march_flat); 1 = general kernel (the shader's unfused sequence, the one
                               march options and border 0 use); 2 = round 1's per-fragment loop (k_march), for comparison */
    int32_t noSkip;         /* 1 = sample every step: ignore the empty-space bitmap                                       */
    int32_t noGray;         /* 1 = keep half4 (r,g,b,density) texels even when the ambient colour is grey                  */
    int32_t noRowPad;       /* 1 = brick rows N texels apart (no 64-byte pad)                                             */
    int32_t marchBands;     /* host-buffer march: 0 = default (6 bands, copies overlapped), 1 = one launch + one copy, n   */
    int32_t marchTileLog2W; /* warp pixel tile: 0 = default (8x4), else 1 + log2(width): 1 = 1x32 ... 6 = 32x1             */
    int32_t linkSpinMs;     /* sheet / image link: how long a kernel waits for a peer, 0 = default (2000)                  */
    int32_t sweepOverlap;   /* 1 = the linked sweep runs as a persistent kernel CONCURRENTLY with the density pass (own stream,
                               per-block flags) instead of after it. Measured (profiles/r02_scaling.md): the particle loop
                               is issue-bound and loses more than the sweep gains at 8 GPUs, so this is off by default    */
    int32_t noTmaSweep;     /* 1 = register-staged sweep kernels (k_sweep_columns / k_sweep_overlapped) instead of the TMA pipeline  */
    int32_t profileSlices;  /* 1 = keep per-slice work figures of the fills and marches that follow (vpe_read_slice_profile)        */
    int32_t noHeadFused;    /* 1 = vpe_fill_linked splits the fill (density + linked sweep) on the head rank too, as round 1 did     */
    int32_t reserved[5];
} VpeDebugOptions;
int vpe_set_debug_options(VpeContext* ctx, const VpeDebugOptions* options);
/* Work per light-axis slice of the last fill / march of this context (NZ entries each, NULL = not wanted; slices outside the
 * context's slab are 0): (particle, metavoxel) pairs and covered metavoxels of the fill, ray samples of the march. The slab
 * renderer balances the slab boundaries with them (VPR.cs:505 walks the slices in order; their cost is far from uniform).
 * Needs VpeDebugOptions.profileSlices; the oracle returns VPE_E_UNSUPPORTED. */
int vpe_read_slice_profile(VpeContext* ctx, int64_t* pairs, int64_t* coveredMetavoxels, int64_t* raySamples);

/* ---- measurement ----
 * Number of distinct volume texels in the union of all samples' 8-texel trilinear footprints for
 * this camera (the march's compulsory read set = 8 B x this + 16 B x pixels).
 * Runs an instrumented march (never timed). */
int vpe_march_footprint(VpeContext* ctx, const VpeCamera* cam, int64_t* uniqueTexels);

/* ---- test hooks ---- */
/* brick (x,y,z) as N*N*N half4 in [slice][row][col] order (≙ mvFillTextures[z,y,x], VPR.cs:312);
 * returns 1 in *covered if the metavoxel has particles (VPR.cs:511), else 0 and no data. */
int vpe_read_brick(VpeContext* ctx, int x, int y, int z, uint16_t* half4, int* covered);
/* The march's empty-space bitmap of metavoxel (x,y,z) (CUDA library only; the reference has no such structure):
 * words [z0][y0][ceil(N/32)], bit x0 set iff one of the 8 texels of the trilinear footprint based at (x0,y0,z0) -
 * texels x0..x0+1, y0..y0+1, z0..z0+1 inside the brick - has a non-zero stored density. The tests compare it with
 * the brick itself: a bit that is wrongly clear would change the image, one that is wrongly set only costs time. */
int vpe_read_sample_bitmap(VpeContext* ctx, int x, int y, int z, uint32_t* words, int* covered);
/* light sheet (≙ lightPropogationUAV, VPR.cs:266): (NY*N) rows x (NX*N) floats. */
int vpe_read_light_sheet(VpeContext* ctx, float* sheet);
/* particle indices binned to metavoxel (x,y,z), in list order; returns count in *n (cap = size
 * of idx). ≙ mvGrid[z,y,x].mParticlesCovered, VPR.cs:453. */
int vpe_read_particle_list(VpeContext* ctx, int x, int y, int z, int32_t* idx, int cap, int* n);
/* world-space centre of metavoxel (x,y,z) ≙ mvGrid[z,y,x].mPos, VPR.cs:390. */
int vpe_read_metavoxel_position(VpeContext* ctx, int x, int y, int z, float pos[3]);

/* q[i] = the fill kernel's inlined a[i] / b[i] (host arrays): the three divisions of the in-sphere body and 1/(1+density)
 * use the fast path of div.rn.f32; the tests fuzz it against IEEE division over the operand ranges that occur.
 * CUDA library only (the oracle divides). */
int vpe_debug_div_rn(VpeContext* ctx, const float* a, const float* b, float* q, int n);

int vpe_get_stats(VpeContext* ctx, VpeStats* stats);
const char* vpe_last_error(VpeContext* ctx);
int vpe_abi_version(void);
/* "cuda" or "oracle" */
const char* vpe_backend(void);

#ifdef __cplusplus
}
#endif
#endif /* VPE_H_ */
