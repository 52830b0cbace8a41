// VolumetricParticleRenderer.hpp — C++ host mirror of the reference's MonoBehaviour
// MetavoxelEngine.VolumetricParticleRenderer (Assets/Main Scene/VolumetricParticleRenderer.cs, "VPR.cs")
// for the Fill Volume + Ray March path, over the C-ABI of include/vpe.h.
//
// The reference's host is C# against UnityEngine; this image has no .NET toolchain, so the host side
// above the C-ABI is written in C++ (and in Python, vpe_b200/renderer.py) with the reference's own
// names, argument meaning and frame structure:
//   public inspector fields  VPR.cs:82-101      -> public members of the same name
//   Start()                  VPR.cs:132-149     -> Start()
//   OnPostRender()           VPR.cs:181-220     -> OnPostRender(particles, camera)  (fill every
//                                                  updateInterval frames, march every frame)
//   FillMetavoxels()         VPR.cs:495-520     -> FillMetavoxels()
//   RenderMetavoxels()       VPR.cs:637-713     -> RenderMetavoxels()
//   Set* GUI setters         VPR.cs:1040-1119   -> same names
// What Unity supplied implicitly (the particle system's particles, the light / camera transforms,
// the particlesRT render target) is passed explicitly.  Errors: Unity logs and carries on
// (VPR.cs:352,790); here every call returns the VPE_* status and lastError() holds the message.
//
// The library (libvpe_cuda.so) is bound at run time through a table of function pointers so that the
// same host code can be pointed at any implementation of include/vpe.h.
#pragma once
#include <dlfcn.h>

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/vpe.h"

namespace MetavoxelEngine {

struct Vector3 {
    float x = 0, y = 0, z = 0;
};

// Entry points of include/vpe.h resolved from a shared object.
struct VpeApi {
    void* handle = nullptr;
#define VPE_FN(name) decltype(&::name) name = nullptr;
    VPE_FN(vpe_default_config) VPE_FN(vpe_create) VPE_FN(vpe_destroy) VPE_FN(vpe_set_config) VPE_FN(vpe_set_light)
    VPE_FN(vpe_set_displacement_cubemap) VPE_FN(vpe_set_light_depth_map) VPE_FN(vpe_fill) VPE_FN(vpe_march)
    VPE_FN(vpe_march_pixels) VPE_FN(vpe_get_stats) VPE_FN(vpe_last_error) VPE_FN(vpe_abi_version) VPE_FN(vpe_backend)
    VPE_FN(vpe_read_brick) VPE_FN(vpe_read_light_sheet)
    VPE_FN(vpe_render_light_depth_map) VPE_FN(vpe_set_march_options) VPE_FN(vpe_composite_scene)
#undef VPE_FN
    bool load(const char* path, std::string* why) {
        handle = dlopen(path, RTLD_NOW | RTLD_LOCAL);
        if (!handle) { if (why) *why = dlerror(); return false; }
#define VPE_BIND(name)                                                        \
    name = reinterpret_cast<decltype(name)>(dlsym(handle, #name));           \
    if (!name) { if (why) *why = std::string("missing symbol ") + #name; return false; }
        VPE_BIND(vpe_default_config) VPE_BIND(vpe_create) VPE_BIND(vpe_destroy) VPE_BIND(vpe_set_config) VPE_BIND(vpe_set_light)
        VPE_BIND(vpe_set_displacement_cubemap) VPE_BIND(vpe_set_light_depth_map) VPE_BIND(vpe_fill) VPE_BIND(vpe_march)
        VPE_BIND(vpe_march_pixels) VPE_BIND(vpe_get_stats) VPE_BIND(vpe_last_error) VPE_BIND(vpe_abi_version) VPE_BIND(vpe_backend)
        VPE_BIND(vpe_read_brick) VPE_BIND(vpe_read_light_sheet)
        VPE_BIND(vpe_render_light_depth_map) VPE_BIND(vpe_set_march_options) VPE_BIND(vpe_composite_scene)
#undef VPE_BIND
        if (vpe_abi_version() != VPE_ABI_VERSION) { if (why) *why = "ABI version mismatch"; return false; }
        return true;
    }
};

class VolumetricParticleRenderer {
public:
    // ---- what the inspector binds in Unity (VPR.cs:72-79), passed explicitly here ----
    VpeTransform dirLight{{0, 0, -44.34f}, {0.185594f, 0, 0, 0.982627f}};  // scene:6763-6764, 6792-6793
    VpeTransform particleSys{{0, 5, 11.2f}, {0, 1, 0, 0}};                 // scene:2271-2276 (emitter transform)
    Vector3 gridCenter;                                                   // scene:3712
    // ---- metavoxel layout/size (VPR.cs:82-85) ----
    int numMetavoxelsX = 10, numMetavoxelsY = 10, numMetavoxelsZ = 10;
    Vector3 mvScale{3, 3, 3};
    int numVoxelsInMetavoxel = 32;
    int numBorderVoxels = 1;
    // ---- rendering vars (VPR.cs:88-101) ----
    int updateInterval = 2;
    int rayMarchSteps = 64;
    Vector3 ambientColor{0.2f, 0.2f, 0.2f};
    float fDisplacementScale = 0.7f;
    bool fadeOutParticles = false;
    float opacityFactor = 0.04f;
    int softParticleStepDistance = 20;
    // ---- counters the reference keeps for its debug uniforms (VPR.cs:124-125) ----
    int numParticlesEmitted = 0;
    int numMetavoxelsCovered = 0;

    explicit VolumetricParticleRenderer(const VpeApi& api, int device = 0) : api_(api), device_(device) {}
    ~VolumetricParticleRenderer() { if (ctx_) api_.vpe_destroy(ctx_); }
    VolumetricParticleRenderer(const VolumetricParticleRenderer&) = delete;
    VolumetricParticleRenderer& operator=(const VolumetricParticleRenderer&) = delete;

    // VPR.cs:132-149: creates the grid and its resources. displacementR8 = the R channel of
    // Assets/Textures/DisplacementTexture.cubemap, 6 x edge x edge.
    int Start(const uint8_t* displacementR8, int edge) {
        fadeOutParticles = false;  // VPR.cs:134
        VpeConfig cfg = config();
        int rc = api_.vpe_create(&cfg, device_, &ctx_);
        if (rc) { err_ = "vpe_create failed (invalid configuration or no CUDA device)"; return rc; }
        if ((rc = check(api_.vpe_set_displacement_cubemap(ctx_, displacementR8, edge)))) return rc;
        frameCount_ = 0;
        return UpdateMetavoxelPositions();
    }

    // VPR.cs:370-394 (+ UpdatePositionOfCameraAtLight, VPR.cs:361-367)
    int UpdateMetavoxelPositions() {
        const float c[3] = {gridCenter.x, gridCenter.y, gridCenter.z};
        return check(api_.vpe_set_light(ctx_, &dirLight, c));
    }

    // lightDepthMap (VPR.cs:184,274): depth01 = (NY*N) x (NX*N) floats or nullptr for "no occluders".
    int SetLightDepthMap(const float* depth01) { return check(api_.vpe_set_light_depth_map(ctx_, depth01)); }
    // lightCamera.RenderWithShader(generateLightDepthMapShader), VPR.cs:184: occluder triangles in world space
    int RenderLightDepthMap(const float* trianglesWorld, int numTriangles) { return check(api_.vpe_render_light_depth_map(ctx_, trianglesWorld, numTriangles)); }
    // debug views (VPR.cs:96-99,744-761), 8-bit particlesRT (VPR.cs:228), scene depth test (VPR.cs:204)
    int SetMarchOptions(const VpeMarchOptions& o) { return check(api_.vpe_set_march_options(ctx_, &o)); }
    // Graphics.Blit(particlesRT, mainSceneRT, matBlendParticles), VPR.cs:210
    int CompositeParticles(const float* particlesRT, float* mainSceneRT, int numPixels, int targetFormat = 0) {
        return check(api_.vpe_composite_scene(ctx_, particlesRT, mainSceneRT, numPixels, targetFormat));
    }

    // VPR.cs:181-220. particles = ParticleSystem.GetParticles(); rgba = particlesRT as float4 (H*W*4).
    int OnPostRender(const VpeParticle* particles, int numParticles, const VpeCamera& camera, float* rgba) {
        int rc = VPE_OK;
        if (updateInterval < 1 || frameCount_ % updateInterval == 0) {  // VPR.cs:186
            if ((rc = check(api_.vpe_set_config(ctx_, ptr(config()))))) return rc;
            if ((rc = UpdateMetavoxelPositions())) return rc;            // VPR.cs:188-195
            if ((rc = FillMetavoxels(particles, numParticles))) return rc;
        }
        frameCount_++;
        return RenderMetavoxels(camera, rgba);                           // VPR.cs:207
    }

    // BinParticlesToMetavoxels + FillMetavoxels, VPR.cs:397-520
    int FillMetavoxels(const VpeParticle* particles, int numParticles) {
        int rc = check(api_.vpe_fill(ctx_, particles, numParticles, &particleSys));
        VpeStats st;
        if (!rc && !api_.vpe_get_stats(ctx_, &st)) { numParticlesEmitted = st.numParticles; numMetavoxelsCovered = st.numMetavoxelsCovered; }
        return rc;
    }

    // VPR.cs:637-713
    int RenderMetavoxels(const VpeCamera& camera, float* rgba, int32_t* samplesPerPixel = nullptr) {
        int rc = check(api_.vpe_set_config(ctx_, ptr(config())));
        if (rc) return rc;
        return check(api_.vpe_march(ctx_, &camera, rgba, samplesPerPixel));
    }

    // ---- GUI callback setters, VPR.cs:1040-1119 ----
    void SetDisplacementScale(float ds) { fDisplacementScale = ds; }
    void SetRayMarchSteps(float steps) { rayMarchSteps = (int)steps; }
    int SetGridScale(float s) { mvScale = Vector3{s, s, s}; int rc = check(api_.vpe_set_config(ctx_, ptr(config()))); return rc ? rc : UpdateMetavoxelPositions(); }
    void SetFadeOutParticles(bool fade) { fadeOutParticles = fade; }
    void SetParticleOpacityFactor(float f) { opacityFactor = f; }
    void SetFadeParticles(bool fade) { fadeOutParticles = fade; }
    void SetUpdateInterval(float interval) { updateInterval = (int)interval; }
    void SetSoftParticleDistance(float stepDistance) { softParticleStepDistance = (int)stepDistance; }

    int GetStats(VpeStats* st) { return check(api_.vpe_get_stats(ctx_, st)); }
    const std::string& lastError() const { return err_; }
    VpeContext* context() { return ctx_; }

private:
    VpeConfig config() const {
        VpeConfig c;
        api_.vpe_default_config(&c);
        c.numMetavoxelsX = numMetavoxelsX; c.numMetavoxelsY = numMetavoxelsY; c.numMetavoxelsZ = numMetavoxelsZ;
        c.mvScale = mvScale.x;  // the reference only supports cubes (VPR.cs:422,425,445)
        c.numVoxelsInMetavoxel = numVoxelsInMetavoxel;
        c.numBorderVoxels = numBorderVoxels;
        c.rayMarchSteps = rayMarchSteps;
        c.ambientColor[0] = ambientColor.x; c.ambientColor[1] = ambientColor.y; c.ambientColor[2] = ambientColor.z;
        c.displacementScale = fDisplacementScale;
        c.fadeOutParticles = fadeOutParticles ? 1 : 0;
        c.opacityFactor = opacityFactor;
        c.softParticleStepDistance = softParticleStepDistance;
        return c;
    }
    static const VpeConfig* ptr(const VpeConfig& c) { return &c; }
    int check(int rc) {
        if (rc) err_ = api_.vpe_last_error(ctx_);
        return rc;
    }
    const VpeApi& api_;
    int device_;
    VpeContext* ctx_ = nullptr;
    int frameCount_ = 0;
    std::string err_;
};

}  // namespace MetavoxelEngine
