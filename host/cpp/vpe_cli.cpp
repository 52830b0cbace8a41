// vpe_cli — headless driver of the C++ host mirror: renders `frames` frames of a particle file through
// MetavoxelEngine::VolumetricParticleRenderer and writes the last frame's RGBA (raw float32) + a summary.
//   vpe_cli <lib.so> <cubemap_r8.bin> <particles.f32 (n x 7)> <grid> <voxels> <mvScale> <width> <height> <camZ> <frames> <out.rgba> [occluders.f32 (n x 9)]
// Used by tests/test_host_cpp.py with libvpe_cuda.so (GPU) and, as a checker of the host logic only,
// with the oracle library (CPU).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "VolumetricParticleRenderer.hpp"

using namespace MetavoxelEngine;

static std::vector<unsigned char> read_file(const char* path) {
    std::vector<unsigned char> v;
    FILE* f = fopen(path, "rb");
    if (!f) return v;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    v.resize(n);
    if (fread(v.data(), 1, n, f) != (size_t)n) v.clear();
    fclose(f);
    return v;
}

int main(int argc, char** argv) {
    // optional 12th argument: a file of occluder triangles (n x 9 float32, world space). The rest of the reference's
    // frame is then run as well (VPR.cs:184,204,210): light depth map, 8-bit particlesRT, blend onto a blue scene.
    if (argc != 12 && argc != 13) { fprintf(stderr, "usage: see the header of vpe_cli.cpp\n"); return 2; }
    VpeApi api;
    std::string why;
    if (!api.load(argv[1], &why)) { fprintf(stderr, "cannot load %s: %s\n", argv[1], why.c_str()); return 3; }
    std::vector<unsigned char> cube = read_file(argv[2]), pbytes = read_file(argv[3]);
    if (cube.size() != 6 * 128 * 128 || pbytes.empty() || pbytes.size() % sizeof(VpeParticle)) { fprintf(stderr, "bad input files\n"); return 4; }
    const int grid = atoi(argv[4]), voxels = atoi(argv[5]);
    const float scale = (float)atof(argv[6]);
    const int width = atoi(argv[7]), height = atoi(argv[8]);
    const float camZ = (float)atof(argv[9]);
    const int frames = atoi(argv[10]);
    VolumetricParticleRenderer r(api, 0);
    r.numMetavoxelsX = r.numMetavoxelsY = r.numMetavoxelsZ = grid;
    r.numVoxelsInMetavoxel = voxels;
    r.mvScale = Vector3{scale, scale, scale};
    r.particleSys = VpeTransform{{0, 0, 0}, {0, 0, 0, 1}};
    int rc = r.Start(cube.data(), 128);
    if (rc) { fprintf(stderr, "Start: %d %s\n", rc, r.lastError().c_str()); return 5; }
    VpeCamera cam{{{0, 0, camZ}, {0, 0, 0, 1}}, 60.0f, width, height};
    std::vector<float> rgba((size_t)width * height * 4);
    const VpeParticle* parts = reinterpret_cast<const VpeParticle*>(pbytes.data());
    const int n = (int)(pbytes.size() / sizeof(VpeParticle));
    int fills = 0;
    if (argc == 13) {
        std::vector<unsigned char> tbytes = read_file(argv[12]);
        if (tbytes.size() % (9 * sizeof(float))) { fprintf(stderr, "bad triangle file\n"); return 4; }
        rc = r.RenderLightDepthMap(reinterpret_cast<const float*>(tbytes.data()), (int)(tbytes.size() / (9 * sizeof(float))));
        if (rc) { fprintf(stderr, "RenderLightDepthMap: %d %s\n", rc, r.lastError().c_str()); return 8; }
        VpeMarchOptions opt{};
        opt.targetFormat = 1;
        if ((rc = r.SetMarchOptions(opt))) { fprintf(stderr, "SetMarchOptions: %d %s\n", rc, r.lastError().c_str()); return 8; }
    }
    for (int f = 0; f < frames; f++) {
        int before = r.numParticlesEmitted;
        r.numParticlesEmitted = -1;
        rc = r.OnPostRender(parts, n, cam, rgba.data());
        if (rc) { fprintf(stderr, "OnPostRender: %d %s\n", rc, r.lastError().c_str()); return 6; }
        if (r.numParticlesEmitted >= 0) fills++; else r.numParticlesEmitted = before;
    }
    VpeStats st;
    r.GetStats(&st);
    if (argc == 13) {  // mainSceneRT: opaque blue; the particles are blended onto it, 8-bit like the reference's targets
        std::vector<float> scene((size_t)width * height * 4);
        for (size_t i = 0; i < scene.size(); i += 4) { scene[i] = 0.0f; scene[i + 1] = 0.0f; scene[i + 2] = 0.5f; scene[i + 3] = 1.0f; }
        if ((rc = r.CompositeParticles(rgba.data(), scene.data(), width * height, 1))) { fprintf(stderr, "CompositeParticles: %d\n", rc); return 8; }
        rgba = scene;
    }
    FILE* o = fopen(argv[11], "wb");
    if (!o || fwrite(rgba.data(), sizeof(float), rgba.size(), o) != rgba.size()) { fprintf(stderr, "cannot write %s\n", argv[11]); return 7; }
    fclose(o);
    printf("backend=%s frames=%d fills=%d particles=%d covered=%d pairs=%lld raySamples=%lld zBoundary=%d\n", api.vpe_backend(), frames, fills,
           st.numParticles, st.numMetavoxelsCovered, (long long)st.numParticlePairs, (long long)st.raySamples, st.zBoundary);
    return 0;
}
