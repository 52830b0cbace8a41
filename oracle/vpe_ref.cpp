// vpe_ref.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY: nothing in the product path may link,
// load or call this file.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs use it, and only as the checker / the reported CPU baseline.
//
// PARITY UNPINNED: the reference (rajabala/Volumetric-Particles-For-Unity) ships no tests,
// golden vectors or fixtures for this path, its shaders are HLSL for Unity/D3D11
// and its host is C# against UnityEngine — none of it can be built or run here.  This file is a
// scalar C++ restatement of the reference's algorithm, written from the sources cited at every
// function, in the reference's own structure (per-metavoxel dispatch, per-column fill fragment,
// per-(pixel, metavoxel) march fragment, ROP blending between metavoxels).  It is validated by
// closed-form cases and an independent numpy twin (tests/), not by reference-held vectors.
//
// Citations are relative to /root/reference:
//   VPR.cs       = Assets/Main Scene/VolumetricParticleRenderer.cs
//   Fill.shader  = Assets/Shaders/Metavoxel/FillVolume.shader
//   March.shader = Assets/Shaders/Metavoxel/RayMarchVoxel.shader
//   MathUtil.cs  = Assets/Main Scene/MathUtil.cs
//
// Arithmetic: IEEE fp32, round-to-nearest-even, NO fused multiply-add (build with
// -ffp-contract=off), expressions evaluated left to right exactly as written here.  Third-party
// arithmetic (UnityEngine Matrix4x4/Quaternion, HLSL intrinsics, D3D11 samplers) is restated as
// documented in DESIGN.md "Normative arithmetic".
#include "../include/vpe.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ------------------------------------------------------------------------------------------
// UnityEngine math, restated
// ------------------------------------------------------------------------------------------
struct V3 {
    float x, y, z;
};
struct M4 {
    float m[4][4];  // m[row][col], column-vector convention (Unity Matrix4x4.mRC)
};

inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
inline V3 operator/(V3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }
inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }

// Unity Quaternion -> rotation matrix (the engine's QuaternionToMatrix).
M4 quat_to_m4(const float q[4]) {
    float x = q[0] * 2.0f, y = q[1] * 2.0f, z = q[2] * 2.0f;
    float xx = q[0] * x, yy = q[1] * y, zz = q[2] * z;
    float xy = q[0] * y, xz = q[0] * z, yz = q[1] * z;
    float wx = q[3] * x, wy = q[3] * y, wz = q[3] * z;
    M4 r;
    r.m[0][0] = 1.0f - (yy + zz); r.m[0][1] = xy - wz;          r.m[0][2] = xz + wy;          r.m[0][3] = 0.0f;
    r.m[1][0] = xy + wz;          r.m[1][1] = 1.0f - (xx + zz); r.m[1][2] = yz - wx;          r.m[1][3] = 0.0f;
    r.m[2][0] = xz - wy;          r.m[2][1] = yz + wx;          r.m[2][2] = 1.0f - (xx + yy); r.m[2][3] = 0.0f;
    r.m[3][0] = 0.0f;             r.m[3][1] = 0.0f;             r.m[3][2] = 0.0f;             r.m[3][3] = 1.0f;
    return r;
}

// Matrix4x4.TRS(t, q, s) = T * R * S.
M4 trs(V3 t, const float q[4], V3 s) {
    M4 r = quat_to_m4(q);
    for (int i = 0; i < 3; i++) {
        r.m[i][0] = r.m[i][0] * s.x;
        r.m[i][1] = r.m[i][1] * s.y;
        r.m[i][2] = r.m[i][2] * s.z;
    }
    r.m[0][3] = t.x; r.m[1][3] = t.y; r.m[2][3] = t.z;
    return r;
}

// Matrix4x4.inverse — general 4x4 inverse by 2x2 sub-determinants (Laplace expansion).
M4 inverse(const M4& a_) {
    const float (*a)[4] = a_.m;
    float s0 = a[0][0] * a[1][1] - a[1][0] * a[0][1];
    float s1 = a[0][0] * a[1][2] - a[1][0] * a[0][2];
    float s2 = a[0][0] * a[1][3] - a[1][0] * a[0][3];
    float s3 = a[0][1] * a[1][2] - a[1][1] * a[0][2];
    float s4 = a[0][1] * a[1][3] - a[1][1] * a[0][3];
    float s5 = a[0][2] * a[1][3] - a[1][2] * a[0][3];
    float c5 = a[2][2] * a[3][3] - a[3][2] * a[2][3];
    float c4 = a[2][1] * a[3][3] - a[3][1] * a[2][3];
    float c3 = a[2][1] * a[3][2] - a[3][1] * a[2][2];
    float c2 = a[2][0] * a[3][3] - a[3][0] * a[2][3];
    float c1 = a[2][0] * a[3][2] - a[3][0] * a[2][2];
    float c0 = a[2][0] * a[3][1] - a[3][0] * a[2][1];
    float det = s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0;
    float inv = 1.0f / det;
    M4 b;
    b.m[0][0] = (a[1][1] * c5 - a[1][2] * c4 + a[1][3] * c3) * inv;
    b.m[0][1] = (-a[0][1] * c5 + a[0][2] * c4 - a[0][3] * c3) * inv;
    b.m[0][2] = (a[3][1] * s5 - a[3][2] * s4 + a[3][3] * s3) * inv;
    b.m[0][3] = (-a[2][1] * s5 + a[2][2] * s4 - a[2][3] * s3) * inv;
    b.m[1][0] = (-a[1][0] * c5 + a[1][2] * c2 - a[1][3] * c1) * inv;
    b.m[1][1] = (a[0][0] * c5 - a[0][2] * c2 + a[0][3] * c1) * inv;
    b.m[1][2] = (-a[3][0] * s5 + a[3][2] * s2 - a[3][3] * s1) * inv;
    b.m[1][3] = (a[2][0] * s5 - a[2][2] * s2 + a[2][3] * s1) * inv;
    b.m[2][0] = (a[1][0] * c4 - a[1][1] * c2 + a[1][3] * c0) * inv;
    b.m[2][1] = (-a[0][0] * c4 + a[0][1] * c2 - a[0][3] * c0) * inv;
    b.m[2][2] = (a[3][0] * s4 - a[3][1] * s2 + a[3][3] * s0) * inv;
    b.m[2][3] = (-a[2][0] * s4 + a[2][1] * s2 - a[2][3] * s0) * inv;
    b.m[3][0] = (-a[1][0] * c3 + a[1][1] * c1 - a[1][2] * c0) * inv;
    b.m[3][1] = (a[0][0] * c3 - a[0][1] * c1 + a[0][2] * c0) * inv;
    b.m[3][2] = (-a[3][0] * s3 + a[3][1] * s1 - a[3][2] * s0) * inv;
    b.m[3][3] = (a[2][0] * s3 - a[2][1] * s1 + a[2][2] * s0) * inv;
    return b;
}

// Matrix4x4 operator*.
M4 mul(const M4& a, const M4& b) {
    M4 r;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            r.m[i][j] = ((a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j]) + a.m[i][2] * b.m[2][j]) +
                        a.m[i][3] * b.m[3][j];
    return r;
}

// Matrix4x4.MultiplyPoint3x4.
V3 mp3x4(const M4& a, V3 p) {
    return v3(((a.m[0][0] * p.x + a.m[0][1] * p.y) + a.m[0][2] * p.z) + a.m[0][3],
              ((a.m[1][0] * p.x + a.m[1][1] * p.y) + a.m[1][2] * p.z) + a.m[1][3],
              ((a.m[2][0] * p.x + a.m[2][1] * p.y) + a.m[2][2] * p.z) + a.m[2][3]);
}

// HLSL mul(float4x4, float4).xyz
V3 mul4(const M4& a, V3 p, float w) {
    return v3(((a.m[0][0] * p.x + a.m[0][1] * p.y) + a.m[0][2] * p.z) + a.m[0][3] * w,
              ((a.m[1][0] * p.x + a.m[1][1] * p.y) + a.m[1][2] * p.z) + a.m[1][3] * w,
              ((a.m[2][0] * p.x + a.m[2][1] * p.y) + a.m[2][2] * p.z) + a.m[2][3] * w);
}

// Transform.forward = rotation * Vector3.forward (third column of the rotation matrix).
V3 forward_of(const float q[4]) {
    M4 r = quat_to_m4(q);
    return v3(r.m[0][2], r.m[1][2], r.m[2][2]);
}

// Vector3.normalized
V3 normalized_unity(V3 v) {
    float mag = sqrtf(dot(v, v));
    if (mag > 1e-5f) return v / mag;
    return v3(0, 0, 0);
}

// HLSL normalize
V3 normalize_hlsl(V3 v) { return v / sqrtf(dot(v, v)); }

// Quaternion.AngleAxis(deg, axis); sin/cos evaluated in double and rounded (Mathf.Sin/Cos).
void angle_axis(float deg, V3 axis, float q[4]) {
    float rad = deg * 0.0174532924f;
    float h = rad * 0.5f;
    float mag = sqrtf(dot(axis, axis));
    float s = (float)sin((double)h);
    float c = (float)cos((double)h);
    q[0] = (axis.x / mag) * s;
    q[1] = (axis.y / mag) * s;
    q[2] = (axis.z / mag) * s;
    q[3] = c;
}

// Mathf.RoundToInt = (int)Math.Round(double) — banker's rounding.
int round_to_int(float f) { return (int)nearbyint((double)f); }  // default FE_TONEAREST (ties-to-even)

// fp32 -> fp16, round-to-nearest-even, IEEE (subnormals, overflow -> inf).
uint16_t f32_to_f16(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t em = x & 0x7fffffffu;
    if (em >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | ((em > 0x7f800000u) ? 0x200u : 0));
    if (em >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);  // rounds to >= 65520 -> inf
    if (em < 0x33000001u) return (uint16_t)sign;               // <= 2^-25 -> 0 (tie to even)
    int e = (int)(em >> 23) - 127;
    uint32_t man = (em & 0x7fffffu) | 0x800000u;
    int shift;
    uint32_t base;
    if (e < -14) { shift = 13 + (-14 - e); base = 0; }          // subnormal half
    else { shift = 13; base = (uint32_t)(e + 15) << 10; man &= 0x7fffffu; }
    uint32_t q = man >> shift;
    uint32_t rem = man & ((1u << shift) - 1);
    uint32_t halfway = 1u << (shift - 1);
    uint32_t h = base + q;
    if (rem > halfway || (rem == halfway && (h & 1u))) h++;
    return (uint16_t)(sign | h);
}

float f16_to_f32(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t e = (h >> 10) & 0x1fu;
    uint32_t m = h & 0x3ffu;
    uint32_t x;
    if (e == 0) {
        if (m == 0) x = sign;
        else {
            int sh = 0;
            while (!(m & 0x400u)) { m <<= 1; sh++; }
            m &= 0x3ffu;
            x = sign | ((uint32_t)(127 - 15 - sh + 1) << 23) | (m << 13);
        }
    } else if (e == 31) x = sign | 0x7f800000u | (m << 13);
    else x = sign | ((e - 15 + 127) << 23) | (m << 13);
    float f;
    memcpy(&f, &x, 4);
    return f;
}

inline float saturate(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }

// float -> int as D3D ftoi: truncate toward zero, saturate, NaN -> 0.
inline int ftoi(float f) {
    if (f != f) return 0;
    if (f >= 2147483648.0f) return 2147483647;
    if (f <= -2147483648.0f) return -2147483647 - 1;
    return (int)f;
}

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

// ------------------------------------------------------------------------------------------
// Context  ≙ the private state of VolumetricParticleRenderer (VPR.cs:106-128)
// ------------------------------------------------------------------------------------------
struct VpeContext {
    VpeConfig cfg;
    std::string err;
    // light / grid (VPR.cs:126,108-109)
    VpeTransform light;
    V3 center;
    bool lightSet = false;
    M4 L2W, W2L;       // dirLight.transform.localToWorldMatrix / worldToLocalMatrix
    V3 lightFwdRaw;    // dirLight.transform.forward
    V3 lightFwd;       // .normalized (VPR.cs:535)
    M4 W2LC;           // lightCamera.transform.worldToLocalMatrix (VPR.cs:365-366,534)
    float sb;          // mvScaleWithBorder (VPR.cs:139)
    std::vector<V3> mvPos;                    // mvGrid[z,y,x].mPos
    std::vector<std::vector<int32_t>> lists;  // mvGrid[z,y,x].mParticlesCovered (indices)
    std::vector<std::vector<uint16_t>> bricks;  // mvFillTextures[z,y,x], half4 [k][y][x]
    std::vector<uint8_t> filled;
    std::vector<float> sheet;     // lightPropogationUAV
    std::vector<float> depthMap;  // lightDepthMap (empty = all 1.0)
    std::vector<uint8_t> cube;
    int cubeEdge = 0;
    // per-fill particle state (VPR.cs:412-418,582-586)
    std::vector<VpeParticle> parts;
    std::vector<V3> partWs;
    std::vector<M4> partW2P;
    std::vector<float> partOpacity;
    bool prepared = false;
    VpeStats stats;
    // march options (vpe_set_march_options): 8-bit target, debug views, scene depth test
    int targetFormat = 0, debugMode = 0;
    std::vector<float> sceneDepth;  // eye-space depth per pixel (empty = no scene geometry)
    int sceneW = 0, sceneH = 0;

    int NX() const { return cfg.numMetavoxelsX; }
    int NY() const { return cfg.numMetavoxelsY; }
    int NZ() const { return cfg.numMetavoxelsZ; }
    int N() const { return cfg.numVoxelsInMetavoxel; }
    int z0() const { return cfg.slabZBegin; }
    int z1() const { return cfg.slabZEnd; }
    size_t mvIndex(int x, int y, int z) const { return ((size_t)z * NY() + y) * NX() + x; }
};

namespace {

int fail(VpeContext* c, int code, const char* msg) {
    if (c) c->err = msg;
    return code;
}

int validate_config(const VpeConfig& c, std::string& why) {
    if (c.numMetavoxelsX < 1 || c.numMetavoxelsY < 1 || c.numMetavoxelsZ < 1) { why = "grid dims must be >= 1"; return -1; }
    if (c.numVoxelsInMetavoxel < 2) { why = "numVoxelsInMetavoxel must be >= 2"; return -1; }
    // the fill clamps the border to [0,N-2] (VPR.cs:528), the march does not
    // (VPR.cs:726); both agree only for 0 <= b <= (N-2)/2.
    if (c.numBorderVoxels < 0 || c.numBorderVoxels > (c.numVoxelsInMetavoxel - 2) / 2) { why = "numBorderVoxels out of range"; return -1; }
    if (!(c.mvScale > 0.0f)) { why = "mvScale must be > 0"; return -1; }
    if (c.rayMarchSteps < 1) { why = "rayMarchSteps must be >= 1"; return -1; }
    if (c.slabZBegin < 0 || c.slabZEnd > c.numMetavoxelsZ || c.slabZBegin >= c.slabZEnd) { why = "bad slab range"; return -1; }
    return 0;
}

void normalise_slab(VpeConfig& c) {
    if (c.slabZBegin == 0 && c.slabZEnd == 0) c.slabZEnd = c.numMetavoxelsZ;
}

// VPR.cs:370-394 UpdateMetavoxelPositions; VPR.cs:361-367 UpdatePositionOfCameraAtLight
void update_metavoxel_positions(VpeContext* c) {
    c->L2W = trs(v3(c->light.position[0], c->light.position[1], c->light.position[2]), c->light.rotation, v3(1, 1, 1));
    c->W2L = inverse(c->L2W);
    c->lightFwdRaw = forward_of(c->light.rotation);
    c->lightFwd = normalized_unity(c->lightFwdRaw);
    V3 lcPos = c->center - c->lightFwdRaw * c->cfg.lightCameraDistance;  // VPR.cs:365
    c->W2LC = inverse(trs(lcPos, c->light.rotation, v3(1, 1, 1)));
    V3 lsWorldOrigin = mp3x4(c->W2L, c->center);  // VPR.cs:380
    float s = c->cfg.mvScale;
    c->mvPos.resize((size_t)c->NX() * c->NY() * c->NZ());
    for (int zz = 0; zz < c->NZ(); zz++)
        for (int yy = 0; yy < c->NY(); yy++)
            for (int xx = 0; xx < c->NX(); xx++) {
                // VPR.cs:388 — integer halves
                V3 lsOffset = v3((float)(c->NX() / 2 - xx) * s, (float)(c->NY() / 2 - yy) * s, (float)(c->NZ() / 2 - zz) * s);
                c->mvPos[c->mvIndex(xx, yy, zz)] = mp3x4(c->L2W, lsWorldOrigin - lsOffset);  // VPR.cs:389
            }
}

// MathUtil.cs:11-25
bool does_box_intersect_sphere(V3 c1, V3 c2, V3 s, float r) {
    float r2 = r * r;
    if (s.x < c1.x) r2 -= (s.x - c1.x) * (s.x - c1.x);
    else if (s.x > c2.x) r2 -= (s.x - c2.x) * (s.x - c2.x);
    if (s.y < c1.y) r2 -= (s.y - c1.y) * (s.y - c1.y);
    else if (s.y > c2.y) r2 -= (s.y - c2.y) * (s.y - c2.y);
    if (s.z < c1.z) r2 -= (s.z - c1.z) * (s.z - c1.z);
    else if (s.z > c2.z) r2 -= (s.z - c2.z) * (s.z - c2.z);
    return r2 > 0;
}

// VPR.cs:397-457 BinParticlesToMetavoxels (restricted to the context's slab; the full grid when
// the slab is the whole grid).  Also builds the per-particle fill input of VPR.cs:582-586.
void bin_particles(VpeContext* c, const VpeParticle* parts, int n, const VpeTransform* em) {
    const int NX = c->NX(), NY = c->NY(), NZ = c->NZ();
    const float s = c->cfg.mvScale;
    for (auto& l : c->lists) l.clear();  // VPR.cs:400-409
    c->parts.assign(parts, parts + n);
    c->partWs.resize(n);
    c->partW2P.resize(n);
    c->partOpacity.resize(n);
    M4 E2W = trs(v3(em->position[0], em->position[1], em->position[2]), em->rotation, v3(1, 1, 1));
    V3 efwd = forward_of(em->rotation);
    V3 sbv = v3(c->sb, c->sb, c->sb);
    int64_t pairs = 0;
    for (int pp = 0; pp < n; pp++) {
        const VpeParticle& p = parts[pp];
        V3 ws = mp3x4(E2W, v3(p.position[0], p.position[1], p.position[2]));  // VPR.cs:418
        V3 ls = mp3x4(c->W2L, ws);                                            // VPR.cs:419
        V3 lsC = mp3x4(c->W2L, c->center);                                    // VPR.cs:420
        V3 off = (ls - lsC) / s;                                              // VPR.cs:422
        V3 idx = off + v3((float)NX * 0.5f, (float)NY * 0.5f, (float)NZ * 0.5f);  // VPR.cs:423
        // per-particle fill input, VPR.cs:582-586
        c->partWs[pp] = ws;
        float q[4];
        angle_axis(p.rotationDeg, efwd, q);
        c->partW2P[pp] = inverse(trs(ws, q, v3(p.size, p.size, p.size)));
        c->partOpacity[pp] = p.lifetime / p.startLifetime;

        float radius = p.size / 2.0f;
        int lox, hix, loy, hiy, loz, hiz;
        if (c->cfg.binMode == VPE_BIN_EXACT) {
            // every cell whose border-enlarged box can reach the sphere (DESIGN.md "bin modes")
            float reach = radius / s + 0.5f * (c->sb / s) + 0.5f;
            lox = std::max(0, (int)floorf(idx.x - reach)); hix = std::min(NX - 1, (int)ceilf(idx.x + reach));
            loy = std::max(0, (int)floorf(idx.y - reach)); hiy = std::min(NY - 1, (int)ceilf(idx.y + reach));
            loz = std::max(0, (int)floorf(idx.z - reach)); hiz = std::min(NZ - 1, (int)ceilf(idx.z + reach));
        } else {
            int ext = round_to_int(radius / s);  // VPR.cs:425
            float e = (float)ext;
            V3 mn = v3(idx.x - e, idx.y - e, idx.z - e), mx = v3(idx.x + e, idx.y + e, idx.z + e);  // VPR.cs:426-427
            mn = v3(fmaxf(0.0f, mn.x), fmaxf(0.0f, mn.y), fmaxf(0.0f, mn.z));                        // VPR.cs:431
            mx = v3(fminf((float)(NX - 1), mx.x), fminf((float)(NY - 1), mx.y), fminf((float)(NZ - 1), mx.z));  // VPR.cs:432
            lox = (int)mn.x; hix = (int)mx.x; loy = (int)mn.y; hiy = (int)mx.y; loz = (int)mn.z; hiz = (int)mx.z;  // VPR.cs:434-438
        }
        loz = std::max(loz, c->z0());
        hiz = std::min(hiz, c->z1() - 1);
        for (int zz = loz; zz <= hiz; zz++)
            for (int yy = loy; yy <= hiy; yy++)
                for (int xx = lox; xx <= hix; xx++) {
                    M4 w2mv = inverse(trs(c->mvPos[c->mvIndex(xx, yy, zz)], c->light.rotation, sbv));  // VPR.cs:440-442
                    V3 mvP = mp3x4(w2mv, ws);                                                          // VPR.cs:444
                    float mvR = radius / c->sb;                                                        // VPR.cs:445
                    if (does_box_intersect_sphere(v3(-0.5f, -0.5f, -0.5f), v3(0.5f, 0.5f, 0.5f), mvP, mvR)) {
                        c->lists[c->mvIndex(xx, yy, zz)].push_back(pp);  // VPR.cs:453
                        pairs++;
                    }
                }
    }
    int covered = 0;
    for (int zz = c->z0(); zz < c->z1(); zz++)
        for (int yy = 0; yy < NY; yy++)
            for (int xx = 0; xx < NX; xx++)
                if (!c->lists[c->mvIndex(xx, yy, zz)].empty()) covered++;
    c->stats.numParticles = n;
    c->stats.numParticlePairs = pairs;
    c->stats.numMetavoxelsCovered = covered;
    c->stats.voxelsFilled = (int64_t)covered * c->N() * c->N() * c->N();
}

// texCUBE(_DisplacementTexture, dir).x — D3D major-axis face selection, bilinear inside the face,
// clamp addressing (Fill.shader:116; sampler: DisplacementTexture.cubemap:21-25).
float sample_cube(const VpeContext* c, V3 d) {
    const int E = c->cubeEdge;
    float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    int face;
    float ma, sc, tc;
    if (ax >= ay && ax >= az) { face = d.x >= 0.0f ? 0 : 1; ma = ax; sc = d.x >= 0.0f ? -d.z : d.z; tc = -d.y; }
    else if (ay >= az)        { face = d.y >= 0.0f ? 2 : 3; ma = ay; sc = d.x; tc = d.y >= 0.0f ? d.z : -d.z; }
    else                      { face = d.z >= 0.0f ? 4 : 5; ma = az; sc = d.z >= 0.0f ? d.x : -d.x; tc = -d.y; }
    float u, v;
    if (ma == 0.0f) { face = 0; u = 0.5f; v = 0.5f; }
    else { u = (sc / ma + 1.0f) * 0.5f; v = (tc / ma + 1.0f) * 0.5f; }
    float fx = u * (float)E - 0.5f, fy = v * (float)E - 0.5f;
    float flx = floorf(fx), fly = floorf(fy);
    float wx = fx - flx, wy = fy - fly;
    int x0 = (int)flx, y0 = (int)fly, x1 = x0 + 1, y1 = y0 + 1;
    x0 = std::min(std::max(x0, 0), E - 1); x1 = std::min(std::max(x1, 0), E - 1);
    y0 = std::min(std::max(y0, 0), E - 1); y1 = std::min(std::max(y1, 0), E - 1);
    const uint8_t* f = c->cube.data() + (size_t)face * E * E;
    float t00 = (float)f[y0 * E + x0] / 255.0f, t10 = (float)f[y0 * E + x1] / 255.0f;
    float t01 = (float)f[y1 * E + x0] / 255.0f, t11 = (float)f[y1 * E + x1] / 255.0f;
    float top = t00 + wx * (t10 - t00);
    float bot = t01 + wx * (t11 - t01);
    return top + wy * (bot - top);
}

// tex2D(_LightDepthMap, uv) — bilinear, clamp (Fill.shader:216).
float sample_depth(const VpeContext* c, float u, float v) {
    if (c->depthMap.empty()) return 1.0f;
    const int W = c->NX() * c->N(), H = c->NY() * c->N();
    float fx = u * (float)W - 0.5f, fy = v * (float)H - 0.5f;
    float flx = floorf(fx), fly = floorf(fy);
    float wx = fx - flx, wy = fy - fly;
    int x0 = (int)flx, y0 = (int)fly, x1 = x0 + 1, y1 = y0 + 1;
    x0 = std::min(std::max(x0, 0), W - 1); x1 = std::min(std::max(x1, 0), W - 1);
    y0 = std::min(std::max(y0, 0), H - 1); y1 = std::min(std::max(y1, 0), H - 1);
    const float* d = c->depthMap.data();
    float t00 = d[(size_t)y0 * W + x0], t10 = d[(size_t)y0 * W + x1];
    float t01 = d[(size_t)y1 * W + x0], t11 = d[(size_t)y1 * W + x1];
    float top = t00 + wx * (t10 - t00);
    float bot = t01 + wx * (t11 - t01);
    return top + wy * (bot - top);
}

struct Voxel {
    float density, ao;
};

// Fill.shader:110-135 compute_voxel_color
bool compute_voxel_color(const VpeContext* c, V3 ps, float opacity, Voxel* v) {
    V3 d = v3(2.0f * ps.x, 2.0f * ps.y, 2.0f * ps.z);
    float raw = sample_cube(c, d);                                                     // :116
    float ds = c->cfg.displacementScale;
    float net = ds * raw + (1.0f - ds);                                                // :119
    float d2 = dot(d, d);                                                              // :121
    float t = saturate((d2 - net) / (0.7f * net - net));                               // :126 smoothstep
    float base = (t * t) * (3.0f - 2.0f * t);
    float density = base * c->cfg.opacityFactor;                                       // :127
    if (c->cfg.fadeOutParticles == 1) density *= opacity;                              // :130-131
    v->density = density;
    v->ao = net;
    return true;
}

// Fill.shader:152-274 frag, for every (x,y) voxel column of metavoxel (xx,yy,zz);
// dispatch state from VPR.cs:559-609 FillMetavoxel and VPR.cs:523-554 SetFillPassConstants.
void fill_metavoxel(VpeContext* c, int xx, int yy, int zz) {
    const int N = c->N();
    const float Nf = (float)N;
    const size_t mi = c->mvIndex(xx, yy, zz);
    const std::vector<int32_t>& list = c->lists[mi];
    std::vector<uint16_t>& brick = c->bricks[mi];
    brick.resize((size_t)N * N * N * 4);
    const M4 mvToWorld = trs(c->mvPos[mi], c->light.rotation, v3(c->sb, c->sb, c->sb));  // VPR.cs:596-598
    const int border = std::min(std::max(c->cfg.numBorderVoxels, 0), N - 2);             // VPR.cs:528
    const float oneVoxelSize = c->sb / Nf;                                               // Fill.shader:160
    const V3 lightStep = v3(c->lightFwd.x * oneVoxelSize, c->lightFwd.y * oneVoxelSize, c->lightFwd.z * oneVoxelSize);
    const int sheetW = c->NX() * N;
    const int numParticles = (int)list.size();
    std::vector<Voxel> column(N);
    for (int py = 0; py < N; py++)
        for (int px = 0; px < N; px++) {
            const float posx = (float)px + 0.5f, posy = (float)py + 0.5f;  // v2f_img pixel centre
            // get_voxel_world_pos(i.pos.xy, 0), Fill.shader:96-107
            V3 norm = v3((posx - Nf / 2.0f) / Nf, (posy - Nf / 2.0f) / Nf, (0.0f - Nf / 2.0f) / Nf);
            V3 voxel0 = mul4(mvToWorld, norm, 1.0f);
            V3 vw = voxel0;
            // Fill.shader:164-184 first particle
            for (int slice = 0; slice < N; slice++) {
                const M4& w2p = c->partW2P[list[0]];
                V3 ps = mul4(w2p, vw, 1.0f);
                float dist2 = dot(ps, ps);
                if (dist2 <= 0.25f) compute_voxel_color(c, ps, c->partOpacity[list[0]], &column[slice]);
                else { column[slice].density = 0.0f; column[slice].ao = 0.0f; }
                vw = vw + lightStep;
            }
            vw = voxel0;
            // Fill.shader:188-208 remaining particles
            for (int slice = 0; slice < N; slice++) {
                for (int pp = 1; pp < numParticles; pp++) {
                    const M4& w2p = c->partW2P[list[pp]];
                    V3 ps = mul4(w2p, vw, 1.0f);
                    float dist2 = dot(ps, ps);
                    if (dist2 <= 0.25f) {
                        Voxel v;
                        compute_voxel_color(c, ps, c->partOpacity[list[pp]], &v);
                        column[slice].density += v.density;
                        column[slice].ao = fmaxf(column[slice].ao, v.ao);
                    }
                }
                vw = vw + lightStep;
            }
            // Fill.shader:211-221 occlusion
            V3 lsVoxel0 = mul4(c->W2LC, voxel0, 1.0f);
            float u = (posx + (float)xx * Nf) / ((float)c->NX() * Nf);
            float v = (posy + (float)yy * Nf) / ((float)c->NY() * Nf);
            float d = sample_depth(c, u, v);
            float a = 1.0f / (c->cfg.lightFar - c->cfg.lightNear), b = -c->cfg.lightNear * a;
            float lsSceneDepth = (d - b) * (1.0f / a);
            float fsi = (lsSceneDepth - lsVoxel0.z) / oneVoxelSize;
            int shadowIndex = ftoi(fsi);
            // Fill.shader:224-250
            const size_t sheetIdx = (size_t)(py + yy * N) * sheetW + (px + xx * N);
            float transmitted = (zz == 0) ? 1.0f : c->sheet[sheetIdx];
            float propagated = transmitted;
            const float diffuse = 0.4f;
            const int borderVoxelIndex = N - border;
            for (int slice = 0; slice < N; slice++) {
                bool inShadow = slice >= shadowIndex;
                if (inShadow) transmitted = 0.0f;
                else if (slice < borderVoxelIndex) propagated = transmitted;  // Fill.shader:239-240 vs 260-261
                float r = diffuse * transmitted + c->cfg.ambientColor[0] * column[slice].ao;
                float g = diffuse * transmitted + c->cfg.ambientColor[1] * column[slice].ao;
                float bl = diffuse * transmitted + c->cfg.ambientColor[2] * column[slice].ao;
                transmitted *= 1.0f / (1.0f + column[slice].density);
                uint16_t* o = &brick[(((size_t)slice * N + py) * N + px) * 4];
                o[0] = f32_to_f16(r); o[1] = f32_to_f16(g); o[2] = f32_to_f16(bl); o[3] = f32_to_f16(column[slice].density);
            }
            c->sheet[sheetIdx] = propagated;  // Fill.shader:250 (propagated is frozen from slice borderVoxelIndex on)
        }
    c->filled[mi] = 1;
}

// VPR.cs:495-520 FillMetavoxels restricted to metavoxel columns [x0,x1) x [y0,y1) and the slab.
void fill_region(VpeContext* c, int x0, int x1, int y0, int y1) {
    const int ncol = (x1 - x0) * (y1 - y0);
#pragma omp parallel for schedule(dynamic, 1)
    for (int col = 0; col < ncol; col++) {
        int xx = x0 + col % (x1 - x0), yy = y0 + col / (x1 - x0);
        for (int zz = c->z0(); zz < c->z1(); zz++)  // nearest the light first, VPR.cs:505
            if (!c->lists[c->mvIndex(xx, yy, zz)].empty()) fill_metavoxel(c, xx, yy, zz);  // VPR.cs:511
    }
}

// ------------------------------------------------------------------------------------------
// Ray march
// ------------------------------------------------------------------------------------------
struct MarchPass {
    // VPR.cs:716-763 SetRaymarchPassConstants
    M4 W2C, C2W;
    float tanHalfFov;
    float W, H;
    float maxGridDim;
    int zBoundary;
    struct Draw {
        int x, y, z; bool over; M4 C2M;
        // Rasteriser stand-in: bounding rectangle (in pixels) of the projected cube, dilated; the
        // per-pixel decision is still the shader's slab test (March.shader:229). Only (pixel,
        // metavoxel) pairs whose ray cannot hit the cube are skipped.
        float px0, px1, py0, py1;
    };
    std::vector<Draw> draws;  // in submission order, VPR.cs:667-711
};

struct SortData {  // VPR.cs:43-62
    int x, y;
    float distance;
};

void build_march_pass(VpeContext* c, const VpeCamera* cam, MarchPass* mp) {
    const int NX = c->NX(), NY = c->NY(), NZ = c->NZ();
    V3 camPos = v3(cam->transform.position[0], cam->transform.position[1], cam->transform.position[2]);
    M4 camL2W = trs(camPos, cam->transform.rotation, v3(1, 1, 1));
    // Camera.cameraToWorldMatrix / worldToCameraMatrix: OpenGL convention, -Z forward
    mp->C2W = camL2W;
    for (int i = 0; i < 3; i++) mp->C2W.m[i][2] = -mp->C2W.m[i][2];
    M4 camW2L = inverse(camL2W);
    mp->W2C = camW2L;
    for (int j = 0; j < 4; j++) mp->W2C.m[2][j] = -mp->W2C.m[2][j];
    float fovRad = 0.0174532924f * cam->fovYDegrees;           // VPR.cs:734
    mp->tanHalfFov = (float)tan((double)(fovRad / 2.0f));      // March.shader:193
    mp->W = (float)cam->width; mp->H = (float)cam->height;     // VPR.cs:737
    mp->maxGridDim = (float)std::max(NX, std::max(NY, NZ));    // March.shader:207-208
    // VPR.cs:613-632 SortMetavoxelSlicesFarToNearFromEye (ties: by (y,x))
    std::vector<SortData> asc;
    for (int yy = 0; yy < NY; yy++)
        for (int xx = 0; xx < NX; xx++) {
            V3 d = c->mvPos[c->mvIndex(xx, yy, 0)] - camPos;
            asc.push_back(SortData{xx, yy, dot(d, d)});
        }
    std::stable_sort(asc.begin(), asc.end(), [](const SortData& a, const SortData& b) { return a.distance < b.distance; });
    std::vector<SortData> farToNear(asc.rbegin(), asc.rend());
    // VPR.cs:642-648
    V3 lsCam = mp3x4(c->W2L, camPos);
    float lsFirst = mp3x4(c->W2L, c->mvPos[c->mvIndex(0, 0, 0)]).z;
    float blendOverIndex = (lsCam.z - lsFirst) / c->cfg.mvScale;
    int zB = std::min(std::max(round_to_int(blendOverIndex), -1), NZ - 1);
    mp->zBoundary = zB;
    mp->draws.clear();
    V3 sv = v3(c->cfg.mvScale, c->cfg.mvScale, c->cfg.mvScale);
    auto submit = [&](int xx, int yy, int zz, bool over) {
        if (zz < c->z0() || zz >= c->z1()) return;
        if (c->lists[c->mvIndex(xx, yy, zz)].empty()) return;  // VPR.cs:674,704
        MarchPass::Draw d;
        d.x = xx; d.y = yy; d.z = zz; d.over = over;
        M4 mvToWorld = trs(c->mvPos[c->mvIndex(xx, yy, zz)], c->light.rotation, sv);  // VPR.cs:774-776
        d.C2M = mul(inverse(mvToWorld), mp->C2W);                                       // VPR.cs:778
        // ≙ the vertex shader + rasteriser (March.shader:148-158, Cull Front): project the 8 corners
        d.px0 = d.py0 = -1e30f; d.px1 = d.py1 = 1e30f;
        double lo[2] = {1e30, 1e30}, hi[2] = {-1e30, -1e30};
        bool behind = false;
        M4 m2c = mul(mp->W2C, mvToWorld);
        for (int k = 0; k < 8 && !behind; k++) {
            V3 corner = v3((k & 1) ? 0.5f : -0.5f, (k & 2) ? 0.5f : -0.5f, (k & 4) ? 0.5f : -0.5f);
            V3 cs = mul4(m2c, corner, 1.0f);
            if (cs.z > -1e-3f * c->cfg.mvScale) { behind = true; break; }
            double sx = (double)cs.x / -(double)cs.z / ((double)mp->tanHalfFov * (double)(mp->W / mp->H));
            double sy = (double)cs.y / -(double)cs.z / (double)mp->tanHalfFov;
            double fx = 0.5 * (double)mp->W * (1.0 + sx), fy = 0.5 * (double)mp->H * (1.0 + sy);
            lo[0] = std::min(lo[0], fx); hi[0] = std::max(hi[0], fx);
            lo[1] = std::min(lo[1], fy); hi[1] = std::max(hi[1], fy);
        }
        if (!behind) {
            d.px0 = (float)(lo[0] - 1.5); d.px1 = (float)(hi[0] + 1.5);
            d.py0 = (float)(lo[1] - 1.5); d.py1 = (float)(hi[1] + 1.5);
        }
        mp->draws.push_back(d);
    };
    for (int zz = 0; zz <= zB; zz++)                       // VPR.cs:667-680
        for (const SortData& vv : farToNear) submit(vv.x, vv.y, zz, true);
    for (int zz = zB + 1; zz < NZ; zz++)                   // VPR.cs:697-711 (list reversed: near to far)
        for (const SortData& vv : asc) submit(vv.x, vv.y, zz, false);
}

// March.shader:95-118 IntersectBox
bool intersect_box(V3 o, V3 d, float* tnear, float* tfar) {
    V3 invR = v3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    V3 tbot = v3(invR.x * (-0.5f - o.x), invR.y * (-0.5f - o.y), invR.z * (-0.5f - o.z));
    V3 ttop = v3(invR.x * (0.5f - o.x), invR.y * (0.5f - o.y), invR.z * (0.5f - o.z));
    V3 tmin = v3(fminf(ttop.x, tbot.x), fminf(ttop.y, tbot.y), fminf(ttop.z, tbot.z));
    V3 tmax = v3(fmaxf(ttop.x, tbot.x), fmaxf(ttop.y, tbot.y), fmaxf(ttop.z, tbot.z));
    float t0x = fmaxf(tmin.x, tmin.y), t0y = fmaxf(tmin.x, tmin.z);
    *tnear = fmaxf(t0x, t0y);
    t0x = fminf(tmax.x, tmax.y); t0y = fminf(tmax.x, tmax.z);
    *tfar = fminf(t0x, t0y);
    return !(*tnear > *tfar);
}

inline int wrap(int i, int n) {
    int r = i % n;
    return r < 0 ? r + n : r;
}

// tex3D(_VolumeTexture, uvw) — trilinear, repeat addressing (VPR.cs:769-770, March.shader:262).
void sample_brick(const uint16_t* brick, int N, V3 uvw, float out[4]) {
    const float Nf = (float)N;
    float fx = uvw.x * Nf - 0.5f, fy = uvw.y * Nf - 0.5f, fz = uvw.z * Nf - 0.5f;
    float flx = floorf(fx), fly = floorf(fy), flz = floorf(fz);
    float wx = fx - flx, wy = fy - fly, wz = fz - flz;
    int x0 = wrap((int)flx, N), x1 = wrap((int)flx + 1, N);
    int y0 = wrap((int)fly, N), y1 = wrap((int)fly + 1, N);
    int z0 = wrap((int)flz, N), z1 = wrap((int)flz + 1, N);
    auto tex = [&](int x, int y, int z, int ch) { return f16_to_f32(brick[((((size_t)z * N + y) * N + x) << 2) + ch]); };
    for (int ch = 0; ch < 4; ch++) {
        float c000 = tex(x0, y0, z0, ch), c100 = tex(x1, y0, z0, ch), c010 = tex(x0, y1, z0, ch), c110 = tex(x1, y1, z0, ch);
        float c001 = tex(x0, y0, z1, ch), c101 = tex(x1, y0, z1, ch), c011 = tex(x0, y1, z1, ch), c111 = tex(x1, y1, z1, ch);
        float c00 = c000 + wx * (c100 - c000), c10 = c010 + wx * (c110 - c010);
        float c01 = c001 + wx * (c101 - c001), c11 = c011 + wx * (c111 - c011);
        float c0 = c00 + wy * (c10 - c00), c1 = c01 + wy * (c11 - c01);
        out[ch] = c0 + wz * (c1 - c0);
    }
}

struct RaySetup {  // per pixel, March.shader:187-224 (independent of the metavoxel)
    V3 dir, start;
    float stepSize;
};

RaySetup ray_setup(const VpeContext* c, const MarchPass* mp, int px, int py) {
    RaySetup r;
    float posx = (float)px + 0.5f, posy = (float)py + 0.5f;  // SV_POSITION pixel centre
    V3 d;
    d.x = (2.0f * posx / mp->W) - 1.0f;                      // :189
    d.y = (2.0f * posy / mp->H) - 1.0f;
    d.x = d.x * (mp->W / mp->H);                             // :190
    d.z = -(1.0f / mp->tanHalfFov);                          // :193
    r.dir = normalize_hlsl(d);                               // :194
    V3 csVolOrigin = mul4(mp->W2C, c->center, 1.0f);         // :206
    float csVolHalfZ = 1.73205f * 0.5f * mp->maxGridDim * c->cfg.mvScale;  // :210
    float csZVolMin = csVolOrigin.z + csVolHalfZ;            // :211
    float k = csZVolMin / r.dir.z;
    r.start = v3(r.dir.x * k, r.dir.y * k, r.dir.z * k);     // :213
    float csRayLength = 2.0f * csVolHalfZ;                   // :214
    float total = mp->maxGridDim * (float)c->cfg.rayMarchSteps;  // :221
    float oneOver = 1.0f / total;                            // :222
    float mvRayLength = csRayLength * (1.0f / c->cfg.mvScale);   // :223
    r.stepSize = mvRayLength * oneOver;                      // :224
    return r;
}

// March.shader:166-302 frag for one (pixel, metavoxel); returns false when the shader returns
// "seethrough" before the loop (no intersection).  src = (rgb, 1 - transmittance).
// sceneEyeDepth: eye-space depth of the opaque scene at this pixel (3e38 = nothing there). orderIndex /
// numCovered: _OrderIndex / _NumMetavoxelsCovered of the draw-order debug view.
bool march_frag(const VpeContext* c, const RaySetup& rs, const MarchPass::Draw& dr, float src[4], int* samples,
                std::vector<uint8_t>* footprint = nullptr, float sceneEyeDepth = 3.0e38f, int orderIndex = 0, int numCovered = 0) {
    const int N = c->N();
    V3 o = mul4(dr.C2M, rs.start, 1.0f);                     // :217
    V3 d = normalize_hlsl(mul4(dr.C2M, rs.dir, 0.0f));       // :218
    float t1, t2;
    if (!intersect_box(o, d, &t1, &t2)) { src[0] = src[1] = src[2] = src[3] = 0.0f; return false; }  // :229-231
    {
        // ≙ `Cull Front ... ZTest Less` against mainSceneRT.depthBuffer (March.shader:14, VPR.cs:204): the fragment
        // exists only where the cube's back face (the ray's exit point, t2) is nearer than the opaque scene.
        // A metavoxel-space length t is a camera-space length t * mvScale along the ray.
        // Faces behind the camera are clipped, not rasterised (such a fragment could only hold samples behind
        // tCamera, i.e. none: this matters for the debug views alone).
        float exitEyeDepth = -(rs.start.z + (t2 * c->cfg.mvScale) * rs.dir.z);
        if (!(exitEyeDepth > 0.0f && exitEyeDepth < sceneEyeDepth)) { src[0] = src[1] = src[2] = src[3] = 0.0f; return false; }
    }
    if (c->debugMode == 1) {  // DrawOrderColoring, March.shader:123-138,170-173
        int numColorsPerChannel = (int)ceilf((float)numCovered / 3.0f);
        int channelSelect = orderIndex / numColorsPerChannel;
        int channelIndex = orderIndex % numColorsPerChannel;
        float channelIntensity = (float)(numColorsPerChannel - channelIndex) / (float)numColorsPerChannel;
        src[0] = src[1] = src[2] = 0.0f; src[3] = 1.0f;
        src[channelSelect == 0 ? 1 : (channelSelect == 1 ? 2 : 0)] = channelIntensity;
        return true;
    }
    if (c->debugMode == 2) {  // March.shader:174-181
        if (dr.over) { src[0] = 0.5f; src[1] = 0.5f; src[2] = 0.0f; src[3] = 1.0f; }
        else { src[0] = 0.0f; src[1] = 0.5f; src[2] = 0.5f; src[3] = 1.0f; }
        return true;
    }
    const float step = rs.stepSize;
    int tEntry = ftoi(ceilf(t1 / step));                     // :236
    int tExit = ftoi(floorf(t2 / step));                     // :237
    V3 camMv = mul4(dr.C2M, v3(0, 0, 0), 1.0f);              // :238
    V3 co = camMv - o;
    int tCamera = ftoi(sqrtf(dot(co, co)) / step);         // :239
    tEntry = std::max(tEntry, tCamera);                      // :240
    float result[3] = {0, 0, 0};
    float transmittance = 1.0f;
    float borderVoxelOffset = (1.0f / (float)N) * (float)c->cfg.numBorderVoxels;  // :245
    V3 rayStep = v3(d.x * step, d.y * step, d.z * step);     // :248
    float fe = (float)tExit;
    V3 pos = v3(o.x + fe * rayStep.x, o.y + fe * rayStep.y, o.z + fe * rayStep.z);  // :249
    const float scale = 1.0f - 2.0f * borderVoxelOffset;
    const float softRcp = 1.0f / (float)c->cfg.softParticleStepDistance;
    const uint16_t* brick = c->bricks[c->mvIndex(dr.x, dr.y, dr.z)].data();
    int n = 0;
    for (int stepIndex = tExit; stepIndex >= tEntry; stepIndex--) {  // :254
        V3 sp = v3(pos.x + 0.5f, pos.y + 0.5f, pos.z + 0.5f);        // :255
        sp = v3(sp.x * scale + borderVoxelOffset, sp.y * scale + borderVoxelOffset, sp.z * scale + borderVoxelOffset);  // :258
        float vc[4];
        sample_brick(brick, N, sp, vc);                              // :262
        if (footprint) {  // measurement: mark the sample's 8-texel trilinear footprint
            const float Nf = (float)N;
            int ix = (int)floorf(sp.x * Nf - 0.5f), iy = (int)floorf(sp.y * Nf - 0.5f), iz = (int)floorf(sp.z * Nf - 0.5f);
            for (int k = 0; k < 8; k++) {
                int x = wrap(ix + (k & 1), N), y = wrap(iy + ((k >> 1) & 1), N), z = wrap(iz + (k >> 2), N);
                (*footprint)[((size_t)z * N + y) * N + x] = 1;
            }
        }
        float density = vc[3];
        if (stepIndex - tCamera < c->cfg.softParticleStepDistance)   // :267
            density *= (float)(stepIndex - tCamera) * softRcp;       // :269
        float blend = 1.0f / (1.0f + density);                       // :272
        for (int ch = 0; ch < 3; ch++) result[ch] = vc[ch] + blend * (result[ch] - vc[ch]);  // :274 lerp(color, result, blend)
        transmittance *= blend;                                      // :275
        pos = pos - rayStep;                                         // :277
        n++;
    }
    *samples += n;
    if (c->debugMode == 3) {  // sample-count bands, March.shader:283-299
        static const float bands[7][4] = {{0.0f, 0.2f, 0.0f, 0.5f}, {0.0f, 0.5f, 0.0f, 0.5f}, {0.5f, 0.5f, 0.0f, 0.5f}, {0.6f, 0.4f, 0.0f, 0.5f},
                                          {0.6f, 0.0f, 0.0f, 0.5f}, {0.8f, 0.0f, 0.0f, 0.5f}, {1.0f, 0.0f, 0.0f, 0.5f}};
        int b = n < 5 ? 0 : n < 10 ? 1 : n < 20 ? 2 : n < 30 ? 3 : n < 40 ? 4 : n < 50 ? 5 : 6;
        for (int ch = 0; ch < 4; ch++) src[ch] = bands[b][ch];
        return true;
    }
    src[0] = result[0]; src[1] = result[1]; src[2] = result[2]; src[3] = 1.0f - transmittance;  // :301
    return true;
}

// float -> UNORM8 -> float, as the ROP of an ARGB32 target stores and re-reads it (particlesRT, VPR.cs:228):
// saturate, scale by 255, round half up.
inline float quantize_unorm8(float x) {
    float v = fminf(fmaxf(x, 0.0f), 1.0f);
    v = floorf(v * 255.0f + 0.5f);
    return v / 255.0f;
}

// Fixed-function blend between metavoxels (March.shader:14-18, VPR.cs:659-662 / 688-691).
inline void rop_blend(float dst[4], const float src[4], bool over) {
    if (over) {  // Blend One OneMinusSrcAlpha
        float k = 1.0f - src[3];
        for (int ch = 0; ch < 4; ch++) dst[ch] = src[ch] + dst[ch] * k;
    } else {     // Blend OneMinusDstAlpha One
        float k = 1.0f - dst[3];
        for (int ch = 0; ch < 4; ch++) dst[ch] = src[ch] * k + dst[ch];
    }
}

int march_pixels_impl(VpeContext* c, const VpeCamera* cam, const int32_t* pixels, int n, float* rgba, int32_t* samples,
                      float* overPart, float* underPart, std::vector<std::vector<uint8_t>>* footprints = nullptr) {
    if (!c->lightSet) return fail(c, VPE_E_NOT_READY, "vpe_set_light has not been called");
    if (!c->prepared) return fail(c, VPE_E_NOT_READY, "vpe_fill has not been called");
    if (cam->width < 1 || cam->height < 1) return fail(c, VPE_E_INVALID_ARG, "bad image size");
    if (!c->sceneDepth.empty() && (c->sceneW != cam->width || c->sceneH != cam->height))
        return fail(c, VPE_E_INVALID_ARG, "scene depth buffer does not match the camera's image size");
    if ((overPart || underPart) && (c->targetFormat != 0 || c->debugMode != 0 || !c->sceneDepth.empty()))
        return fail(c, VPE_E_UNSUPPORTED, "march options are not available for slab partial images");
    double t0 = now_ms();
    MarchPass mp;
    build_march_pass(c, cam, &mp);
    const int total = cam->width * cam->height;
    const int count = pixels ? n : total;
    int64_t totalSamples = 0;
    int bad = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : totalSamples) if (!footprints)
    for (int i = 0; i < count; i++) {
        int pix = pixels ? pixels[i] : i;
        if (pix < 0 || pix >= total) { bad = 1; continue; }
        int px = pix % cam->width, py = pix / cam->width;
        RaySetup rs = ray_setup(c, &mp, px, py);
        float dst[4] = {0, 0, 0, 0};      // VPR.cs:171-172 clear colour
        float dOver[4] = {0, 0, 0, 0}, dUnder[4] = {0, 0, 0, 0};
        int ns = 0;
        const float pcx = (float)px + 0.5f, pcy = (float)py + 0.5f;
        const float sceneEye = c->sceneDepth.empty() ? 3.0e38f : c->sceneDepth[(size_t)py * cam->width + px];
        int orderIndex = -1;
        for (const MarchPass::Draw& dr : mp.draws) {
            float src[4];
            orderIndex++;  // RenderMetavoxel(xx, yy, zz, mvCount++), VPR.cs:675,705
            if (pcx < dr.px0 || pcx > dr.px1 || pcy < dr.py0 || pcy > dr.py1) continue;  // not rasterised
            if (!c->filled[c->mvIndex(dr.x, dr.y, dr.z)]) {
                // region-filled oracle (vpe_fill_region on a subset): a ray may only enter filled metavoxels
                V3 o = mul4(dr.C2M, rs.start, 1.0f);
                V3 d = normalize_hlsl(mul4(dr.C2M, rs.dir, 0.0f));
                float t1, t2;
                if (intersect_box(o, d, &t1, &t2)) bad = 2;
                continue;
            }
            std::vector<uint8_t>* fp = nullptr;
            if (footprints) {
                fp = &(*footprints)[c->mvIndex(dr.x, dr.y, dr.z)];
                if (fp->empty()) fp->assign((size_t)c->N() * c->N() * c->N(), 0);
            }
            if (!march_frag(c, rs, dr, src, &ns, fp, sceneEye, orderIndex, (int)mp.draws.size())) continue;  // seethrough: blending (0,0,0,0) is the identity
            rop_blend(dst, src, dr.over);
            if (c->targetFormat == 1)
                for (int ch = 0; ch < 4; ch++) dst[ch] = quantize_unorm8(dst[ch]);
            if (overPart && dr.over) rop_blend(dOver, src, true);
            if (underPart && !dr.over) rop_blend(dUnder, src, false);
        }
        for (int ch = 0; ch < 4; ch++) {
            if (rgba) rgba[(size_t)i * 4 + ch] = dst[ch];
            if (overPart) overPart[(size_t)i * 4 + ch] = dOver[ch];
            if (underPart) underPart[(size_t)i * 4 + ch] = dUnder[ch];
        }
        if (samples) samples[i] = ns;
        totalSamples += ns;
    }
    if (bad == 2) return fail(c, VPE_E_NOT_READY, "a ray entered a covered metavoxel that has not been filled");
    if (bad) return fail(c, VPE_E_INVALID_ARG, "pixel index out of range");
    c->stats.raySamples = totalSamples;
    c->stats.zBoundary = mp.zBoundary;
    c->stats.marchLaunches = 0;
    c->stats.marchMs = (float)(now_ms() - t0);
    return VPE_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------
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
    (void)device;
    if (!cfg || !out) return VPE_E_INVALID_ARG;
    VpeConfig c2 = *cfg;
    normalise_slab(c2);
    std::string why;
    if (validate_config(c2, why)) return VPE_E_INVALID_ARG;
    VpeContext* c = new VpeContext();
    c->cfg = c2;
    memset(&c->stats, 0, sizeof(c->stats));
    c->sb = c2.mvScale * (float)c2.numVoxelsInMetavoxel / (float)(c2.numVoxelsInMetavoxel - 2 * c2.numBorderVoxels);  // VPR.cs:139
    size_t nmv = (size_t)c2.numMetavoxelsX * c2.numMetavoxelsY * c2.numMetavoxelsZ;
    c->lists.resize(nmv);
    c->bricks.resize(nmv);
    c->filled.assign(nmv, 0);
    c->sheet.assign((size_t)c2.numMetavoxelsX * c2.numMetavoxelsY * c2.numVoxelsInMetavoxel * c2.numVoxelsInMetavoxel, 1.0f);
    c->center = v3(0, 0, 0);
    *out = c;
    return VPE_OK;
}

int vpe_destroy(VpeContext* c) {
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
        c2.numMetavoxelsZ != c->cfg.numMetavoxelsZ || c2.numVoxelsInMetavoxel != c->cfg.numVoxelsInMetavoxel ||
        c2.slabZBegin != c->cfg.slabZBegin || c2.slabZEnd != c->cfg.slabZEnd)
        return fail(c, VPE_E_INVALID_ARG, "grid dims, voxel count and slab are fixed at create");
    bool geom = c2.mvScale != c->cfg.mvScale || c2.numBorderVoxels != c->cfg.numBorderVoxels ||
                c2.lightCameraDistance != c->cfg.lightCameraDistance;
    c->cfg = c2;
    c->sb = c2.mvScale * (float)c2.numVoxelsInMetavoxel / (float)(c2.numVoxelsInMetavoxel - 2 * c2.numBorderVoxels);
    if (geom && c->lightSet) update_metavoxel_positions(c);  // SetGridScale, VPR.cs:1059-1063
    return VPE_OK;
}

int vpe_set_light(VpeContext* c, const VpeTransform* light, const float gridCenter[3]) {
    if (!c || !light || !gridCenter) return fail(c, VPE_E_INVALID_ARG, "null argument");
    c->light = *light;
    c->center = v3(gridCenter[0], gridCenter[1], gridCenter[2]);
    c->lightSet = true;
    update_metavoxel_positions(c);
    return VPE_OK;
}

int vpe_set_displacement_cubemap(VpeContext* c, const uint8_t* r8, int edge) {
    if (!c || !r8 || edge < 1) return fail(c, VPE_E_INVALID_ARG, "bad cubemap");
    c->cube.assign(r8, r8 + (size_t)6 * edge * edge);
    c->cubeEdge = edge;
    return VPE_OK;
}

int vpe_set_light_depth_map(VpeContext* c, const float* depth01) {
    if (!c) return VPE_E_INVALID_ARG;
    if (!depth01) { c->depthMap.clear(); return VPE_OK; }
    size_t n = (size_t)c->NX() * c->N() * c->NY() * c->N();
    c->depthMap.assign(depth01, depth01 + n);
    return VPE_OK;
}

// ≙ lightCamera.RenderWithShader(generateLightDepthMapShader) (VPR.cs:184) with the camera of InitCameraAtLight
// (VPR.cs:320-367: orthographic, extents +-NX*s/2 x +-NY*s/2, near/far 0.3/1000, at centre - forward*200, the
// light's rotation) and GenerateLightDepthMap.shader:6 (Cull Front, ZWrite On, ZTest Less; depth only).
// Rasteriser restated: pixel centres, top-left rule, fp32 edge functions, affine depth (orthographic).
// Unity front faces are clockwise; with Cull Front the counter-clockwise (back) faces are drawn.
int vpe_render_light_depth_map(VpeContext* c, const float* tris, int numTriangles) {
    if (!c || (!tris && numTriangles > 0) || numTriangles < 0) return fail(c, VPE_E_INVALID_ARG, "bad triangle list");
    if (!c->lightSet) return fail(c, VPE_E_NOT_READY, "vpe_set_light has not been called");
    const int W = c->NX() * c->N(), H = c->NY() * c->N();
    c->depthMap.assign((size_t)W * H, 1.0f);  // CameraClearFlags.Depth, VPR.cs:347
    const float r = (float)c->NX() * c->cfg.mvScale * 0.5f, t = (float)c->NY() * c->cfg.mvScale * 0.5f;  // VPR.cs:340
    const float zn = c->cfg.lightNear, zf = c->cfg.lightFar;
    for (int i = 0; i < numTriangles; i++) {
        float sx[3], sy[3], sz[3];
        for (int k = 0; k < 3; k++) {
            V3 p = mul4(c->W2LC, v3(tris[i * 9 + k * 3], tris[i * 9 + k * 3 + 1], tris[i * 9 + k * 3 + 2]), 1.0f);
            sx[k] = (p.x / r * 0.5f + 0.5f) * (float)W;   // pixel coordinates; row 0 = light-space y = -t (uv.y = 0)
            sy[k] = (p.y / t * 0.5f + 0.5f) * (float)H;
            sz[k] = (p.z - zn) / (zf - zn);               // D3D orthographic depth, linear in [0,1]
        }
        const float area = (sx[1] - sx[0]) * (sy[2] - sy[0]) - (sy[1] - sy[0]) * (sx[2] - sx[0]);
        if (!(area > 0.0f)) continue;  // clockwise = front face: culled (Cull Front); degenerate: nothing to draw
        const float minx = fminf(sx[0], fminf(sx[1], sx[2])), maxx = fmaxf(sx[0], fmaxf(sx[1], sx[2]));
        const float miny = fminf(sy[0], fminf(sy[1], sy[2])), maxy = fmaxf(sy[0], fmaxf(sy[1], sy[2]));
        const int x0 = std::max(0, (int)floorf(minx - 0.5f)), x1 = std::min(W - 1, (int)ceilf(maxx - 0.5f));
        const int y0 = std::max(0, (int)floorf(miny - 0.5f)), y1 = std::min(H - 1, (int)ceilf(maxy - 0.5f));
        for (int y = y0; y <= y1; y++)
            for (int x = x0; x <= x1; x++) {
                const float px = (float)x + 0.5f, py = (float)y + 0.5f;
                float w[3];
                bool inside = true;
                for (int k = 0; k < 3 && inside; k++) {
                    const int a = (k + 1) % 3, b = (k + 2) % 3;  // edge opposite vertex k
                    const float ex = sx[b] - sx[a], ey = sy[b] - sy[a];
                    w[k] = ex * (py - sy[a]) - ey * (px - sx[a]);
                    // top-left rule for a counter-clockwise triangle in a y-up raster: an edge owns its points
                    // if it is a left edge (going down) or a horizontal top edge (going left)
                    const bool owns = ey < 0.0f || (ey == 0.0f && ex < 0.0f);
                    if (w[k] < 0.0f || (w[k] == 0.0f && !owns)) inside = false;
                }
                if (!inside) continue;
                const float z = ((w[0] / area) * sz[0] + (w[1] / area) * sz[1]) + (w[2] / area) * sz[2];
                if (!(z >= 0.0f && z <= 1.0f)) continue;  // near / far clip
                float& dst = c->depthMap[(size_t)y * W + x];
                if (z < dst) dst = z;  // ZTest Less, ZWrite On
            }
    }
    return VPE_OK;
}

int vpe_read_light_depth_map(VpeContext* c, float* depth01) {
    if (!c || !depth01) return VPE_E_INVALID_ARG;
    const size_t n = (size_t)c->NX() * c->N() * c->NY() * c->N();
    for (size_t i = 0; i < n; i++) depth01[i] = c->depthMap.empty() ? 1.0f : c->depthMap[i];
    return VPE_OK;
}

// debug switches select kernels / layouts of the CUDA library; the oracle has one path and ignores them
int vpe_set_debug_options(VpeContext* c, const VpeDebugOptions* o) { return (c && o) ? VPE_OK : VPE_E_INVALID_ARG; }
int vpe_debug_div_rn(VpeContext* c, const float* a, const float* b, float* q, int n) {
    if (!c || !a || !b || !q || n < 0) return VPE_E_INVALID_ARG;
    for (int i = 0; i < n; i++) q[i] = a[i] / b[i];
    return VPE_OK;
}
int vpe_read_slice_profile(VpeContext* c, int64_t*, int64_t*, int64_t*) { return fail(c, VPE_E_UNSUPPORTED, "slice profile: CUDA library only"); }

int vpe_set_march_options(VpeContext* c, const VpeMarchOptions* o) {
    if (!c || !o) return VPE_E_INVALID_ARG;
    if (o->targetFormat < 0 || o->targetFormat > 1 || o->debugMode < 0 || o->debugMode > 3) return fail(c, VPE_E_INVALID_ARG, "bad march option");
    if (o->sceneDepth && (o->sceneWidth < 1 || o->sceneHeight < 1)) return fail(c, VPE_E_INVALID_ARG, "bad scene depth size");
    c->targetFormat = o->targetFormat;
    c->debugMode = o->debugMode;
    c->sceneDepth.clear();
    if (o->sceneDepth) {
        c->sceneW = o->sceneWidth; c->sceneH = o->sceneHeight;
        c->sceneDepth.assign(o->sceneDepth, o->sceneDepth + (size_t)o->sceneWidth * o->sceneHeight);
    }
    return VPE_OK;
}

// ≙ Graphics.Blit(particlesRT, mainSceneRT, matBlendParticles) (VPR.cs:210) with CompositeParticles.shader:10
// `Blend One OneMinusSrcAlpha, One One`.
int vpe_composite_scene(VpeContext* c, const float* particles, float* scene, int numPixels, int targetFormat) {
    if (!c || !particles || !scene || numPixels < 0 || targetFormat < 0 || targetFormat > 1) return fail(c, VPE_E_INVALID_ARG, "bad argument");
    for (size_t i = 0; i < (size_t)numPixels; i++) {
        const float* s = particles + i * 4;
        float* d = scene + i * 4;
        const float k = 1.0f - s[3];
        d[0] = s[0] + d[0] * k; d[1] = s[1] + d[1] * k; d[2] = s[2] + d[2] * k;
        d[3] = s[3] + d[3];
        if (targetFormat == 1)
            for (int ch = 0; ch < 4; ch++) d[ch] = quantize_unorm8(d[ch]);
    }
    return VPE_OK;
}

int vpe_fill_prepare(VpeContext* c, const VpeParticle* particles, int n, const VpeTransform* emitter, int onDevice) {
    if (!c || (!particles && n > 0) || n < 0 || !emitter) return fail(c, VPE_E_INVALID_ARG, "bad particle input");
    if (onDevice) return fail(c, VPE_E_UNSUPPORTED, "the oracle has no device path");
    if (!c->lightSet) return fail(c, VPE_E_NOT_READY, "vpe_set_light has not been called");
    if (c->cubeEdge == 0) return fail(c, VPE_E_NOT_READY, "vpe_set_displacement_cubemap has not been called");
    bin_particles(c, particles, n, emitter);
    std::fill(c->filled.begin(), c->filled.end(), 0);
    std::fill(c->sheet.begin(), c->sheet.end(), 1.0f);  // VPR.cs:498-499 clear to Color.red (R = 1)
    c->prepared = true;
    return VPE_OK;
}

int vpe_fill_region(VpeContext* c, int x0, int x1, int y0, int y1) {
    if (!c) return VPE_E_INVALID_ARG;
    if (!c->prepared) return fail(c, VPE_E_NOT_READY, "vpe_fill_prepare has not been called");
    if (x0 < 0 || y0 < 0 || x1 > c->NX() || y1 > c->NY() || x0 >= x1 || y0 >= y1) return fail(c, VPE_E_INVALID_ARG, "bad region");
    double t0 = now_ms();
    fill_region(c, x0, x1, y0, y1);
    c->stats.fillMs = (float)(now_ms() - t0);
    c->stats.fillLaunches = 0;
    return VPE_OK;
}

// The split entry points of the multi-GPU fill. The oracle keeps the reference's fused per-column
// fragment: the density pass has nothing to do ahead of time, the sweep of a region is the fused fill.
int vpe_fill_density(VpeContext* c) {
    if (!c) return VPE_E_INVALID_ARG;
    if (!c->prepared) return fail(c, VPE_E_NOT_READY, "vpe_fill_prepare has not been called");
    return VPE_OK;
}
int vpe_fill_sweep_region(VpeContext* c, int x0, int x1, int y0, int y1) { return vpe_fill_region(c, x0, x1, y0, y1); }

int vpe_fill(VpeContext* c, const VpeParticle* particles, int n, const VpeTransform* emitter) {
    double t0 = now_ms();
    int rc = vpe_fill_prepare(c, particles, n, emitter, 0);
    if (rc) return rc;
    rc = vpe_fill_region(c, 0, c->NX(), 0, c->NY());
    c->stats.fillMs = (float)(now_ms() - t0);
    return rc;
}

int vpe_march(VpeContext* c, const VpeCamera* cam, float* rgba, int32_t* samples) {
    if (!c || !cam || !rgba) return fail(c, VPE_E_INVALID_ARG, "null argument");
    return march_pixels_impl(c, cam, nullptr, 0, rgba, samples, nullptr, nullptr);
}

int vpe_march_pixels(VpeContext* c, const VpeCamera* cam, const int32_t* pixels, int n, float* rgba, int32_t* samples) {
    if (!c || !cam || !rgba || !pixels || n < 0) return fail(c, VPE_E_INVALID_ARG, "null argument");
    return march_pixels_impl(c, cam, pixels, n, rgba, samples, nullptr, nullptr);
}

int vpe_march_footprint(VpeContext* c, const VpeCamera* cam, int64_t* uniqueTexels) {
    if (!c || !cam || !uniqueTexels) return fail(c, VPE_E_INVALID_ARG, "null argument");
    std::vector<std::vector<uint8_t>> fps(c->lists.size());
    std::vector<float> rgba((size_t)cam->width * cam->height * 4);
    int rc = march_pixels_impl(c, cam, nullptr, 0, rgba.data(), nullptr, nullptr, nullptr, &fps);
    int64_t n = 0;
    for (auto& f : fps)
        for (uint8_t b : f) n += b;
    *uniqueTexels = n;
    return rc;
}

// Oracle-only: slab partial images on the host (same semantics as vpe_march_partial_device).
int vpe_ref_march_partial(VpeContext* c, const VpeCamera* cam, float* over, float* under, int32_t* samples) {
    if (!c || !cam || !over || !under) return fail(c, VPE_E_INVALID_ARG, "null argument");
    return march_pixels_impl(c, cam, nullptr, 0, nullptr, samples, over, under);
}

// Oracle-only: which metavoxels (flat z,y,x index) do the listed pixels' rays enter?  Lets a test
// fill only the metavoxel columns a pixel subset needs (lazy fill at BASELINE's full sizes).
int vpe_ref_touched_metavoxels(VpeContext* c, const VpeCamera* cam, const int32_t* pixels, int n, uint8_t* touched) {
    if (!c || !cam || !pixels || !touched) return fail(c, VPE_E_INVALID_ARG, "null argument");
    if (!c->lightSet || !c->prepared) return fail(c, VPE_E_NOT_READY, "fill_prepare first");
    MarchPass mp;
    build_march_pass(c, cam, &mp);
    size_t nmv = c->lists.size();
    memset(touched, 0, nmv);
    for (int i = 0; i < n; i++) {
        int px = pixels[i] % cam->width, py = pixels[i] / cam->width;
        RaySetup rs = ray_setup(c, &mp, px, py);
        for (const MarchPass::Draw& dr : mp.draws) {
            V3 o = mul4(dr.C2M, rs.start, 1.0f);
            V3 d = normalize_hlsl(mul4(dr.C2M, rs.dir, 0.0f));
            float t1, t2;
            if (intersect_box(o, d, &t1, &t2)) touched[c->mvIndex(dr.x, dr.y, dr.z)] = 1;
        }
    }
    return VPE_OK;
}

int vpe_set_stream(VpeContext* c, void*) { return fail(c, VPE_E_UNSUPPORTED, "oracle"); }
int vpe_fill_device(VpeContext* c, const VpeParticle*, int, const VpeTransform*) { return fail(c, VPE_E_UNSUPPORTED, "oracle"); }
int vpe_march_device(VpeContext* c, const VpeCamera*, float*, int32_t*) { return fail(c, VPE_E_UNSUPPORTED, "oracle"); }
float* vpe_light_sheet_device(VpeContext*) { return nullptr; }
int vpe_sheet_link_create(VpeContext* c, void*, void**) { return fail(c, VPE_E_UNSUPPORTED, "oracle"); }
int vpe_sheet_link_connect(VpeContext* c, const void*, const void*, int) { return fail(c, VPE_E_UNSUPPORTED, "oracle"); }
int vpe_fill_sweep_linked(VpeContext* c) { return fail(c, VPE_E_UNSUPPORTED, "oracle"); }
int vpe_fill_linked(VpeContext* c) { return fail(c, VPE_E_UNSUPPORTED, "oracle"); }
int vpe_read_sample_bitmap(VpeContext* c, int, int, int, uint32_t*, int*) { return fail(c, VPE_E_UNSUPPORTED, "oracle"); }
int vpe_sheet_link_status(VpeContext* c, int*) { return fail(c, VPE_E_UNSUPPORTED, "oracle"); }
int vpe_march_partial_device(VpeContext* c, const VpeCamera*, float*, float*, int32_t*) { return fail(c, VPE_E_UNSUPPORTED, "oracle"); }
int vpe_image_link_create(VpeContext* c, int, int, int, int, void*, void**) { return fail(c, VPE_E_UNSUPPORTED, "oracle"); }
int vpe_image_link_connect(VpeContext* c, const void* const*, int) { return fail(c, VPE_E_UNSUPPORTED, "oracle"); }
int vpe_march_linked(VpeContext* c, const VpeCamera*, int32_t*) { return fail(c, VPE_E_UNSUPPORTED, "oracle"); }
int vpe_composite_linked(VpeContext* c, float*) { return fail(c, VPE_E_UNSUPPORTED, "oracle"); }
int vpe_image_link_status(VpeContext* c, int*) { return fail(c, VPE_E_UNSUPPORTED, "oracle"); }
int vpe_composite_device(VpeContext* c, const float* const*, int, int, float*) { return fail(c, VPE_E_UNSUPPORTED, "oracle"); }

// Oracle-only: host-side light sheet access for the slab hand-off tests.
int vpe_ref_write_light_sheet(VpeContext* c, const float* sheet) {
    if (!c || !sheet) return VPE_E_INVALID_ARG;
    std::copy(sheet, sheet + c->sheet.size(), c->sheet.begin());
    return VPE_OK;
}

int vpe_read_brick(VpeContext* c, int x, int y, int z, uint16_t* half4, int* covered) {
    if (!c || !covered) return VPE_E_INVALID_ARG;
    if (x < 0 || y < 0 || z < 0 || x >= c->NX() || y >= c->NY() || z >= c->NZ()) return fail(c, VPE_E_INVALID_ARG, "metavoxel index out of range");
    size_t mi = c->mvIndex(x, y, z);
    *covered = (!c->lists[mi].empty() && c->filled[mi]) ? 1 : 0;
    if (*covered && half4) memcpy(half4, c->bricks[mi].data(), c->bricks[mi].size() * sizeof(uint16_t));
    return VPE_OK;
}

int vpe_read_light_sheet(VpeContext* c, float* sheet) {
    if (!c || !sheet) return VPE_E_INVALID_ARG;
    memcpy(sheet, c->sheet.data(), c->sheet.size() * sizeof(float));
    return VPE_OK;
}

int vpe_read_particle_list(VpeContext* c, int x, int y, int z, int32_t* idx, int cap, int* n) {
    if (!c || !n) return VPE_E_INVALID_ARG;
    if (x < 0 || y < 0 || z < 0 || x >= c->NX() || y >= c->NY() || z >= c->NZ()) return fail(c, VPE_E_INVALID_ARG, "metavoxel index out of range");
    const std::vector<int32_t>& l = c->lists[c->mvIndex(x, y, z)];
    *n = (int)l.size();
    for (int i = 0; i < (int)l.size() && i < cap && idx; i++) idx[i] = l[i];
    return VPE_OK;
}

int vpe_read_metavoxel_position(VpeContext* c, int x, int y, int z, float pos[3]) {
    if (!c || !pos) return VPE_E_INVALID_ARG;
    if (!c->lightSet) return fail(c, VPE_E_NOT_READY, "vpe_set_light has not been called");
    if (x < 0 || y < 0 || z < 0 || x >= c->NX() || y >= c->NY() || z >= c->NZ()) return fail(c, VPE_E_INVALID_ARG, "metavoxel index out of range");
    V3 p = c->mvPos[c->mvIndex(x, y, z)];
    pos[0] = p.x; pos[1] = p.y; pos[2] = p.z;
    return VPE_OK;
}

int vpe_get_stats(VpeContext* c, VpeStats* s) {
    if (!c || !s) return VPE_E_INVALID_ARG;
    *s = c->stats;
    int64_t bytes = 0;
    for (auto& b : c->bricks) bytes += (int64_t)b.size() * 2;
    s->brickPoolBytes = bytes;
    s->fillKernelMs = s->fillMs;
    s->marchKernelMs = s->marchMs;
    return VPE_OK;
}

const char* vpe_last_error(VpeContext* c) { return c ? c->err.c_str() : "null context"; }
int vpe_abi_version(void) { return VPE_ABI_VERSION; }
const char* vpe_backend(void) { return "oracle"; }

// Oracle-only helpers exposed for the tests that pin the oracle's own primitives.
uint16_t vpe_ref_f32_to_f16(float f) { return f32_to_f16(f); }
float vpe_ref_f16_to_f32(uint16_t h) { return f16_to_f32(h); }
int vpe_ref_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
// Launchers such as torch.distributed.run export OMP_NUM_THREADS=1; a CPU baseline that claims "all host cores"
// sets its thread count itself. n <= 0: the number of processors. Returns the count in force.
int vpe_ref_set_num_threads(int n) {
#ifdef _OPENMP
    if (n <= 0) n = omp_get_num_procs();
    omp_set_dynamic(0);
    omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

}  // extern "C"
