// vpe_math.cuh — the engine's "normative arithmetic" (DESIGN.md §Normative arithmetic).
//
// Everything in here is IEEE fp32 with NO fused multiply-add: the translation unit is compiled
// with `-fmad=false` (device) and `-ffp-contract=off` (host), and expressions are evaluated left
// to right as written.  Host and device share these functions so that a quantity computed once on
// the host (pass constants) and per thread on the device (per-metavoxel, per-particle terms) has
// the same bits, and so that the discontinuous decisions of the reference — binning truncation
// (VPR.cs:434-438), `dist2 <= 0.25` (Fill.shader:172,198), `(int)` shadow index
// (Fill.shader:221), ceil/floor of t/step (March.shader:236-239) — resolve identically to the
// CPU oracle.  The oracle has its own, independently written general-matrix versions of these.
//
// Where the reference calls a general 4x4 routine on an affine matrix (last row 0,0,0,1), the
// functions below evaluate the same Laplace-expansion formula with the terms that are exactly
// zero / one removed; removing `x*0`, `+0` and `*1` does not change any rounding.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define VPE_HD __host__ __device__ __forceinline__
#else
#define VPE_HD inline
#endif

namespace vpe {

struct F3 {
    float x, y, z;
};
// rows 0..2 of an affine 4x4 (column-vector convention): m[r][c], c = 3 is the translation
struct Affine {
    float m[3][4];
};
struct M3 {
    float m[3][3];
};

VPE_HD F3 f3(float x, float y, float z) { F3 r; r.x = x; r.y = y; r.z = z; return r; }
VPE_HD F3 add(F3 a, F3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
VPE_HD F3 sub(F3 a, F3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
VPE_HD float dot3(F3 a, F3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }

// Unity Quaternion (x,y,z,w) -> rotation matrix.
VPE_HD M3 quat_to_m3(const float q[4]) {
    float x = q[0] * 2.0f, y = q[1] * 2.0f, z = q[2] * 2.0f;
    float xx = q[0] * x, yy = q[1] * y, zz = q[2] * z;
    float xy = q[0] * y, xz = q[0] * z, yz = q[1] * z;
    float wx = q[3] * x, wy = q[3] * y, wz = q[3] * z;
    M3 r;
    r.m[0][0] = 1.0f - (yy + zz); r.m[0][1] = xy - wz;          r.m[0][2] = xz + wy;
    r.m[1][0] = xy + wz;          r.m[1][1] = 1.0f - (xx + zz); r.m[1][2] = yz - wx;
    r.m[2][0] = xz - wy;          r.m[2][1] = yz + wx;          r.m[2][2] = 1.0f - (xx + yy);
    return r;
}

// Matrix4x4.TRS(t, q, (s,s,s)) — rows 0..2.
VPE_HD Affine trs(F3 t, const M3& r, float s) {
    Affine a;
    for (int i = 0; i < 3; i++) {
        a.m[i][0] = r.m[i][0] * s;
        a.m[i][1] = r.m[i][1] * s;
        a.m[i][2] = r.m[i][2] * s;
    }
    a.m[0][3] = t.x; a.m[1][3] = t.y; a.m[2][3] = t.z;
    return a;
}

// The translation-independent part of Matrix4x4.inverse of an affine matrix: the 2x2
// sub-determinants s0,s1,s3, 1/det and the inverse's 3x3 block.
struct AffineInvLin {
    float s0, s1, s3, inv;
    float b[3][3];
};

VPE_HD AffineInvLin affine_inverse_linear(const Affine& A) {
    const float (*a)[4] = A.m;
    AffineInvLin L;
    L.s0 = a[0][0] * a[1][1] - a[1][0] * a[0][1];
    L.s1 = a[0][0] * a[1][2] - a[1][0] * a[0][2];
    L.s3 = a[0][1] * a[1][2] - a[1][1] * a[0][2];
    float det = (L.s0 * a[2][2] - L.s1 * a[2][1]) + L.s3 * a[2][0];
    L.inv = 1.0f / det;
    L.b[0][0] = (a[1][1] * a[2][2] - a[1][2] * a[2][1]) * L.inv;
    L.b[0][1] = (-a[0][1] * a[2][2] + a[0][2] * a[2][1]) * L.inv;
    L.b[0][2] = L.s3 * L.inv;
    L.b[1][0] = (-a[1][0] * a[2][2] + a[1][2] * a[2][0]) * L.inv;
    L.b[1][1] = (a[0][0] * a[2][2] - a[0][2] * a[2][0]) * L.inv;
    L.b[1][2] = -L.s1 * L.inv;
    L.b[2][0] = (a[1][0] * a[2][1] - a[1][1] * a[2][0]) * L.inv;
    L.b[2][1] = (-a[0][0] * a[2][1] + a[0][1] * a[2][0]) * L.inv;
    L.b[2][2] = L.s0 * L.inv;
    return L;
}

// The translation column of Matrix4x4.inverse of the affine matrix with linear part `A`
// (only A's 3x3 block is read) and translation t.
VPE_HD F3 affine_inverse_translation(const Affine& A, const AffineInvLin& L, F3 t) {
    const float (*a)[4] = A.m;
    float s2 = a[0][0] * t.y - a[1][0] * t.x;
    float s4 = a[0][1] * t.y - a[1][1] * t.x;
    float s5 = a[0][2] * t.y - a[1][2] * t.x;
    F3 r;
    r.x = (-a[2][1] * s5 + a[2][2] * s4 - t.z * L.s3) * L.inv;
    r.y = (a[2][0] * s5 - a[2][2] * s2 + t.z * L.s1) * L.inv;
    r.z = (-a[2][0] * s4 + a[2][1] * s2 - t.z * L.s0) * L.inv;
    return r;
}

VPE_HD Affine affine_inverse(const Affine& A) {
    AffineInvLin L = affine_inverse_linear(A);
    F3 t = affine_inverse_translation(A, L, f3(A.m[0][3], A.m[1][3], A.m[2][3]));
    Affine B;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) B.m[i][j] = L.b[i][j];
    B.m[0][3] = t.x; B.m[1][3] = t.y; B.m[2][3] = t.z;
    return B;
}

// Matrix4x4.MultiplyPoint3x4 / HLSL mul(M, float4(p,1)).xyz
VPE_HD F3 xform_point(const Affine& a, F3 p) {
    return f3(((a.m[0][0] * p.x + a.m[0][1] * p.y) + a.m[0][2] * p.z) + a.m[0][3],
              ((a.m[1][0] * p.x + a.m[1][1] * p.y) + a.m[1][2] * p.z) + a.m[1][3],
              ((a.m[2][0] * p.x + a.m[2][1] * p.y) + a.m[2][2] * p.z) + a.m[2][3]);
}
// HLSL mul(M, float4(v,0)).xyz
VPE_HD F3 xform_dir(const Affine& a, F3 p) {
    return f3((a.m[0][0] * p.x + a.m[0][1] * p.y) + a.m[0][2] * p.z,
              (a.m[1][0] * p.x + a.m[1][1] * p.y) + a.m[1][2] * p.z,
              (a.m[2][0] * p.x + a.m[2][1] * p.y) + a.m[2][2] * p.z);
}

// Affine product (Matrix4x4 operator* restricted to rows 0..2 of affine operands).
VPE_HD Affine affine_mul(const Affine& a, const Affine& b) {
    Affine r;
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++)
            r.m[i][j] = (a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j]) + a.m[i][2] * b.m[2][j];
        r.m[i][3] = ((a.m[i][0] * b.m[0][3] + a.m[i][1] * b.m[1][3]) + a.m[i][2] * b.m[2][3]) + a.m[i][3];
    }
    return r;
}

// float -> int as D3D ftoi / CUDA __float2int_rz: truncate, saturate, NaN -> 0.
VPE_HD int ftoi_sat(float f) {
#if defined(__CUDA_ARCH__)
    return __float2int_rz(f);
#else
    if (f != f) return 0;
    if (f >= 2147483648.0f) return 2147483647;
    if (f <= -2147483648.0f) return -2147483647 - 1;
    return (int)f;
#endif
}

VPE_HD float min_f(float a, float b) { return fminf(a, b); }
VPE_HD float max_f(float a, float b) { return fmaxf(a, b); }

}  // namespace vpe
