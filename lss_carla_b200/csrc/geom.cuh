// Per-point geometry arithmetic shared by the plan kernels (plan.cu) and the run-plan kernels (runplan.cu).
//
// All per-point arithmetic uses the round-to-nearest intrinsics (__fmul_rn/__fadd_rn/__fsub_rn/__fdiv_rn), which nvcc
// never contracts into FMAs, so results are bit-identical to the reference's evaluation of src/models.py:179-188,:212
// (oracle/lss_oracle.py documents the measured association of the 3x3 products).
#pragma once
#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// per-point arithmetic
// ------------------------------------------------------------------------------------------------

struct CalibPtrs {
    const float *frustum;     // [D,fH,fW,3]
    const float *post_trans;  // [B*N,3]
    const float *M1;          // [B*N,3,3] inverse(post_rots)
    const float *M2;          // [B*N,3,3] rots @ inverse(intrins)
    const float *trans;       // [B*N,3]
    const float *rots, *intrins, *post_rots;   // raw calibration [B*N,3,3] (fused plan build: M1/M2 made on the fly)
};

__device__ __forceinline__ float row_dot_unfused(const float *__restrict__ m, float v0, float v1, float v2) {
    // (a0*v0 + a1*v1) + a2*v2, one rounding per operation  (models.py:180,187 as ATen's CPU bmm evaluates it)
    return __fadd_rn(__fadd_rn(__fmul_rn(m[0], v0), __fmul_rn(m[1], v1)), __fmul_rn(m[2], v2));
}

__device__ __forceinline__ void ego_point(const CalibPtrs &c, int cam, int in_cam, float out[3],
                                          const float *m1 = nullptr, const float *m2 = nullptr) {
    if (m1 == nullptr) { m1 = c.M1 + cam * 9; m2 = c.M2 + cam * 9; }
    const float *fr = c.frustum + (size_t)in_cam * 3;
    const float *pt = c.post_trans + cam * 3;
    const float p0 = __fsub_rn(__ldg(fr + 0), __ldg(pt + 0));   // models.py:179
    const float p1 = __fsub_rn(__ldg(fr + 1), __ldg(pt + 1));
    const float p2 = __fsub_rn(__ldg(fr + 2), __ldg(pt + 2));
    const float q0 = row_dot_unfused(m1 + 0, p0, p1, p2);       // models.py:180
    const float q1 = row_dot_unfused(m1 + 3, p0, p1, p2);
    const float q2 = row_dot_unfused(m1 + 6, p0, p1, p2);
    const float r0 = __fmul_rn(q0, q2);                         // models.py:183-185
    const float r1 = __fmul_rn(q1, q2);
    const float *tr = c.trans + cam * 3;
    out[0] = __fadd_rn(row_dot_unfused(m2 + 0, r0, r1, q2), __ldg(tr + 0));   // models.py:187-188
    out[1] = __fadd_rn(row_dot_unfused(m2 + 3, r0, r1, q2), __ldg(tr + 1));
    out[2] = __fadd_rn(row_dot_unfused(m2 + 6, r0, r1, q2), __ldg(tr + 2));
}

// ((g - lo) / dx).long()  with x86 semantics for values a 64-bit integer cannot hold (models.py:212)
__device__ __forceinline__ long long quantise(float g, float lo, float dx, float inv_dx) {
    // a power-of-two voxel size divides exactly like a multiplication by its reciprocal (same bits, no IEEE division)
    const float t = __fsub_rn(g, lo);
    const float u = inv_dx != 0.0f ? __fmul_rn(t, inv_dx) : __fdiv_rn(t, dx);
    if (!(fabsf(u) < 9223372036854775808.0f)) return (long long)0x8000000000000000ULL;  // NaN, inf, overflow
    return __float2ll_rz(u);
}

__device__ __forceinline__ int voxel_of_point(const Dims &d, int b, const float g[3], long long ii[3]) {
    ii[0] = quantise(g[0], d.lo[0], d.dx[0], d.inv_dx[0]);
    ii[1] = quantise(g[1], d.lo[1], d.dx[1], d.inv_dx[1]);
    ii[2] = quantise(g[2], d.lo[2], d.dx[2], d.inv_dx[2]);
    const bool kept = ii[0] >= 0 && ii[0] < d.nx && ii[1] >= 0 && ii[1] < d.ny && ii[2] >= 0 && ii[2] < d.nz;  // :219-221
    if (!kept) return -1;
    return ((b * d.nz + (int)ii[2]) * d.nx + (int)ii[0]) * d.ny + (int)ii[1];
}


// Closed-form inverse with one explicit rounding per operation (no FMA contraction), so that every kernel that
// inlines it produces the same bits.
__device__ __forceinline__ float det2(float a, float b, float c, float d) { return __fsub_rn(__fmul_rn(a, d), __fmul_rn(b, c)); }
__device__ __forceinline__ void inv3x3(const float *a, float *o) {
    const float c00 = det2(a[4], a[5], a[7], a[8]), c01 = det2(a[5], a[3], a[8], a[6]), c02 = det2(a[3], a[4], a[6], a[7]);
    const float det = __fadd_rn(__fadd_rn(__fmul_rn(a[0], c00), __fmul_rn(a[1], c01)), __fmul_rn(a[2], c02));
    const float r = __frcp_rn(det);
    o[0] = __fmul_rn(c00, r); o[1] = __fmul_rn(det2(a[2], a[1], a[8], a[7]), r); o[2] = __fmul_rn(det2(a[1], a[2], a[4], a[5]), r);
    o[3] = __fmul_rn(c01, r); o[4] = __fmul_rn(det2(a[0], a[2], a[6], a[8]), r); o[5] = __fmul_rn(det2(a[2], a[0], a[5], a[3]), r);
    o[6] = __fmul_rn(c02, r); o[7] = __fmul_rn(det2(a[1], a[0], a[7], a[6]), r); o[8] = __fmul_rn(det2(a[0], a[1], a[3], a[4]), r);
}

// M1 = inverse(post_rots), M2 = rots @ inverse(intrins) of one camera (models.py:180,186 without the host round trip)
__device__ __forceinline__ void calib_matrices_of(const float *rots, const float *intrins, const float *post_rots, int cam,
                                                  float *M1, float *M2) {
    float a[9], inv[9];
    for (int i = 0; i < 9; ++i) a[i] = post_rots[cam * 9 + i];
    inv3x3(a, inv);
    for (int i = 0; i < 9; ++i) M1[i] = inv[i];
    for (int i = 0; i < 9; ++i) a[i] = intrins[cam * 9 + i];
    inv3x3(a, inv);
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            const float *R = rots + cam * 9 + r * 3;
            M2[r * 3 + c] = __fadd_rn(__fadd_rn(__fmul_rn(R[0], inv[c]), __fmul_rn(R[1], inv[3 + c])), __fmul_rn(R[2], inv[6 + c]));
        }
}


#define LSS_RAW_CAMS 8    // cameras a CTA may span when the calibration inverses are made on the fly (RAW builds)
