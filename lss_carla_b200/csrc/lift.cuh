// Lift operand preparation, the work of one CTA (shared by k_lift_prepare in lift.cu and the fused prologue in runplan.cu).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

#define LIFT_PX 32        // pixels per CTA (one 128-byte line of every channel row)
#define LIFT_THREADS 256  // 8 warps: warp w owns depth bins / channels w, w+8, ...
#define LIFT_WARPS 8

__device__ __forceinline__ float lift_load(const float *p) { return __ldg(p); }
__device__ __forceinline__ float lift_load(const __nv_bfloat16 *p) { return __bfloat162float(__ldg(p)); }

// The work of one CTA (LIFT_THREADS threads, `smem` = [D+C][33] floats); `cta` in [0, B*N*ceil(HW/32)).
template <typename T>
__device__ __forceinline__ void lift_prepare_cta(const Dims &d, const T *__restrict__ dn, float *__restrict__ prob,
                                                 float *__restrict__ ctx_t, float *__restrict__ prob_col, int cta, float *smem) {
    __shared__ float s_red[LIFT_WARPS][LIFT_PX];
    const int chunks = (d.HW + LIFT_PX - 1) / LIFT_PX;
    const int bn = cta / chunks;
    const int hw0 = (cta - bn * chunks) * LIFT_PX;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int hw = hw0 + lane;
    const bool live = hw < d.HW;
    const int DC = d.D + d.C;
    const T *src = dn + (size_t)bn * DC * d.HW + hw;
    constexpr int S = LIFT_PX + 1;

    for (int r = warp; r < DC; r += LIFT_WARPS) smem[r * S + lane] = live ? lift_load(src + (size_t)r * d.HW) : 0.f;
    __syncthreads();

    // ---- softmax over depth (models.py:50,58): max, exp, sum, normalise -- one exp per element
    float m = -INFINITY;
    for (int dd = warp; dd < d.D; dd += LIFT_WARPS) m = fmaxf(m, smem[dd * S + lane]);
    s_red[warp][lane] = m;
    __syncthreads();
#pragma unroll
    for (int w = 0; w < LIFT_WARPS; ++w) m = fmaxf(m, s_red[w][lane]);
    __syncthreads();
    float sum = 0.f;
    for (int dd = warp; dd < d.D; dd += LIFT_WARPS) {
        const float e = expf(smem[dd * S + lane] - m);
        smem[dd * S + lane] = e;
        sum += e;
    }
    s_red[warp][lane] = sum;
    __syncthreads();
    sum = 0.f;
#pragma unroll
    for (int w = 0; w < LIFT_WARPS; ++w) sum += s_red[w][lane];
    if (live) {
        float *pdst = prob + (size_t)bn * d.D * d.HW + hw;
        // optional second copy, camera-column major [bn][w][D][fH]: the operand block the forward gather stages
        const int h = hw / d.fW, w = hw - h * d.fW;
        float *cdst = prob_col ? prob_col + ((size_t)(bn * d.fW + w) * d.D) * d.fH + h : nullptr;
        for (int dd = warp; dd < d.D; dd += LIFT_WARPS) {
            const float pv = smem[dd * S + lane] / sum;
            pdst[(size_t)dd * d.HW] = pv;
            if (cdst) cdst[dd * d.fH] = pv;
        }
    }

    // ---- context transpose [C][HW] -> [HW][C]
    const float *ct = smem + d.D * S;
    float *cdst = ctx_t + ((size_t)bn * d.HW + hw0) * d.C;
    const int npx = min(LIFT_PX, d.HW - hw0);
    for (int px = warp; px < npx; px += LIFT_WARPS)
        for (int c = lane; c < d.C; c += 32) cdst[(size_t)px * d.C + c] = ct[c * S + px];
}

