// Lift operand preparation: depth softmax + pixel-major context.
//
// Replaces CamEncode.get_depth_dist and the operand half of get_depth_feat (src/models.py:49-61) and the
// view/permute of get_cam_feats (src/models.py:199-200).  The outer product depth (x) context itself is
// never formed: the splat kernels multiply prob[p] * ctx_t[pixel(p), c] on the fly.
//
//   depthnet_out f32[B*N, D+C, fH, fW]  ->  prob  f32[B*N, D, fH*fW]   softmax over D per pixel
//                                            ctx_t f32[B*N, fH*fW, C]   context transposed to pixel-major
//
// One CTA owns 32 consecutive pixels of one camera: all D+C channel rows of those pixels are read ONCE
// (128-byte lines) into shared memory, the softmax runs out of shared memory, and the context tile is
// written back transposed.
// The depthnet output may also arrive as bfloat16 (autocast training): it is widened on load, everything after the
// load -- softmax, the splat, the accumulation -- is the float32 path bit for bit (lss_lift_prepare_bf16).
#include <cuda_bf16.h>

#include "common.cuh"

#define LIFT_PX 32        // pixels per CTA (one 128-byte line of every channel row)
#define LIFT_THREADS 256  // 8 warps: warp w owns depth bins / channels w, w+8, ...
#define LIFT_WARPS 8

__device__ __forceinline__ float lift_load(const float *p) { return __ldg(p); }
__device__ __forceinline__ float lift_load(const __nv_bfloat16 *p) { return __bfloat162float(__ldg(p)); }

template <typename T>
__global__ void __launch_bounds__(LIFT_THREADS)
k_lift_prepare(Dims d, const T *__restrict__ dn, float *__restrict__ prob, float *__restrict__ ctx_t,
               float *__restrict__ prob_col) {
    extern __shared__ float smem[];                 // [D+C][33]
    __shared__ float s_red[LIFT_WARPS][LIFT_PX];
    const int chunks = (d.HW + LIFT_PX - 1) / LIFT_PX;
    const int bn = blockIdx.x / chunks;
    const int hw0 = (blockIdx.x - bn * chunks) * LIFT_PX;
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

template <typename T>
static int lift_prepare_impl(const lss_problem *p, const T *depthnet_out, float *prob, float *ctx_t, float *prob_col, void *stream) {
    int st = lss_check_problem(p);
    if (st != LSS_OK) return st;
    LSS_REQUIRE(depthnet_out && prob && ctx_t, LSS_ERR_BAD_ARG);
    const Dims d = make_dims(p);
    const int chunks = (d.HW + LIFT_PX - 1) / LIFT_PX;
    const size_t smem = (size_t)(d.D + d.C) * (LIFT_PX + 1) * sizeof(float);
    if (smem > 227 * 1024) return LSS_ERR_UNSUPPORTED;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(k_lift_prepare<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return LSS_ERR_CUDA;
    k_lift_prepare<T><<<d.B * d.N * chunks, LIFT_THREADS, smem, (cudaStream_t)stream>>>(d, depthnet_out, prob, ctx_t, prob_col);
    LSS_CHECK_LAUNCH();
    return LSS_OK;
}

extern "C" int lss_lift_prepare(const lss_problem *p, const float *depthnet_out, float *prob, float *ctx_t,
                                float *prob_col, void *stream) {
    return lift_prepare_impl<float>(p, depthnet_out, prob, ctx_t, prob_col, stream);
}

extern "C" int lss_lift_prepare_bf16(const lss_problem *p, const void *depthnet_out_bf16, float *prob, float *ctx_t,
                                     float *prob_col, void *stream) {
    return lift_prepare_impl<__nv_bfloat16>(p, (const __nv_bfloat16 *)depthnet_out_bf16, prob, ctx_t, prob_col, stream);
}
