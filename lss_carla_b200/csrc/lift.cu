// Lift operand preparation: depth softmax + pixel-major context.
//
// Replaces CamEncode.get_depth_dist and the operand half of get_depth_feat (src/models.py:49-61) and the
// view/permute of get_cam_feats (src/models.py:199-200).  The outer product depth (x) context itself is
// never formed: the splat kernels multiply prob[p] * ctx_t[pixel(p), c] on the fly.
//
//   depthnet_out f32[B*N, D+C, fH, fW]  ->  prob  f32[B*N, D, fH*fW]   softmax over D per pixel
//                                            ctx_t f32[B*N, fH*fW, C]   context transposed to pixel-major
#include "common.cuh"

#define LIFT_PX 32        // pixels per CTA (one 128-byte line of every channel row)
#define LIFT_THREADS 256  // 8 warps: warp w owns depth bins / channels w, w+8, ...

__global__ void __launch_bounds__(LIFT_THREADS)
k_lift_prepare(Dims d, const float *__restrict__ dn, float *__restrict__ prob, float *__restrict__ ctx_t) {
    extern __shared__ float smem[];          // [C][33] transpose tile, then reused: [8][32] reductions
    __shared__ float s_red[8][LIFT_PX];
    const int chunks = (d.HW + LIFT_PX - 1) / LIFT_PX;
    const int bn = blockIdx.x / chunks;
    const int hw0 = (blockIdx.x % chunks) * LIFT_PX;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int hw = hw0 + lane;
    const bool live = hw < d.HW;
    const float *src = dn + (size_t)bn * (d.D + d.C) * d.HW;

    // ---- softmax over depth (models.py:50,58): max, sum of exp, normalise
    float m = -INFINITY;
    for (int dd = warp; dd < d.D; dd += 8)
        if (live) m = fmaxf(m, __ldg(src + (size_t)dd * d.HW + hw));
    s_red[warp][lane] = m;
    __syncthreads();
#pragma unroll
    for (int w = 0; w < 8; ++w) m = fmaxf(m, s_red[w][lane]);
    __syncthreads();
    float sum = 0.f;
    for (int dd = warp; dd < d.D; dd += 8)
        if (live) sum += expf(__ldg(src + (size_t)dd * d.HW + hw) - m);
    s_red[warp][lane] = sum;
    __syncthreads();
    sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += s_red[w][lane];
    float *pdst = prob + (size_t)bn * d.D * d.HW;
    for (int dd = warp; dd < d.D; dd += 8)
        if (live) pdst[(size_t)dd * d.HW + hw] = expf(__ldg(src + (size_t)dd * d.HW + hw) - m) / sum;

    // ---- context transpose [C][HW] -> [HW][C] through shared memory
    const float *csrc = src + (size_t)d.D * d.HW;
    for (int c = warp; c < d.C; c += 8) smem[c * (LIFT_PX + 1) + lane] = live ? __ldg(csrc + (size_t)c * d.HW + hw) : 0.f;
    __syncthreads();
    float *cdst = ctx_t + ((size_t)bn * d.HW + hw0) * d.C;
    const int npx = min(LIFT_PX, d.HW - hw0);
    for (int px = warp; px < npx; px += 8)
        for (int c = lane; c < d.C; c += 32) cdst[(size_t)px * d.C + c] = smem[c * (LIFT_PX + 1) + px];
}

extern "C" int lss_lift_prepare(const lss_problem *p, const float *depthnet_out, float *prob, float *ctx_t,
                                void *stream) {
    int st = lss_check_problem(p);
    if (st != LSS_OK) return st;
    LSS_REQUIRE(depthnet_out && prob && ctx_t, LSS_ERR_BAD_ARG);
    const Dims d = make_dims(p);
    const int chunks = (d.HW + LIFT_PX - 1) / LIFT_PX;
    const size_t smem = (size_t)d.C * (LIFT_PX + 1) * sizeof(float);
    k_lift_prepare<<<d.B * d.N * chunks, LIFT_THREADS, smem, (cudaStream_t)stream>>>(d, depthnet_out, prob, ctx_t);
    LSS_CHECK_LAUNCH();
    return LSS_OK;
}
