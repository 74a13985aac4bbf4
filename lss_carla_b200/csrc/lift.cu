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
#include "lift.cuh"

template <typename T>
__global__ void __launch_bounds__(LIFT_THREADS)
k_lift_prepare(Dims d, const T *__restrict__ dn, float *__restrict__ prob, float *__restrict__ ctx_t,
               float *__restrict__ prob_col) {
    extern __shared__ float smem[];                 // [D+C][33]
    lift_prepare_cta<T>(d, dn, prob, ctx_t, prob_col, (int)blockIdx.x, smem);
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
