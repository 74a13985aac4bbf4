// Operator-level drop-ins for tools.QuickCumsum / tools.cumsum_trick (src/tools.py:182-219) and library
// introspection.
//
// The reference computes per-run sums as differences of a GLOBAL prefix sum (tools.py:196-201), which
// carries cancellation error proportional to the prefix magnitude.  Here every run of
// equal rank is summed on its own, sequentially in input order -- same value within the north_star
// tolerance, closer to the exact sum, and independent of what precedes the run.
#include "common.cuh"

#define QC_BLOCK 1024

// flag[i] = 1 where a new run starts (i > 0 and ranks[i] != ranks[i-1]); per-block flag counts
__global__ void __launch_bounds__(QC_BLOCK)
k_qc_block_counts(int64_t n, const long long *__restrict__ ranks, int32_t *__restrict__ block_counts) {
    const int64_t i = (int64_t)blockIdx.x * QC_BLOCK + threadIdx.x;
    const int f = (i > 0 && i < n && ranks[i] != ranks[i - 1]) ? 1 : 0;
    const int c = __syncthreads_count(f);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = c;
}

// run_id[i] = number of run starts in (0, i]; run_start[r] = first element of run r
__global__ void __launch_bounds__(QC_BLOCK)
k_qc_run_ids(int64_t n, const long long *__restrict__ ranks, const int32_t *__restrict__ block_offsets,
             int32_t *__restrict__ run_id, int32_t *__restrict__ run_start, int32_t *__restrict__ n_runs) {
    __shared__ int s_warp[32];
    const int64_t i = (int64_t)blockIdx.x * QC_BLOCK + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = (i > 0 && i < n && ranks[i] != ranks[i - 1]) ? 1 : 0;
    int inc = f;
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(LSS_FULL_MASK, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < warp; ++w) woff += s_warp[w];
    const int id = block_offsets[blockIdx.x] + woff + inc;
    if (i < n) {
        run_id[i] = id;
        if (f || i == 0) run_start[id] = (int32_t)i;
        if (i == n - 1) { *n_runs = id + 1; run_start[id + 1] = (int32_t)n; }
    }
}

template <int KC>
__global__ void __launch_bounds__(256)
k_qc_fwd(int64_t n, int C, const float *__restrict__ x, int64_t xrs, const long long *__restrict__ geom,
         const int32_t *__restrict__ run_start, int n_runs, float *__restrict__ sums, long long *__restrict__ geom_out) {
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= n_runs) return;
    const int lane = threadIdx.x & 31;
    const int s = __ldg(run_start + r), e = __ldg(run_start + r + 1);
    float acc[KC];
#pragma unroll
    for (int k = 0; k < KC; ++k) acc[k] = 0.f;
#pragma unroll 4
    for (int i = s; i < e; ++i) {
        const float *row = x + (size_t)i * xrs;
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            const int c = lane + 32 * k;
            if (c < C) acc[k] = __fadd_rn(acc[k], __ldg(row + c));
        }
    }
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        const int c = lane + 32 * k;
        if (c < C) sums[(size_t)r * C + c] = acc[k];
    }
    if (geom_out && lane < 4) geom_out[(size_t)r * 4 + lane] = geom[(size_t)(e - 1) * 4 + lane];   // last point of the run (tools.py:200)
}

__global__ void __launch_bounds__(256)
k_qc_bwd(int64_t n, int C, const float *__restrict__ gsums, const int32_t *__restrict__ run_id, float *__restrict__ gx) {
    const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= n) return;
    const int lane = threadIdx.x & 31;
    const float *row = gsums + (size_t)__ldg(run_id + i) * C;
    for (int c = lane; c < C; c += 32) gx[(size_t)i * C + c] = __ldg(row + c);
}

extern "C" size_t lss_quickcumsum_scratch_elems(int64_t n) {
    if (n < 0) n = 0;
    const size_t nb = (size_t)((n + QC_BLOCK - 1) / QC_BLOCK);
    return nb + 2 + (size_t)n + 2;   // [block offsets: nb+1][pad][run_start: n+1]
}

extern "C" int lss_quickcumsum_runs(int64_t n, const int64_t *ranks, int32_t *run_id, int32_t *n_runs,
                                    int32_t *scratch, void *stream) {
    LSS_REQUIRE(n >= 0 && n < ((int64_t)1 << 31), LSS_ERR_UNSUPPORTED);
    LSS_REQUIRE(n_runs && scratch, LSS_ERR_BAD_ARG);
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) return cudaMemsetAsync(n_runs, 0, 4, s) == cudaSuccess ? LSS_OK : LSS_ERR_CUDA;
    LSS_REQUIRE(ranks && run_id, LSS_ERR_BAD_ARG);
    const int nb = (int)((n + QC_BLOCK - 1) / QC_BLOCK);
    int32_t *block_off = scratch;
    int32_t *run_start = scratch + nb + 2;
    k_qc_block_counts<<<nb, QC_BLOCK, 0, s>>>(n, (const long long *)ranks, block_off);
    k_scan_single<<<1, 1024, 0, s>>>(block_off, nb, nullptr);
    k_qc_run_ids<<<nb, QC_BLOCK, 0, s>>>(n, (const long long *)ranks, block_off, run_id, run_start, n_runs);
    LSS_CHECK_LAUNCH();
    return LSS_OK;
}

// `scratch` is the buffer filled by lss_quickcumsum_runs (it holds the run offsets).
extern "C" int lss_quickcumsum_fwd(int64_t n, int32_t C, const float *x, int64_t x_row_stride,
                                           const int64_t *geom_feats, const int32_t *scratch, int32_t n_runs,
                                           float *sums, int64_t *geom_out, void *stream) {
    LSS_REQUIRE(n >= 0 && C > 0 && n_runs >= 0, LSS_ERR_BAD_ARG);
    LSS_REQUIRE(C <= LSS_MAX_CHANNELS, LSS_ERR_UNSUPPORTED);
    if (n == 0 || n_runs == 0) return LSS_OK;
    LSS_REQUIRE(x && sums && scratch, LSS_ERR_BAD_ARG);
    LSS_REQUIRE(geom_out == nullptr || geom_feats != nullptr, LSS_ERR_BAD_ARG);
    const int nb = (int)((n + QC_BLOCK - 1) / QC_BLOCK);
    const int32_t *run_start = scratch + nb + 2;
    cudaStream_t s = (cudaStream_t)stream;
    const int grid = (n_runs + 7) / 8;
    const int kc = lss_kc_for(C);
#define QC_CASE(K) k_qc_fwd<K><<<grid, 256, 0, s>>>(n, C, x, x_row_stride, (const long long *)geom_feats, run_start, n_runs, sums, (long long *)geom_out)
    if (kc <= 1) QC_CASE(1); else if (kc <= 2) QC_CASE(2); else if (kc <= 4) QC_CASE(4); else QC_CASE(8);
#undef QC_CASE
    LSS_CHECK_LAUNCH();
    return LSS_OK;
}

extern "C" int lss_quickcumsum_bwd(int64_t n, int32_t C, const float *grad_sums, const int32_t *run_id, float *grad_x,
                                   void *stream) {
    LSS_REQUIRE(n >= 0 && C > 0, LSS_ERR_BAD_ARG);
    if (n == 0) return LSS_OK;
    LSS_REQUIRE(grad_sums && run_id && grad_x, LSS_ERR_BAD_ARG);
    k_qc_bwd<<<(unsigned)((n + 7) / 8), 256, 0, (cudaStream_t)stream>>>(n, C, grad_sums, run_id, grad_x);
    LSS_CHECK_LAUNCH();
    return LSS_OK;
}

// ------------------------------------------------------------------------------------------------
// introspection
// ------------------------------------------------------------------------------------------------

extern "C" int lss_version(void) { return LSS_B200_VERSION; }

extern "C" const char *lss_status_string(int status) {
    switch (status) {
        case LSS_OK: return "ok";
        case LSS_ERR_BAD_ARG: return "bad argument (null pointer, non-positive dimension or inconsistent sizes)";
        case LSS_ERR_ALIGN: return "misaligned pointer";
        case LSS_ERR_UNSUPPORTED: return "dimensions outside the compiled limits";
        case LSS_ERR_CUDA: return "CUDA runtime error (launch failed or no device)";
        case LSS_ERR_WORKSPACE: return "workspace missing or sized for a different problem";
        default: return "unknown status";
    }
}

extern "C" void lss_get_limits(lss_limits *out) {
    if (!out) return;
    out->max_points_per_sample = (int32_t)(LSS_PIDX_MASK + 1);
    out->max_tile_cols = LSS_MAX_TILE_COLS;
    out->max_depth_bins = LSS_MAX_DEPTH;
    out->max_channels = LSS_MAX_CHANNELS;
}

// ------------------------------------------------------------------------------------------------
// host-buffer pipeline helpers (api.StepPipeline): a whole stage -- waits, copies, record -- per call, so that the
// host side of a step is a handful of foreign-function calls instead of a dozen framework calls
// ------------------------------------------------------------------------------------------------

extern "C" void *lss_pipe_event_create(void) {
    cudaEvent_t e = nullptr;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    return (void *)e;
}

extern "C" int lss_pipe_event_destroy(void *event) {
    return (event == nullptr || cudaEventDestroy((cudaEvent_t)event) == cudaSuccess) ? LSS_OK : LSS_ERR_CUDA;
}

extern "C" int lss_pipe_event_synchronize(void *event) {
    LSS_REQUIRE(event != nullptr, LSS_ERR_BAD_ARG);
    return cudaEventSynchronize((cudaEvent_t)event) == cudaSuccess ? LSS_OK : LSS_ERR_CUDA;
}

extern "C" int lss_pipe_stage(void *stream, void *wait_a, void *wait_b, int32_t n_copies, void *const *dst,
                              const void *const *src, const size_t *bytes, void *record) {
    LSS_REQUIRE(n_copies >= 0 && n_copies <= 4 && (n_copies == 0 || (dst && src && bytes)), LSS_ERR_BAD_ARG);
    cudaStream_t s = (cudaStream_t)stream;
    if (wait_a && cudaStreamWaitEvent(s, (cudaEvent_t)wait_a, 0) != cudaSuccess) return LSS_ERR_CUDA;
    if (wait_b && cudaStreamWaitEvent(s, (cudaEvent_t)wait_b, 0) != cudaSuccess) return LSS_ERR_CUDA;
    for (int i = 0; i < n_copies; ++i)
        if (cudaMemcpyAsync(dst[i], src[i], bytes[i], cudaMemcpyDefault, s) != cudaSuccess) return LSS_ERR_CUDA;
    if (record && cudaEventRecord((cudaEvent_t)record, s) != cudaSuccess) return LSS_ERR_CUDA;
    return LSS_OK;
}

extern "C" int lss_pipe_step(void *copy_in_stream, void *compute_stream, void *copy_out_stream, void *graph_exec,
                             void *ev_in, void *ev_compute, void *ev_done, void *in_dev, const void *in_host, size_t in_bytes,
                             void *out_host, const void *out_dev, size_t out_bytes) {
    LSS_REQUIRE(graph_exec && ev_in && ev_compute && ev_done && in_dev && in_host && out_host && out_dev, LSS_ERR_BAD_ARG);
    cudaStream_t ci = (cudaStream_t)copy_in_stream, cs = (cudaStream_t)compute_stream, co = (cudaStream_t)copy_out_stream;
    cudaEvent_t e_in = (cudaEvent_t)ev_in, e_c = (cudaEvent_t)ev_compute, e_done = (cudaEvent_t)ev_done;
    bool ok = cudaStreamWaitEvent(ci, e_c, 0) == cudaSuccess
           && cudaMemcpyAsync(in_dev, in_host, in_bytes, cudaMemcpyHostToDevice, ci) == cudaSuccess
           && cudaEventRecord(e_in, ci) == cudaSuccess;
    ok = ok && cudaStreamWaitEvent(cs, e_in, 0) == cudaSuccess && cudaStreamWaitEvent(cs, e_done, 0) == cudaSuccess
            && cudaGraphLaunch((cudaGraphExec_t)graph_exec, cs) == cudaSuccess && cudaEventRecord(e_c, cs) == cudaSuccess;
    ok = ok && cudaStreamWaitEvent(co, e_c, 0) == cudaSuccess
            && cudaMemcpyAsync(out_host, out_dev, out_bytes, cudaMemcpyDeviceToHost, co) == cudaSuccess
            && cudaEventRecord(e_done, co) == cudaSuccess;
    return ok ? LSS_OK : LSS_ERR_CUDA;
}

// ------------------------------------------------------------------------------------------------
// process-wide options (the only global state of the library)
// ------------------------------------------------------------------------------------------------

static int g_lss_options[LSS_OPT_COUNT] = {1};      // LSS_OPT_PDL: on

int lss_option_value(int option) { return option >= 0 && option < LSS_OPT_COUNT ? g_lss_options[option] : 0; }

extern "C" int lss_set_option(int option, int value) {
    LSS_REQUIRE(option >= 0 && option < LSS_OPT_COUNT, LSS_ERR_BAD_ARG);
    g_lss_options[option] = value;
    return LSS_OK;
}

extern "C" int lss_get_option(int option) { return lss_option_value(option); }
