// Shared device/host helpers for liblss_b200 (sm_100a).  See include/lss_b200.h for the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <math.h>

#include "../../include/lss_b200.h"

#define LSS_FULL_MASK 0xffffffffu
#define LSS_PIDX_BITS 20                     // point-in-sample bits of a bucket entry
#define LSS_PIDX_MASK ((1u << LSS_PIDX_BITS) - 1u)
#define LSS_MAX_TILE_COLS 4096               // 12 bits of column per entry
#define LSS_MAX_DEPTH 256                    // fused backward keeps per-depth state in registers
#define LSS_MAX_CHANNELS 256
#define LSS_LONG_VOXEL 64                    // voxels with at least this many points are summed by a whole CTA

#define LSS_CHECK_LAUNCH()                                                \
    do {                                                                  \
        cudaError_t e__ = cudaGetLastError();                             \
        if (e__ != cudaSuccess) return LSS_ERR_CUDA;                      \
    } while (0)

#define LSS_REQUIRE(cond, code) \
    do {                        \
        if (!(cond)) return (code); \
    } while (0)

int lss_option_value(int option);     // ops.cu: process-wide options of lss_set_option

// Kernel launch with optional programmatic dependent launch (PDL): with `pdl`, the grid may start while the
// previous kernel of the stream is still draining; the kernel must execute lss_pdl_wait() before it reads anything that
// kernel wrote, and producers call lss_pdl_trigger() early.  Captured into CUDA graphs as programmatic edges.
template <typename... KArgs, typename... Args>
static inline cudaError_t lss_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl,
                                     Args... args) {
    const bool no_pdl = lss_option_value(LSS_OPT_PDL) == 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl && !no_pdl) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#define lss_pdl_wait() asm volatile("griddepcontrol.wait;" ::: "memory")
#define lss_pdl_trigger() asm volatile("griddepcontrol.launch_dependents;")

static inline bool lss_aligned(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

static inline int lss_check_problem(const lss_problem *p) {
    if (!p) return LSS_ERR_BAD_ARG;
    if (p->B <= 0 || p->N <= 0 || p->D <= 0 || p->fH <= 0 || p->fW <= 0 || p->C <= 0) return LSS_ERR_BAD_ARG;
    if (p->nx <= 0 || p->ny <= 0 || p->nz <= 0) return LSS_ERR_BAD_ARG;
    for (int k = 0; k < 3; ++k)
        if (!(p->dx[k] > 0.f)) return LSS_ERR_BAD_ARG;
    const int64_t P = (int64_t)p->N * p->D * p->fH * p->fW;
    if (P > (int64_t)LSS_PIDX_MASK + 1) return LSS_ERR_UNSUPPORTED;
    if ((int64_t)p->B * P >= (int64_t)1 << 31) return LSS_ERR_UNSUPPORTED;
    if ((int64_t)p->B * p->nx * p->ny * p->nz >= (int64_t)1 << 31) return LSS_ERR_UNSUPPORTED;
    if (p->C > LSS_MAX_CHANNELS) return LSS_ERR_UNSUPPORTED;
    return LSS_OK;
}

// Device copy of the problem with the derived sizes every kernel needs.
struct Dims {
    int B, N, D, fH, fW, C;
    int nx, ny, nz;
    int HW;        // fH*fW
    int DHW;       // D*fH*fW
    int P;         // points per sample N*D*fH*fW
    int n_points;  // B*P
    float dx[3], lo[3];
    float inv_dx[3];   // 1/dx where dx is a power of two (x / dx == x * (1/dx) bit for bit), else 0: true division
    unsigned long long mDHW, mHW, mfW;   // ceil(2^40 / DHW), ceil(2^40 / HW), ceil(2^40 / fW): exact division of a 20-bit point index
};

// x / divisor for x < 2^20, divisor < 2^20, with m = ceil(2^40 / divisor)
__host__ __device__ __forceinline__ unsigned lss_div20(unsigned x, unsigned long long m) {
    return (unsigned)(((unsigned long long)x * m) >> 40);
}

static inline Dims make_dims(const lss_problem *p) {
    Dims d;
    d.B = p->B; d.N = p->N; d.D = p->D; d.fH = p->fH; d.fW = p->fW; d.C = p->C;
    d.nx = p->nx; d.ny = p->ny; d.nz = p->nz;
    d.HW = p->fH * p->fW;
    d.DHW = p->D * d.HW;
    d.P = p->N * d.DHW;
    d.n_points = p->B * d.P;
    for (int k = 0; k < 3; ++k) {
        d.dx[k] = p->dx[k]; d.lo[k] = p->lo[k];
        int e = 0;
        const float m = frexpf(p->dx[k], &e);          // dx = m * 2^e, m in [0.5, 1)
        d.inv_dx[k] = (m == 0.5f && e > -100 && e < 100) ? 1.0f / p->dx[k] : 0.0f;
    }
    d.mDHW = ((1ull << 40) + (unsigned)d.DHW - 1) / (unsigned)d.DHW;
    d.mHW = ((1ull << 40) + (unsigned)d.HW - 1) / (unsigned)d.HW;
    d.mfW = ((1ull << 40) + (unsigned)d.fW - 1) / (unsigned)d.fW;
    return d;
}

// Tiling of the BEV grid: one CTA tile = (b, iz, ix, TY consecutive iy).
struct Tiling {
    int TY;       // columns per tile
    int nty;      // tiles per (b,iz,ix) row
    int n_tiles;
};

__host__ __device__ __forceinline__ int lss_kc_for(int C) { return (C + 31) / 32; }

// Camera-column-major position of the point-in-sample index pidx = ((n*D + d)*fH + h)*fW + w:
// ((n*fW + w)*D + d)*fH + h.  All points of one camera column (n, w) are contiguous, depth-major.
__device__ __forceinline__ unsigned lss_column_major(const Dims &d, unsigned pidx) {
    const unsigned cam = lss_div20(pidx, d.mDHW);
    const unsigned r = pidx - cam * d.DHW;
    const unsigned dd = lss_div20(r, d.mHW);
    const unsigned hw = r - dd * d.HW;
    const unsigned h = lss_div20(hw, d.mfW), w = hw - h * d.fW;
    return ((cam * d.fW + w) * d.D + dd) * d.fH + h;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(LSS_FULL_MASK, v, o);
    return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(LSS_FULL_MASK, v, o));
    return v;
}

// dense voxel id -> element offset of its C-channel row in a BEV-shaped tensor
//   rows workspace / NCHW-transposed rows:   v * C
//   channels_last BEV [B, nx, ny, nz*C]:      ((b*nx+ix)*ny+iy)*(nz*C) + iz*C
__device__ __forceinline__ size_t voxel_row_offset_cl(int v, const Dims &d) {
    int iy = v % d.ny; int t = v / d.ny;
    int ix = t % d.nx; t /= d.nx;
    int iz = t % d.nz; int b = t / d.nz;
    return ((size_t)((b * d.nx + ix) * (size_t)d.ny + iy)) * (size_t)(d.nz * d.C) + (size_t)iz * d.C;
}

// Pixel-owner backward gather over channel-contiguous gradient rows (defined in splat.cu, shared with runplan.cu):
// `prow` int32[B,N,fW,D,fH] = row of the point's voxel or -1, `rows` = base of the rows (row r starts at element r*C) --
// the compact rows of a sorted tile plan, or a channels_last BEV gradient itself (run plan).  C in {32, 64, 128}, fH <= 32.
int lss_bwd_gather_rows(const Dims &d, const int32_t *prow, const float *prob_col, const float *ctx_t, const float *rows,
                        float *grad_dn, int b0, int b1, bool pdl, cudaStream_t s);

// Single-CTA exclusive scan of a[0..n) in place, a[n] = total (cold paths only: parity dump, run offsets).
static __global__ void __launch_bounds__(1024) k_scan_single(int32_t *a, int n, int32_t *total) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const int x = i < n ? a[i] : 0;
        int inc = x;
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(LSS_FULL_MASK, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < warp; ++w) woff += s_warp[w];
        const int excl = s_carry + woff + inc - x;
        if (i < n) a[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) { a[n] = s_carry; if (total) *total = s_carry; }
}

// Bitonic sort of a[0..n) ascending by the whole CTA (shared or global memory; keys unique): all comparators point the same way ("flip" bitonic
// network), so the virtual +inf padding beyond n never moves and n need not be a power of two.
// Index arithmetic uses shifts only (k, j are powers of two).
template <typename Keys>
__device__ __forceinline__ void bitonic_sort_block(Keys a, int n) {
    int m = 1, lm = 0;
    while (m < n) { m <<= 1; ++lm; }
    const int half = m >> 1;
    for (int lk = 1; lk <= lm; ++lk) {
        const int k = 1 << lk, hk = k >> 1;
        for (int q = threadIdx.x; q < half; q += blockDim.x) {   // flip stage: i <-> i ^ (k-1)
            const int i = ((q >> (lk - 1)) << lk) | (q & (hk - 1));
            const int l = i ^ (k - 1);
            if (l < n) { const uint32_t x = a[i], y = a[l]; if (x > y) { a[i] = y; a[l] = x; } }
        }
        __syncthreads();
        for (int lj = lk - 2; lj >= 0; --lj) {
            const int j = 1 << lj;
            for (int q = threadIdx.x; q < half; q += blockDim.x) {
                const int i = ((q >> lj) << (lj + 1)) | (q & (j - 1));
                const int l = i + j;
                if (l < n) { const uint32_t x = a[i], y = a[l]; if (x > y) { a[i] = y; a[l] = x; } }
            }
            __syncthreads();
        }
    }
}

