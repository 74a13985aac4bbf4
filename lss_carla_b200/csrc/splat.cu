// Splat: voxel pooling of lifted camera features into the BEV grid, forward and backward.
//
// Replaces, for shdragron/LSS-Carla:
//   the outer product of CamEncode.get_depth_feat    src/models.py:59      (never materialised)
//   the value half of LiftSplatShoot.voxel_pooling    src/models.py:209, 222-246
//   QuickCumsum.forward / backward, cumsum_trick      src/tools.py:182-219
//
// Forward is a TILE-OWNER kernel: one CTA owns the BEV tile (b, iz, ix, TY consecutive iy) and reads the
// bucket of points the plan assigned to it (plan.cu).  Every BEV element is written exactly once, zeros
// included, so there is no memset and no global atomic.  In SORTED mode a warp walks one voxel's points in
// ascending flat index and adds float32(prob * ctx) sequentially -- a fixed summation order; in
// SMEM_ATOMIC mode warps walk the (unsorted) bucket and accumulate with shared-memory atomics.
// RED_GLOBAL is the classical pixel-owner red.global.add splat, kept for measurement.
//
// Backward is a pure gather (tools.py:212-219): for NCHW gradients a tile-owner kernel first transposes the
// rows of hit voxels into channel-contiguous rows; a pixel-owner kernel then gathers one row per frustum
// point and fuses the outer-product and softmax backward.
#include "common.cuh"

#define SPLAT_THREADS 256
#define SPLAT_WARPS (SPLAT_THREADS / 32)

struct SrcArgs {
    const float *base;   // LIFT: ctx_t [B*N, HW, C]        DENSE: x with strides s[0..5]
    const float *prob;   // LIFT: prob [B*N, D, HW]         DENSE: unused
    long long s[6];
};

template <bool DENSE>
__device__ __forceinline__ void entry_source(const Dims &d, const SrcArgs &a, int b, uint32_t e, float &w, long long &off) {
    const int pidx = (int)(e & LSS_PIDX_MASK);
    const int n = pidx / d.DHW;
    if (DENSE) {
        int r = pidx - n * d.DHW;
        const int dd = r / d.HW; r -= dd * d.HW;
        const int h = r / d.fW; const int wv = r - h * d.fW;
        off = b * a.s[0] + n * a.s[1] + dd * a.s[2] + h * a.s[3] + wv * a.s[4];
        w = 1.0f;
    } else {
        const int hw = pidx % d.HW;
        off = ((long long)(b * d.N + n) * d.HW + hw) * d.C;
        w = __ldg(a.prob + (size_t)b * d.P + pidx);
    }
}

struct TileCoord { int b, iz, ix, y0, cols; };

__device__ __forceinline__ TileCoord tile_coord(const Dims &d, const Tiling &tl, int tile) {
    TileCoord t;
    const int ty = tile % tl.nty; int r = tile / tl.nty;
    t.ix = r % d.nx; r /= d.nx;
    t.iz = r % d.nz; t.b = r / d.nz;
    t.y0 = ty * tl.TY;
    t.cols = min(tl.TY, d.ny - t.y0);
    return t;
}

// ------------------------------------------------------------------------------------------------
// forward, tile-owner
// ------------------------------------------------------------------------------------------------
// smem (floats): NCHW            acc[C][TYP]   TYP = TY|1 (odd stride: conflict-free column writes)
//                CL + ATOMIC     acc[TY][C]
//                then (SORTED)   colstart[TY+1] as int
template <int KC, bool ATOMIC, bool CL, bool DENSE>
__global__ void __launch_bounds__(SPLAT_THREADS)
k_splat_fwd_tile(Dims d, Tiling tl, const int32_t *__restrict__ tile_start, const uint32_t *__restrict__ entries,
                 SrcArgs src, float *__restrict__ bev) {
    extern __shared__ float smem[];
    const int tile = blockIdx.x;
    const TileCoord tc = tile_coord(d, tl, tile);
    const int s = __ldg(tile_start + tile), n = __ldg(tile_start + tile + 1) - s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int C = d.C;
    const int TYP = tl.TY | 1;
    const size_t plane = (size_t)d.nx * d.ny;
    // NCHW: element (c, y) of this tile;  CL: element (y, c)
    const size_t nchw_base = ((size_t)(tc.b * d.nz + tc.iz) * C) * plane + (size_t)tc.ix * d.ny + tc.y0;
    const size_t cl_base = ((size_t)(tc.b * d.nx + tc.ix) * d.ny + tc.y0) * (size_t)(d.nz * C) + (size_t)tc.iz * C;
    const size_t cl_stride = (size_t)d.nz * C;
    const long long cs = DENSE ? src.s[5] : 1;

    if (n == 0) {   // nothing lands here: stream zeros
        if (CL) {
            for (int i = threadIdx.x; i < tc.cols * C; i += SPLAT_THREADS) bev[cl_base + (size_t)(i / C) * cl_stride + (i % C)] = 0.f;
        } else {
            for (int c = warp; c < C; c += SPLAT_WARPS)
                for (int y = lane; y < tc.cols; y += 32) bev[nchw_base + (size_t)c * plane + y] = 0.f;
        }
        return;
    }

    float *acc_tile = smem;
    const uint32_t *ent = entries + s;

    if (!ATOMIC) {
        int *colstart = (int *)(smem + (CL ? 0 : (size_t)C * TYP));
        // first bucket position of every column (bucket is sorted by column)
        for (int i = threadIdx.x; i <= n; i += SPLAT_THREADS) {
            const int c_prev = i == 0 ? -1 : (int)(__ldg(ent + i - 1) >> LSS_PIDX_BITS);
            const int c_cur = i == n ? tl.TY : (int)(__ldg(ent + i) >> LSS_PIDX_BITS);
            for (int c = c_prev + 1; c <= c_cur; ++c) colstart[c] = i;
        }
        __syncthreads();
        for (int col = warp; col < tc.cols; col += SPLAT_WARPS) {
            const int cb = colstart[col], ce = colstart[col + 1];
            float acc[KC];
#pragma unroll
            for (int k = 0; k < KC; ++k) acc[k] = 0.f;
            for (int j0 = cb; j0 < ce; j0 += 32) {
                float w = 0.f; long long off = 0;
                if (j0 + lane < ce) entry_source<DENSE>(d, src, tc.b, __ldg(ent + j0 + lane), w, off);
                const int cnt = min(32, ce - j0);
#pragma unroll 4
                for (int jj = 0; jj < cnt; ++jj) {
                    const float wj = __shfl_sync(LSS_FULL_MASK, w, jj);
                    const long long oj = __shfl_sync(LSS_FULL_MASK, off, jj);
                    const float *row = src.base + oj;
#pragma unroll
                    for (int k = 0; k < KC; ++k) {
                        const int c = lane + 32 * k;
                        if (c < C) acc[k] = __fadd_rn(acc[k], __fmul_rn(wj, __ldg(row + c * cs)));   // ascending point order
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                const int c = lane + 32 * k;
                if (c < C) {
                    if (CL) bev[cl_base + (size_t)col * cl_stride + c] = acc[k];
                    else acc_tile[c * TYP + col] = acc[k];
                }
            }
        }
        if (CL) return;
        __syncthreads();
    } else {
        const int tile_elems = CL ? tl.TY * C : C * TYP;
        for (int i = threadIdx.x; i < tile_elems; i += SPLAT_THREADS) acc_tile[i] = 0.f;
        __syncthreads();
        for (int j0 = warp * 32; j0 < n; j0 += SPLAT_THREADS) {
            float w = 0.f; long long off = 0; int col = 0;
            if (j0 + lane < n) {
                const uint32_t e = __ldg(ent + j0 + lane);
                col = (int)(e >> LSS_PIDX_BITS);
                entry_source<DENSE>(d, src, tc.b, e, w, off);
            }
            const int cnt = min(32, n - j0);
#pragma unroll 4
            for (int jj = 0; jj < cnt; ++jj) {
                const float wj = __shfl_sync(LSS_FULL_MASK, w, jj);
                const long long oj = __shfl_sync(LSS_FULL_MASK, off, jj);
                const int cj = __shfl_sync(LSS_FULL_MASK, col, jj);
                const float *row = src.base + oj;
#pragma unroll
                for (int k = 0; k < KC; ++k) {
                    const int c = lane + 32 * k;
                    if (c < C) atomicAdd(acc_tile + (CL ? cj * C + c : c * TYP + cj), __fmul_rn(wj, __ldg(row + c * cs)));
                }
            }
        }
        __syncthreads();
    }

    // ---- stream the tile out
    if (CL) {
        for (int i = threadIdx.x; i < tc.cols * C; i += SPLAT_THREADS) bev[cl_base + (size_t)(i / C) * cl_stride + (i % C)] = acc_tile[i];
    } else {
        for (int c = warp; c < C; c += SPLAT_WARPS)
            for (int y = lane; y < tc.cols; y += 32) bev[nchw_base + (size_t)c * plane + y] = acc_tile[c * TYP + y];
    }
}

// ------------------------------------------------------------------------------------------------
// forward, pixel-owner red.global (measurement mode): bev must be zero on entry
// ------------------------------------------------------------------------------------------------
template <int KC, bool CL>
__global__ void __launch_bounds__(SPLAT_THREADS)
k_splat_fwd_red(Dims d, const int32_t *__restrict__ vox, const float *__restrict__ prob, const float *__restrict__ ctx_t,
                float *__restrict__ bev) {
    const int pix = blockIdx.x * SPLAT_WARPS + (threadIdx.x >> 5);   // (bn, hw)
    if (pix >= d.B * d.N * d.HW) return;
    const int lane = threadIdx.x & 31;
    const int bn = pix / d.HW, hw = pix - bn * d.HW;
    const size_t plane = (size_t)d.nx * d.ny;
    float ctx[KC];
#pragma unroll
    for (int k = 0; k < KC; ++k) ctx[k] = (lane + 32 * k < d.C) ? __ldg(ctx_t + (size_t)pix * d.C + lane + 32 * k) : 0.f;
    for (int dd = 0; dd < d.D; ++dd) {
        const size_t p = ((size_t)bn * d.D + dd) * d.HW + hw;
        const int v = __ldg(vox + p);
        if (v < 0) continue;
        const float w = __ldg(prob + p);
        const size_t base = CL ? voxel_row_offset_cl(v, d) : (size_t)(v / plane) * d.C * plane + (v % plane);
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            const int c = lane + 32 * k;
            if (c < d.C) atomicAdd(bev + base + (CL ? (size_t)c : (size_t)c * plane), __fmul_rn(w, ctx[k]));
        }
    }
}

// dense-x variant of the red.global mode: one warp per frustum point
template <bool CL>
__global__ void __launch_bounds__(SPLAT_THREADS)
k_vp_fwd_red(Dims d, const int32_t *__restrict__ vox, SrcArgs src, float *__restrict__ bev) {
    const int p = blockIdx.x * SPLAT_WARPS + (threadIdx.x >> 5);
    if (p >= d.n_points) return;
    const int v = __ldg(vox + p);
    if (v < 0) return;
    const int lane = threadIdx.x & 31;
    const int b = p / d.P;
    float w; long long off;
    entry_source<true>(d, src, b, (uint32_t)(p - b * d.P), w, off);
    const size_t plane = (size_t)d.nx * d.ny;
    const size_t base = CL ? voxel_row_offset_cl(v, d) : (size_t)(v / plane) * d.C * plane + (v % plane);
    for (int c = lane; c < d.C; c += 32) atomicAdd(bev + base + (CL ? (size_t)c : (size_t)c * plane), __ldg(src.base + off + c * src.s[5]));
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------

// NCHW gradient -> channel-contiguous rows of the voxels that received points: rows[v, 0:C].
// One CTA per non-empty tile: coalesced row loads into shared memory, transposed 4*C-byte row stores.
__global__ void __launch_bounds__(SPLAT_THREADS)
k_bwd_rows_nchw(Dims d, Tiling tl, const int32_t *__restrict__ tile_start, const uint32_t *__restrict__ entries,
                const float *__restrict__ grad_bev, float *__restrict__ rows) {
    extern __shared__ float smem[];
    const int tile = blockIdx.x;
    const int s = __ldg(tile_start + tile), n = __ldg(tile_start + tile + 1) - s;
    if (n == 0) return;
    const TileCoord tc = tile_coord(d, tl, tile);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int C = d.C, TYP = tl.TY | 1;
    float *g_tile = smem;                       // [C][TYP]
    int *hit = (int *)(smem + (size_t)C * TYP); // [TY]
    const size_t plane = (size_t)d.nx * d.ny;
    const size_t nchw_base = ((size_t)(tc.b * d.nz + tc.iz) * C) * plane + (size_t)tc.ix * d.ny + tc.y0;
    for (int i = threadIdx.x; i < tl.TY; i += SPLAT_THREADS) hit[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += SPLAT_THREADS) hit[__ldg(entries + s + i) >> LSS_PIDX_BITS] = 1;
    for (int c = warp; c < C; c += SPLAT_WARPS)
        for (int y = lane; y < tc.cols; y += 32) g_tile[c * TYP + y] = __ldg(grad_bev + nchw_base + (size_t)c * plane + y);
    __syncthreads();
    const int v0 = ((tc.b * d.nz + tc.iz) * d.nx + tc.ix) * d.ny + tc.y0;
    for (int col = warp; col < tc.cols; col += SPLAT_WARPS) {
        if (!hit[col]) continue;
        float *dst = rows + (size_t)(v0 + col) * C;
        for (int c = lane; c < C; c += 32) dst[c] = g_tile[c * TYP + col];
    }
}

// Pixel-owner gather: one warp per camera pixel (bn, h, w) walks its D frustum points.
//   g      = rows[voxel(p)]                               (QuickCumsum.backward + griddify backward)
//   gp_d   = <g, ctx>            d_ctx += prob_d * g        (outer product backward, models.py:59)
//   d_logit_d = prob_d * (gp_d - sum_d' prob_d' gp_d')      (softmax backward, models.py:50)
template <int KC, int NCH, bool CL>
__global__ void __launch_bounds__(SPLAT_THREADS)
k_bwd_gather(Dims d, const int32_t *__restrict__ vox, const float *__restrict__ prob, const float *__restrict__ ctx_t,
             const float *__restrict__ rows, float *__restrict__ grad_dn) {
    const int pix = blockIdx.x * SPLAT_WARPS + (threadIdx.x >> 5);
    if (pix >= d.B * d.N * d.HW) return;
    const int lane = threadIdx.x & 31;
    const int bn = pix / d.HW, hw = pix - bn * d.HW;
    float ctx[KC], dctx[KC];
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        ctx[k] = (lane + 32 * k < d.C) ? __ldg(ctx_t + (size_t)pix * d.C + lane + 32 * k) : 0.f;
        dctx[k] = 0.f;
    }
    int vv[NCH]; float pr[NCH], gp[NCH];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
        const int dd = ch * 32 + lane;
        vv[ch] = -1; pr[ch] = 0.f; gp[ch] = 0.f;
        if (dd < d.D) {
            const size_t p = ((size_t)bn * d.D + dd) * d.HW + hw;
            vv[ch] = __ldg(vox + p);
            pr[ch] = __ldg(prob + p);
        }
    }
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
        const int cnt = min(32, d.D - ch * 32);
        for (int dl = 0; dl < cnt; ++dl) {
            const int v = __shfl_sync(LSS_FULL_MASK, vv[ch], dl);
            if (v < 0) continue;                                   // warp-uniform
            const float pj = __shfl_sync(LSS_FULL_MASK, pr[ch], dl);
            const float *row = rows + (CL ? voxel_row_offset_cl(v, d) : (size_t)v * d.C);
            float dot = 0.f;
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                const int c = lane + 32 * k;
                const float g = c < d.C ? __ldg(row + c) : 0.f;
                dot = fmaf(g, ctx[k], dot);
                dctx[k] = fmaf(pj, g, dctx[k]);
            }
            dot = warp_sum(dot);
            if (lane == dl) gp[ch] = dot;
        }
    }
    float sdot = 0.f;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) sdot = fmaf(pr[ch], gp[ch], sdot);
    sdot = warp_sum(sdot);
    float *out = grad_dn + (size_t)bn * (d.D + d.C) * d.HW + hw;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
        const int dd = ch * 32 + lane;
        if (dd < d.D) out[(size_t)dd * d.HW] = pr[ch] * (gp[ch] - sdot);
    }
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        const int c = lane + 32 * k;
        if (c < d.C) out[(size_t)(d.D + c) * d.HW] = dctx[k];
    }
}

// operator-level backward of voxel_pooling: grad_x[p, :] = row of p's voxel, or 0
template <bool CL>
__global__ void __launch_bounds__(SPLAT_THREADS)
k_vp_bwd(Dims d, const int32_t *__restrict__ vox, const float *__restrict__ rows, float *__restrict__ grad_x) {
    const int p = blockIdx.x * SPLAT_WARPS + (threadIdx.x >> 5);
    if (p >= d.n_points) return;
    const int lane = threadIdx.x & 31;
    const int v = __ldg(vox + p);
    float *dst = grad_x + (size_t)p * d.C;
    if (v < 0) { for (int c = lane; c < d.C; c += 32) dst[c] = 0.f; return; }
    const float *row = rows + (CL ? voxel_row_offset_cl(v, d) : (size_t)v * d.C);
    for (int c = lane; c < d.C; c += 32) dst[c] = __ldg(row + c);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------

static inline Tiling make_tiling(const lss_plan_layout *L) {
    Tiling t; t.TY = L->tile_cols; t.nty = L->tiles_per_row; t.n_tiles = L->n_tiles; return t;
}

template <int KC, bool ATOMIC, bool CL, bool DENSE>
static int launch_fwd_tile(const Dims &d, const Tiling &tl, const int32_t *tile_start, const uint32_t *entries,
                           const SrcArgs &src, float *bev, cudaStream_t s) {
    const int TYP = tl.TY | 1;
    size_t smem = 0;
    if (ATOMIC) smem = (size_t)(CL ? tl.TY * d.C : d.C * TYP) * 4;
    else smem = (size_t)(CL ? 0 : d.C * TYP) * 4 + (size_t)(tl.TY + 1) * 4;
    auto kern = k_splat_fwd_tile<KC, ATOMIC, CL, DENSE>;
    if (smem > 227 * 1024) return LSS_ERR_UNSUPPORTED;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return LSS_ERR_CUDA;
    kern<<<tl.n_tiles, SPLAT_THREADS, smem, s>>>(d, tl, tile_start, entries, src, bev);
    LSS_CHECK_LAUNCH();
    return LSS_OK;
}

template <int KC, bool DENSE>
static int dispatch_fwd_tile(bool atomic, bool cl, const Dims &d, const Tiling &tl, const int32_t *ts, const uint32_t *en,
                             const SrcArgs &src, float *bev, cudaStream_t s) {
    if (atomic) return cl ? launch_fwd_tile<KC, true, true, DENSE>(d, tl, ts, en, src, bev, s)
                          : launch_fwd_tile<KC, true, false, DENSE>(d, tl, ts, en, src, bev, s);
    return cl ? launch_fwd_tile<KC, false, true, DENSE>(d, tl, ts, en, src, bev, s)
              : launch_fwd_tile<KC, false, false, DENSE>(d, tl, ts, en, src, bev, s);
}

template <bool DENSE>
static int dispatch_fwd_kc(bool atomic, bool cl, const Dims &d, const Tiling &tl, const int32_t *ts, const uint32_t *en,
                           const SrcArgs &src, float *bev, cudaStream_t s) {
    const int kc = lss_kc_for(d.C);
    if (kc <= 1) return dispatch_fwd_tile<1, DENSE>(atomic, cl, d, tl, ts, en, src, bev, s);
    if (kc <= 2) return dispatch_fwd_tile<2, DENSE>(atomic, cl, d, tl, ts, en, src, bev, s);
    if (kc <= 4) return dispatch_fwd_tile<4, DENSE>(atomic, cl, d, tl, ts, en, src, bev, s);
    return dispatch_fwd_tile<8, DENSE>(atomic, cl, d, tl, ts, en, src, bev, s);
}

static size_t bev_elems(const Dims &d) { return (size_t)d.B * d.nz * d.C * d.nx * d.ny; }

extern "C" int lss_splat_fwd(const lss_problem *p, const lss_plan_layout *L, const void *workspace, const float *prob,
                             const float *ctx_t, float *bev, int mode, int layout, void *stream) {
    int st = lss_check_problem(p);
    if (st != LSS_OK) return st;
    LSS_REQUIRE(L && workspace, LSS_ERR_WORKSPACE);
    LSS_REQUIRE(prob && ctx_t && bev, LSS_ERR_BAD_ARG);
    LSS_REQUIRE(layout == LSS_LAYOUT_NCHW || layout == LSS_LAYOUT_CHANNELS_LAST, LSS_ERR_BAD_ARG);
    const Dims d = make_dims(p);
    LSS_REQUIRE(L->n_points == d.n_points, LSS_ERR_WORKSPACE);
    const Tiling tl = make_tiling(L);
    const char *w = (const char *)workspace;
    const int32_t *vox = (const int32_t *)(w + L->off_vox);
    const uint32_t *entries = (const uint32_t *)(w + L->off_entries);
    const int32_t *tile_start = (const int32_t *)(w + L->off_tile_start);
    cudaStream_t s = (cudaStream_t)stream;
    const bool cl = layout == LSS_LAYOUT_CHANNELS_LAST;
    SrcArgs src{};
    src.base = ctx_t; src.prob = prob;
    if (mode == LSS_SPLAT_SORTED || mode == LSS_SPLAT_SMEM_ATOMIC)
        return dispatch_fwd_kc<false>(mode == LSS_SPLAT_SMEM_ATOMIC, cl, d, tl, tile_start, entries, src, bev, s);
    if (mode == LSS_SPLAT_RED_GLOBAL) {
        if (cudaMemsetAsync(bev, 0, bev_elems(d) * 4, s) != cudaSuccess) return LSS_ERR_CUDA;
        const int npix = d.B * d.N * d.HW;
        const int grid = (npix + SPLAT_WARPS - 1) / SPLAT_WARPS;
        const int kc = lss_kc_for(d.C);
#define RED_CASE(K)                                                                                        \
    if (cl) k_splat_fwd_red<K, true><<<grid, SPLAT_THREADS, 0, s>>>(d, vox, prob, ctx_t, bev);             \
    else k_splat_fwd_red<K, false><<<grid, SPLAT_THREADS, 0, s>>>(d, vox, prob, ctx_t, bev)
        if (kc <= 1) { RED_CASE(1); } else if (kc <= 2) { RED_CASE(2); } else if (kc <= 4) { RED_CASE(4); } else { RED_CASE(8); }
#undef RED_CASE
        LSS_CHECK_LAUNCH();
        return LSS_OK;
    }
    return LSS_ERR_BAD_ARG;
}

extern "C" int lss_voxel_pooling_fwd(const lss_problem *p, const lss_plan_layout *L, const void *workspace,
                                     const float *x, const int64_t *xs_host, float *bev, int mode, int layout,
                                     void *stream) {
    int st = lss_check_problem(p);
    if (st != LSS_OK) return st;
    LSS_REQUIRE(L && workspace, LSS_ERR_WORKSPACE);
    LSS_REQUIRE(x && xs_host && bev, LSS_ERR_BAD_ARG);
    LSS_REQUIRE(layout == LSS_LAYOUT_NCHW || layout == LSS_LAYOUT_CHANNELS_LAST, LSS_ERR_BAD_ARG);
    const Dims d = make_dims(p);
    LSS_REQUIRE(L->n_points == d.n_points, LSS_ERR_WORKSPACE);
    const Tiling tl = make_tiling(L);
    const char *w = (const char *)workspace;
    const int32_t *vox = (const int32_t *)(w + L->off_vox);
    const uint32_t *entries = (const uint32_t *)(w + L->off_entries);
    const int32_t *tile_start = (const int32_t *)(w + L->off_tile_start);
    cudaStream_t s = (cudaStream_t)stream;
    const bool cl = layout == LSS_LAYOUT_CHANNELS_LAST;
    SrcArgs src{};
    src.base = x; src.prob = nullptr;
    for (int i = 0; i < 6; ++i) src.s[i] = xs_host[i];
    if (mode == LSS_SPLAT_SORTED || mode == LSS_SPLAT_SMEM_ATOMIC)
        return dispatch_fwd_kc<true>(mode == LSS_SPLAT_SMEM_ATOMIC, cl, d, tl, tile_start, entries, src, bev, s);
    if (mode == LSS_SPLAT_RED_GLOBAL) {
        if (cudaMemsetAsync(bev, 0, bev_elems(d) * 4, s) != cudaSuccess) return LSS_ERR_CUDA;
        const int grid = (d.n_points + SPLAT_WARPS - 1) / SPLAT_WARPS;
        if (cl) k_vp_fwd_red<true><<<grid, SPLAT_THREADS, 0, s>>>(d, vox, src, bev);
        else k_vp_fwd_red<false><<<grid, SPLAT_THREADS, 0, s>>>(d, vox, src, bev);
        LSS_CHECK_LAUNCH();
        return LSS_OK;
    }
    return LSS_ERR_BAD_ARG;
}

static int launch_bwd_rows(const Dims &d, const Tiling &tl, const int32_t *tile_start, const uint32_t *entries,
                           const float *grad_bev, float *rows, cudaStream_t s) {
    const size_t smem = (size_t)d.C * (tl.TY | 1) * 4 + (size_t)tl.TY * 4;
    if (smem > 227 * 1024) return LSS_ERR_UNSUPPORTED;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(k_bwd_rows_nchw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return LSS_ERR_CUDA;
    k_bwd_rows_nchw<<<tl.n_tiles, SPLAT_THREADS, smem, s>>>(d, tl, tile_start, entries, grad_bev, rows);
    LSS_CHECK_LAUNCH();
    return LSS_OK;
}

template <int KC, int NCH>
static void launch_gather(bool cl, int grid, cudaStream_t s, const Dims &d, const int32_t *vox, const float *prob,
                          const float *ctx_t, const float *rows, float *grad_dn) {
    if (cl) k_bwd_gather<KC, NCH, true><<<grid, SPLAT_THREADS, 0, s>>>(d, vox, prob, ctx_t, rows, grad_dn);
    else k_bwd_gather<KC, NCH, false><<<grid, SPLAT_THREADS, 0, s>>>(d, vox, prob, ctx_t, rows, grad_dn);
}

template <int KC>
static void dispatch_gather_nch(bool cl, int grid, cudaStream_t s, const Dims &d, const int32_t *vox, const float *prob,
                                const float *ctx_t, const float *rows, float *grad_dn) {
    const int nch = (d.D + 31) / 32;
    if (nch <= 1) launch_gather<KC, 1>(cl, grid, s, d, vox, prob, ctx_t, rows, grad_dn);
    else if (nch <= 2) launch_gather<KC, 2>(cl, grid, s, d, vox, prob, ctx_t, rows, grad_dn);
    else if (nch <= 4) launch_gather<KC, 4>(cl, grid, s, d, vox, prob, ctx_t, rows, grad_dn);
    else launch_gather<KC, 8>(cl, grid, s, d, vox, prob, ctx_t, rows, grad_dn);
}

extern "C" int lss_splat_bwd(const lss_problem *p, const lss_plan_layout *L, const void *workspace,
                             const float *grad_bev, int layout, const float *prob, const float *ctx_t,
                             float *grad_rows, float *grad_depthnet, void *stream) {
    int st = lss_check_problem(p);
    if (st != LSS_OK) return st;
    LSS_REQUIRE(L && workspace, LSS_ERR_WORKSPACE);
    LSS_REQUIRE(grad_bev && prob && ctx_t && grad_depthnet, LSS_ERR_BAD_ARG);
    LSS_REQUIRE(layout == LSS_LAYOUT_NCHW || layout == LSS_LAYOUT_CHANNELS_LAST, LSS_ERR_BAD_ARG);
    LSS_REQUIRE(p->D <= LSS_MAX_DEPTH, LSS_ERR_UNSUPPORTED);
    const Dims d = make_dims(p);
    LSS_REQUIRE(L->n_points == d.n_points, LSS_ERR_WORKSPACE);
    const Tiling tl = make_tiling(L);
    const char *w = (const char *)workspace;
    const int32_t *vox = (const int32_t *)(w + L->off_vox);
    const uint32_t *entries = (const uint32_t *)(w + L->off_entries);
    const int32_t *tile_start = (const int32_t *)(w + L->off_tile_start);
    cudaStream_t s = (cudaStream_t)stream;
    const bool cl = layout == LSS_LAYOUT_CHANNELS_LAST;
    const float *rows = grad_bev;
    if (!cl) {
        LSS_REQUIRE(grad_rows != nullptr, LSS_ERR_WORKSPACE);
        st = launch_bwd_rows(d, tl, tile_start, entries, grad_bev, grad_rows, s);
        if (st != LSS_OK) return st;
        rows = grad_rows;
    }
    const int npix = d.B * d.N * d.HW;
    const int grid = (npix + SPLAT_WARPS - 1) / SPLAT_WARPS;
    const int kc = lss_kc_for(d.C);
    if (kc <= 1) dispatch_gather_nch<1>(cl, grid, s, d, vox, prob, ctx_t, rows, grad_depthnet);
    else if (kc <= 2) dispatch_gather_nch<2>(cl, grid, s, d, vox, prob, ctx_t, rows, grad_depthnet);
    else if (kc <= 4) dispatch_gather_nch<4>(cl, grid, s, d, vox, prob, ctx_t, rows, grad_depthnet);
    else dispatch_gather_nch<8>(cl, grid, s, d, vox, prob, ctx_t, rows, grad_depthnet);
    LSS_CHECK_LAUNCH();
    return LSS_OK;
}

extern "C" int lss_voxel_pooling_bwd(const lss_problem *p, const lss_plan_layout *L, const void *workspace,
                                     const float *grad_bev, int layout, float *grad_rows, float *grad_x, void *stream) {
    int st = lss_check_problem(p);
    if (st != LSS_OK) return st;
    LSS_REQUIRE(L && workspace, LSS_ERR_WORKSPACE);
    LSS_REQUIRE(grad_bev && grad_x, LSS_ERR_BAD_ARG);
    LSS_REQUIRE(layout == LSS_LAYOUT_NCHW || layout == LSS_LAYOUT_CHANNELS_LAST, LSS_ERR_BAD_ARG);
    const Dims d = make_dims(p);
    LSS_REQUIRE(L->n_points == d.n_points, LSS_ERR_WORKSPACE);
    const Tiling tl = make_tiling(L);
    const char *w = (const char *)workspace;
    const int32_t *vox = (const int32_t *)(w + L->off_vox);
    const uint32_t *entries = (const uint32_t *)(w + L->off_entries);
    const int32_t *tile_start = (const int32_t *)(w + L->off_tile_start);
    cudaStream_t s = (cudaStream_t)stream;
    const bool cl = layout == LSS_LAYOUT_CHANNELS_LAST;
    const float *rows = grad_bev;
    if (!cl) {
        LSS_REQUIRE(grad_rows != nullptr, LSS_ERR_WORKSPACE);
        st = launch_bwd_rows(d, tl, tile_start, entries, grad_bev, grad_rows, s);
        if (st != LSS_OK) return st;
        rows = grad_rows;
    }
    const int grid = (d.n_points + SPLAT_WARPS - 1) / SPLAT_WARPS;
    if (cl) k_vp_bwd<true><<<grid, SPLAT_THREADS, 0, s>>>(d, vox, rows, grad_x);
    else k_vp_bwd<false><<<grid, SPLAT_THREADS, 0, s>>>(d, vox, rows, grad_x);
    LSS_CHECK_LAUNCH();
    return LSS_OK;
}
