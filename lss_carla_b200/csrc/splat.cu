// Splat: voxel pooling of lifted camera features into the BEV grid, forward and backward.
//
// Replaces, for shdragron/LSS-Carla:
//   the outer product of CamEncode.get_depth_feat    src/models.py:59      (never materialised)
//   the value half of LiftSplatShoot.voxel_pooling    src/models.py:209, 222-246
//   QuickCumsum.forward / backward, cumsum_trick      src/tools.py:182-219
//
// Forward is a TILE-OWNER kernel: one CTA owns the BEV tile (b, iz, ix, TY consecutive iy) and reads the
// bucket of points the plan assigned to it (plan.cu).  Every BEV element is written exactly once, zeros
// included, so there is no memset and no global atomic.  The tile is staged in shared memory (zero-filled
// with 16-byte stores, non-empty voxels overwrite their channels) and streamed out with 16-byte stores.
//   SORTED       the bucket is sorted by (column, point): warps take 32-entry chunks, find the voxel
//                segments that START in their chunk by ballot, and add float32(prob*ctx) sequentially in
//                ascending point order (a fixed summation order; a segment is never split between warps).
//   SMEM_ATOMIC  warps walk the (unsorted) bucket and accumulate with shared-memory atomics.
//   RED_GLOBAL   classical pixel-owner red.global.add splat into a zeroed BEV, kept for measurement.
//
// Backward is a pure gather (tools.py:212-219): for NCHW gradients a tile-owner kernel first gathers the
// channels of every hit voxel into a channel-contiguous row; a pixel-owner kernel then reads one row per
// frustum point and fuses the outer-product and softmax backward.
#include <cstdlib>

#include "common.cuh"

#define SPLAT_THREADS 256
#define SPLAT_WARPS (SPLAT_THREADS / 32)
#define GATHER_THREADS 128

struct SrcArgs {
    const float *base;   // LIFT: ctx_t [B*N, HW, C]        DENSE: x with strides s[0..5]
    const float *prob;   // LIFT: prob [B*N, D, HW]         DENSE: unused
    const float *prob_col;   // LIFT, optional: prob in camera-column order [B*N, fW, D, fH]
    long long s[6];
};

// Channel ownership of a lane.  VW > 0: C == 32*VW, the lane owns the VW contiguous channels
// [lane*VW, lane*VW+VW) and moves them with one VW*4-byte access.  VW == 0: generic, the lane owns
// channels lane + 32*k (k < KC), predicated on c < C, with an arbitrary element stride.
template <int VW, int KC> struct ChanMap {
    static constexpr int NA = VW ? VW : KC;
    __device__ static __forceinline__ int ch(int lane, int i) { return VW ? lane * VW + i : lane + 32 * i; }
};

template <int VW, int KC>
__device__ __forceinline__ void load_row(const float *__restrict__ row, long long cs, int lane, int C, float *x) {
    if (VW == 4) { const float4 t = __ldg(reinterpret_cast<const float4 *>(row) + lane); x[0] = t.x; x[1] = t.y; x[2] = t.z; x[3] = t.w; }
    else if (VW == 2) { const float2 t = __ldg(reinterpret_cast<const float2 *>(row) + lane); x[0] = t.x; x[1] = t.y; }
    else if (VW == 1) { x[0] = __ldg(row + lane); }
    else {
#pragma unroll
        for (int k = 0; k < KC; ++k) { const int c = lane + 32 * k; x[k] = c < C ? __ldg(row + c * cs) : 0.f; }
    }
}

// Source of one bucket entry: weight and element offset of its C-channel row.
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

// The tile as a 2-D block: NR rows of RL floats; shared-memory row stride SRS, global row stride GRS.
//   NCHW           rows = channels, row = `cols` consecutive iy      (SRS = TY+4, GRS = nx*ny)
//   channels_last  rows = columns,  row = C consecutive channels     (SRS = C,    GRS = nz*C)
struct Tile2D { int NR, RL, SRS; size_t GRS, gbase; };

template <bool CL>
__device__ __forceinline__ Tile2D tile_2d(const Dims &d, const Tiling &tl, const TileCoord &tc) {
    Tile2D t;
    if (CL) {
        t.NR = tc.cols; t.RL = d.C; t.SRS = d.C; t.GRS = (size_t)d.nz * d.C;
        t.gbase = ((size_t)(tc.b * d.nx + tc.ix) * d.ny + tc.y0) * t.GRS + (size_t)tc.iz * d.C;
    } else {
        t.NR = d.C; t.RL = tc.cols; t.SRS = tl.TY + 4; t.GRS = (size_t)d.nx * d.ny;
        t.gbase = ((size_t)(tc.b * d.nz + tc.iz) * d.C) * t.GRS + (size_t)tc.ix * d.ny + tc.y0;
    }
    return t;
}

// Stream a tile to global memory (src == nullptr: zeros).  A warp walks whole rows (or 32/vpr rows at once when a
// row has fewer than 32 vector slots), so the inner loop is one shared load, one global store and two pointer
// increments: the store stream, not index arithmetic, has to be the limit here.
template <bool VEC4, int NT = SPLAT_THREADS>
__device__ __forceinline__ void store_tile(const Tile2D &t, const float *__restrict__ src, float *__restrict__ bev) {
    float *g = bev + t.gbase;
    if (!VEC4) {
        const int total = t.NR * t.RL;
        int row = threadIdx.x / t.RL, v = threadIdx.x - row * t.RL;
        const int dr = NT / t.RL, dv = NT - dr * t.RL;
        for (int i = threadIdx.x; i < total; i += NT) {
            g[(size_t)row * t.GRS + v] = src ? src[row * t.SRS + v] : 0.f;
            v += dv; row += dr;
            if (v >= t.RL) { v -= t.RL; ++row; }
        }
        return;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int vpr = t.RL >> 2;                              // 16-byte slots per row
    const int rpw = vpr >= 32 ? 1 : 32 / vpr;               // rows a warp covers per pass
    const int sub = rpw == 1 ? 0 : lane / vpr;
    const int v0 = rpw == 1 ? lane : lane - sub * vpr;
    if (sub >= rpw) return;
    const int row0 = warp * rpw + sub, rstep = (NT / 32) * rpw;
    float4 *gp = reinterpret_cast<float4 *>(g + (size_t)row0 * t.GRS) + v0;
    const size_t gstep = (size_t)rstep * t.GRS / 4;         // GRS % 4 == 0 on the vector path
    if (src == nullptr) {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int row = row0; row < t.NR; row += rstep, gp += gstep)
            for (int v = v0; v < vpr; v += 32) gp[v - v0] = z;
        return;
    }
    const float4 *sp = reinterpret_cast<const float4 *>(src + row0 * t.SRS) + v0;
    const int sstep = rstep * t.SRS / 4;                    // SRS % 4 == 0
    for (int row = row0; row < t.NR; row += rstep, gp += gstep, sp += sstep)
        for (int v = v0; v < vpr; v += 32) gp[v - v0] = sp[v - v0];
}

// ------------------------------------------------------------------------------------------------
// forward, tile-owner
// ------------------------------------------------------------------------------------------------

// Walk the bucket `ent[0..n)` of one tile and deliver every voxel sum to out[c*cstr + col*colstr]
// (shared-memory staging tile or the BEV tensor itself).
//
// SORTED: a warp takes the 32-entry chunk c0 and owns every voxel segment that STARTS in it (head bits
// found by ballot); it requests UB context rows before consuming them, adds float32(prob*ctx) in
// ascending point order and closes the running sum at every head.  A segment is never split between
// warps: the last segment of a chunk is followed into the next chunks by the same warp.
// ATOMIC: no order, every entry is added with an atomic (shared or global).
template <int VW, int KC, bool ATOMIC, bool DENSE>
__device__ __forceinline__ void walk_bucket(const Dims &d, const SrcArgs &src, int b, const uint32_t *__restrict__ ent,
                                            int n, float *__restrict__ out, long long cstr, long long colstr) {
    using CM = ChanMap<VW, KC>;
    constexpr int NA = CM::NA;
    constexpr int UB = NA >= 8 ? 2 : (NA >= 4 ? 4 : 8);   // rows in flight per warp (register budget)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int C = d.C;
    const long long cs = DENSE ? src.s[5] : 1;

#define LSS_FLUSH(col)                                                                  \
    _Pragma("unroll") for (int a = 0; a < NA; ++a) {                                    \
        const int c = CM::ch(lane, a);                                                  \
        if (VW || c < C) out[c * cstr + (col) * colstr] = acc[a];                       \
    }

    for (int c0 = warp * 32; c0 < n; c0 += SPLAT_THREADS) {
        const int i = c0 + lane;
        const bool valid = i < n;
        const uint32_t e = valid ? __ldg(ent + i) : 0u;
        float w = 0.f; long long off = 0;
        if (valid) entry_source<DENSE>(d, src, b, e, w, off);
        const int lim = min(32, n - c0);
        const int col = (int)(e >> LSS_PIDX_BITS);
        if (ATOMIC) {
            for (int j0 = 0; j0 < lim; j0 += UB) {
                float x[UB][NA];
#pragma unroll
                for (int u = 0; u < UB; ++u)       // lanes >= lim hold off = 0: a valid (unused) row
                    load_row<VW, KC>(src.base + __shfl_sync(LSS_FULL_MASK, off, (j0 + u) & 31), cs, lane, C, x[u]);
#pragma unroll
                for (int u = 0; u < UB; ++u) {
                    if (j0 + u >= lim) break;
                    const float wj = __shfl_sync(LSS_FULL_MASK, w, j0 + u);
                    const int cj = __shfl_sync(LSS_FULL_MASK, col, j0 + u);
#pragma unroll
                    for (int a = 0; a < NA; ++a) {
                        const int c = CM::ch(lane, a);
                        if (VW || c < C) atomicAdd(out + c * cstr + cj * colstr, __fmul_rn(wj, x[u][a]));
                    }
                }
            }
            continue;
        }
        const uint32_t ep = (valid && i > 0) ? __ldg(ent + i - 1) : 0u;
        const bool head = valid && (i == 0 || (e >> LSS_PIDX_BITS) != (ep >> LSS_PIDX_BITS));
        const unsigned heads = __ballot_sync(LSS_FULL_MASK, head);
        if (heads == 0u) continue;               // the whole chunk continues a segment owned by an earlier warp
        const int first = __ffs(heads) - 1;      // entries before it belong to that earlier segment too
        float acc[NA];
#pragma unroll
        for (int a = 0; a < NA; ++a) acc[a] = 0.f;
        int cur = -1;
        for (int j0 = first & ~(UB - 1); j0 < lim; j0 += UB) {
            float x[UB][NA];
#pragma unroll
            for (int u = 0; u < UB; ++u)         // row addresses depend on the entry only, not on prob
                load_row<VW, KC>(src.base + __shfl_sync(LSS_FULL_MASK, off, (j0 + u) & 31), cs, lane, C, x[u]);
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                const int jj = j0 + u;
                if (jj >= lim) break;
                const float wj = __shfl_sync(LSS_FULL_MASK, w, jj);
                if ((heads >> jj) & 1u) {
                    if (cur >= 0) { LSS_FLUSH(cur) }
#pragma unroll
                    for (int a = 0; a < NA; ++a) acc[a] = 0.f;
                    cur = __shfl_sync(LSS_FULL_MASK, col, jj);
                }
                if (jj >= first) {
#pragma unroll
                    for (int a = 0; a < NA; ++a) acc[a] = __fadd_rn(acc[a], __fmul_rn(wj, x[u][a]));
                }
            }
        }
        if (lim == 32) {                         // the last segment of a full chunk may run on
            for (int k = c0 + 32; k < n; k += 32) {
                const int i2 = k + lane;
                const uint32_t e2 = i2 < n ? __ldg(ent + i2) : 0u;
                const bool same = i2 < n && (int)(e2 >> LSS_PIDX_BITS) == cur;
                const unsigned m = __ballot_sync(LSS_FULL_MASK, same);
                const int cnt = m == LSS_FULL_MASK ? 32 : __ffs(~m) - 1;
                if (cnt == 0) break;
                float w2 = 0.f; long long off2 = 0;
                if (lane < cnt) entry_source<DENSE>(d, src, b, e2, w2, off2);
                for (int jj = 0; jj < cnt; ++jj) {
                    const float wj = __shfl_sync(LSS_FULL_MASK, w2, jj);
                    float x[NA];
                    load_row<VW, KC>(src.base + __shfl_sync(LSS_FULL_MASK, off2, jj), cs, lane, C, x);
#pragma unroll
                    for (int a = 0; a < NA; ++a) acc[a] = __fadd_rn(acc[a], __fmul_rn(wj, x[a]));
                }
                if (cnt < 32) break;
            }
        }
        LSS_FLUSH(cur)
    }
#undef LSS_FLUSH
}

template <int NT = SPLAT_THREADS>
__device__ __forceinline__ void zero_smem(float *buf, int n_floats) {
    float4 *z = reinterpret_cast<float4 *>(buf);
    for (int i = threadIdx.x; i < (n_floats >> 2); i += NT) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// One CTA per tile, LSU stores.  Generic path: any shape / alignment.
template <int VW, int KC, bool ATOMIC, bool CL, bool DENSE, bool VEC4>
__global__ void __launch_bounds__(SPLAT_THREADS)
k_splat_fwd_tile(Dims d, Tiling tl, const int32_t *__restrict__ tile_start, const uint32_t *__restrict__ entries,
                 SrcArgs src, float *__restrict__ bev) {
    extern __shared__ __align__(16) float smem[];
    const int tile = blockIdx.x;
    const TileCoord tc = tile_coord(d, tl, tile);
    const Tile2D t2 = tile_2d<CL>(d, tl, tc);
    const int s = __ldg(tile_start + tile), n = __ldg(tile_start + tile + 1) - s;
    if (n == 0) { store_tile<VEC4>(t2, nullptr, bev); return; }   // nothing lands here: stream zeros
    zero_smem(smem, CL ? tl.TY * d.C : d.C * t2.SRS);
    __syncthreads();
    walk_bucket<VW, KC, ATOMIC, DENSE>(d, src, tc.b, entries + s, n, smem, CL ? 1 : t2.SRS, CL ? d.C : 1);
    __syncthreads();
    store_tile<VEC4>(t2, smem, bev);
}

// Fast path of the SORTED mode for C = 8 * CPL (32, 64, 128 channels), fused lift operands: two kernels.
//
// (1) k_fwd_gather -- no shared memory, no barriers, all registers: many CTAs stay resident and their
// dependent load chains (segment table -> entries -> softmax weight / context rows) overlap.  A GROUP of
// 8 lanes owns one voxel segment at a time: lane gl keeps the channel quads gl, gl+8, ... (CPL channels) of
// the running sum in registers and moves them with 16-byte loads (8 lanes = one 128-byte line), so one warp instruction advances four
// voxels.  One CTA walks the voxel records of ONE camera column (b, n, w) -- the plan buckets every
// non-empty voxel by the column of its first point -- and first stages the column's operands in shared
// memory: its fH context rows and D*fH softmax weights (a few KB).  Almost every point of those voxels
// belongs to the column, so the inner loop reads shared memory only; a point of another column or camera
// takes its row and weight from global memory.  Groups look 8 entries ahead (lane gl resolves entry
// base+gl) and add float32(prob*ctx) in ascending point order.  The sum goes to the voxel's compact row
// of vsum (L2-resident, 4*C*V_hit bytes).
template <int CPL>
__global__ void __launch_bounds__(GATHER_THREADS, 1024 / GATHER_THREADS)
k_fwd_gather(Dims d, int key_lo, int n_keys, const int32_t *__restrict__ key_count, const int4 *__restrict__ seg_recs,
             const int32_t *__restrict__ counters, const int4 *__restrict__ mixed_recs, long long n_rows_cap,
             const uint32_t *__restrict__ entries, const float *__restrict__ prob, const float *__restrict__ prob_col,
             const float *__restrict__ ctx_t, float *__restrict__ vsum) {
    extern __shared__ __align__(16) float s_col[];       // [fH][C] context rows of the column, [D][fH] softmax weights
    constexpr int NG = GATHER_THREADS / 8;               // groups per CTA
    constexpr int LF = CPL <= 8 ? 4 : 2;                 // context rows in flight per group (generic voxels)
    lss_pdl_wait();                                      // launched programmatically behind the plan build (or lift_prepare):
                                                         // everything below reads what those kernels wrote
    lss_pdl_trigger();                                   // PDL: the store kernel may be scheduled in our tail.  The trigger comes
                                                         // AFTER the wait: a dependent that starts early then knows that every
                                                         // kernel before this one is complete (transitivity of the chain)
    // The FIRST CTAs of the grid drain the queue of mixed voxels (group per voxel, all operands from global memory):
    // they start at once, live a few microseconds and hand their slots to the column CTAs that did not fit the first
    // wave, instead of forming the tail of the kernel.
    const int n_queue = (int)gridDim.x - n_keys;
    const int bid = (int)blockIdx.x < n_queue ? n_keys + (int)blockIdx.x : (int)blockIdx.x - n_queue;
    const bool column = bid < n_keys;                    // (measured: a warp per mixed voxel -- 16 rows in flight, sum
    const int key = column ? key_lo + bid : 0;           // handed on by shuffle -- is slower than a group per voxel)
    const int n_rec = column ? __ldg(key_count + key) : __ldg(counters + 1);
    const int w0 = column ? key % d.fW : -1, bn = key / d.fW;
    const int n0 = column ? bn % d.N : -1;
    const int C = d.C;
    float *s_prob = s_col + d.fH * C;
    if (column) {   // ---- stage the column: fH context rows and D*fH weights, the operands of (almost) all its points
        const float4 *src = reinterpret_cast<const float4 *>(ctx_t + ((size_t)bn * d.HW + w0) * C);
        const int c4 = C >> 2;
        for (int i = threadIdx.x; i < d.fH * c4; i += GATHER_THREADS) {
            const int h = i / c4, q = i - h * c4;
            reinterpret_cast<float4 *>(s_col)[i] = __ldg(src + (size_t)h * d.fW * c4 + q);
        }
        if (prob_col != nullptr) {                       // contiguous block [D][fH] of this column
            const float *psrc = prob_col + (size_t)key * d.D * d.fH;
            for (int i = threadIdx.x; i < d.D * d.fH; i += GATHER_THREADS) s_prob[i] = __ldg(psrc + i);
        } else {
            const float *psrc = prob + (size_t)bn * d.DHW + w0;
            for (int i = threadIdx.x; i < d.D * d.fH; i += GATHER_THREADS) s_prob[i] = __ldg(psrc + (size_t)i * d.fW);
        }
    }
    const int lane = threadIdx.x & 31;
    const int gl = lane & 7;
    if (!column) {
        // ---- long voxels (>= LSS_LONG_VOXEL points, e.g. the cells right in front of a camera): one at a time by the
        // whole CTA.  NG points per pass: every group fetches one point's context row and writes float32(prob*ctx) to
        // shared memory (all loads in flight together); thread c then adds the NG products of channel c in ascending
        // point order -- the same sequence of float32 additions as everywhere else, without the serial load chain.
        const int n_long = __ldg(counters + 2);
        float *s_prod = s_col;                                // [NG][C]
        for (int rl = bid - n_keys; rl < n_long; rl += n_queue) {   // CTA-uniform
            const int4 lrec = __ldg(mixed_recs + (n_rows_cap - 1 - rl));
            const uint32_t *ent = entries + lrec.x;
            const float *prob_b = prob + (size_t)lrec.z * d.P;
            const float *ctx_b = ctx_t + (size_t)lrec.z * d.N * d.HW * C + gl * 4;
            const int j = threadIdx.x >> 3;                   // this group's point within the pass
            float accc = 0.f;                                 // running sum of channel threadIdx.x
            for (int base = 0; base < lrec.y; base += NG) {
                const int cnt = min(NG, lrec.y - base);
                if (j < cnt) {
                    const unsigned pidx = __ldg(ent + base + j) & LSS_PIDX_MASK;
                    const unsigned cam = lss_div20(pidx, d.mDHW);
                    const unsigned rr = pidx - cam * d.DHW;
                    const unsigned hw = rr - lss_div20(rr, d.mHW) * d.HW;
                    const float w = __ldg(prob_b + pidx);
                    const float4 *rowp = reinterpret_cast<const float4 *>(ctx_b + (size_t)(cam * d.HW + hw) * C);
                    float4 *dst = reinterpret_cast<float4 *>(s_prod + j * C) + gl;
#pragma unroll
                    for (int q = 0; q < CPL / 4; ++q) {
                        const float4 v = __ldg(rowp + 8 * q);
                        dst[8 * q] = make_float4(__fmul_rn(w, v.x), __fmul_rn(w, v.y), __fmul_rn(w, v.z), __fmul_rn(w, v.w));
                    }
                }
                __syncthreads();
                if ((int)threadIdx.x < C)
                    for (int jj = 0; jj < cnt; ++jj) accc = __fadd_rn(accc, s_prod[jj * C + threadIdx.x]);
                __syncthreads();
            }
            if ((int)threadIdx.x < C) vsum[(size_t)lrec.w * C + threadIdx.x] = accc;
        }
    }
    const int4 *recs = column ? seg_recs + (size_t)key * (d.D * d.fH) : mixed_recs;
    const float *s_ctx = s_col + gl * 4;                 // lane gl: float4 slots gl, gl+8, ...
    const int stride = column ? NG : NG * n_queue;
    int r = (column ? 0 : NG * (bid - n_keys)) + (threadIdx.x >> 3);
    // software pipeline over the group's records: the record and the first 8 entries of the NEXT voxel are
    // requested before the current one is consumed.  All shuffles use the full mask (a lane-dependent mask
    // costs a MATCH per shuffle), so every loop below is warp-uniform and the groups are predicated.
    int4 rec = make_int4(0, 0, 0, 0);                    // {first entry, length, batch index, compact row}
    uint32_t e0 = 0;
    if (r < n_rec) rec = __ldg(recs + r);
    if (rec.y > 0 && gl < rec.y) e0 = __ldg(entries + rec.x + gl);
    if (column) __syncthreads();
    while (__any_sync(LSS_FULL_MASK, r < n_rec)) {
        const int nr = r + stride;
        const float *prob_b = prob + (size_t)rec.z * d.P;
        const float *ctx_b = ctx_t + (size_t)rec.z * d.N * d.HW * C + gl * 4;
        int4 nrec = make_int4(0, 0, 0, 0);
        if (nr < n_rec) nrec = __ldg(recs + nr);
        const bool live = r < n_rec;
        float acc[CPL];
#pragma unroll
        for (int a = 0; a < CPL; ++a) acc[a] = 0.f;
        if (live && rec.y < 0) {
            // PURE voxel: the fH image rows of this column at depth bin dd, in order -- operands are all staged
            const float *wp = s_prob + (-rec.y - 1) * d.fH;
            for (int h = 0; h < d.fH; ++h) {
                const float wj = wp[h];
                const float4 *rowp = reinterpret_cast<const float4 *>(s_ctx + h * C);
#pragma unroll
                for (int q = 0; q < CPL / 4; ++q) {
                    const float4 v = rowp[8 * q];
                    acc[4 * q] = __fadd_rn(acc[4 * q], __fmul_rn(wj, v.x));
                    acc[4 * q + 1] = __fadd_rn(acc[4 * q + 1], __fmul_rn(wj, v.y));
                    acc[4 * q + 2] = __fadd_rn(acc[4 * q + 2], __fmul_rn(wj, v.z));
                    acc[4 * q + 3] = __fadd_rn(acc[4 * q + 3], __fmul_rn(wj, v.w));
                }
            }
        }
        const int len = live && rec.y > 0 ? rec.y : 0;
        const int maxlen = __reduce_max_sync(LSS_FULL_MASK, len);
        const uint32_t *ent = entries + rec.x;
        for (int base = 0; base < maxlen; base += 8) {     // generic voxels of the warp, in lockstep
            const int cnt = min(8, len - base);            // <= 0 for a group that has nothing (more) to do
            const int maxcnt = min(8, maxlen - base);
            float w = 0.f;
            int ro = -1;                                   // >= 0: row offset in ctx_t; < 0: ~offset in s_col (row 0 if unused)
            if (gl < cnt) {
                const unsigned pidx = (base == 0 ? e0 : __ldg(ent + base + gl)) & LSS_PIDX_MASK;
                const unsigned cam = lss_div20(pidx, d.mDHW);
                const unsigned rr = pidx - cam * d.DHW;
                const unsigned dd = lss_div20(rr, d.mHW);
                const unsigned hw = rr - dd * d.HW;
                const unsigned h = lss_div20(hw, d.mfW), ww = hw - h * d.fW;
                if ((int)cam == n0 && (int)ww == w0) { ro = ~(int)(h * C); w = s_prob[dd * d.fH + h]; }
                else { ro = (int)((cam * d.HW + hw) * C); w = __ldg(prob_b + pidx); }
            }
#pragma unroll
            for (int j0 = 0; j0 < 8; j0 += LF) {
                if (j0 >= maxcnt) break;
                float x[LF][CPL];
#pragma unroll
                for (int u = 0; u < LF; ++u) {             // LF rows in flight; a generic pointer reaches both the
                    const int oj = __shfl_sync(LSS_FULL_MASK, ro, j0 + u, 8);      // staged rows and global memory
                    const float4 *rowp = oj < 0 ? reinterpret_cast<const float4 *>(s_ctx + ~oj)
                                                : reinterpret_cast<const float4 *>(ctx_b + oj);
#pragma unroll
                    for (int q = 0; q < CPL / 4; ++q) {
                        const float4 v = rowp[8 * q];
                        x[u][4 * q] = v.x; x[u][4 * q + 1] = v.y; x[u][4 * q + 2] = v.z; x[u][4 * q + 3] = v.w;
                    }
                }
#pragma unroll
                for (int u = 0; u < LF; ++u) {
                    const float wj = __shfl_sync(LSS_FULL_MASK, w, j0 + u, 8);
                    if (j0 + u < cnt) {
#pragma unroll
                        for (int a = 0; a < CPL; ++a) acc[a] = __fadd_rn(acc[a], __fmul_rn(wj, x[u][a]));
                    }
                }
            }
        }
        if (live) {
            float4 *dst = reinterpret_cast<float4 *>(vsum + (size_t)rec.w * C) + gl;
#pragma unroll
            for (int q = 0; q < CPL / 4; ++q) dst[8 * q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
        }
        r = nr;
        rec = nrec;
        if (rec.y > 0 && gl < rec.y) e0 = __ldg(entries + rec.x + gl);    // generic voxel: its first 8 entries
    }
}

// (2) k_fwd_store -- tile owner, pure streaming: the compact rows of the tile (one contiguous block) are
// transposed into a zero-filled shared-memory tile, which is written out with 16-byte stores.  Every BEV
// element is written exactly once, zeros included.  (A variant without the staging tile -- zeros from
// registers, non-empty slots looked up through a column map -- was measured slower: 28 us vs 19 us.)
template <bool CL, bool VEC4, int NT>
__global__ void __launch_bounds__(NT)
k_fwd_store(Dims d, Tiling tl, int tile_lo, int CH, const int32_t *__restrict__ tile_start, const int32_t *__restrict__ tile_nseg,
            const int32_t *__restrict__ tile_row0, const uint32_t *__restrict__ segs, const float *__restrict__ vsum,
            float *__restrict__ bev) {
    extern __shared__ __align__(16) float smem[];
    const int tile = tile_lo + blockIdx.x;
    const int c0 = blockIdx.y * (CH < 0 ? -CH : CH);      // this CTA owns channels [c0, c0 + CH) of the tile
    const TileCoord tc = tile_coord(d, tl, tile);
    Tile2D t2 = tile_2d<CL>(d, tl, tc);
    { const int ch = CH < 0 ? -CH : CH;
      if (CL) { t2.RL = ch; t2.SRS = ch; t2.gbase += c0; }
      else { t2.NR = ch; t2.gbase += (size_t)c0 * t2.GRS; } }
    const int nseg = CH < 0 ? 0 : __ldg(tile_nseg + tile);      // CH < 0: measurement aid, zero tiles only
    if (CH < 0) CH = -CH;
    if (nseg == 0) { store_tile<VEC4, NT>(t2, nullptr, bev); return; }
    const int s = __ldg(tile_start + tile), row0 = __ldg(tile_row0 + tile);
    const int C = d.C, c4 = CH >> 2;                      // C % 4 == 0 and CH % 4 == 0 on this path
    const int per_pass = NT / c4;              // rows per pass
    const int q = threadIdx.x % c4, r0 = threadIdx.x / c4;
    // PDL: everything above reads plan data only (written by earlier, completed launches); the compact rows come
    // from the gather kernel, which may still be running if this kernel was launched programmatically
    lss_pdl_wait();
    // request the first rows before zero-filling the staging tile
    int col0 = 0;
    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool has0 = r0 < per_pass && r0 < nseg;
    if (has0) {
        col0 = (int)(__ldg(segs + s + r0) >> LSS_PIDX_BITS);
        v0 = __ldg(reinterpret_cast<const float4 *>(vsum + (size_t)(row0 + r0) * C + c0) + q);
    }
    zero_smem<NT>(smem, CL ? tl.TY * CH : CH * t2.SRS);
    __syncthreads();
    if (r0 < per_pass) {
        for (int r = r0; r < nseg; r += per_pass) {
            int col; float4 v;
            if (r == r0) { col = col0; v = v0; }
            else {
                col = (int)(__ldg(segs + s + r) >> LSS_PIDX_BITS);
                v = __ldg(reinterpret_cast<const float4 *>(vsum + (size_t)(row0 + r) * C + c0) + q);
            }
            if (CL) {
                *reinterpret_cast<float4 *>(smem + col * CH + 4 * q) = v;
            } else {
                float *dst = smem + (4 * q) * t2.SRS + col;
                dst[0] = v.x; dst[t2.SRS] = v.y; dst[2 * t2.SRS] = v.z; dst[3 * t2.SRS] = v.w;
            }
        }
    }
    __syncthreads();
    store_tile<VEC4, NT>(t2, smem, bev);
}

// (2a) k_fwd_store_rows -- NCHW, 16-byte aligned rows: the same streaming store WITHOUT a staging tile.  All C channel
// rows of a tile share one hit pattern (the tile's non-empty columns), so every lane resolves its two 16-byte slots
// ONCE -- which compact row, if any, feeds each of its 8 columns -- and then walks the channels: a slot without a
// hit is a zero store straight from registers, the others read the tile's compact rows, which are staged in
// shared memory (a few KB, odd stride).  Shared memory per CTA drops from 52 KB to ~8 KB, so the SM runs its full
// complement of CTAs, and the 51 KB zero-fill / read-back of the staging tile disappears.
#define ROWS_CAP 64      // compact rows staged in shared memory (tiles with more non-empty columns read the rest from global)
// The prologue is two global round trips: the three words of tile meta data are requested together, and the tile's compact
// rows are requested (into registers) together with its column list, before the column map is built; rows and map then go
// to shared memory behind one barrier.  A CTA lives ~6 us of which ~2 us are stores, so the dependent loads in front of the
// store stream are what keeps the kernel away from the write floor.  (Measured and dropped in round 1: a four-round-trip
// prologue, 5 CTAs per SM, a persistent cp.async-pipelined variant, zero sectors written by spare CTAs of the gather grid.)
__global__ void __launch_bounds__(SPLAT_THREADS, 4)
k_fwd_store_rows(Dims d, Tiling tl, int tile_lo, const int32_t *__restrict__ tile_start, const int32_t *__restrict__ tile_nseg,
                 const int32_t *__restrict__ tile_row0, const uint32_t *__restrict__ segs, const float *__restrict__ vsum,
                 float *__restrict__ bev) {
    extern __shared__ __align__(16) float s_rows[];               // [ROWS_CAP][C + 1], then short map[TY rounded to 8]
    lss_pdl_wait();                                               // plan data and compact rows come from the kernels before this one
    lss_pdl_trigger();                                            // a dependent launch (the backward's row gather) may be scheduled in our tail
    const int tile = tile_lo + blockIdx.x;
    const int C = d.C, SR = C + 1, c4 = C >> 2;
    short *s_map = reinterpret_cast<short *>(s_rows + ROWS_CAP * SR + (ROWS_CAP & 1));
    const int nseg = __ldg(tile_nseg + tile), s = __ldg(tile_start + tile), row0 = __ldg(tile_row0 + tile);   // one round trip
    for (int i = threadIdx.x; i < (tl.TY + 7) / 8 * 4; i += SPLAT_THREADS) reinterpret_cast<unsigned *>(s_map)[i] = 0xFFFFFFFFu;
    const TileCoord tc = tile_coord(d, tl, tile);
    const Tile2D t2 = tile_2d<false>(d, tl, tc);
    if (nseg == 0) { store_tile<true>(t2, nullptr, bev); return; }
    const int nst = min(nseg, ROWS_CAP);
    {
        const uint32_t e0 = threadIdx.x < nseg ? __ldg(segs + s + threadIdx.x) : 0u;
        float4 r[4];                                              // this thread's share of the first 4 * 256 row quads
        const float4 *rsrc = reinterpret_cast<const float4 *>(vsum + (size_t)row0 * C);   // the tile's rows are one block
        const int nq = nst * c4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = threadIdx.x + j * SPLAT_THREADS;
            r[j] = i < nq ? __ldg(rsrc + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncthreads();                                          // the map is initialised
        if (threadIdx.x < nseg) s_map[e0 >> LSS_PIDX_BITS] = (short)threadIdx.x;
        for (int k = threadIdx.x + SPLAT_THREADS; k < nseg; k += SPLAT_THREADS) s_map[__ldg(segs + s + k) >> LSS_PIDX_BITS] = (short)k;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = threadIdx.x + j * SPLAT_THREADS;
            if (i < nq) {
                const int k = i / c4, q = i - k * c4;
                float *dst = s_rows + k * SR + 4 * q;
                dst[0] = r[j].x; dst[1] = r[j].y; dst[2] = r[j].z; dst[3] = r[j].w;
            }
        }
        for (int i = threadIdx.x + 4 * SPLAT_THREADS; i < nq; i += SPLAT_THREADS) {     // C > 64 only
            const int k = i / c4, q = i - k * c4;
            const float4 v = __ldg(rsrc + i);
            float *dst = s_rows + k * SR + 4 * q;
            dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
        }
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int vpr = t2.RL >> 2;                                   // 16-byte slots per channel row (<= 64 handled per lane pair)
    const float *rows_g = vsum + (size_t)row0 * C;
    for (int vb = 0; vb < vpr; vb += 32) {                        // usually 2 slots per lane (warp-uniform loop)
        const int v0 = vb + lane;
        if (v0 >= vpr) continue;
        const uint2 m = *reinterpret_cast<const uint2 *>(s_map + 4 * v0);
        const int k0 = (short)(m.x & 0xFFFFu), k1 = (short)(m.x >> 16), k2 = (short)(m.y & 0xFFFFu), k3 = (short)(m.y >> 16);
        float4 *gp = reinterpret_cast<float4 *>(bev + t2.gbase + (size_t)warp * t2.GRS) + v0;
        const size_t gstep = (size_t)SPLAT_WARPS * t2.GRS / 4;
        if (nseg <= ROWS_CAP) {                                   // CTA-uniform: every compact row is staged
            // one loop for all lanes (no divergence between lanes with and without hits): predicated shared loads
            const float *p0 = s_rows + max(k0, 0) * SR, *p1 = s_rows + max(k1, 0) * SR;
            const float *p2 = s_rows + max(k2, 0) * SR, *p3 = s_rows + max(k3, 0) * SR;
#pragma unroll 4
            for (int c = warp; c < C; c += SPLAT_WARPS, gp += gstep) {
                float4 o;
                o.x = k0 >= 0 ? p0[c] : 0.f; o.y = k1 >= 0 ? p1[c] : 0.f;
                o.z = k2 >= 0 ? p2[c] : 0.f; o.w = k3 >= 0 ? p3[c] : 0.f;
                *gp = o;
            }
        } else {
            auto val = [&](int k, int c) { return k < 0 ? 0.f : (k < ROWS_CAP ? s_rows[k * SR + c] : __ldg(rows_g + (size_t)k * C + c)); };
            for (int c = warp; c < C; c += SPLAT_WARPS, gp += gstep) *gp = make_float4(val(k0, c), val(k1, c), val(k2, c), val(k3, c));
        }
    }
}

// (2b) k_fwd_store_tma (contiguous channels_last tiles of TILE plans) -- the same streaming store as a PERSISTENT kernel: every CTA walks tiles with two staging
// buffers.  The rows of the staged tile leave through the bulk-copy engine (cp.async.bulk shared -> global, issued by
// one warp), so the CTA does not wait for its stores: while tile k drains it zero-fills the other buffer and
// transposes tile k+1 into it; the meta data and the first compact rows of tile k+1 are requested one iteration
// ahead, which takes the dependent global loads off the critical path.  Needs 16-byte aligned rows (VEC4 shapes).
__device__ __forceinline__ void bulk_store(float *gdst, const float *ssrc, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gdst), "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <bool CL>
__global__ void __launch_bounds__(SPLAT_THREADS)
k_fwd_store_tma(Dims d, Tiling tl, int tile_lo, int n_tiles, const int32_t *__restrict__ tile_start,
                const int32_t *__restrict__ tile_nseg, const int32_t *__restrict__ tile_row0,
                const uint32_t *__restrict__ segs, const float *__restrict__ vsum, float *__restrict__ bev) {
    extern __shared__ __align__(16) float smem[];    // (dynamic shared memory starts 128-byte aligned)
    const int C = d.C, c4 = C >> 2;
    const int SRS = CL ? C : tl.TY + 4;
    const int tile_floats = CL ? tl.TY * C : C * SRS;
    const int per_pass = SPLAT_THREADS / c4;              // compact rows per pass (C = 64: 16)
    const int q = threadIdx.x % c4, r0 = threadIdx.x / c4;
    const bool loader = r0 < per_pass;
    // prefetched state of the NEXT tile: meta data and this thread's share of its first 2*per_pass compact rows
    int n_nseg = 0, n_s = 0, n_row0 = 0, n_col[2] = {0, 0};
    float4 n_v[2];
    auto prefetch = [&](int t) {
        n_nseg = 0;
        if (t >= n_tiles || tile_nseg == nullptr) return;     // (null: measurement aid, zero tiles only)
        const int tile = tile_lo + t;
        n_nseg = __ldg(tile_nseg + tile);
        if (n_nseg == 0) return;
        n_s = __ldg(tile_start + tile);
        n_row0 = __ldg(tile_row0 + tile);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int r = r0 + u * per_pass;
            if (loader && r < n_nseg) {
                n_col[u] = (int)(__ldg(segs + n_s + r) >> LSS_PIDX_BITS);
                n_v[u] = __ldg(reinterpret_cast<const float4 *>(vsum + (size_t)(n_row0 + r) * C) + q);
            }
        }
    };
    prefetch(blockIdx.x);
    int it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        float *buf = smem + (it & 1) * tile_floats;
        const int nseg = n_nseg, s = n_s, row0 = n_row0;
        const int col[2] = {n_col[0], n_col[1]};
        const float4 v[2] = {n_v[0], n_v[1]};
        prefetch(t + gridDim.x);                          // dependent loads of the next tile start now
        if (threadIdx.x < 32) bulk_wait_read<1>();        // this buffer was handed to the copy engine two tiles ago
        __syncthreads();
        zero_smem(buf, tile_floats);
        __syncthreads();
        if (loader) {
            for (int r = r0, u = 0; r < nseg; r += per_pass, ++u) {
                int cc; float4 vv;
                if (u < 2) { cc = col[u]; vv = v[u]; }
                else {
                    cc = (int)(__ldg(segs + s + r) >> LSS_PIDX_BITS);
                    vv = __ldg(reinterpret_cast<const float4 *>(vsum + (size_t)(row0 + r) * C) + q);
                }
                if (CL) *reinterpret_cast<float4 *>(buf + cc * C + 4 * q) = vv;
                else {
                    float *dst = buf + (4 * q) * SRS + cc;
                    dst[0] = vv.x; dst[SRS] = vv.y; dst[2 * SRS] = vv.z; dst[3 * SRS] = vv.w;
                }
            }
        }
        fence_async_smem();                               // generic-proxy writes -> visible to the async proxy
        __syncthreads();
        if (threadIdx.x < 32) {
            const TileCoord tc = tile_coord(d, tl, tile_lo + t);
            const Tile2D t2 = tile_2d<CL>(d, tl, tc);
            float *g = bev + t2.gbase;
            if (CL && t2.GRS == (size_t)C) {              // the whole tile is one contiguous run
                if (threadIdx.x == 0) bulk_store(g, buf, (unsigned)(t2.NR * C) * 4u);
            } else {
                for (int r = threadIdx.x; r < t2.NR; r += 32) bulk_store(g + (size_t)r * t2.GRS, buf + r * SRS, (unsigned)t2.RL * 4u);
            }
            bulk_commit();                                // one group per tile and lane (possibly empty)
        }
    }
    if (threadIdx.x < 32) bulk_wait_all();
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
// One CTA per non-empty tile marks its hit columns in shared memory; warps then gather the C strided
// channels of each hit voxel (4-byte loads C*nx*ny apart; neighbouring columns share 32-byte sectors
// through L1) and write one contiguous 4*C-byte row.  Only sectors that contain a hit voxel are read.
__global__ void __launch_bounds__(SPLAT_THREADS)
k_bwd_rows_nchw(Dims d, Tiling tl, const int32_t *__restrict__ tile_start, const uint32_t *__restrict__ entries,
                const float *__restrict__ grad_bev, float *__restrict__ rows) {
    extern __shared__ __align__(16) float smem[];
    const int tile = blockIdx.x;
    const int s = __ldg(tile_start + tile), n = __ldg(tile_start + tile + 1) - s;
    if (n == 0) return;
    const TileCoord tc = tile_coord(d, tl, tile);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int C = d.C;
    unsigned *hit = reinterpret_cast<unsigned *>(smem);      // bitmap over the tile's columns
    const int nwords = (tl.TY + 31) >> 5;
    for (int i = threadIdx.x; i < nwords; i += SPLAT_THREADS) hit[i] = 0u;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += SPLAT_THREADS) {
        const int col = (int)(__ldg(entries + s + i) >> LSS_PIDX_BITS);
        atomicOr(hit + (col >> 5), 1u << (col & 31));
    }
    __syncthreads();
    const size_t plane = (size_t)d.nx * d.ny;
    const float *gsrc = grad_bev + ((size_t)(tc.b * d.nz + tc.iz) * C) * plane + (size_t)tc.ix * d.ny + tc.y0;
    const int v0 = ((tc.b * d.nz + tc.iz) * d.nx + tc.ix) * d.ny + tc.y0;
    // hit columns are dealt round-robin to the warps
    int k = 0;
    for (int wd = 0; wd < nwords; ++wd) {
        unsigned bits = hit[wd];
        while (bits) {
            const int col = (wd << 5) + __ffs(bits) - 1;
            bits &= bits - 1;
            if ((k++ & (SPLAT_WARPS - 1)) != warp) continue;
            float *dst = rows + (size_t)(v0 + col) * C;
            for (int c = lane; c < C; c += 32) dst[c] = __ldg(gsrc + (size_t)c * plane + col);
        }
    }
}

// Pixel-owner gather: one warp per camera pixel (bn, h, w) walks its D frustum points.
//   g      = rows[voxel(p)]                               (QuickCumsum.backward + griddify backward)
//   gp_d   = <g, ctx>            d_ctx += prob_d * g        (outer product backward, models.py:59)
//   d_logit_d = prob_d * (gp_d - sum_d' prob_d' gp_d')      (softmax backward, models.py:50)
// Lane `dl` of a 32-depth chunk owns depth dl: it loads voxel id / prob / row offset once; rows are then
// broadcast by shuffle and loaded by all lanes (dropped points read row 0 with weight 0, so the loop is
// branch-free and loads of consecutive depths overlap).  The per-depth dot products are reduced through
// shared memory (one column sum per lane) instead of 5 shuffles per depth.  The 8 warps of a CTA own 8
// consecutive pixels; outputs are staged in shared memory and written as 32-byte runs.
template <int VW, int KC, int NCH, bool CL>
__global__ void __launch_bounds__(SPLAT_THREADS)
k_bwd_gather(Dims d, const int32_t *__restrict__ vox, const float *__restrict__ prob, const float *__restrict__ ctx_t,
             const float *__restrict__ rows, float *__restrict__ grad_dn) {
    extern __shared__ __align__(16) float smem[];
    using CM = ChanMap<VW, KC>;
    constexpr int NA = CM::NA;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int npix = d.B * d.N * d.HW;
    const int pix0 = blockIdx.x * SPLAT_WARPS;
    const int pix = pix0 + warp;
    const int DC = d.D + d.C;
    float *part = smem + warp * (32 * 33);                 // [32 depths][33] partial dot products
    float *stage = smem + SPLAT_WARPS * (32 * 33);         // [D+C][8] outputs of the CTA's 8 pixels
    if (pix < npix) {
        const int bn = pix / d.HW, hw = pix - bn * d.HW;
        float ctx[NA], dctx[NA];
        load_row<VW, KC>(ctx_t + (size_t)pix * d.C, 1, lane, d.C, ctx);
#pragma unroll
        for (int a = 0; a < NA; ++a) dctx[a] = 0.f;
        float pr[NCH], gp[NCH];
        long long ro[NCH];
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            const int dd = ch * 32 + lane;
            pr[ch] = 0.f; gp[ch] = 0.f; ro[ch] = 0;
            if (dd < d.D) {
                const size_t p = ((size_t)bn * d.D + dd) * d.HW + hw;
                const int v = __ldg(vox + p);
                pr[ch] = __ldg(prob + p);      // dropped points still take part in the softmax backward
                ro[ch] = v < 0 ? -1 : (CL ? (long long)voxel_row_offset_cl(v, d) : (long long)v * d.C);
            }
        }
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            const int cnt = min(32, d.D - ch * 32);
            if (cnt <= 0) break;
#pragma unroll 4
            for (int dl = 0; dl < cnt; ++dl) {
                const long long o = __shfl_sync(LSS_FULL_MASK, ro[ch], dl);
                const float pj = o < 0 ? 0.f : __shfl_sync(LSS_FULL_MASK, pr[ch], dl);   // dropped: weight 0
                float g[NA];
                // dropped point: read this pixel's own (finite) context row instead, weight 0
                load_row<VW, KC>(o < 0 ? ctx_t + (size_t)pix * d.C : rows + o, 1, lane, d.C, g);
                float dot = 0.f;
#pragma unroll
                for (int a = 0; a < NA; ++a) {
                    dot = fmaf(g[a], ctx[a], dot);
                    dctx[a] = fmaf(pj, g[a], dctx[a]);
                }
                part[dl * 33 + lane] = o < 0 ? 0.f : dot;
            }
            __syncwarp();
            if (lane < cnt) {
                float sacc = 0.f;
#pragma unroll 8
                for (int j = 0; j < 32; ++j) sacc += part[lane * 33 + j];
                gp[ch] = sacc;
            }
            __syncwarp();
        }
        float sdot = 0.f;
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) sdot = fmaf(pr[ch], gp[ch], sdot);
        sdot = warp_sum(sdot);
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            const int dd = ch * 32 + lane;
            if (dd < d.D) stage[dd * SPLAT_WARPS + warp] = pr[ch] * (gp[ch] - sdot);
        }
#pragma unroll
        for (int a = 0; a < NA; ++a) {
            const int c = CM::ch(lane, a);
            if (VW || c < d.C) stage[(d.D + c) * SPLAT_WARPS + warp] = dctx[a];
        }
    }
    __syncthreads();
    // write [D+C] x 8 pixels: 8 consecutive threads cover 8 consecutive pixels of one channel row
    const int pl = threadIdx.x & (SPLAT_WARPS - 1);
    const int opix = pix0 + pl;
    if (opix < npix) {
        const int bn = opix / d.HW, hw = opix - bn * d.HW;
        float *out = grad_dn + (size_t)bn * DC * d.HW + hw;
        for (int c = threadIdx.x >> 3; c < DC; c += SPLAT_THREADS / SPLAT_WARPS) out[(size_t)c * d.HW] = stage[c * SPLAT_WARPS + pl];
    }
}

// ---- backward through the plan's compact rows (sorted plans, C in {32, 64, 128})
//
// (1) k_bwd_rows_compact: gradient rows of the non-empty voxels, channel-contiguous, in the plan's compact
// row order (the mirror image of k_fwd_store).  NCHW: a warp owns 8 channels, a lane one voxel of the
// tile: 8 independent 4-byte loads per lane (neighbouring columns share 32-byte sectors), one full
// 32-byte sector written per lane.  channels_last: plain 16-byte copies.
template <bool CL>
__global__ void __launch_bounds__(SPLAT_THREADS, 4)
k_bwd_rows_compact(Dims d, Tiling tl, int tile_lo, const int32_t *__restrict__ tile_start, const int32_t *__restrict__ tile_nseg,
                   const int32_t *__restrict__ tile_row0, const uint32_t *__restrict__ segs,
                   const float *__restrict__ grad_bev, float *__restrict__ grows) {
    lss_pdl_wait();                                // the gradient (and, in a captured step, the plan) come from kernels before this one
    lss_pdl_trigger();
    const int tile = tile_lo + blockIdx.x;
    const int nseg = __ldg(tile_nseg + tile);
    const int s = __ldg(tile_start + tile), row0 = __ldg(tile_row0 + tile);
    if (nseg == 0) return;
    const TileCoord tc = tile_coord(d, tl, tile);
    const int C = d.C;
    if (CL) {
        const int c4 = C >> 2, per_pass = SPLAT_THREADS / c4;
        const int q = threadIdx.x % c4, r0 = threadIdx.x / c4;
        const size_t GRS = (size_t)d.nz * C;
        const float *gsrc = grad_bev + ((size_t)(tc.b * d.nx + tc.ix) * d.ny + tc.y0) * GRS + (size_t)tc.iz * C;
        if (r0 < per_pass)
            for (int k = r0; k < nseg; k += per_pass) {
                const int col = (int)(__ldg(segs + s + k) >> LSS_PIDX_BITS);
                reinterpret_cast<float4 *>(grows + (size_t)(row0 + k) * C)[q] =
                    __ldg(reinterpret_cast<const float4 *>(gsrc + (size_t)col * GRS) + q);
            }
        return;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t plane = (size_t)d.nx * d.ny;
    const float *gsrc = grad_bev + ((size_t)(tc.b * d.nz + tc.iz) * C) * plane + (size_t)tc.ix * d.ny + tc.y0;
    for (int k0 = 0; k0 < nseg; k0 += 32) {
        const int k = k0 + lane;
        const bool live = k < nseg;
        const int col = live ? (int)(__ldg(segs + s + k) >> LSS_PIDX_BITS) : 0;
        for (int oct = warp; oct < (C >> 3); oct += SPLAT_WARPS) {
            const float *src = gsrc + (size_t)(oct * 8) * plane + col;
            float v[8];
#pragma unroll
            for (int a = 0; a < 8; ++a) v[a] = live ? __ldg(src + (size_t)a * plane) : 0.f;
            if (live) {
                float4 *dst = reinterpret_cast<float4 *>(grows + (size_t)(row0 + k) * C + oct * 8);
                dst[0] = make_float4(v[0], v[1], v[2], v[3]);
                dst[1] = make_float4(v[4], v[5], v[6], v[7]);
            }
        }
    }
}

// (2) k_bwd_gather_px: pixel owner, like k_bwd_gather, but a GROUP of 8 lanes owns a camera pixel (lane gl
// keeps channel quads gl, gl+8, ... of the context and of its gradient), and one CTA owns all fH image
// rows of WC neighbouring columns: the fH pixels of a column hit the same voxel at every depth, so their
// gradient rows are L1 hits for all but the first.  The compact row index and the softmax weight of all
// the CTA's points are staged in shared memory first (one round of coalesced loads); the depth loop then
// reads shared memory, keeps LF gradient rows in flight and needs no control flow (dropped points read
// row 0 with weight 0).  Outputs are staged and written as runs of WC floats.
//
// (Measured and dropped in round 1: fetching the rows straight from an NCHW gradient inside this kernel, 83 us.)
struct GpxMagic { unsigned long long per, fH, npx, WC; };   // ceil(2^40 / x) of D*fH, fH, fH*WC, WC (lss_div20)
template <int CPL>
__global__ void __launch_bounds__(SPLAT_THREADS)
k_bwd_gather_px(Dims d, int bn_lo, int WC, int stage_rows, GpxMagic mg, const int32_t *__restrict__ prow, const float *__restrict__ prob_col,
                const float *__restrict__ ctx_t, const float *__restrict__ grows, float *__restrict__ grad_dn) {
    extern __shared__ __align__(16) float smem[];
    constexpr int LF = CPL <= 8 ? 4 : 2;
    constexpr int NT = SPLAT_THREADS;
    constexpr int C = 8 * CPL;                            // the dispatch guarantees d.C == 8 * CPL
    const int bn = bn_lo + blockIdx.y, w0 = blockIdx.x * WC;
    const int D = d.D, fH = d.fH, DC = D + C;
    const int npx = fH * WC;                              // pixels of the CTA (<= 32), group g <-> pixel (h, wl)
    float *s_p = smem;                                    // [32][D] softmax weight
    int *s_row = reinterpret_cast<int *>(smem + 32 * D);  // [32][D] compact row or -1
    float *s_gp = smem + 64 * D;                          // [32][D] <grad row, ctx>
    float *s_out = smem + 96 * D;                         // [D + C][npx] staged outputs
    float *s_g = smem + ((96 * D + DC * npx + 3) & ~3);   // [WC][D][C] rows of the columns' primary voxels (16-byte aligned)
    {   // the CTA's columns are one contiguous block [wl][D][fH] in the column-major arrays
        const int ncol = min(WC, d.fW - w0), per = D * fH;
        const size_t base = ((size_t)bn * d.fW + w0) * per;
        for (int i = threadIdx.x; i < WC * per; i += NT) {
            const int wl = (int)lss_div20((unsigned)i, mg.per), r = i - wl * per;
            const int dd = (int)lss_div20((unsigned)r, mg.fH), h = r - dd * fH;
            const bool ok = wl < ncol;
            s_p[(h * WC + wl) * D + dd] = ok ? __ldg(prob_col + base + i) : 0.f;
            s_row[(h * WC + wl) * D + dd] = ok ? __ldg(prow + base + i) : -1;
        }
    }
    const int lane = threadIdx.x & 31, gl = lane & 7;
    const int g = threadIdx.x >> 3;
    const int h = (int)lss_div20((unsigned)g, mg.WC), wl = g - h * WC;
    const bool active = g < npx && w0 + wl < d.fW;
    const size_t my_ctx = ((size_t)bn * d.HW + (active ? h * d.fW + w0 + wl : 0)) * C;
    float ctx[CPL], dctx[CPL];
    {
        const float4 *cp = reinterpret_cast<const float4 *>(ctx_t + my_ctx) + gl;
#pragma unroll
        for (int q = 0; q < CPL / 4; ++q) {
            const float4 v = __ldg(cp + 8 * q);
            ctx[4 * q] = v.x; ctx[4 * q + 1] = v.y; ctx[4 * q + 2] = v.z; ctx[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int a = 0; a < CPL; ++a) dctx[a] = 0.f;
    }
    lss_pdl_wait();                                       // the gradient rows come from k_bwd_rows_compact
    __syncthreads();
    const int gg = g < npx ? g : 0;                       // idle groups mirror group 0 (shuffles stay warp-uniform)
    const float *my_p = s_p + gg * D;
    const int *my_row = s_row + gg * D;
    const float4 *rows4 = reinterpret_cast<const float4 *>(grows) + gl;
    constexpr int c4 = C >> 2;
    if (stage_rows) {
        // Almost always the fH pixels of a column hit the same voxel at a given depth.  Fetch that "primary" row
        // (the one of pixel h = 0) ONCE per (column, depth), all of them in flight together, zeros where there is
        // none; the depth loop then reads shared memory only and has no control flow.
        // (cp.async, global -> shared without a register round trip: ALL the CTA's rows are in flight together; a
        // source size of 0 zero-fills the slot of a dropped point)
        const unsigned s_g_addr = (unsigned)__cvta_generic_to_shared(s_g);
        for (int i = threadIdx.x; i < WC * D * c4; i += NT) {
            const int cd = i / c4, q = i - cd * c4;       // cd = wl * D + dd; pixel (h = 0, wl) has index wl
            const int r = s_row[cd];
            const float4 *src = reinterpret_cast<const float4 *>(grows) + (size_t)max(r, 0) * c4 + q;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(s_g_addr + 16u * (unsigned)i), "l"(src), "r"(r >= 0 ? 16 : 0) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const int *col_row = s_row + (g < npx ? wl : 0) * D;
        const float4 *my_sg = reinterpret_cast<const float4 *>(s_g + (size_t)(g < npx ? wl : 0) * D * C) + gl;
        unsigned exc = 0;                                 // does this pixel hit other voxels than the primary ones?
#pragma unroll 4
        for (int dd = 0; dd < D; ++dd) {
            const int rj = my_row[dd];
            const bool same = rj == col_row[dd];
            exc |= (active && rj >= 0 && !same) ? 1u : 0u;
            const float pj = (active && same) ? my_p[dd] : 0.f;       // dropped: the staged row is zero anyway
            float dot = 0.f;
#pragma unroll
            for (int q = 0; q < CPL / 4; ++q) {
                const float4 v = my_sg[dd * c4 + 8 * q];
                dot = fmaf(v.x, ctx[4 * q], dot); dot = fmaf(v.y, ctx[4 * q + 1], dot);
                dot = fmaf(v.z, ctx[4 * q + 2], dot); dot = fmaf(v.w, ctx[4 * q + 3], dot);
                dctx[4 * q] = fmaf(pj, v.x, dctx[4 * q]); dctx[4 * q + 1] = fmaf(pj, v.y, dctx[4 * q + 1]);
                dctx[4 * q + 2] = fmaf(pj, v.z, dctx[4 * q + 2]); dctx[4 * q + 3] = fmaf(pj, v.w, dctx[4 * q + 3]);
            }
            dot += __shfl_xor_sync(LSS_FULL_MASK, dot, 1);
            dot += __shfl_xor_sync(LSS_FULL_MASK, dot, 2);
            dot += __shfl_xor_sync(LSS_FULL_MASK, dot, 4);
            if (gl == 0 && g < npx) s_gp[g * D + dd] = (active && same) ? dot : 0.f;
        }
        // the exceptions (voxels shared with another column / camera at the frustum borders): rows from global memory
        if (__any_sync(LSS_FULL_MASK, exc != 0u)) {
            for (int dd = 0; dd < D; ++dd) {
                const int rj = my_row[dd];
                const bool mine = active && rj >= 0 && rj != col_row[dd];
                if (!__any_sync(LSS_FULL_MASK, mine)) continue;
                const float pj = mine ? my_p[dd] : 0.f;
                float dot = 0.f;
                if (mine) {
#pragma unroll
                    for (int q = 0; q < CPL / 4; ++q) {
                        const float4 v = __ldg(rows4 + (size_t)rj * c4 + 8 * q);
                        dot = fmaf(v.x, ctx[4 * q], dot); dot = fmaf(v.y, ctx[4 * q + 1], dot);
                        dot = fmaf(v.z, ctx[4 * q + 2], dot); dot = fmaf(v.w, ctx[4 * q + 3], dot);
                        dctx[4 * q] = fmaf(pj, v.x, dctx[4 * q]); dctx[4 * q + 1] = fmaf(pj, v.y, dctx[4 * q + 1]);
                        dctx[4 * q + 2] = fmaf(pj, v.z, dctx[4 * q + 2]); dctx[4 * q + 3] = fmaf(pj, v.w, dctx[4 * q + 3]);
                    }
                }
                dot += __shfl_xor_sync(LSS_FULL_MASK, dot, 1);
                dot += __shfl_xor_sync(LSS_FULL_MASK, dot, 2);
                dot += __shfl_xor_sync(LSS_FULL_MASK, dot, 4);
                if (mine && gl == 0) s_gp[g * D + dd] = dot;
            }
        }
    } else {
        // no room to stage the rows: LF rows in flight from global memory, dropped points read a finite stand-in
        const float4 *safe4 = reinterpret_cast<const float4 *>(ctx_t + my_ctx) + gl;
        for (int d0 = 0; d0 < D; d0 += LF) {
            float x[LF][CPL];
            float pj[LF];
            bool on[LF];
#pragma unroll
            for (int u = 0; u < LF; ++u) {
                const int dd = min(d0 + u, D - 1);
                const int rj = (active && d0 + u < D) ? my_row[dd] : -1;
                on[u] = rj >= 0;
                pj[u] = on[u] ? my_p[dd] : 0.f;
                const float4 *rp = on[u] ? rows4 + (size_t)rj * c4 : safe4;
#pragma unroll
                for (int q = 0; q < CPL / 4; ++q) {
                    const float4 v = __ldg(rp + 8 * q);
                    x[u][4 * q] = v.x; x[u][4 * q + 1] = v.y; x[u][4 * q + 2] = v.z; x[u][4 * q + 3] = v.w;
                }
            }
#pragma unroll
            for (int u = 0; u < LF; ++u) {
                float dot = 0.f;
#pragma unroll
                for (int a = 0; a < CPL; ++a) { dot = fmaf(x[u][a], ctx[a], dot); dctx[a] = fmaf(pj[u], x[u][a], dctx[a]); }
                dot += __shfl_xor_sync(LSS_FULL_MASK, dot, 1);
                dot += __shfl_xor_sync(LSS_FULL_MASK, dot, 2);
                dot += __shfl_xor_sync(LSS_FULL_MASK, dot, 4);
                if (gl == 0 && g < npx && d0 + u < D) s_gp[g * D + d0 + u] = on[u] ? dot : 0.f;
            }
        }
    }
    __syncwarp();
    // softmax backward (models.py:50): d_logit_d = p_d * (gp_d - sum_d' p_d' gp_d'); dropped points take part
    {
        const float *pp = s_p + gg * D, *gp = s_gp + gg * D;
        float sd = 0.f;
        for (int dd = gl; dd < D; dd += 8) sd = fmaf(pp[dd], gp[dd], sd);
        sd += __shfl_xor_sync(LSS_FULL_MASK, sd, 1);
        sd += __shfl_xor_sync(LSS_FULL_MASK, sd, 2);
        sd += __shfl_xor_sync(LSS_FULL_MASK, sd, 4);
        if (g < npx) {
            for (int dd = gl; dd < D; dd += 8) s_out[dd * npx + g] = pp[dd] * (gp[dd] - sd);
#pragma unroll
            for (int q = 0; q < CPL / 4; ++q)
#pragma unroll
                for (int a = 0; a < 4; ++a) s_out[(D + 4 * (gl + 8 * q) + a) * npx + g] = dctx[4 * q + a];
        }
    }
    __syncthreads();
    float *out = grad_dn + (size_t)bn * DC * d.HW + w0;
    for (int i = threadIdx.x; i < DC * npx; i += NT) {
        const int c = (int)lss_div20((unsigned)i, mg.npx), px = i - c * npx;
        const int hh = (int)lss_div20((unsigned)px, mg.WC), ww = px - hh * WC;
        if (w0 + ww < d.fW) out[(size_t)c * d.HW + hh * d.fW + ww] = s_out[i];
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

static int num_sms() {     // of the CURRENT device (one process may drive several GPUs)
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    return n;
}

// cudaFuncSetAttribute applies to the current device and is cheap: no per-process cache
template <typename K>
static int opt_in_smem_dev(K kern, size_t smem) {
    if (smem > 227 * 1024) return LSS_ERR_UNSUPPORTED;
    if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return LSS_ERR_CUDA;
    return LSS_OK;
}

template <int VW, int KC, bool ATOMIC, bool CL, bool DENSE, bool VEC4>
static int launch_fwd_tile(const Dims &d, const Tiling &tl, const int32_t *tile_start, const uint32_t *entries,
                           const SrcArgs &src, float *bev, cudaStream_t s) {
    const size_t smem = (size_t)(CL ? tl.TY * d.C : d.C * (tl.TY + 4)) * 4;
    auto kern = k_splat_fwd_tile<VW, KC, ATOMIC, CL, DENSE, VEC4>;
    int st = opt_in_smem_dev(kern, smem);
    if (st != LSS_OK) return st;
    kern<<<tl.n_tiles, SPLAT_THREADS, smem, s>>>(d, tl, tile_start, entries, src, bev);
    LSS_CHECK_LAUNCH();
    return LSS_OK;
}

struct PlanPtrs { const int32_t *vox; const uint32_t *entries; const int32_t *tile_start; const uint32_t *segs; const int32_t *tile_nseg, *tile_row0, *counters, *key_count; const int4 *seg_recs, *mixed_recs; };

static inline PlanPtrs plan_ptrs(const lss_plan_layout *L, const void *workspace) {
    const char *w = (const char *)workspace;
    PlanPtrs pp;
    pp.vox = (const int32_t *)(w + L->off_vox);
    pp.entries = (const uint32_t *)(w + L->off_entries);
    pp.tile_start = (const int32_t *)(w + L->off_tile_start);
    pp.segs = (const uint32_t *)(w + L->off_segs);
    pp.tile_nseg = (const int32_t *)(w + L->off_tile_nseg);
    pp.tile_row0 = (const int32_t *)(w + L->off_tile_row0);
    pp.counters = (const int32_t *)(w + L->off_counters);
    pp.key_count = (const int32_t *)(w + L->off_key_count);
    pp.seg_recs = (const int4 *)(w + L->off_seg_recs);
    pp.mixed_recs = (const int4 *)(w + L->off_mixed_recs);
    return pp;
}

template <bool CL, bool VEC4>
static int launch_fwd_store(const Dims &d, const Tiling &tl, const PlanPtrs &pp, const float *vsum, float *bev, int b0, int b1,
                            bool pdl, cudaStream_t s) {
    const int tps = tl.n_tiles / d.B;                     // tiles per sample
    // NCHW with 16-byte rows: the staging-free kernel (36.1 vs 37.9 us forward at cfg 2)
    if (!CL && VEC4 && tl.TY <= 32767) {
        const size_t rsm = (size_t)(ROWS_CAP * (d.C + 1) + 1) * 4 + (size_t)((tl.TY + 7) / 8) * 16;
        if (lss_launch(k_fwd_store_rows, dim3((b1 - b0) * tps), dim3(SPLAT_THREADS), rsm, s, pdl, d, tl, b0 * tps,
                       pp.tile_start, pp.tile_nseg, pp.tile_row0, pp.segs, vsum, bev) != cudaSuccess) return LSS_ERR_CUDA;
        LSS_CHECK_LAUNCH();
        return LSS_OK;
    }
    const size_t smem = (size_t)(CL ? tl.TY * d.C : d.C * (tl.TY + 4)) * 4;
    auto kern = k_fwd_store<CL, VEC4, SPLAT_THREADS>;
    int st = opt_in_smem_dev(kern, smem);
    if (st != LSS_OK) return st;
    if (lss_launch(kern, dim3((b1 - b0) * tps, 1), dim3(SPLAT_THREADS), smem, s, pdl, d, tl, b0 * tps, d.C,
                   pp.tile_start, pp.tile_nseg, pp.tile_row0, pp.segs, vsum, bev) != cudaSuccess) return LSS_ERR_CUDA;
    LSS_CHECK_LAUNCH();
    return LSS_OK;
}

template <bool CL>
static int launch_fwd_store_tma(const Dims &d, const Tiling &tl, const PlanPtrs &pp, const float *vsum, float *bev, int b0, int b1,
                                cudaStream_t s) {
    const int tile_floats = CL ? tl.TY * d.C : d.C * (tl.TY + 4);
    const size_t smem = (size_t)2 * tile_floats * 4;
    auto kern = k_fwd_store_tma<CL>;
    int st = opt_in_smem_dev(kern, smem);
    if (st != LSS_OK) return st;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SPLAT_THREADS, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    const int tps = tl.n_tiles / d.B, n_tiles = (b1 - b0) * tps;
    const int grid = min(n_tiles, num_sms() * per_sm);
    kern<<<grid, SPLAT_THREADS, smem, s>>>(d, tl, b0 * tps, n_tiles, pp.tile_start, pp.tile_nseg, pp.tile_row0, pp.segs, vsum, bev);
    LSS_CHECK_LAUNCH();
    return LSS_OK;
}

// GROUP variant of the deterministic forward: gather into compact rows, then stream the tiles out
static int run_fwd_group(bool cl, bool vec4, const Dims &d, const Tiling &tl, const PlanPtrs &pp, long long L_rows_cap,
                         const float *prob, const float *prob_col, const float *ctx_t, float *vsum, float *bev,
                         int variant, int b0, int b1, cudaStream_t s) {
    if (variant != LSS_VARIANT_GROUP_STORE) {
        const int key_lo = b0 * d.N * d.fW;
        const int n_keys = (b1 - b0) * d.N * d.fW;       // one CTA per camera column of the samples [b0, b1) ...
        // ... plus the CTAs that drain the queue of mixed voxels (of ALL samples): they ride with the part that starts at 0
        const int grid = n_keys + (b0 == 0 ? 2 * num_sms() : 0);
        const size_t gsm = max((size_t)(d.fH * d.C + d.D * d.fH), (size_t)(GATHER_THREADS / 8) * d.C) * 4;   // column operands / long-voxel products
        if (gsm > 48 * 1024) return LSS_ERR_UNSUPPORTED;
#define GATHER_ARGS d, key_lo, n_keys, pp.key_count, pp.seg_recs, pp.counters, pp.mixed_recs, L_rows_cap, pp.entries, prob, prob_col, ctx_t, vsum
        // programmatic launch: the grid is scheduled while its predecessor drains and waits at its top (launch latency only)
        cudaError_t ge;
        if (d.C == 32) ge = lss_launch(k_fwd_gather<4>, dim3(grid), dim3(GATHER_THREADS), gsm, s, true, GATHER_ARGS);
        else if (d.C == 64) ge = lss_launch(k_fwd_gather<8>, dim3(grid), dim3(GATHER_THREADS), gsm, s, true, GATHER_ARGS);
        else ge = lss_launch(k_fwd_gather<16>, dim3(grid), dim3(GATHER_THREADS), gsm, s, true, GATHER_ARGS);
        if (ge != cudaSuccess) return LSS_ERR_CUDA;
#undef GATHER_ARGS
        LSS_CHECK_LAUNCH();
        if (variant == LSS_VARIANT_GROUP_GATHER) return LSS_OK;
    }
    // persistent bulk-copy variant: measured faster only where a tile is ONE contiguous run (channels_last, nz = 1:
    // 36.7 vs 38.7 us forward at cfg 2); with one 800-byte bulk copy per channel row (NCHW) it is slower (42.7 us)
    const size_t tma_smem = (size_t)2 * tl.TY * d.C * 4;
    if (cl && d.nz == 1 && vec4 && tma_smem <= 227 * 1024) return launch_fwd_store_tma<true>(d, tl, pp, vsum, bev, b0, b1, s);
    const bool pdl = variant != LSS_VARIANT_GROUP_STORE;  // only right behind its gather
    if (cl) return vec4 ? launch_fwd_store<true, true>(d, tl, pp, vsum, bev, b0, b1, pdl, s) : launch_fwd_store<true, false>(d, tl, pp, vsum, bev, b0, b1, pdl, s);
    return vec4 ? launch_fwd_store<false, true>(d, tl, pp, vsum, bev, b0, b1, pdl, s) : launch_fwd_store<false, false>(d, tl, pp, vsum, bev, b0, b1, pdl, s);
}

template <int VW, int KC, bool DENSE>
static int dispatch_fwd_flags(bool atomic, bool cl, bool vec4, const Dims &d, const Tiling &tl, const int32_t *ts,
                              const uint32_t *en, const SrcArgs &src, float *bev, cudaStream_t s) {
#define FWD_CASE(A, L, V) return launch_fwd_tile<VW, KC, A, L, DENSE, V>(d, tl, ts, en, src, bev, s)
    if (atomic) { if (cl) { if (vec4) FWD_CASE(true, true, true); else FWD_CASE(true, true, false); }
                  else    { if (vec4) FWD_CASE(true, false, true); else FWD_CASE(true, false, false); } }
    else        { if (cl) { if (vec4) FWD_CASE(false, true, true); else FWD_CASE(false, true, false); }
                  else    { if (vec4) FWD_CASE(false, false, true); else FWD_CASE(false, false, false); } }
#undef FWD_CASE
}

// vector stores of the tile need 16-byte aligned rows in global memory
static bool tile_vec4_ok(const Dims &d, const Tiling &tl, bool cl, const float *bev) {
    if (!lss_aligned(bev, 16)) return false;
    if (cl) return d.C % 4 == 0;
    return d.ny % 4 == 0 && tl.TY % 4 == 0;
}

template <bool DENSE>
static int dispatch_fwd(bool atomic, bool cl, int variant, const Dims &d, const Tiling &tl, const PlanPtrs &pp,
                        long long rows_cap, const SrcArgs &src, float *vsum, float *bev, int b0, int b1, cudaStream_t s) {
    const bool vec4 = tile_vec4_ok(d, tl, cl, bev);
    const bool rows16 = lss_aligned(src.base, 16);
    // GROUP kernels: deterministic mode, fused operands, 8 lanes x C/8 channels, 32-bit row offsets
    const bool group_ok = !DENSE && !atomic && vsum != nullptr && rows16 && lss_aligned(vsum, 16) &&
                          (d.C == 32 || d.C == 64 || d.C == 128) && (long long)d.N * d.HW * d.C < (1ll << 31) &&
                          (size_t)(d.fH * d.C + d.D * d.fH) * 4 <= 48 * 1024;
    if (variant >= LSS_VARIANT_GROUP && !group_ok) return LSS_ERR_UNSUPPORTED;
    if (group_ok && variant != LSS_VARIANT_WARP) return run_fwd_group(cl, vec4, d, tl, pp, rows_cap, src.prob, src.prob_col, src.base, vsum, bev, variant, b0, b1, s);
    if (b0 != 0 || b1 != d.B) return LSS_ERR_UNSUPPORTED;   // only the GROUP kernels take a sample range
    const int32_t *ts = pp.tile_start;
    const uint32_t *en = pp.entries;
    if (!DENSE && rows16) {
        if (d.C == 32) return dispatch_fwd_flags<1, 1, DENSE>(atomic, cl, vec4, d, tl, ts, en, src, bev, s);
        if (d.C == 64) return dispatch_fwd_flags<2, 1, DENSE>(atomic, cl, vec4, d, tl, ts, en, src, bev, s);
        if (d.C == 128) return dispatch_fwd_flags<4, 1, DENSE>(atomic, cl, vec4, d, tl, ts, en, src, bev, s);
    }
    const int kc = lss_kc_for(d.C);
    if (kc <= 2) return dispatch_fwd_flags<0, 2, DENSE>(atomic, cl, vec4, d, tl, ts, en, src, bev, s);
    if (kc <= 4) return dispatch_fwd_flags<0, 4, DENSE>(atomic, cl, vec4, d, tl, ts, en, src, bev, s);
    return dispatch_fwd_flags<0, 8, DENSE>(atomic, cl, vec4, d, tl, ts, en, src, bev, s);
}

static size_t bev_elems(const Dims &d) { return (size_t)d.B * d.nz * d.C * d.nx * d.ny; }

extern "C" int lss_bev_clear(const lss_problem *p, float *bev, void *stream) {
    int st = lss_check_problem(p);
    if (st != LSS_OK) return st;
    LSS_REQUIRE(bev != nullptr, LSS_ERR_BAD_ARG);
    const Dims d = make_dims(p);
    return cudaMemsetAsync(bev, 0, bev_elems(d) * 4, (cudaStream_t)stream) == cudaSuccess ? LSS_OK : LSS_ERR_CUDA;
}

extern "C" int lss_splat_fwd(const lss_problem *p, const lss_plan_layout *L, const void *workspace, const float *prob,
                             const float *ctx_t, const float *prob_col, float *voxel_sums, float *bev, int mode,
                             int layout, int variant, int precleared, int b0, int b1, void *stream) {
    int st = lss_check_problem(p);
    if (st != LSS_OK) return st;
    LSS_REQUIRE(L && workspace, LSS_ERR_WORKSPACE);
    LSS_REQUIRE(prob && ctx_t && bev, LSS_ERR_BAD_ARG);
    LSS_REQUIRE(layout == LSS_LAYOUT_NCHW || layout == LSS_LAYOUT_CHANNELS_LAST, LSS_ERR_BAD_ARG);
    const Dims d = make_dims(p);
    if (b1 <= 0) { b0 = 0; b1 = d.B; }                    // (0, 0) = all samples
    LSS_REQUIRE(0 <= b0 && b0 < b1 && b1 <= d.B, LSS_ERR_BAD_ARG);
    LSS_REQUIRE(L->n_points == d.n_points, LSS_ERR_WORKSPACE);
    const Tiling tl = make_tiling(L);
    const char *w = (const char *)workspace;
    const int32_t *vox = (const int32_t *)(w + L->off_vox);
    const uint32_t *entries = (const uint32_t *)(w + L->off_entries);
    const int32_t *tile_start = (const int32_t *)(w + L->off_tile_start);
    cudaStream_t s = (cudaStream_t)stream;
    const bool cl = layout == LSS_LAYOUT_CHANNELS_LAST;
    SrcArgs src{};
    src.base = ctx_t; src.prob = prob; src.prob_col = prob_col;
    if (mode == LSS_SPLAT_SORTED || mode == LSS_SPLAT_SMEM_ATOMIC)
        return dispatch_fwd<false>(mode == LSS_SPLAT_SMEM_ATOMIC, cl, variant, d, tl, plan_ptrs(L, workspace), L->n_rows_cap, src, voxel_sums, bev, b0, b1, s);
    if (mode == LSS_SPLAT_RED_GLOBAL) {
        LSS_REQUIRE(b0 == 0 && b1 == d.B, LSS_ERR_UNSUPPORTED);
        if (!precleared && cudaMemsetAsync(bev, 0, bev_elems(d) * 4, s) != cudaSuccess) return LSS_ERR_CUDA;
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
                                     int variant, int precleared, void *stream) {
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
        return dispatch_fwd<true>(mode == LSS_SPLAT_SMEM_ATOMIC, cl, variant, d, tl, plan_ptrs(L, workspace), L->n_rows_cap, src, nullptr, bev, 0, d.B, s);
    if (mode == LSS_SPLAT_RED_GLOBAL) {
        if (!precleared && cudaMemsetAsync(bev, 0, bev_elems(d) * 4, s) != cudaSuccess) return LSS_ERR_CUDA;
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
    const size_t smem = (size_t)((tl.TY + 31) / 32) * 4;
    k_bwd_rows_nchw<<<tl.n_tiles, SPLAT_THREADS, smem, s>>>(d, tl, tile_start, entries, grad_bev, rows);
    LSS_CHECK_LAUNCH();
    return LSS_OK;
}

template <int VW, int KC, int NCH, bool CL>
static int launch_gather_one(int grid, size_t smem, cudaStream_t s, const Dims &d, const int32_t *vox, const float *prob,
                             const float *ctx_t, const float *rows, float *grad_dn) {
    auto kern = k_bwd_gather<VW, KC, NCH, CL>;
    const int st = opt_in_smem_dev(kern, smem);
    if (st != LSS_OK) return st;
    kern<<<grid, SPLAT_THREADS, smem, s>>>(d, vox, prob, ctx_t, rows, grad_dn);
    LSS_CHECK_LAUNCH();
    return LSS_OK;
}

template <int VW, int KC>
static int dispatch_gather(bool cl, cudaStream_t s, const Dims &d, const int32_t *vox, const float *prob,
                           const float *ctx_t, const float *rows, float *grad_dn) {
    const int npix = d.B * d.N * d.HW;
    const int grid = (npix + SPLAT_WARPS - 1) / SPLAT_WARPS;
    const size_t smem = ((size_t)SPLAT_WARPS * 32 * 33 + (size_t)(d.D + d.C) * SPLAT_WARPS) * 4;
    const int nch = (d.D + 31) / 32;
#define G_CASE(N) return cl ? launch_gather_one<VW, KC, N, true>(grid, smem, s, d, vox, prob, ctx_t, rows, grad_dn) \
                            : launch_gather_one<VW, KC, N, false>(grid, smem, s, d, vox, prob, ctx_t, rows, grad_dn)
    if (nch <= 2) G_CASE(2);
    if (nch <= 4) G_CASE(4);
    G_CASE(8);
#undef G_CASE
}

static int run_gather(bool cl, cudaStream_t s, const Dims &d, const int32_t *vox, const float *prob, const float *ctx_t,
                      const float *rows, float *grad_dn) {
    const bool vec_ok = lss_aligned(ctx_t, 16) && lss_aligned(rows, 16) && (!cl || (d.C % 4 == 0));
    if (vec_ok && d.C == 32) return dispatch_gather<1, 1>(cl, s, d, vox, prob, ctx_t, rows, grad_dn);
    if (vec_ok && d.C == 64) return dispatch_gather<2, 1>(cl, s, d, vox, prob, ctx_t, rows, grad_dn);
    if (vec_ok && d.C == 128) return dispatch_gather<4, 1>(cl, s, d, vox, prob, ctx_t, rows, grad_dn);
    const int kc = lss_kc_for(d.C);
    if (kc <= 2) return dispatch_gather<0, 2>(cl, s, d, vox, prob, ctx_t, rows, grad_dn);
    if (kc <= 4) return dispatch_gather<0, 4>(cl, s, d, vox, prob, ctx_t, rows, grad_dn);
    return dispatch_gather<0, 8>(cl, s, d, vox, prob, ctx_t, rows, grad_dn);
}

static bool bwd_compact_ok(const Dims &d, const float *ctx_t, const float *grows) {
    return (d.C == 32 || d.C == 64 || d.C == 128) && d.fH <= 32 && lss_aligned(ctx_t, 16) && lss_aligned(grows, 16);
}

int lss_bwd_gather_rows(const Dims &d, const int32_t *prow, const float *prob_col, const float *ctx_t, const float *rows,
                        float *grad_dn, int b0, int b1, bool pdl, cudaStream_t s) {
    if (!bwd_compact_ok(d, ctx_t, rows)) return LSS_ERR_UNSUPPORTED;
    const int WC = max(1, 32 / d.fH);
    const dim3 grid((d.fW + WC - 1) / WC, (b1 - b0) * d.N);
    auto magic = [](unsigned x) { return ((1ull << 40) + x - 1) / x; };
    const GpxMagic mg{magic((unsigned)(d.D * d.fH)), magic((unsigned)d.fH), magic((unsigned)(d.fH * WC)), magic((unsigned)WC)};
    size_t smem = ((size_t)96 * d.D + (size_t)(d.D + d.C) * d.fH * WC) * 4;
    if (smem > 227 * 1024) return LSS_ERR_UNSUPPORTED;
    const size_t rows_smem = (size_t)WC * d.D * d.C * 4 + 16;    // staged gradient rows (+ alignment), if they fit next to a second CTA
    const int stage_rows = smem + rows_smem <= 100 * 1024;
    if (stage_rows) smem += rows_smem;
#define GPX(CPL)                                                                                                  \
    do {                                                                                                          \
        if (smem > 48 * 1024 && cudaFuncSetAttribute(k_bwd_gather_px<CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) \
            return LSS_ERR_CUDA;                                                                                  \
        if (lss_launch(k_bwd_gather_px<CPL>, grid, dim3(SPLAT_THREADS), smem, s, pdl, d, b0 * d.N, WC, stage_rows, mg, \
                       prow, prob_col, ctx_t, rows, grad_dn) != cudaSuccess) return LSS_ERR_CUDA;                 \
    } while (0)
    if (d.C == 32) GPX(4); else if (d.C == 64) GPX(8); else GPX(16);
#undef GPX
    LSS_CHECK_LAUNCH();
    return LSS_OK;
}

// compact-row backward (sorted plans): gradient rows of the non-empty voxels, then the pixel-owner gather
static int run_bwd_compact(bool cl, const Dims &d, const Tiling &tl, const PlanPtrs &pp, const int32_t *prow,
                           const float *grad_bev, const float *prob_col, const float *ctx_t, float *grows, float *grad_dn,
                           int stage, int b0, int b1, cudaStream_t s) {
    const int tps = tl.n_tiles / d.B;                     // tiles per sample
    const int WC = max(1, 32 / d.fH);
    const dim3 grid((d.fW + WC - 1) / WC, (b1 - b0) * d.N);
    auto magic = [](unsigned x) { return ((1ull << 40) + x - 1) / x; };
    const GpxMagic mg{magic((unsigned)(d.D * d.fH)), magic((unsigned)d.fH), magic((unsigned)(d.fH * WC)), magic((unsigned)WC)};
    size_t smem = ((size_t)96 * d.D + (size_t)(d.D + d.C) * d.fH * WC) * 4;
    if (smem > 227 * 1024) return LSS_ERR_UNSUPPORTED;
    const size_t rows_smem = (size_t)WC * d.D * d.C * 4 + 16 + (size_t)WC * d.D * 4;  // staged gradient rows (+ alignment, + offsets), if they fit next to a second CTA
    const int stage_rows = smem + rows_smem <= 100 * 1024;
    if (stage_rows) smem += rows_smem;
    if (stage != 2) {
        const cudaError_t e = cl
            ? lss_launch(k_bwd_rows_compact<true>, dim3((b1 - b0) * tps), dim3(SPLAT_THREADS), 0, s, true, d, tl, b0 * tps,
                         pp.tile_start, pp.tile_nseg, pp.tile_row0, pp.segs, grad_bev, grows)
            : lss_launch(k_bwd_rows_compact<false>, dim3((b1 - b0) * tps), dim3(SPLAT_THREADS), 0, s, true, d, tl, b0 * tps,
                         pp.tile_start, pp.tile_nseg, pp.tile_row0, pp.segs, grad_bev, grows);
        if (e != cudaSuccess) return LSS_ERR_CUDA;
        LSS_CHECK_LAUNCH();
        if (stage == 1) return LSS_OK;
    }
    return lss_bwd_gather_rows(d, prow, prob_col, ctx_t, grows, grad_dn, b0, b1, stage != 2, s);
}

extern "C" int lss_splat_bwd(const lss_problem *p, const lss_plan_layout *L, const void *workspace,
                             const float *grad_bev, int layout, const float *prob, const float *ctx_t,
                             const float *prob_col, float *grad_rows, float *grad_depthnet, int plan_sorted, int stage,
                             int b0, int b1, void *stream) {
    int st = lss_check_problem(p);
    if (st != LSS_OK) return st;
    LSS_REQUIRE(L && workspace, LSS_ERR_WORKSPACE);
    LSS_REQUIRE(grad_bev && prob && ctx_t && grad_depthnet, LSS_ERR_BAD_ARG);
    LSS_REQUIRE(layout == LSS_LAYOUT_NCHW || layout == LSS_LAYOUT_CHANNELS_LAST, LSS_ERR_BAD_ARG);
    LSS_REQUIRE(p->D <= LSS_MAX_DEPTH, LSS_ERR_UNSUPPORTED);
    const Dims d = make_dims(p);
    if (b1 <= 0) { b0 = 0; b1 = d.B; }                    // (0, 0) = all samples
    LSS_REQUIRE(0 <= b0 && b0 < b1 && b1 <= d.B && stage >= 0 && stage <= 2, LSS_ERR_BAD_ARG);
    LSS_REQUIRE(L->n_points == d.n_points, LSS_ERR_WORKSPACE);
    const Tiling tl = make_tiling(L);
    const PlanPtrs pp = plan_ptrs(L, workspace);
    cudaStream_t s = (cudaStream_t)stream;
    const bool cl = layout == LSS_LAYOUT_CHANNELS_LAST;
    if (plan_sorted && prob_col != nullptr && grad_rows != nullptr && bwd_compact_ok(d, ctx_t, grad_rows))
        return run_bwd_compact(cl, d, tl, pp, (const int32_t *)((const char *)workspace + L->off_prow), grad_bev, prob_col,
                               ctx_t, grad_rows, grad_depthnet, stage, b0, b1, s);
    LSS_REQUIRE(stage == 0 && b0 == 0 && b1 == d.B, LSS_ERR_UNSUPPORTED);   // only the compact-row kernels take parts
    const float *rows = grad_bev;
    if (!cl) {
        LSS_REQUIRE(grad_rows != nullptr, LSS_ERR_WORKSPACE);
        st = launch_bwd_rows(d, tl, pp.tile_start, pp.entries, grad_bev, grad_rows, s);
        if (st != LSS_OK) return st;
        rows = grad_rows;
    }
    return run_gather(cl, s, d, pp.vox, prob, ctx_t, rows, grad_depthnet);
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
