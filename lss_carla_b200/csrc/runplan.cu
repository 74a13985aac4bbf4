// Run plan + channels_last forward: the fast path of the fused lift-splat (default of bench.py and of the API).
//
// Replaces, for shdragron/LSS-Carla (paths under the reference root):
//   LiftSplatShoot.get_geometry                      src/models.py:170-190   (geometry in registers)
//   voxel_pooling: quantise, mask, rank, argsort     src/models.py:212-231   (no sort: runs + per-voxel counts)
//   the lift outer product + QuickCumsum + griddify  src/models.py:59, :234-244, src/tools.py:193-209
//
// A RUN is the fH image rows of one (camera, feature column, depth bin): consecutive points of the camera-column-major
// order cm = ((bn*fW + w)*D + d)*fH + h.  A warp owns 32/fH whole runs, lane = image row.  The points of a run that share
// a voxel form a SUB-RUN; its first lane is the leader.  With the BEV in channels_last a voxel is one contiguous C-float
// row, so whoever owns a voxel writes it directly: no tile-owner store pass, no compact rows, and nothing to sort --
//   k_run_index     voxel row per point (bit-exact arithmetic of geom.cuh) -> prow; every leader pushes its sub-run on the
//                   voxel's list (one atomicExch: sub[leader] = {previous head, row mask}) and adds its size to cnt[voxel]
//   k_run_classify  no atomics on the common path, no fences: a leader that pushed FIRST (previous head empty) and is still
//                   the head owns the whole voxel (EXCLUSIVE, emask[leader] = its row mask); a first pusher that is no
//                   longer the head appends the voxel to the queue of shared voxels {list head, points, row, batch}
//   k_fwd_gather_cl one CTA per camera column sums its exclusive sub-runs out of staged operands (8-lane groups, lane =
//                   C/8 channels); queue CTAs walk the short lists of the shared voxels, sort their few points by flat
//                   index and sum them from global operands; voxels with >= 64 points are summed by a whole CTA
// Per voxel the result is acc = 0; for p ascending in flat (b,n,d,h,w) index: acc = fl32(acc + fl32(prob[p]*ctx[p])) --
// the definition of LSS_SPLAT_SORTED (the reference's stable argsort order), bit for bit.
#include "common.cuh"
#include "geom.cuh"
#include "lift.cuh"

#define RP_THREADS 256
#define RP_WARPS (RP_THREADS / 32)
#define RP_U 2                     // warp-rounds per thread (measured at cfg 2: 2 -> 61.0 us per step, 3 -> 63.2, 4 -> 64.1)
#define GCL_THREADS 128
#define GCL_NG (GCL_THREADS / 8)   // 8-lane groups per gather CTA
#define GCL_SHORT_CAP 64           // == LSS_LONG_VOXEL: shared voxels below this are summed by a group
#define GCL_SORT_CAP 1024          // long voxels up to this many points are sorted in shared memory

// Profiling aid, compiled in with -DLSS_RP_TIMELINE only (scripts/bench_runplan_quick.py): earliest start / latest end
// (globaltimer ns) of the zero role (0), the index role (1), k_run_classify (2) and the column CTAs of the gather (3).
#ifdef LSS_RP_TIMELINE
__device__ unsigned long long g_rp_tl[8] = {~0ull, 0, ~0ull, 0, ~0ull, 0, ~0ull, 0};
__device__ int g_rp_tl_on = 0;
__device__ __forceinline__ void tl_stamp(int k, bool end) {
    if (!g_rp_tl_on || threadIdx.x != 0) return;
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (end) atomicMax(g_rp_tl + 2 * k + 1, t); else atomicMin(g_rp_tl + 2 * k, t);
}
extern "C" int lss_debug_runplan_timeline(int on, unsigned long long *out_host) {
    unsigned long long h[8];
    if (out_host) { if (cudaMemcpyFromSymbol(h, g_rp_tl, sizeof(h)) != cudaSuccess) return -4; for (int i = 0; i < 8; ++i) out_host[i] = h[i]; }
    const unsigned long long init[8] = {~0ull, 0, ~0ull, 0, ~0ull, 0, ~0ull, 0};
    if (cudaMemcpyToSymbol(g_rp_tl, init, sizeof(init)) != cudaSuccess) return -4;
    return cudaMemcpyToSymbol(g_rp_tl_on, &on, sizeof(on)) == cudaSuccess ? 0 : -4;
}
#else
#define tl_stamp(k, end) ((void)0)
#endif

// Bounds / invariant checks of the run-plan kernels, compiled in with -DLSS_DEVICE_ASSERTS only (scripts/run_with_asserts.py:
// compute-sanitizer is closed on this GPU pool, so index ranges, list integrity and capacity limits are checked by the kernels
// themselves on the small cases; a violated check traps and the next CUDA call fails).
#ifdef LSS_DEVICE_ASSERTS
#include <cstdio>
#define LSS_DASSERT(cond) do { if (!(cond)) { printf("LSS_DASSERT failed: %s (%s:%d) block %d thread %d\n", #cond, __FILE__, __LINE__, (int)blockIdx.x, (int)threadIdx.x); __trap(); } } while (0)
#else
#define LSS_DASSERT(cond) ((void)0)
#endif

// thread -> point mapping shared by k_run_index and k_run_classify (they must agree on what a sub-run is)
struct RunDims {
    int fH, RPW;        // runs per warp = 32 / fH
    int R;              // runs = B*N*fW*D
    int fWD;            // fW*D: runs per camera
    unsigned fmask;     // fH low bits
};

struct RunLane { int r, h, rw; bool valid; unsigned run_mask; };

__device__ __forceinline__ RunLane run_lane(const RunDims &rd, int cta, int u) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    RunLane q;
    q.rw = lane / rd.fH;
    q.h = lane - q.rw * rd.fH;
    const int wg = (cta * RP_WARPS + warp) * RP_U + u;
    q.r = wg * rd.RPW + q.rw;
    q.valid = q.rw < rd.RPW && q.r < rd.R;
    q.run_mask = q.rw < rd.RPW ? rd.fmask << (q.rw * rd.fH) : 0u;
    return q;
}

// ------------------------------------------------------------------------------------------------
// plan kernels
// ------------------------------------------------------------------------------------------------

// The work of one CTA of the index pass (RP_THREADS threads); `cta` in [0, ceil(R / runs per CTA)).
template <bool RAW>
__device__ __forceinline__ void run_index_cta(const Dims &d, const RunDims &rd, const CalibPtrs &c, int32_t *__restrict__ prow,
                                              int32_t *__restrict__ cnt, int32_t *__restrict__ head, int2 *__restrict__ sub,
                                              int32_t *__restrict__ counters, int cta) {
    tl_stamp(1, false);
    if (cta == 0 && threadIdx.x < 4) counters[threadIdx.x] = 0;
    __shared__ float s_m[RAW ? LSS_RAW_CAMS : 1][18];
    const int cam0 = (int)(((long long)cta * RP_WARPS * RP_U * rd.RPW) / rd.fWD);
    if (RAW) {      // the calibration matrices of the few cameras this CTA touches, made on the fly (no extra launch)
        const int cam = cam0 + (int)threadIdx.x;
        if (threadIdx.x < LSS_RAW_CAMS && cam < d.B * d.N) calib_matrices_of(c.rots, c.intrins, c.post_rots, cam, s_m[threadIdx.x], s_m[threadIdx.x] + 9);
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int u = 0; u < RP_U; ++u) {
        const RunLane q = run_lane(rd, cta, u);
        int row = -1;
        if (q.valid) {
            const int bn = q.r / rd.fWD, rem = q.r - bn * rd.fWD;
            const int w = rem / d.D, dd = rem - w * d.D;
            const int in_cam = (dd * d.fH + q.h) * d.fW + w;
            float g[3];
            if (RAW) ego_point(c, bn, in_cam, g, s_m[bn - cam0], s_m[bn - cam0] + 9);
            else ego_point(c, bn, in_cam, g);
            long long ii[3];
            const int b = bn / d.N;
            if (voxel_of_point(d, b, g, ii) >= 0)
                row = ((b * d.nx + (int)ii[0]) * d.ny + (int)ii[1]) * d.nz + (int)ii[2];
            prow[(size_t)q.r * d.fH + q.h] = row;
        }
        const unsigned peers = __match_any_sync(LSS_FULL_MASK, row >= 0 ? row : -1 - lane) & q.run_mask;
        if (row >= 0 && lane == __ffs(peers) - 1) {      // sub-run leader: push on the voxel's list, count its points
            const size_t cm = (size_t)q.r * d.fH + q.h;
            LSS_DASSERT(row < d.B * d.nx * d.ny * d.nz && cm < (size_t)d.n_points);
            const int old = atomicExch(head + row, (int)cm + 1);
            LSS_DASSERT(old >= 0 && old <= d.n_points && old != (int)cm + 1);
            atomicAdd(cnt + row, __popc(peers));         // result unused: red.global
            sub[cm] = make_int2(old, (int)(peers >> lane));
        }
    }
    tl_stamp(1, true);
}

// Expand a sub-run (leader `cm` in camera-column-major order incl. the batch part, row mask relative to the leader) into
// the point-in-sample flat indices ((n*D + d)*fH + h)*fW + w of its points.  `nl` (a power of two) lanes cooperate: lane
// `jl` writes points jl, jl+nl, ...  Returns the number of points.
__device__ __forceinline__ int expand_subrun(const Dims &d, int fWD, int cm, unsigned mask, uint32_t *out, int jl, int nl) {
    const int r = cm / d.fH, h0 = cm - r * d.fH;
    const int bn = r / fWD, rem = r - bn * fWD;
    const int w = rem / d.D, dd = rem - w * d.D;
    const int n = bn % d.N;
    const unsigned base = (unsigned)((n * d.D + dd) * d.fH + h0) * (unsigned)d.fW + (unsigned)w;
    int j = 0;
    for (unsigned m = mask; m; m &= m - 1, ++j)
        if ((j & (nl - 1)) == jl) out[j] = base + (unsigned)(__ffs(m) - 1) * (unsigned)d.fW;
    return j;
}

__global__ void __launch_bounds__(RP_THREADS)
k_run_classify(Dims d, RunDims rd, const int32_t *__restrict__ prow, uint32_t *__restrict__ emask, int32_t *__restrict__ cnt,
               int32_t *__restrict__ head, const int2 *__restrict__ sub, uint32_t *__restrict__ pool,
               int4 *__restrict__ mixed_recs, int32_t *__restrict__ counters, long long n_mixed_cap) {
    lss_pdl_trigger();
    lss_pdl_wait();                                       // prow, cnt, head, sub and the cleared counters come from k_run_index
    tl_stamp(2, false);
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int u = 0; u < RP_U; ++u) {
        const RunLane q = run_lane(rd, (int)blockIdx.x, u);
        const size_t cm = (size_t)q.r * d.fH + q.h;
        const int row = q.valid ? __ldg(prow + cm) : -1;
        const unsigned peers = __match_any_sync(LSS_FULL_MASK, row >= 0 ? row : -1 - lane) & q.run_mask;
        unsigned em = 0u;
        if (row >= 0 && lane == __ffs(peers) - 1) {
            const int2 node = __ldcg(sub + cm);            // {previous head of the voxel's list, row mask}
            if (node.x == 0) {                            // this sub-run pushed first: it answers for the voxel
                const int hd = __ldcg(head + row);
                if (hd == (int)cm + 1) {                  // ... and nobody pushed after it: the voxel is this sub-run alone
                    em = (unsigned)node.y;
                } else {                                  // shared voxel: one queue record, made by the tail of its list
                    const int c = __ldcg(cnt + row);
                    const int b = row / (d.nx * d.ny * d.nz);
                    if (c >= GCL_SHORT_CAP) {             // long voxel: its point set is written out for the CTA path
                        const int pos = atomicAdd(counters + 1, c);
                        LSS_DASSERT(pos >= 0 && pos + c <= d.n_points);
                        int i = 0;
                        for (int cur = hd; cur != 0;) {
                            LSS_DASSERT(cur >= 1 && cur <= d.n_points);
                            const int2 nd = __ldcg(sub + (cur - 1));
                            LSS_DASSERT(nd.y != 0 && __ldg(prow + (cur - 1)) == row && i + __popc((unsigned)nd.y) <= c);
                            i += expand_subrun(d, rd.fWD, cur - 1, (unsigned)nd.y, pool + pos + i, 0, 1);
                            cur = nd.x;
                        }
                        LSS_DASSERT(i == c);
                        mixed_recs[n_mixed_cap - 1 - atomicAdd(counters + 2, 1)] = make_int4(pos, c, row, b);
                    } else {
                        LSS_DASSERT(c >= 2 && hd >= 1 && hd <= d.n_points && hd != (int)cm + 1);
                        const int slot = atomicAdd(counters, 1);
                        LSS_DASSERT(slot + __ldcg(counters + 2) < n_mixed_cap);
                        mixed_recs[slot] = make_int4(hd, c, row, b);
                    }
                }
                head[row] = 0;                            // scratch grids are left clean for the next build: only the
                cnt[row] = 0;                             // first pusher of a voxel reads them here
            }
        }
        if (q.valid) emask[cm] = em;
    }
    tl_stamp(2, true);
}

// ------------------------------------------------------------------------------------------------
// zero-fill through the bulk-copy engine
// ------------------------------------------------------------------------------------------------

// Zero role of one CTA: thread 0 streams its share of [dst, dst + bytes) out of a zeroed shared-memory chunk with bulk copies
// (cp.async.bulk shared -> global: the copies need no registers and no issue slots, the source is read-only, so all of them
// stay in flight).  With `evict_first` the lines are marked evict-first in L2: 82 MB of zeros should not push the plan and the
// lift operands, which the gather is about to read, out of the cache.
#define ZERO_CHUNK (16 * 1024)
__device__ __forceinline__ void zero_role(float *__restrict__ dst, size_t bytes, int cta, int n_cta, float *s_zero, int evict_first) {
    tl_stamp(0, false);
    for (int i = threadIdx.x; i < ZERO_CHUNK / 16; i += blockDim.x) reinterpret_cast<float4 *>(s_zero)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the async proxy
    __syncthreads();
    if (threadIdx.x != 0) return;
    const size_t n_chunks = (bytes + ZERO_CHUNK - 1) / ZERO_CHUNK;
    const unsigned src = (unsigned)__cvta_generic_to_shared(s_zero);
    unsigned long long pol = 0;
    if (evict_first) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    for (size_t ch = cta; ch < n_chunks; ch += n_cta) {
        const size_t off = ch * ZERO_CHUNK;
        const unsigned sz = (unsigned)min((size_t)ZERO_CHUNK, bytes - off);
        if (evict_first)
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                         :: "l"(reinterpret_cast<char *>(dst) + off), "r"(src), "r"(sz), "l"(pol) : "memory");
        else
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         :: "l"(reinterpret_cast<char *>(dst) + off), "r"(src), "r"(sz) : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    tl_stamp(0, true);
}

// Fused prologue of a step: ONE launch whose CTAs take one of three independent roles --
//   [0, n_zero)                    zero-fill of the BEV tensor (models.py:240), the only bandwidth-bound piece of the path
//   [n_zero, n_zero + n_lift)      lift operands: depth softmax + pixel-major context (models.py:49-61)
//   [.., + n_index)                run index: geometry + voxel rows + list pushes (models.py:170-190, :212-221)
// (measured at cfg 2: lift before index 61.0 us per step, index before lift 63.1 -- the grid is a little more than one wave)
// The index and lift roles are chains of memory latencies that would otherwise queue behind the zero-fill on other streams
// (measured: on separate graph branches each of them ends when the zero-fill ends, plus ~10 us of fork / join overhead);
// inside one grid they share the SMs with the zero role's single issuing thread and finish within its shadow.
struct PrologueArgs {
    float *bev; size_t bev_bytes; int n_zero, evict_first;            // zero role (n_zero = 0: off)
    int n_index; RunDims rd; CalibPtrs c;                             // index role (n_index = 0: off)
    int32_t *prow, *cnt, *head, *counters; int2 *sub;
    int n_lift; const float *dn; float *prob, *ctx_t, *prob_col;      // lift role (n_lift = 0: off)
};

template <bool RAW>
__global__ void __launch_bounds__(RP_THREADS)
k_prologue(Dims d, PrologueArgs a) {
    extern __shared__ __align__(128) float s_pro[];
    lss_pdl_trigger();                                    // k_run_classify may be scheduled while this grid drains
    int cta = (int)blockIdx.x;
    if (cta < a.n_zero) { zero_role(a.bev, a.bev_bytes, cta, a.n_zero, s_pro, a.evict_first); return; }
    cta -= a.n_zero;
    if (cta < a.n_lift) { lift_prepare_cta<float>(d, a.dn, a.prob, a.ctx_t, a.prob_col, cta, s_pro); return; }
    cta -= a.n_lift;
    run_index_cta<RAW>(d, a.rd, a.c, a.prow, a.cnt, a.head, a.sub, a.counters, cta);
}

// ------------------------------------------------------------------------------------------------
// forward gather
// ------------------------------------------------------------------------------------------------

template <int CPL>
__global__ void __launch_bounds__(GCL_THREADS, 1024 / GCL_THREADS)
k_fwd_gather_cl(Dims d, int n_keys, unsigned long long mfH, const int32_t *__restrict__ prow, const uint32_t *__restrict__ emask,
                const int32_t *__restrict__ counters, const int4 *__restrict__ mixed_recs, long long n_mixed_cap,
                uint32_t *__restrict__ pool, const int2 *__restrict__ sub, int fWD, const float *__restrict__ prob_col,
                const float *__restrict__ ctx_t, float *__restrict__ bev) {
    extern __shared__ __align__(16) float s_dyn[];
    constexpr int C = 8 * CPL;
    constexpr int LF = CPL <= 8 ? 4 : 2;                  // context rows in flight per group (shared voxels)
    lss_pdl_wait();
    tl_stamp(3, false);
    const int n_queue = (int)gridDim.x - n_keys;          // the FIRST CTAs drain the queue of shared voxels (they form no tail)
    const bool column = (int)blockIdx.x >= n_queue;
    const int lane = threadIdx.x & 31, gl = lane & 7, g = threadIdx.x >> 3;
    const int per = d.D * d.fH;
    if (column) {
        const int key = (int)blockIdx.x - n_queue;        // (bn, w)
        const int bn = key / d.fW, w0 = key - bn * d.fW;
        float *s_ctx = s_dyn;                             // [fH][C]
        float *s_prob = s_dyn + d.fH * C;                 // [D][fH]
        int *s_row = reinterpret_cast<int *>(s_prob + per);
        unsigned *s_em = reinterpret_cast<unsigned *>(s_row + per);
        unsigned short *s_list = reinterpret_cast<unsigned short *>(s_em + per);
        __shared__ int s_n;
        if (threadIdx.x == 0) s_n = 0;
        {
            const float4 *src = reinterpret_cast<const float4 *>(ctx_t + ((size_t)bn * d.HW + w0) * C);
            constexpr int c4 = C >> 2;
            for (int i = threadIdx.x; i < d.fH * c4; i += GCL_THREADS) {
                const int h = i / c4, q = i - h * c4;
                reinterpret_cast<float4 *>(s_ctx)[i] = __ldg(src + (size_t)h * d.fW * c4 + q);
            }
        }
        __syncthreads();
        const size_t base = (size_t)key * per;
        for (int i0 = 0; i0 < per; i0 += GCL_THREADS) {   // stage the column's plan + weights, compact its leaders
            const int i = i0 + threadIdx.x;
            unsigned em = 0u;
            if (i < per) {
                em = __ldg(emask + base + i);
                s_em[i] = em;
                s_row[i] = __ldg(prow + base + i);
                s_prob[i] = __ldg(prob_col + base + i);
            }
            const unsigned hb = __ballot_sync(LSS_FULL_MASK, em != 0u);
            int wbase = 0;
            if (lane == 0 && hb) wbase = atomicAdd(&s_n, __popc(hb));
            wbase = __shfl_sync(LSS_FULL_MASK, wbase, 0);
            LSS_DASSERT(em == 0u || wbase + __popc(hb) <= per);
            if (em != 0u) s_list[wbase + __popc(hb & ((1u << lane) - 1u))] = (unsigned short)i;
        }
        __syncthreads();
        const int n_list = s_n;
        const float *my_ctx = s_ctx + gl * 4;
        for (int i = g; __any_sync(LSS_FULL_MASK, i < n_list); i += GCL_NG) {
            const bool live = i < n_list;
            const int s = live ? (int)s_list[i] : 0;
            const unsigned m = live ? s_em[s] : 0u;
            const int h0 = s - (int)lss_div20((unsigned)s, mfH) * d.fH;
            float acc[CPL];
#pragma unroll
            for (int a = 0; a < CPL; ++a) acc[a] = 0.f;
            const float *wp = s_prob + s;
            const float *rowp0 = my_ctx + h0 * C;
            for (int j = 0; j < d.fH; ++j) {
                if ((m >> j) & 1u) {                      // bit j set => image row h0 + j < fH of the same run
                    const float wj = wp[j];
                    const float4 *rowp = reinterpret_cast<const float4 *>(rowp0 + j * C);
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
            if (live) {
                LSS_DASSERT(s_row[s] >= 0 && s_row[s] < d.B * d.nx * d.ny * d.nz && (m >> (d.fH - h0)) == 0u);
                float4 *dst = reinterpret_cast<float4 *>(bev + (size_t)s_row[s] * C) + gl;
#pragma unroll
                for (int q = 0; q < CPL / 4; ++q) dst[8 * q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
            }
        }
        tl_stamp(3, true);
        return;
    }

    // ---- queue CTAs: shared voxels.  Shared-memory layout: [NG][64] unsorted, [NG][64] sorted, [SORT_CAP] keys, [NG][C] products
    uint32_t *s_un = reinterpret_cast<uint32_t *>(s_dyn) + g * GCL_SHORT_CAP;
    uint32_t *s_so = reinterpret_cast<uint32_t *>(s_dyn) + (GCL_NG + g) * GCL_SHORT_CAP;
    uint32_t *s_keys = reinterpret_cast<uint32_t *>(s_dyn) + 2 * GCL_NG * GCL_SHORT_CAP;
    float *s_prod = s_dyn + 2 * GCL_NG * GCL_SHORT_CAP + GCL_SORT_CAP;
    const int qid = (int)blockIdx.x;
    const int HWC = d.HW * C;
    auto decode = [&](unsigned pidx, int b, int &ro, size_t &wi) {   // context row offset (floats) and prob_col index of a point
        const unsigned cam = lss_div20(pidx, d.mDHW);
        const unsigned rr = pidx - cam * d.DHW;
        const unsigned dd = lss_div20(rr, d.mHW);
        const unsigned hw = rr - dd * d.HW;
        const unsigned h = lss_div20(hw, d.mfW), ww = hw - h * d.fW;
        const unsigned bnn = (unsigned)b * d.N + cam;
        ro = (int)(cam * HWC + hw * C);
        wi = ((size_t)(bnn * d.fW + ww) * d.D + dd) * d.fH + h;
    };
    {   // long voxels (>= 64 points): one at a time by the whole CTA.  Keys are sorted by the CTA (shared memory up to SORT_CAP
        // points, else in place in the pool), then NG points per pass: every group fetches one point's context row and writes
        // float32(prob*ctx) to shared memory (all loads in flight together); thread c adds the NG products of channel c in
        // ascending point order -- the same sequence of float32 additions as everywhere else, without the serial load chain.
        const int n_long = __ldg(counters + 2);
        for (int rl = qid; rl < n_long; rl += n_queue) {
            const int4 rec = __ldg(mixed_recs + (n_mixed_cap - 1 - rl));
            uint32_t *gk = pool + rec.x;
            const bool in_smem = rec.y <= GCL_SORT_CAP;
            if (in_smem) {
                for (int i = threadIdx.x; i < rec.y; i += GCL_THREADS) s_keys[i] = __ldcg(gk + i);
                __syncthreads();
                bitonic_sort_block(s_keys, rec.y);
            } else {
                bitonic_sort_block((volatile uint32_t *)gk, rec.y);
            }
            __syncthreads();
            const float *ctx_b = ctx_t + (size_t)rec.w * d.N * HWC + gl * 4;
            float accc = 0.f;
            for (int base = 0; base < rec.y; base += GCL_NG) {
                const int cntp = min(GCL_NG, rec.y - base);
                if (g < cntp) {
                    const unsigned pidx = in_smem ? s_keys[base + g] : ((volatile uint32_t *)gk)[base + g];
                    int ro; size_t wi;
                    decode(pidx, rec.w, ro, wi);
                    const float w = __ldg(prob_col + wi);
                    const float4 *rowp = reinterpret_cast<const float4 *>(ctx_b + ro);
                    float4 *dst = reinterpret_cast<float4 *>(s_prod + g * C) + gl;
#pragma unroll
                    for (int q = 0; q < CPL / 4; ++q) {
                        const float4 v = __ldg(rowp + 8 * q);
                        dst[8 * q] = make_float4(__fmul_rn(w, v.x), __fmul_rn(w, v.y), __fmul_rn(w, v.z), __fmul_rn(w, v.w));
                    }
                }
                __syncthreads();
                if ((int)threadIdx.x < C)
                    for (int jj = 0; jj < cntp; ++jj) accc = __fadd_rn(accc, s_prod[jj * C + threadIdx.x]);
                __syncthreads();
            }
            if ((int)threadIdx.x < C) bev[(size_t)rec.z * C + threadIdx.x] = accc;
        }
    }
    const int n_rec = __ldg(counters);
    for (int r = GCL_NG * qid + g; __any_sync(LSS_FULL_MASK, r < n_rec); r += GCL_NG * n_queue) {
        const bool live = r < n_rec;
        int4 rec = make_int4(0, 0, 0, 0);                 // {head of the voxel's list, points, voxel row, batch}
        if (live) rec = __ldg(mixed_recs + r);
        const int len = rec.y;
        const int maxlen = __reduce_max_sync(LSS_FULL_MASK, len);
        {   // walk the voxel's list (a few sub-runs), the groups of the warp in lockstep; the 8 lanes expand each node together
            int cur = rec.x, filled = 0;
            while (__any_sync(LSS_FULL_MASK, cur != 0)) {
                int2 nd = make_int2(0, 0);
                if (cur != 0) {
                    LSS_DASSERT(cur >= 1 && cur <= d.n_points);
                    nd = __ldcg(sub + (cur - 1));
                    LSS_DASSERT(filled + __popc((unsigned)nd.y) <= len && len < GCL_SHORT_CAP);
                    filled += expand_subrun(d, fWD, cur - 1, (unsigned)nd.y, s_un + filled, gl, 8);
                }
                cur = nd.x;
            }
        }
        __syncwarp();
        for (int i = gl; i < maxlen; i += 8) {            // rank sort by flat point index (keys unique, len < 64)
            const uint32_t e = i < len ? s_un[i] : 0u;
            int rank = 0;
            for (int j = 0; j < maxlen; ++j) rank += (j < len && s_un[j] < e) ? 1 : 0;
            if (i < len) s_so[rank] = e;
        }
        __syncwarp();
        const float *ctx_b = ctx_t + (size_t)rec.w * d.N * HWC + gl * 4;
        float acc[CPL];
#pragma unroll
        for (int a = 0; a < CPL; ++a) acc[a] = 0.f;
        for (int base = 0; base < maxlen; base += 8) {    // the groups of the warp in lockstep, 8 points per step
            const int cntp = min(8, len - base);          // <= 0 for a group that has nothing (more) to do
            const int maxcnt = min(8, maxlen - base);
            float w = 0.f;
            int ro = 0;
            if (gl < cntp) {
                size_t wi;
                decode(s_so[base + gl], rec.w, ro, wi);
                w = __ldg(prob_col + wi);
            }
#pragma unroll
            for (int j0 = 0; j0 < 8; j0 += LF) {
                if (j0 >= maxcnt) break;
                float x[LF][CPL];
#pragma unroll
                for (int u = 0; u < LF; ++u) {            // LF rows in flight (idle slots re-read row 0 of the sample: finite, unused)
                    const int oj = __shfl_sync(LSS_FULL_MASK, ro, j0 + u, 8);
                    const float4 *rowp = reinterpret_cast<const float4 *>(ctx_b + oj);
#pragma unroll
                    for (int q = 0; q < CPL / 4; ++q) {
                        const float4 v = __ldg(rowp + 8 * q);
                        x[u][4 * q] = v.x; x[u][4 * q + 1] = v.y; x[u][4 * q + 2] = v.z; x[u][4 * q + 3] = v.w;
                    }
                }
#pragma unroll
                for (int u = 0; u < LF; ++u) {
                    const float wj = __shfl_sync(LSS_FULL_MASK, w, j0 + u, 8);
                    if (j0 + u < cntp) {
#pragma unroll
                        for (int a = 0; a < CPL; ++a) acc[a] = __fadd_rn(acc[a], __fmul_rn(wj, x[u][a]));
                    }
                }
            }
        }
        if (live) {
            LSS_DASSERT(rec.z >= 0 && rec.z < d.B * d.nx * d.ny * d.nz && rec.w == rec.z / (d.nx * d.ny * d.nz));
            float4 *dst = reinterpret_cast<float4 *>(bev + (size_t)rec.z * C) + gl;
#pragma unroll
            for (int q = 0; q < CPL / 4; ++q) dst[8 * q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------

static int rp_num_sms() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    return n;
}

static inline RunDims make_run_dims(const lss_problem *p) {
    RunDims rd;
    rd.fH = p->fH; rd.RPW = 32 / p->fH;
    rd.R = p->B * p->N * p->fW * p->D;
    rd.fWD = p->fW * p->D;
    rd.fmask = p->fH == 32 ? 0xFFFFFFFFu : ((1u << p->fH) - 1u);
    return rd;
}

static int runplan_supported(const lss_problem *p) {
    int st = lss_check_problem(p);
    if (st != LSS_OK) return st;
    if (p->fH > 32) return LSS_ERR_UNSUPPORTED;                                   // a run must fit a warp
    if (!(p->C == 32 || p->C == 64 || p->C == 128)) return LSS_ERR_UNSUPPORTED;    // 8 lanes x C/8 channels
    if ((long long)p->D * p->fH > 65535) return LSS_ERR_UNSUPPORTED;               // 16-bit slot numbers inside a column
    if ((long long)p->N * p->fH * p->fW * p->C >= (1ll << 31)) return LSS_ERR_UNSUPPORTED;
    return LSS_OK;
}

extern "C" int lss_runplan_layout_init(const lss_problem *p, lss_runplan_layout *out) {
    int st = runplan_supported(p);
    if (st != LSS_OK) return st;
    LSS_REQUIRE(out != nullptr, LSS_ERR_BAD_ARG);
    const Dims d = make_dims(p);
    out->n_points = d.n_points;
    out->n_runs = (int64_t)p->B * p->N * p->fW * p->D;
    out->n_voxels = (int64_t)p->B * p->nx * p->ny * p->nz;
    out->n_mixed_cap = (int64_t)d.n_points / 2 + 2;       // a shared voxel holds at least two points
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    size_t off = 0;
    out->off_prow = off;       off += up((size_t)d.n_points * 4);
    out->off_emask = off;      off += up((size_t)d.n_points * 4);
    out->off_sub = off;        off += up((size_t)d.n_points * 8);
    out->off_pool = off;       off += up((size_t)d.n_points * 4);
    out->off_mixed_recs = off; off += up((size_t)out->n_mixed_cap * 16);
    out->off_counters = off;   off += up(64 * 4);
    out->off_cnt = off;        off += up((size_t)out->n_voxels * 4);
    out->off_head = off;       off += up((size_t)out->n_voxels * 4);
    out->bytes = off;
    return LSS_OK;
}

extern "C" int lss_runplan_reset(const lss_runplan_layout *L, void *ws, void *stream) {
    LSS_REQUIRE(L && ws, LSS_ERR_WORKSPACE);
    char *w = (char *)ws;    // counters, cnt and head are contiguous
    if (cudaMemsetAsync(w + L->off_counters, 0, L->bytes - L->off_counters, (cudaStream_t)stream) != cudaSuccess) return LSS_ERR_CUDA;
    return LSS_OK;
}

static int launch_prologue(const lss_problem *p, const lss_runplan_layout *L, void *workspace, bool index, const CalibPtrs &c, bool raw,
                           const float *dn, float *prob, float *ctx_t, float *prob_col, float *bev, size_t bev_bytes,
                           cudaStream_t s) {
    const Dims d = make_dims(p);
    PrologueArgs a = {};
    a.rd = make_run_dims(p);
    char *w = (char *)workspace;
    if (bev != nullptr && bev_bytes > 0) {
        // two issuing CTAs per SM (one: 67.5 us per step at cfg 2 against 63); evict-first lines: the zeros should not push
        // the plan and the lift operands out of L2 (measured neutral to slightly better)
        const size_t n_chunks = (bev_bytes + ZERO_CHUNK - 1) / ZERO_CHUNK;
        a.bev = bev; a.bev_bytes = bev_bytes; a.evict_first = 1;
        a.n_zero = (int)min((size_t)2 * rp_num_sms(), n_chunks);
    }
    if (index) {
        const int runs_per_cta = RP_WARPS * RP_U * a.rd.RPW;
        if (raw) LSS_REQUIRE(runs_per_cta / a.rd.fWD + 2 <= LSS_RAW_CAMS, LSS_ERR_UNSUPPORTED);   // cameras one CTA may span
        a.n_index = (a.rd.R + runs_per_cta - 1) / runs_per_cta;
        a.c = c;
        a.prow = (int32_t *)(w + L->off_prow); a.cnt = (int32_t *)(w + L->off_cnt); a.head = (int32_t *)(w + L->off_head);
        a.counters = (int32_t *)(w + L->off_counters); a.sub = (int2 *)(w + L->off_sub);
    }
    size_t smem = a.n_zero ? ZERO_CHUNK : 0;
    if (dn != nullptr) {
        a.n_lift = p->B * p->N * ((d.HW + LIFT_PX - 1) / LIFT_PX);
        a.dn = dn; a.prob = prob; a.ctx_t = ctx_t; a.prob_col = prob_col;
        smem = max(smem, (size_t)(d.D + d.C) * (LIFT_PX + 1) * sizeof(float));
    }
    const int grid = a.n_zero + a.n_index + a.n_lift;
    if (grid == 0) return LSS_OK;
    if (smem > 227 * 1024) return LSS_ERR_UNSUPPORTED;
    auto kern = raw ? k_prologue<true> : k_prologue<false>;
    if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return LSS_ERR_CUDA;
    kern<<<grid, RP_THREADS, smem, s>>>(d, a);
    LSS_CHECK_LAUNCH();
    if (index) {
        if (lss_launch(k_run_classify, dim3(a.n_index), dim3(RP_THREADS), 0, s, true, d, a.rd, (const int32_t *)a.prow,
                       (uint32_t *)(w + L->off_emask), a.cnt, a.head, (const int2 *)a.sub, (uint32_t *)(w + L->off_pool),
                       (int4 *)(w + L->off_mixed_recs), a.counters, (long long)L->n_mixed_cap) != cudaSuccess) return LSS_ERR_CUDA;
        LSS_CHECK_LAUNCH();
    }
    return LSS_OK;
}

extern "C" int lss_runplan_build(const lss_problem *p, const lss_runplan_layout *L, void *workspace, const float *frustum,
                                 const float *post_trans, const float *M1, const float *M2, const float *trans,
                                 const float *rots, const float *intrins, const float *post_rots, void *stream) {
    return lss_liftsplat_prologue(p, L, workspace, frustum, post_trans, M1, M2, trans, rots, intrins, post_rots, nullptr, nullptr,
                                  nullptr, nullptr, nullptr, stream);
}

extern "C" int lss_liftsplat_prologue(const lss_problem *p, const lss_runplan_layout *L, void *workspace, const float *frustum,
                                      const float *post_trans, const float *M1, const float *M2, const float *trans,
                                      const float *rots, const float *intrins, const float *post_rots, const float *depthnet_out,
                                      float *prob, float *ctx_t, float *prob_col, float *bev, void *stream) {
    int st = runplan_supported(p);
    if (st != LSS_OK) return st;
    const bool index = frustum != nullptr;
    const bool raw = M1 == nullptr || M2 == nullptr;
    if (index) {
        LSS_REQUIRE(L && workspace, LSS_ERR_WORKSPACE);
        LSS_REQUIRE(post_trans && trans, LSS_ERR_BAD_ARG);
        LSS_REQUIRE(!raw || (rots && intrins && post_rots), LSS_ERR_BAD_ARG);
        LSS_REQUIRE(L->n_points == (int64_t)p->B * p->N * p->D * p->fH * p->fW, LSS_ERR_WORKSPACE);
    }
    if (depthnet_out != nullptr) LSS_REQUIRE(prob && ctx_t && prob_col, LSS_ERR_BAD_ARG);
    const size_t bytes = (size_t)p->B * p->nz * p->C * p->nx * p->ny * 4;
    if (bev != nullptr) LSS_REQUIRE(lss_aligned(bev, 16) && bytes % 16 == 0, LSS_ERR_ALIGN);
    const CalibPtrs c{frustum, post_trans, M1, M2, trans, rots, intrins, post_rots};
    return launch_prologue(p, L, workspace, index, c, index && raw, depthnet_out, prob, ctx_t, prob_col, bev, bytes, (cudaStream_t)stream);
}

static int launch_bev_zero(float *bev, size_t bytes, cudaStream_t s) {
    if (bytes % 16 != 0 || !lss_aligned(bev, 16))
        return cudaMemsetAsync(bev, 0, bytes, s) == cudaSuccess ? LSS_OK : LSS_ERR_CUDA;
    lss_problem dummy = {};
    dummy.B = dummy.N = dummy.D = dummy.fH = dummy.fW = 1; dummy.C = 32; dummy.nx = dummy.ny = dummy.nz = 1;
    dummy.dx[0] = dummy.dx[1] = dummy.dx[2] = 1.f;
    return launch_prologue(&dummy, nullptr, nullptr, false, CalibPtrs{}, false, nullptr, nullptr, nullptr, nullptr, bev, bytes, s);
}

extern "C" int lss_bev_zero(const lss_problem *p, float *bev, int part, int n_parts, void *stream) {
    int st = lss_check_problem(p);
    if (st != LSS_OK) return st;
    LSS_REQUIRE(bev != nullptr && n_parts >= 1 && part >= 0 && part < n_parts, LSS_ERR_BAD_ARG);
    const size_t rows = (size_t)p->B * p->nz * p->nx * p->ny, row_bytes = (size_t)p->C * 4;   // split on voxel-row boundaries
    const size_t r0 = rows * part / n_parts, r1 = rows * (part + 1) / n_parts;
    if (r1 == r0) return LSS_OK;
    return launch_bev_zero(reinterpret_cast<float *>(reinterpret_cast<char *>(bev) + r0 * row_bytes), (r1 - r0) * row_bytes, (cudaStream_t)stream);
}

static size_t gather_cl_smem(const lss_problem *p) {
    const size_t per = (size_t)p->D * p->fH;
    const size_t col = ((size_t)p->fH * p->C + 3 * per) * 4 + per * 2 + 16;
    const size_t que = ((size_t)2 * GCL_NG * GCL_SHORT_CAP + GCL_SORT_CAP + (size_t)GCL_NG * p->C) * 4;
    return col > que ? col : que;
}

extern "C" int lss_liftsplat_fwd_cl(const lss_problem *p, const lss_runplan_layout *L, const void *workspace,
                                    const float *prob_col, const float *ctx_t, float *bev, int precleared, void *stream) {
    int st = runplan_supported(p);
    if (st != LSS_OK) return st;
    LSS_REQUIRE(L && workspace, LSS_ERR_WORKSPACE);
    LSS_REQUIRE(prob_col && ctx_t && bev, LSS_ERR_BAD_ARG);
    LSS_REQUIRE(lss_aligned(ctx_t, 16) && lss_aligned(bev, 16), LSS_ERR_ALIGN);
    const Dims d = make_dims(p);
    LSS_REQUIRE(L->n_points == d.n_points, LSS_ERR_WORKSPACE);
    cudaStream_t s = (cudaStream_t)stream;
    if (!precleared) {
        st = launch_bev_zero(bev, (size_t)L->n_voxels * p->C * 4, s);
        if (st != LSS_OK) return st;
    }
    const size_t smem = gather_cl_smem(p);
    if (smem > 48 * 1024) return LSS_ERR_UNSUPPORTED;
    const char *w = (const char *)workspace;
    const int n_keys = p->B * p->N * p->fW;
    const int grid = n_keys + 2 * rp_num_sms();
    const unsigned long long mfH = ((1ull << 40) + (unsigned)p->fH - 1) / (unsigned)p->fH;
#define GCL_ARGS d, n_keys, mfH, (const int32_t *)(w + L->off_prow), (const uint32_t *)(w + L->off_emask),                \
                 (const int32_t *)(w + L->off_counters), (const int4 *)(w + L->off_mixed_recs), (long long)L->n_mixed_cap, \
                 (uint32_t *)(const_cast<char *>(w) + L->off_pool), (const int2 *)(w + L->off_sub), p->fW * p->D, prob_col, ctx_t, bev
    cudaError_t e;
    if (p->C == 32) e = lss_launch(k_fwd_gather_cl<4>, dim3(grid), dim3(GCL_THREADS), smem, s, false, GCL_ARGS);
    else if (p->C == 64) e = lss_launch(k_fwd_gather_cl<8>, dim3(grid), dim3(GCL_THREADS), smem, s, false, GCL_ARGS);
    else e = lss_launch(k_fwd_gather_cl<16>, dim3(grid), dim3(GCL_THREADS), smem, s, false, GCL_ARGS);
#undef GCL_ARGS
    if (e != cudaSuccess) return LSS_ERR_CUDA;
    LSS_CHECK_LAUNCH();
    return LSS_OK;
}

extern "C" int lss_liftsplat_bwd_cl(const lss_problem *p, const lss_runplan_layout *L, const void *workspace,
                                    const float *grad_bev, const float *prob_col, const float *ctx_t, float *grad_depthnet,
                                    void *stream) {
    int st = runplan_supported(p);
    if (st != LSS_OK) return st;
    LSS_REQUIRE(L && workspace, LSS_ERR_WORKSPACE);
    LSS_REQUIRE(grad_bev && prob_col && ctx_t && grad_depthnet, LSS_ERR_BAD_ARG);
    LSS_REQUIRE(lss_aligned(ctx_t, 16) && lss_aligned(grad_bev, 16), LSS_ERR_ALIGN);
    LSS_REQUIRE(p->D <= LSS_MAX_DEPTH, LSS_ERR_UNSUPPORTED);
    const Dims d = make_dims(p);
    LSS_REQUIRE(L->n_points == d.n_points, LSS_ERR_WORKSPACE);
    return lss_bwd_gather_rows(d, (const int32_t *)((const char *)workspace + L->off_prow), prob_col, ctx_t, grad_bev,
                               grad_depthnet, 0, p->B, false, (cudaStream_t)stream);
}
