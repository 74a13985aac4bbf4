// Run plan + channels_last forward: the fast path of the fused lift-splat (default of bench.py and of the API).
//
// Replaces, for shdragron/LSS-Carla (paths under the reference root):
//   LiftSplatShoot.get_geometry                      src/models.py:170-190   (geometry in registers)
//   voxel_pooling: quantise, mask, rank, argsort     src/models.py:212-231   (no sort: runs + per-voxel lists)
//   the lift outer product + QuickCumsum + griddify  src/models.py:59, :234-244, src/tools.py:193-209
//
// A RUN is the fH image rows of one (camera, feature column, depth bin): consecutive points of the camera-column-major
// order cm = ((bn*fW + w)*D + d)*fH + h.  A warp owns 32/fH whole runs, lane = image row.  The points of a run that share
// a voxel form a SUB-RUN; its first lane is the leader.  With the BEV in channels_last a voxel is one contiguous C-float
// row, so whoever owns a voxel writes it directly: no tile-owner store pass, no compact rows, and nothing to sort --
//   index role (k_prologue)   voxel row per point (bit-exact arithmetic of geom.cuh) -> prow; every leader pushes its sub-run
//                   on the voxel's list with ONE 64-bit atomicExch on head[voxel] = {build epoch, leader + 1} and records
//                   sub[leader] = {previous head of this build (0: it pushed first), row mask}.  Heads of older builds carry
//                   an older epoch and read as empty: nothing is cleared between builds and the forward only READS the plan.
//                   A pusher that finds a sub-run of this build on the list also queues the voxel {row, previous, itself, mask}.
//   k_fwd_columns   one CTA per camera column stages its operands, finds its first pushers and looks at their voxels' heads:
//                   still the head = the sub-run is alone = EXCLUSIVE, summed from the staged operands by an 8-lane group
//                   (lane = C/8 channels) and written as one voxel row.  Voxels shared by several sub-runs (about 5 %) come
//                   from the queue; the CTAs at the END of the same grid take one record per warp: the two sub-runs' nodes and
//                   the head in one go, the few points sorted by flat index, weights and context rows in one go, added in key
//                   order; >= 64 points: a whole CTA.
//   zero-fill       cp.async.bulk from a zeroed shared-memory chunk, sample by sample, with a progress counter per sample that
//                   the writers of voxel rows wait for: as the first CTAs of k_fwd_columns (lss_liftsplat_fwd_cl) or as its own
//                   small grid in front of prologue and columns (lss_liftsplat_forward: three launches running side by side,
//                   the columns polling a READY flag instead of waiting for the grids before them to complete).
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
// (globaltimer ns) of the zero CTAs (0), the index role (1), the lift role (2), the column CTAs (3) and k_fwd_shared (4).
#ifdef LSS_RP_TIMELINE
__device__ unsigned long long g_rp_tl[10] = {~0ull, 0, ~0ull, 0, ~0ull, 0, ~0ull, 0, ~0ull, 0};
__device__ int g_rp_tl_on = 0;
__device__ __forceinline__ void tl_stamp(int k, bool end) {
    if (!g_rp_tl_on || threadIdx.x != 0) return;
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (end) atomicMax(g_rp_tl + 2 * k + 1, t); else atomicMin(g_rp_tl + 2 * k, t);
}
#define TL_MARKS 8
__device__ unsigned long long g_rp_marks[4096 * TL_MARKS];
__device__ __forceinline__ void tl_mark(int cta, int phase) {      // per-CTA phase stamps of the forward's column CTAs
    if (!g_rp_tl_on || threadIdx.x != 0 || cta >= 4096) return;
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_rp_marks[cta * TL_MARKS + phase] = t;
}
extern "C" int lss_debug_runplan_marks(unsigned long long *out_host, int n_cta) {
    return cudaMemcpyFromSymbol(out_host, g_rp_marks, sizeof(unsigned long long) * TL_MARKS * n_cta) == cudaSuccess ? 0 : -4;
}
extern "C" int lss_debug_runplan_timeline(int on, unsigned long long *out_host) {
    unsigned long long h[10];
    if (out_host) { if (cudaMemcpyFromSymbol(h, g_rp_tl, sizeof(h)) != cudaSuccess) return -4; for (int i = 0; i < 10; ++i) out_host[i] = h[i]; }
    const unsigned long long init[10] = {~0ull, 0, ~0ull, 0, ~0ull, 0, ~0ull, 0, ~0ull, 0};
    if (cudaMemcpyToSymbol(g_rp_tl, init, sizeof(init)) != cudaSuccess) return -4;
    return cudaMemcpyToSymbol(g_rp_tl_on, &on, sizeof(on)) == cudaSuccess ? 0 : -4;
}
#else
#define tl_stamp(k, end) ((void)0)
#define tl_mark(cta, phase) ((void)0)
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

// counters[] of the workspace (int32[64])
#define RPC_EPOCH 0        // epoch of the last completed build (the plan the forward reads)
#define RPC_PRO_DONE 1     // lift / index CTAs of the running prologue that have finished (the last one publishes epoch + READY)
#define RPC_FWD_DONE 2     // shared-voxel CTAs of the running forward that have finished (the last one sums the long voxels and resets the scratch)
#define RPC_POOL 3         // pool cursor of the running forward
#define RP_QSHARDS 32      // the queue of shared voxels is 32 interleaved sub-queues (record k of shard s sits in slot k*32 + s) with one
                           // counter each, (build epoch << 32) | count, on its own 128-byte line: index CTA i queues on shard i % 32,
                           // so that the ~2 000 reservations of a build do not serialise on one L2 address
#define RPC_NLONG 12       // long voxels (>= 64 points) that did not fit the list of the CTA that met them (running forward)
#define RPC_STAT 6         // [6], [7]: shared / long voxels summed by the last completed forward
#define RPC_NSHARED 9      // shared voxels summed by the running forward
#define RPC_NLONG_CTA 10   // long voxels summed by the CTA that met them (running forward)
#define RP_FLAG_STRIDE 32  // ints between two polled flags: every flag has its own 128-byte line (and L2 slice)
#define RP_READY_LINES 32  // copies of READY (lss_liftsplat_forward: 1 once the plan and the lift operands of the step are
                           // complete); column CTA i polls copy i % 32, so that a thousand pollers do not queue on one L2 line

// floor(x / divisor) for any 32-bit x, with m = ceil(2^64 / divisor), divisor >= 2: x*e < 2^64 for e = m*divisor - 2^64 < divisor.
// (The index pass makes four runtime divisions per point; as library calls they were a quarter of its instructions.)
__device__ __forceinline__ unsigned div_magic(unsigned x, unsigned long long m) { return m ? (unsigned)__umul64hi((unsigned long long)x, m) : x; }
static inline unsigned long long magic_of(unsigned divisor) {      // 0 stands for divisor 1
    return divisor < 2 ? 0ull : (unsigned long long)(((unsigned __int128)1 << 64) / divisor) + 1ull;
}

// thread -> point mapping of the index pass
struct RunDims {
    int fH, RPW;        // runs per warp = 32 / fH
    int R;              // runs = B*N*fW*D
    int fWD;            // fW*D: runs per camera
    unsigned fmask;     // fH low bits
    unsigned long long m_fH, m_fWD, m_D, m_N;      // magic_of(fH), (fW*D), (D), (N)
};

struct RunLane { int r, h, rw; bool valid; unsigned run_mask; };

__device__ __forceinline__ RunLane run_lane(const RunDims &rd, int cta, int u) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    RunLane q;
    q.rw = (int)div_magic((unsigned)lane, rd.m_fH);
    q.h = lane - q.rw * rd.fH;
    const int wg = (cta * RP_WARPS + warp) * RP_U + u;
    q.r = wg * rd.RPW + q.rw;
    q.valid = q.rw < rd.RPW && q.r < rd.R;
    q.run_mask = q.rw < rd.RPW ? rd.fmask << (q.rw * rd.fH) : 0u;
    return q;
}

// ------------------------------------------------------------------------------------------------
// plan kernel
// ------------------------------------------------------------------------------------------------

// The work of one CTA of the index pass (RP_THREADS threads); `cta` in [0, n_index).
template <bool RAW>
__device__ __forceinline__ void run_index_cta(const Dims &d, const RunDims &rd, const CalibPtrs &c, int32_t *__restrict__ prow,
                                              unsigned long long *__restrict__ head, int2 *__restrict__ sub, int2 *__restrict__ sub2,
                                              int4 *__restrict__ recs, unsigned long long *__restrict__ qcount,
                                              const int32_t *__restrict__ counters, int cta, float *__restrict__ clear_bev) {
    tl_stamp(1, false);
    __shared__ float s_m[RAW ? LSS_RAW_CAMS : 1][18];
    __shared__ unsigned s_epoch;
    const int cam0 = (int)div_magic((unsigned)(cta * RP_WARPS * RP_U * rd.RPW), rd.m_fWD);       // first run of the CTA < R < 2^31
    if (threadIdx.x == 32) s_epoch = (unsigned)__ldcg(counters + RPC_EPOCH) + 1u;     // this build's epoch (published by prologue_cta_done)
    if (RAW) {      // the calibration matrices of the few cameras this CTA touches, made on the fly (no extra launch)
        const int cam = cam0 + (int)threadIdx.x;
        if (threadIdx.x < LSS_RAW_CAMS && cam < d.B * d.N) calib_matrices_of(c.rots, c.intrins, c.post_rots, cam, s_m[threadIdx.x], s_m[threadIdx.x] + 9);
    }
    __syncthreads();
    const unsigned long long tag = (unsigned long long)s_epoch << 32;
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int u = 0; u < RP_U; ++u) {
        const RunLane q = run_lane(rd, cta, u);
        int row = -1;
        const size_t cm = (size_t)q.r * d.fH + q.h;
        if (clear_bev != nullptr && q.valid) {            // persistent output: the previous build's sub-run leaders name the rows its
            const int old_row = __ldcg(prow + cm);        // forward wrote -- zero them (the rest of the tensor is still zero).
            const int old_mask = __ldcg(&sub[cm].y);      // (Measured: the warp clearing its leaders' rows together, 128 / C rows per
            if (old_mask != 0 && old_row >= 0) {          // store instruction, is slower -- and slows the regular build by 0.5 us.)
                float4 *dst = reinterpret_cast<float4 *>(clear_bev + (size_t)old_row * d.C);
                for (int k = 0; k < (d.C >> 2); ++k) dst[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        if (q.valid) {
            const int bn = (int)div_magic((unsigned)q.r, rd.m_fWD), rem = q.r - bn * rd.fWD;
            const int w = (int)div_magic((unsigned)rem, rd.m_D), dd = rem - w * d.D;
            const int in_cam = (dd * d.fH + q.h) * d.fW + w;
            float g[3];
            if (RAW) ego_point(c, bn, in_cam, g, s_m[bn - cam0], s_m[bn - cam0] + 9);
            else ego_point(c, bn, in_cam, g);
            long long ii[3];
            const int b = (int)div_magic((unsigned)bn, rd.m_N);
            if (voxel_of_point(d, b, g, ii) >= 0)
                row = ((b * d.nx + (int)ii[0]) * d.ny + (int)ii[1]) * d.nz + (int)ii[2];
            prow[cm] = row;
            sub2[cm] = make_int2(bn * d.HW + q.h * d.fW + w, (bn - b * d.N) * d.DHW + in_cam);      // {pixel = context row, flat index in the sample}
        }
        const unsigned peers = __match_any_sync(LSS_FULL_MASK, row >= 0 ? row : -1 - lane) & q.run_mask;
        int2 node = make_int2(0, 0);
        if (row >= 0 && lane == __ffs(peers) - 1) {      // sub-run leader: push on the voxel's list
            LSS_DASSERT(row < d.B * d.nx * d.ny * d.nz && cm < (size_t)d.n_points);
            const unsigned long long old = atomicExch(head + row, tag | (unsigned long long)(cm + 1));
            const int prev = (old >> 32) == (tag >> 32) ? (int)(unsigned)old : 0;      // heads of older builds read as empty
            LSS_DASSERT(prev >= 0 && prev <= d.n_points && prev != (int)cm + 1);
            node = make_int2(prev, (int)(peers >> lane));
            if (prev != 0) {    // the voxel is shared: queue it for the shared-voxel CTAs of the forward.  Every pusher but the first
                                // queues one record; only the one made by the SECOND pusher (whose `prev` pushed first) is taken up.
                const int shard = cta & (RP_QSHARDS - 1);
                unsigned long long *nrec = qcount + shard * (RP_FLAG_STRIDE / 2);
                atomicMax(nrec, tag);                     // a count of an older build reads as zero (same-address order)
                const unsigned k = (unsigned)atomicAdd(nrec, 1ull);
                recs[(size_t)k * RP_QSHARDS + shard] = make_int4(row, prev, (int)cm + 1, node.y);
            }
        }
        if (q.valid) sub[cm] = node;                      // {0, 0} wherever no sub-run starts
    }
    tl_stamp(1, true);
}

// Expand a sub-run into the flat point-in-sample indices ((n*D + d)*fH + h)*fW + w of its points, given the flat index `key0`
// of its leader and its row mask (bit j = image row h0 + j): point j goes to out[at + j] if that is below `cap`.
__device__ __forceinline__ int expand_keys(unsigned key0, unsigned mask, int fW, uint32_t *out, int at, int cap) {
    int j = 0;
    for (unsigned m = mask; m; m &= m - 1, ++j)
        if (at + j < cap) out[at + j] = key0 + (unsigned)(__ffs(m) - 1) * (unsigned)fW;
    return j;
}

// ------------------------------------------------------------------------------------------------
// zero-fill through the bulk-copy engine
// ------------------------------------------------------------------------------------------------

// Zero role of one CTA: thread 0 streams its share of the tensor out of a zeroed shared-memory chunk with bulk copies
// (cp.async.bulk shared -> global: the copies need no registers and no issue slots, the source is read-only, so all of them
// stay in flight).  The lines are marked evict-first in L2: 82 MB of zeros should not push the plan and the lift operands,
// which the gather reads, out of the cache.
#define ZERO_CHUNK (16 * 1024)
#define ZERO_DEPTH 3               // segments whose copies may be in flight behind the one being issued

__device__ __forceinline__ void zero_smem_init(float *s_zero) {
    for (int i = threadIdx.x; i < ZERO_CHUNK / 16; i += blockDim.x) reinterpret_cast<float4 *>(s_zero)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the async proxy
    __syncthreads();
}

__device__ __forceinline__ void zero_issue(char *dst, size_t bytes, int cta, int n_cta, unsigned src, unsigned long long pol) {
    const size_t n_chunks = (bytes + ZERO_CHUNK - 1) / ZERO_CHUNK;
    for (size_t ch = cta; ch < n_chunks; ch += n_cta) {
        const size_t off = ch * ZERO_CHUNK;
        const unsigned sz = (unsigned)min((size_t)ZERO_CHUNK, bytes - off);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                     :: "l"(dst + off), "r"(src), "r"(sz), "l"(pol) : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// whole tensor, no progress flags (k_prologue, lss_bev_zero)
__device__ __forceinline__ void zero_role(float *__restrict__ dst, size_t bytes, int cta, int n_cta, float *s_zero) {
    tl_stamp(0, false);
    zero_smem_init(s_zero);
    if (threadIdx.x != 0) return;
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    zero_issue(reinterpret_cast<char *>(dst), bytes, cta, n_cta, (unsigned)__cvta_generic_to_shared(s_zero), pol);
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    tl_stamp(0, true);
}

__device__ __forceinline__ int ld_acquire(const int32_t *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Poll until *p >= target.  The producers are CTAs that are already running (zero CTAs of this or of the preceding grid, the
// prologue's last CTA), so the wait is bounded by their work; after 2 s something is wrong with the calling sequence (e.g.
// lss_liftsplat_fwd_cl on a workspace whose counters were overwritten) and the kernel traps instead of hanging the GPU.
__device__ __forceinline__ int spin_until(const int32_t *p, int target, unsigned sleep_ns) {
    int v = ld_acquire(p);
    if (v >= target) return v;
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while ((v = ld_acquire(p)) < target) {
        __nanosleep(sleep_ns);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > 2000000000ull) __trap();
    }
    return v;
}

// The zeros of segment `seg` issued by this CTA have landed: make them visible device-wide, then count the CTA in.
__device__ __forceinline__ void zero_publish(int32_t *zero_done, int seg) {
    asm volatile("fence.proxy.async;" ::: "memory");      // async-proxy writes before the generic-proxy release below
    __threadfence();                                      // (publishing without the fences was measured 2 us SLOWER per forward)
    atomicAdd(zero_done + seg * RP_FLAG_STRIDE, 1);
}

// Segment by segment (one segment = the BEV slab of one sample) with a progress counter per segment: the column CTAs of
// sample b write their voxel rows as soon as zero_done[b] == n_cta, while the zeros of the later samples are still in flight.
// The chunk a CTA starts with rotates from segment to segment, so that every CTA issues the same number of chunks (+-1).
// At most `window` chunks of `chunk` bytes are in flight per CTA: the copies of all CTAs together should cover the
// bandwidth-delay product of the memory system and no more -- a deeper queue of zeros only delays the loads and atomics of
// the latency-bound CTAs that run next to the zero-fill.
__device__ __forceinline__ void zero_wait(int pending_allowed) {
    switch (pending_allowed) {
        case 0: asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); break;
        case 1: asm volatile("cp.async.bulk.wait_group 1;" ::: "memory"); break;
        case 2: asm volatile("cp.async.bulk.wait_group 2;" ::: "memory"); break;
        case 3: asm volatile("cp.async.bulk.wait_group 3;" ::: "memory"); break;
        default: asm volatile("cp.async.bulk.wait_group 7;" ::: "memory"); break;
    }
}

__device__ __forceinline__ void zero_role_segments(float *__restrict__ dst, size_t seg_bytes, int n_seg, int cta, int n_cta, float *s_zero,
                                                   int32_t *__restrict__ zero_done, int chunk, int window) {
    tl_stamp(0, false);
    zero_smem_init(s_zero);
    if (threadIdx.x != 0) return;
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    const unsigned src = (unsigned)__cvta_generic_to_shared(s_zero);
    const size_t cps = (seg_bytes + chunk - 1) / chunk;   // chunks per segment
    const int rot = (int)(cps % (size_t)n_cta);           // chunks of the last, partial round
    const int pend = window > 4 ? 7 : window - 1;         // groups that may stay pending behind the newest one
    int first = cta, issued = 0, pub = 0;                 // pub: next segment to publish
    int end_at[8];                                        // chunks issued up to the end of segment s, ring over s & 7 (every
                                                          // segment holds a chunk of this CTA, so at most pend + 1 <= 8 are open)
    for (int sgm = 0; sgm < n_seg; ++sgm) {
        char *base = reinterpret_cast<char *>(dst) + (size_t)sgm * seg_bytes;
        for (size_t ch = first; ch < cps; ch += n_cta) {
            const size_t off = ch * chunk;
            const unsigned sz = (unsigned)min((size_t)chunk, seg_bytes - off);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                         :: "l"(base + off), "r"(src), "r"(sz), "l"(pol) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            ++issued;
            zero_wait(pend);                              // the first issued - pend chunks are down
            while (pub < sgm && end_at[pub & 7] <= issued - pend) zero_publish(zero_done, pub++);
        }
        end_at[sgm & 7] = issued;
        while (pub <= sgm && end_at[pub & 7] <= issued - pend) zero_publish(zero_done, pub++);
        first -= rot;                                     // the CTAs right behind those of the partial round take the next one
        if (first < 0) first += n_cta;
    }
    zero_wait(0);
    while (pub < n_seg) zero_publish(zero_done, pub++);
    tl_stamp(0, true);
}

// Fused prologue of a step: ONE launch whose CTAs take one of three independent roles --
//   [0, n_zero)                    zero-fill of the BEV tensor (models.py:240) for lss_liftsplat_prologue(bev) (lss_liftsplat_forward
//                                  zero-fills with k_zero_flags)
//   [n_zero, n_zero + n_lift)      lift operands: depth softmax + pixel-major context (models.py:49-61)
//   [.., + n_index)                run index: geometry + voxel rows + list pushes (models.py:170-190, :212-221)
// (measured at cfg 2: lift before index 61.0 us per step, index before lift 63.1 -- the grid is a little more than one wave)
// The last lift / index CTA to finish publishes the build's epoch and, with `ready`, READY = 1 for the column CTAs of
// k_fwd_columns, which are launched programmatically behind this grid and poll the flag.
struct PrologueArgs {
    float *bev; size_t bev_bytes; int n_zero;                         // zero role (n_zero = 0: off)
    int n_index; RunDims rd; CalibPtrs c;                             // index role (n_index = 0: off)
    int32_t *prow, *counters; unsigned long long *head, *qcount; int2 *sub, *sub2; int4 *recs;
    int n_lift; const float *dn; float *prob, *ctx_t, *prob_col;      // lift role (n_lift = 0: off)
    int32_t *ready;                                                   // != null: raise READY when lift + index are complete
    float *clear_bev;                                                 // != null: the index role zeroes the rows the previous build's forward wrote
};

// `epoch0` (thread 0): the epoch of the plan in the workspace when this CTA started.
__device__ __forceinline__ void prologue_cta_done(const PrologueArgs &a, int epoch0) {
    if (a.counters == nullptr) return;
    __syncthreads();                                      // the CTA's writes are done ...
    if (threadIdx.x != 0) return;
    __threadfence();                                      // ... and ordered before the count
    if (atomicAdd(a.counters + RPC_PRO_DONE, 1) != a.n_index + a.n_lift - 1) return;
    __threadfence();                                      // the last CTA: everybody's writes are visible to it, and through
    const int epoch = (int)((unsigned)epoch0 + (a.n_index ? 1u : 0u));   // READY carries the epoch of the plan to use, as (epoch & 0x7fffffff) + 1:
    // non-zero (raised) for every epoch, an unbuilt plan (epoch 0: all-empty lists, a forward of zeros) included; the forward compares
    // 31 bits of the epoch with the list heads' 32 -- enough to tell a build from the ones before it
    if (a.ready) for (int i = 0; i < RP_READY_LINES; ++i) atomicExch(a.ready + i * RP_FLAG_STRIDE, (epoch & 0x7fffffff) + 1);
    a.counters[RPC_PRO_DONE] = 0;
    a.counters[RPC_EPOCH] = epoch;                        // every index CTA has read the old epoch
}

template <bool RAW>
__global__ void __launch_bounds__(RP_THREADS, 2048 / RP_THREADS)      // 32 registers: 8 CTAs per SM
k_prologue(Dims d, PrologueArgs a) {
    extern __shared__ __align__(128) float s_pro[];
    lss_pdl_trigger();                                    // k_fwd_columns may be scheduled while this grid runs
    int cta = (int)blockIdx.x;
    if (cta < a.n_zero) { zero_role(a.bev, a.bev_bytes, cta, a.n_zero, s_pro); return; }
    const int epoch0 = (threadIdx.x == 0 && a.counters != nullptr) ? __ldcg(a.counters + RPC_EPOCH) : 0;
    cta -= a.n_zero;
    if (cta < a.n_lift) {
        tl_stamp(2, false);
        lift_prepare_cta<float>(d, a.dn, a.prob, a.ctx_t, a.prob_col, cta, s_pro);
        tl_stamp(2, true);
    } else {
        run_index_cta<RAW>(d, a.rd, a.c, a.prow, a.head, a.sub, a.sub2, a.recs, a.qcount, a.counters, cta - a.n_lift, a.clear_bev);
    }
    prologue_cta_done(a, epoch0);
}

// The zero-fill of lss_liftsplat_forward: one warp per CTA, two CTAs per SM, launched FIRST so that they are spread evenly over
// the SMs (a grid that is launched programmatically into a busy GPU gets its CTAs wherever room appears; several zero CTAs on
// one SM make its bulk-copy engine the bottleneck: 57 us instead of 15 measured).  It lets its successors start at once.
__global__ void __launch_bounds__(32)
k_zero_flags(float *bev, size_t seg_bytes, int n_seg, int32_t *zero_done, int chunk, int window) {
    extern __shared__ __align__(128) float s_z[];
    lss_pdl_trigger();
    zero_role_segments(bev, seg_bytes, n_seg, (int)blockIdx.x, (int)gridDim.x, s_z, zero_done, chunk, window);
}

// ------------------------------------------------------------------------------------------------
// forward: zero-fill + classify + gather in one launch
// ------------------------------------------------------------------------------------------------

struct FwdArgs {
    int n_zero, z_chunk, z_window;   // zero CTAs of the column grid (LSS_ZERO_ORDERED), their chunk size and window
    int zero_target;                 // a sample's slab is clear when zero_done[b] reaches this (0: pre-cleared, no waiting)
    int wait_ready;                  // lss_liftsplat_forward: poll READY instead of waiting for the preceding grid to complete
    size_t seg_bytes;                // bytes of one sample's BEV slab
    int32_t *zero_done, *ready;      // [B][32], [32][32]
    int n_keys, fWD, n_cons, cons_first;   // camera columns B*N*fW; fW*D; shared-voxel CTAs; their place in the grid
    unsigned long long mfH;
    const int32_t *prow; const int2 *sub, *sub2; const unsigned long long *head;
    int32_t *counters; uint32_t *pool;
    const unsigned long long *qcount; // [32] sub-queue counters, 128 bytes apart
    const int4 *recs; long long n_rec_cap; // shared voxels queued by the index pass {voxel row, previous pusher, this pusher (point + 1), its row mask}
    int4 *longs;                     // forward scratch: overflow list of long voxels {voxel row} (end of the record array)
    const float *prob_col, *ctx_t; float *bev;
};


#define GCL_LONG_CAP 64            // long voxels a shared-voxel CTA keeps for itself (more: the last CTA of the grid takes them)
#define GCL_PRE 3                  // exclusive voxels per 8-lane group summed before the CTA waits for its zeros
#define GCL_ROWS 16                // context rows a warp has in flight for a shared voxel (cp.async into shared memory)

// fl32(acc + fl32(w * v)) on two channels: scalar products, ONE packed add (sm_100 FADD2) -- the same two roundings per channel
// as __fmul_rn + __fadd_rn with three instructions instead of four.  A packed product feeding the packed add (__fmul2_rn into
// __fadd2_rn, or mul.rn.f32x2 into add.rn.f32x2 in PTX, with or without --fmad=false) is contracted by ptxas 12.9 into one
// FFMA2, i.e. a fused multiply-add with a single rounding (measured: 1-ulp differences against the sequential definition;
// profiles/r02_sass_excerpt.txt); FMUL + FADD2 is not.
__device__ __forceinline__ void mul_add2(float &a0, float &a1, float w, float v0, float v1) {
    const float2 r = __fadd2_rn(make_float2(a0, a1), make_float2(__fmul_rn(w, v0), __fmul_rn(w, v1)));
    a0 = r.x; a1 = r.y;
}

// A voxel with >= 64 points, summed by a whole CTA.  Scratch (s_dyn): [SORT_CAP] keys, [NG][C] products.  Keys are sorted by the CTA
// (shared memory up to SORT_CAP points, else in the pool), then NG points per pass: every group fetches one point's context
// row and writes float32(prob*ctx) to shared memory (all loads in flight together); thread c adds the NG products of channel c
// in ascending point order -- the same sequence of float32 additions as everywhere else.
template <int CPL>
__device__ __forceinline__ void sum_long_voxel(const Dims &d, const FwdArgs &a, const int row, float *s_dyn) {
    constexpr int C = 8 * CPL;
    __shared__ int s_len, s_pos;
    const int lane = threadIdx.x & 31, gl = lane & 7, g = threadIdx.x >> 3;
    const int HWC = d.HW * C;
    const int b = row / (d.nx * d.ny * d.nz);
    auto decode = [&](unsigned pidx, int &ro, size_t &wi) {   // context row offset (floats) in the sample and prob_col index of a point
        const unsigned cam = lss_div20(pidx, d.mDHW);
        const unsigned rr = pidx - cam * d.DHW;
        const unsigned dd = lss_div20(rr, d.mHW);
        const unsigned hw = rr - dd * d.HW;
        const unsigned h = lss_div20(hw, d.mfW), ww = hw - h * d.fW;
        const unsigned bnn = (unsigned)b * d.N + cam;
        ro = (int)(cam * HWC + hw * C);
        wi = ((size_t)(bnn * d.fW + ww) * d.D + dd) * d.fH + h;
    };
    uint32_t *s_keys = reinterpret_cast<uint32_t *>(s_dyn);
    float *s_prod = s_dyn + GCL_SORT_CAP;
    const int hd = (int)(unsigned)__ldcg(a.head + row);
    if (threadIdx.x == 0) {                               // points of the voxel; room in the pool if they do not fit shared memory
        int c = 0;
        for (int cur = hd; cur != 0;) { const int2 nd = __ldcg(a.sub + (cur - 1)); c += __popc((unsigned)nd.y); cur = nd.x; }
        s_len = c;
        s_pos = c > GCL_SORT_CAP ? atomicAdd(a.counters + RPC_POOL, c) : 0;
        LSS_DASSERT(c >= GCL_SHORT_CAP && s_pos >= 0 && s_pos + c <= d.n_points);
        if (a.zero_target) spin_until(a.zero_done + b * RP_FLAG_STRIDE, a.zero_target, 200);
    }
    __syncthreads();
    const int npt = s_len;
    const bool in_smem = npt <= GCL_SORT_CAP;
    uint32_t *gk = a.pool + s_pos;
    if (threadIdx.x == 0) {
        uint32_t *out = in_smem ? s_keys : gk;
        int i = 0;
        for (int cur = hd; cur != 0;) {
            const int2 nd = __ldcg(a.sub + (cur - 1));
            i += expand_keys((unsigned)__ldcg(a.sub2 + (cur - 1)).y, (unsigned)nd.y, d.fW, out, i, npt);
            cur = nd.x;
        }
    }
    __syncthreads();
    if (in_smem) bitonic_sort_block(s_keys, npt);
    else bitonic_sort_block((volatile uint32_t *)gk, npt);
    __syncthreads();
    const float *ctx_s = a.ctx_t + (size_t)b * d.N * HWC;
    float accc = 0.f;
    for (int p0 = 0; p0 < npt; p0 += GCL_NG) {
        const int cntp = min(GCL_NG, npt - p0);
        if (g < cntp) {
            const unsigned pidx = in_smem ? s_keys[p0 + g] : ((volatile uint32_t *)gk)[p0 + g];
            int ro; size_t wi;
            decode(pidx, ro, wi);
            const float w = __ldcg(a.prob_col + wi);
            const float4 *rowp = reinterpret_cast<const float4 *>(ctx_s + ro + gl * 4);
            float4 *dst = reinterpret_cast<float4 *>(s_prod + g * C) + gl;
#pragma unroll
            for (int q = 0; q < CPL / 4; ++q) {
                const float4 v = __ldcg(rowp + 8 * q);
                dst[8 * q] = make_float4(__fmul_rn(w, v.x), __fmul_rn(w, v.y), __fmul_rn(w, v.z), __fmul_rn(w, v.w));
            }
        }
        __syncthreads();
        if ((int)threadIdx.x < C)
            for (int jj = 0; jj < cntp; ++jj) accc = __fadd_rn(accc, s_prod[jj * C + threadIdx.x]);
        __syncthreads();
    }
    if ((int)threadIdx.x < C) a.bev[(size_t)row * C + threadIdx.x] = accc;
    __syncthreads();
}

// End of a column / shared-voxel CTA.  The last CTA of the grid to get here sums the long voxels that did not fit a CTA's own
// list and leaves the scratch of the forward clean: by then every CTA has seen the flags it waits for.
template <int CPL>
__device__ __forceinline__ void forward_cta_done(const Dims &d, const FwdArgs &a, float *s_dyn, int n_done, int n_long_cta) {
    __shared__ int s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        if (n_done) atomicAdd(a.counters + RPC_NSHARED, n_done);
        if (n_long_cta) atomicAdd(a.counters + RPC_NLONG_CTA, n_long_cta);
        __threadfence();
        s_last = atomicAdd(a.counters + RPC_FWD_DONE, 1) == a.n_keys + a.n_cons - 1;
        if (s_last) __threadfence();
    }
    __syncthreads();
    if (!s_last) return;
    const int n_long = __ldcg(a.counters + RPC_NLONG);
    for (int l = 0; l < n_long; ++l) sum_long_voxel<CPL>(d, a, __ldcg(a.longs + l).x, s_dyn);
    if (threadIdx.x == 0) {
        const int n_long_all = n_long + atomicExch(a.counters + RPC_NLONG_CTA, 0);
        a.counters[RPC_STAT] = atomicExch(a.counters + RPC_NSHARED, 0) + n_long_all;
        a.counters[RPC_STAT + 1] = n_long_all;
        a.counters[RPC_NLONG] = 0; a.counters[RPC_POOL] = 0; a.counters[RPC_FWD_DONE] = 0;
        if (a.wait_ready) for (int i = 0; i < RP_READY_LINES; ++i) a.ready[i * RP_FLAG_STRIDE] = 0;
        if (a.zero_target) for (int i = 0; i < d.B; ++i) a.zero_done[i * RP_FLAG_STRIDE] = 0;
    }
}

// Shared voxels (several sub-runs on the voxel's list, about 5 % of the voxels: neighbouring columns at close range, the
// overlap of neighbouring cameras), spread evenly over the last `n_cons` CTAs of the forward grid: one warp per voxel, from the
// queue the index pass made -- every pusher that found a sub-run on the list before it queued {voxel, previous, itself, its row
// mask}; the record whose `previous` pushed first is the voxel's, the others are skipped (96 % of the shared voxels hold two
// sub-runs and one record).  The warp requests the previous sub-run's node, the two sub-runs' pixel / key words and the list
// head together; sub-runs pushed later hang between the head and the record's maker.  Lane j takes point j of a sub-run; the warp
// rank-sorts the few keys by flat point index, fetches the weights (lane = point) and up to GCL_ROWS context rows at once
// (cp.async: all of them in flight, no registers), waits for the zeros of the voxel's sample and adds the products in key order,
// lane = C/32 channels: the same sequence of float32 operations per channel as everywhere else.  Voxels with >= 64 points are
// summed by the whole CTA afterwards.  The last of these CTAs to finish leaves the scratch of the forward clean.
template <int CPL>
__device__ __forceinline__ void shared_voxels_cta(const Dims &d, const FwdArgs &a, int cons, int epoch, float *s_dyn) {
    constexpr int C = 8 * CPL;
    constexpr int CH = C / 32;                            // channels per lane
    constexpr int PR = C / 4;                             // 16-byte pieces per context row
    constexpr int NW = GCL_THREADS / 32;
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    const int vps = d.nx * d.ny * d.nz;                   // voxels per sample
    __shared__ int s_nlong, s_ndone;
    __shared__ int s_longrow[GCL_LONG_CAP];
    if (threadIdx.x == 0) { s_nlong = 0; s_ndone = 0; }
    __syncthreads();
    tl_stamp(4, false);
    // the sub-queues were filled by the index pass of the build this forward reads: lane s holds the count of shard s
    const unsigned long long q64 = __ldcg(a.qcount + lane * (RP_FLAG_STRIDE / 2));
    const int q_cnt = ((unsigned)(q64 >> 32) & 0x7fffffffu) == ((unsigned)epoch & 0x7fffffffu) ? (int)(unsigned)q64 : 0;
    const int n_rec = __reduce_max_sync(LSS_FULL_MASK, q_cnt) * RP_QSHARDS;      // slots to look at (the tail of a short shard is skipped)
    {
        // per warp: keys, context-row numbers and weight indices of the voxel's points in list order [3][64], their order
        // by key [64], then the context rows [GCL_ROWS][C]
        uint32_t *w_key = reinterpret_cast<uint32_t *>(s_dyn) + wp * (4 * GCL_SHORT_CAP);
        uint32_t *w_pix = w_key + GCL_SHORT_CAP, *w_widx = w_pix + GCL_SHORT_CAP, *w_ord = w_widx + GCL_SHORT_CAP;
        float *w_rows = s_dyn + NW * 4 * GCL_SHORT_CAP + wp * (GCL_ROWS * C);
        const int n_warps = a.n_cons * NW;
        auto take = [&](int cm, unsigned m, int2 n2, int len) {      // lane j takes point j of a sub-run (leader cm, row mask m)
            const int cnt = __popc(m);
            if (lane < cnt && len + lane < GCL_SHORT_CAP) {
                const unsigned k = __fns(m, 0, lane + 1);             // image-row offset of this lane's point
                w_key[len + lane] = (unsigned)n2.y + k * (unsigned)d.fW;
                w_pix[len + lane] = (unsigned)n2.x + k * (unsigned)d.fW;
                w_widx[len + lane] = (unsigned)cm + k;
            }
            return cnt;
        };
        for (int r = cons * NW + wp; r < n_rec; r += n_warps) {
            if ((r >> 5) >= __shfl_sync(LSS_FULL_MASK, q_cnt, r & (RP_QSHARDS - 1))) continue;      // beyond the end of its shard
            const int4 rec = __ldcg(a.recs + r);          // {voxel row, previous pusher, the pusher that made the record, its row mask}
            LSS_DASSERT(rec.x >= 0 && rec.x < d.B * vps && rec.y >= 1 && rec.y <= d.n_points && rec.z >= 1 && rec.z <= d.n_points);
            const int2 nf = __ldcg(a.sub + (rec.y - 1));  // all of these are requested together
            const int2 nf2 = __ldcg(a.sub2 + (rec.y - 1));
            const int2 nc2 = __ldcg(a.sub2 + (rec.z - 1));
            const int hd = (int)(unsigned)__ldcg(a.head + rec.x);
            if (nf.x != 0) continue;                      // `previous` did not push first: the record of the second pusher covers the voxel
            int len = take(rec.y - 1, (unsigned)nf.y, nf2, 0);
            len += take(rec.z - 1, (unsigned)rec.w, nc2, len);
            for (int cur = hd; cur != rec.z;) {           // sub-runs pushed after the second one (4 % of the shared voxels)
                LSS_DASSERT(cur >= 1 && cur <= d.n_points);
                const int2 nd = __ldcg(a.sub + (cur - 1)), nd2 = __ldcg(a.sub2 + (cur - 1));
                LSS_DASSERT(nd.y != 0 && nd.x != 0 && __ldcg(a.prow + (cur - 1)) == rec.x);
                len += take(cur - 1, (unsigned)nd.y, nd2, len);
                cur = nd.x;
            }
            if (len >= GCL_SHORT_CAP) {                   // long voxel: for the whole CTA, after its warps are through
                if (lane == 0) {
                    const int pos = atomicAdd(&s_nlong, 1);
                    if (pos < GCL_LONG_CAP) s_longrow[pos] = rec.x;
                    else a.longs[atomicAdd(a.counters + RPC_NLONG, 1)] = make_int4(rec.x, 0, 0, 0);      // (overflow: the last CTA)
                }
                continue;
            }
            if (lane == 0) atomicAdd(&s_ndone, 1);
            __syncwarp();
            for (int i = lane; i < len; i += 32) {        // rank sort by flat point index (keys unique, len < 64)
                const uint32_t e = w_key[i];
                int rank = 0;
                for (int j = 0; j < len; ++j) rank += w_key[j] < e ? 1 : 0;
                w_ord[rank] = (uint32_t)i;
            }
            __syncwarp();
            float acc[CH];
#pragma unroll
            for (int k = 0; k < CH; ++k) acc[k] = 0.f;
            for (int p0 = 0; p0 < len; p0 += GCL_ROWS) {
                const int n = min(GCL_ROWS, len - p0);
                float w = 0.f;
                unsigned pix = 0u;
                if (lane < n) {
                    const uint32_t i = w_ord[p0 + lane];
                    pix = w_pix[i];
                    w = __ldcg(a.prob_col + w_widx[i]);
                }
                for (int k0 = 0; k0 < n; k0 += 32 / PR) { // 32 / PR rows per instruction, lane = (row, 16-byte piece)
                    const int k = k0 + lane / PR, piece = lane % PR;
                    const unsigned pix_k = __shfl_sync(LSS_FULL_MASK, pix, k & 31);
                    if (k < n) {
                        const unsigned dst = (unsigned)__cvta_generic_to_shared(w_rows + k * C + piece * 4);
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(a.ctx_t + (size_t)pix_k * C + piece * 4) : "memory");
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncwarp();
                for (int k = 0; k < n; ++k) {
                    const float wk = __shfl_sync(LSS_FULL_MASK, w, k);
                    const float *rp = w_rows + k * C + lane * CH;
                    if (CH >= 2) {
#pragma unroll
                        for (int c = 0; c < CH; c += 2) mul_add2(acc[c], acc[c + (CH >= 2 ? 1 : 0)], wk, rp[c], rp[c + (CH >= 2 ? 1 : 0)]);
                    } else {
                        acc[0] = __fadd_rn(acc[0], __fmul_rn(wk, rp[0]));
                    }
                }
                __syncwarp();
            }
            LSS_DASSERT(len >= 2);
            if (a.zero_target && lane == 0) spin_until(a.zero_done + (rec.x / vps) * RP_FLAG_STRIDE, a.zero_target, 200);
            __syncwarp();
            float *dst = a.bev + (size_t)rec.x * C + lane * CH;
#pragma unroll
            for (int c = 0; c < CH; ++c) dst[c] = acc[c];
        }
    }
    tl_stamp(4, true);
    __syncthreads();
    const int n_long_cta = min(s_nlong, GCL_LONG_CAP);    // long voxels met by this CTA's warps: one at a time by the whole CTA
    for (int l = 0; l < n_long_cta; ++l) sum_long_voxel<CPL>(d, a, s_longrow[l], s_dyn);
    forward_cta_done<CPL>(d, a, s_dyn, s_ndone, n_long_cta);
}

// The forward grid.  [0, n_zero): zero CTAs (LSS_ZERO_ORDERED only).  [.., + n_cons): the shared voxels, from the queue the index
// pass made (shared_voxels_cta): they depend on nothing the column CTAs do, are few and take the longest, so they get their SM
// slots first.  [.., + n_keys): one CTA per camera column (bn, w): stage the column's operands and plan, find the sub-runs that
// are alone on their voxel's list (EXCLUSIVE), sum them from the staged operands and write their voxel rows.
template <int CPL>
__global__ void __launch_bounds__(GCL_THREADS, 9)        // <= 56 registers: 8 column CTAs per SM next to the two zero CTAs
k_fwd_columns(Dims d, FwdArgs a) {
    extern __shared__ __align__(128) float s_dyn[];
    constexpr int C = 8 * CPL;
    if ((int)blockIdx.x < a.n_zero) {                     // ---- zero CTAs (scheduled first: nobody they depend on comes later)
        lss_pdl_wait();
        zero_role_segments(a.bev, a.seg_bytes, d.B, (int)blockIdx.x, a.n_zero, s_dyn, a.zero_done, a.z_chunk, a.z_window);
        return;
    }
    __shared__ int s_epoch;
    if (a.wait_ready) {                                   // the zero-fill grid is still streaming: only the plan + lift grid counts
        if (threadIdx.x == 0) s_epoch = spin_until(a.ready + (blockIdx.x % RP_READY_LINES) * RP_FLAG_STRIDE, 1, 100) - 1;   // READY = the plan's epoch + 1
    } else {
        lss_pdl_wait();                                   // plan and lift operands come from the preceding kernel(s)
        if (threadIdx.x == 0) s_epoch = __ldcg(a.counters + RPC_EPOCH);
    }
    __syncthreads();
    const int rel = (int)blockIdx.x - a.n_zero;
    if (a.cons_first ? rel < a.n_cons : rel >= a.n_keys) { shared_voxels_cta<CPL>(d, a, a.cons_first ? rel : rel - a.n_keys, s_epoch, s_dyn); return; }
    tl_stamp(3, false);
    const int lane = threadIdx.x & 31, gl = lane & 7, g = threadIdx.x >> 3;
    const int per = d.D * d.fH;
    const int key = a.cons_first ? rel - a.n_cons : rel;  // camera column (bn, w)
    const int bn = key / d.fW, w0 = key - bn * d.fW;
    const int b = bn / d.N;
    tl_mark(key, 0);
    float *s_ctx = s_dyn;                                 // [fH][C]
    float *s_prob = s_dyn + d.fH * C;                     // [D][fH]
    int *s_row = reinterpret_cast<int *>(s_prob + per);   // [per] voxel row
    unsigned *s_aux = reinterpret_cast<unsigned *>(s_row + per);     // [per] row mask of an EXCLUSIVE leader / 0
    unsigned short *s_list = reinterpret_cast<unsigned short *>(s_aux + per);      // [per] slots of the EXCLUSIVE leaders
    __shared__ int s_n;
    if (threadIdx.x == 0) s_n = 0;
    const unsigned long long tag = (unsigned long long)((unsigned)s_epoch & 0x7fffffffu) << 32;      // (31 bits: see prologue_cta_done)
    const unsigned long long tag_mask = 0x7fffffffffffffffull;
    const size_t base = (size_t)key * per;
    constexpr int SU = 3;                                 // slots per thread and round: their loads are all in flight together
    constexpr int c4 = C >> 2;
    const float4 *ctx_src = reinterpret_cast<const float4 *>(a.ctx_t + ((size_t)bn * d.HW + w0) * C);
    const int n_c4 = d.fH * c4;
    float4 cv[2];                                         // the column's context rows (<= 2 pieces per thread: requested first, stored last)
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int i = threadIdx.x + u * GCL_THREADS;
        if (i < n_c4) { const int h = i / c4, q = i - h * c4; cv[u] = __ldcg(ctx_src + (size_t)h * d.fW * c4 + q); }
    }
    __syncthreads();                                      // s_n
    for (int i0 = 0; i0 < per; i0 += SU * GCL_THREADS) {  // stage the column's plan + weights, classify and compact its leaders
        int2 node[SU]; int row[SU]; float pw[SU]; unsigned long long hd[SU];
#pragma unroll
        for (int u = 0; u < SU; ++u) {
            const int i = i0 + u * GCL_THREADS + threadIdx.x;
            node[u] = make_int2(0, 0); row[u] = -1; pw[u] = 0.f;
            if (i < per) {
                node[u] = __ldcg(a.sub + base + i);       // {previous head of the voxel's list, row mask}
                row[u] = __ldcg(a.prow + base + i);
                pw[u] = __ldcg(a.prob_col + base + i);
            }
        }
#pragma unroll
        for (int u = 0; u < SU; ++u) {                    // a sub-run that pushed first answers for its voxel: is it still the head?
            hd[u] = 0ull;
            if (node[u].y != 0 && node[u].x == 0) {
                LSS_DASSERT(row[u] >= 0 && row[u] < d.B * d.nx * d.ny * d.nz);
                hd[u] = __ldcg(a.head + row[u]);
            }
        }
        unsigned hb[SU];                                  // which lanes hold an EXCLUSIVE leader in slot u
#pragma unroll
        for (int u = 0; u < SU; ++u) {
            const int i = i0 + u * GCL_THREADS + threadIdx.x;
            unsigned em = 0u;
            if (i < per) {
                s_row[i] = row[u];
                s_prob[i] = pw[u];
                unsigned aux = 0u;
                if (hd[u] != 0ull) {
                    LSS_DASSERT(((hd[u] & tag_mask) >> 32) == (tag >> 32) && (unsigned)hd[u] >= 1u && (unsigned)hd[u] <= (unsigned)d.n_points);
                    // nobody pushed after it: EXCLUSIVE.  (Otherwise the voxel is shared: the index pass has queued it for the
                    // shared-voxel CTAs at the end of this grid.)
                    if ((hd[u] & tag_mask) == (tag | (unsigned long long)(base + i + 1))) aux = em = (unsigned)node[u].y;
                }
                s_aux[i] = aux;
            }
            hb[u] = __ballot_sync(LSS_FULL_MASK, em != 0u);
        }
        int total = 0;                                    // one reservation in the CTA's list for the warp's leaders of all SU slots
#pragma unroll
        for (int u = 0; u < SU; ++u) total += __popc(hb[u]);
        int wbase = 0;
        if (lane == 0 && total) wbase = atomicAdd(&s_n, total);
        wbase = __shfl_sync(LSS_FULL_MASK, wbase, 0);
        LSS_DASSERT(wbase + total <= per);
#pragma unroll
        for (int u = 0; u < SU; ++u) {
            if ((hb[u] >> lane) & 1u) s_list[wbase + __popc(hb[u] & ((1u << lane) - 1u))] = (unsigned short)(i0 + u * GCL_THREADS + threadIdx.x);
            wbase += __popc(hb[u]);
        }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int i = threadIdx.x + u * GCL_THREADS;
        if (i < n_c4) reinterpret_cast<float4 *>(s_ctx)[i] = cv[u];
    }
    for (int i = threadIdx.x + 2 * GCL_THREADS; i < n_c4; i += GCL_THREADS) {      // (tall feature maps: the rest of the rows)
        const int h = i / c4, q = i - h * c4;
        reinterpret_cast<float4 *>(s_ctx)[i] = __ldcg(ctx_src + (size_t)h * d.fW * c4 + q);
    }
    tl_mark(key, 1);
    __syncthreads();
    // The sums of the first GCL_PRE exclusive voxels of every group are made BEFORE the CTA waits for its sample's zeros: when
    // the zero-fill is the last thing to finish, only the stores are left behind it.
    const int n_list = s_n;
    const float *my_ctx = s_ctx + gl * 4;
    auto sum_exclusive = [&](int i, float (&acc)[CPL]) -> int {      // voxel row of leader i of the list (or -1), its sum in acc
        const bool live = i < n_list;
        const int s = live ? (int)s_list[i] : 0;
        const unsigned m = live ? s_aux[s] : 0u;
        const int h0 = s - (int)lss_div20((unsigned)s, a.mfH) * d.fH;
#pragma unroll
        for (int k = 0; k < CPL; ++k) acc[k] = 0.f;
        const float *wp_ = s_prob + s;
        const float *rowp0 = my_ctx + h0 * C;
#pragma unroll 8
        for (int j = 0; j < d.fH; ++j) {
            if ((m >> j) & 1u) {                          // bit j set => image row h0 + j < fH of the same run
                const float wj = wp_[j];
                const float4 *rowp = reinterpret_cast<const float4 *>(rowp0 + j * C);
#pragma unroll
                for (int q = 0; q < CPL / 4; ++q) {
                    const float4 v = rowp[8 * q];
                    mul_add2(acc[4 * q], acc[4 * q + 1], wj, v.x, v.y);
                    mul_add2(acc[4 * q + 2], acc[4 * q + 3], wj, v.z, v.w);
                }
            }
        }
        LSS_DASSERT(!live || (s_row[s] >= 0 && s_row[s] < d.B * d.nx * d.ny * d.nz && (m >> (d.fH - h0)) == 0u));
        return live ? s_row[s] : -1;
    };
    auto store_row = [&](int row, const float (&acc)[CPL]) {
        if (row < 0) return;
        float4 *dst = reinterpret_cast<float4 *>(a.bev + (size_t)row * C) + gl;
#pragma unroll
        for (int q = 0; q < CPL / 4; ++q) dst[8 * q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
    };
    float pre[GCL_PRE][CPL];
    int pre_row[GCL_PRE];
#pragma unroll
    for (int t = 0; t < GCL_PRE; ++t) pre_row[t] = sum_exclusive(g + t * GCL_NG, pre[t]);
    if (threadIdx.x == 0 && a.zero_target)                // the zeros of this sample's slab must be down before any row is written
        spin_until(a.zero_done + b * RP_FLAG_STRIDE, a.zero_target, 400);
    __syncthreads();
    tl_mark(key, 2);
#pragma unroll
    for (int t = 0; t < GCL_PRE; ++t) store_row(pre_row[t], pre[t]);
    for (int i = g + GCL_PRE * GCL_NG; __any_sync(LSS_FULL_MASK, i < n_list); i += GCL_NG) {
        float acc[CPL];
        store_row(sum_exclusive(i, acc), acc);
    }
    tl_mark(key, 3);
    tl_stamp(3, true);
    forward_cta_done<CPL>(d, a, s_dyn, 0, 0);
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
    rd.m_fH = magic_of((unsigned)p->fH); rd.m_fWD = magic_of((unsigned)rd.fWD); rd.m_D = magic_of((unsigned)p->D); rd.m_N = magic_of((unsigned)p->N);
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
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    size_t off = 0;
    out->off_prow = off;       off += up((size_t)d.n_points * 4);
    out->off_sub = off;        off += up((size_t)d.n_points * 8);
    out->off_sub2 = off;       off += up((size_t)d.n_points * 8);
    out->off_pool = off;       off += up((size_t)d.n_points * 4);
    // one record per sub-run that is not the first on its voxel's list, in RP_QSHARDS interleaved sub-queues: room for the
    // shard that gets one index CTA (<= 512 points) more than its share
    out->n_rec_cap = (int64_t)d.n_points + RP_QSHARDS * 512 + 2;
    out->off_recs = off;       off += up((size_t)out->n_rec_cap * 16);
    out->off_longs = off;      off += up(((size_t)d.n_points / GCL_SHORT_CAP + 2) * 16);
    out->off_counters = off;   off += up(64 * 4);
    out->off_qcount = off;     off += up((size_t)RP_QSHARDS * RP_FLAG_STRIDE * 4);
    out->off_zero_done = off;  off += up((size_t)p->B * RP_FLAG_STRIDE * 4);
    out->off_ready = off;      off += up((size_t)RP_READY_LINES * RP_FLAG_STRIDE * 4);
    out->off_head = off;       off += up((size_t)out->n_voxels * 8);
    out->bytes = off;
    return LSS_OK;
}

extern "C" int lss_runplan_reset(const lss_runplan_layout *L, void *ws, void *stream) {
    LSS_REQUIRE(L && ws, LSS_ERR_WORKSPACE);
    char *w = (char *)ws;    // counters, zero_done and head are contiguous
    if (cudaMemsetAsync(w + L->off_counters, 0, L->bytes - L->off_counters, (cudaStream_t)stream) != cudaSuccess) return LSS_ERR_CUDA;
    return LSS_OK;
}

// cameras one index CTA may span when it makes the calibration inverses itself
static bool raw_build_fits(const lss_problem *p) {
    const RunDims rd = make_run_dims(p);
    return (RP_WARPS * RP_U * rd.RPW) / rd.fWD + 2 <= LSS_RAW_CAMS;
}

extern "C" int lss_runplan_raw_supported(const lss_problem *p) {
    int st = runplan_supported(p);
    if (st != LSS_OK) return st;
    return raw_build_fits(p) ? LSS_OK : LSS_ERR_UNSUPPORTED;
}

// Pacing of the zero-fill with progress counters: CTAs, bytes per bulk copy, copies in flight per CTA.  Measured at cfg 2 next to
// the other kernels of the forward: between (1 CTA per SM, 2 in flight) and (2 per SM, 4 in flight) the
// forward is flat within 0.5 us; more in flight (4 per SM, or 8 deep) costs 3-5 us -- a deeper queue of zeros delays the loads
// and atomics of the latency-bound CTAs next to it -- and half a CTA per SM stretches the zero-fill itself by as much.
// `own_grid`: k_zero_flags in front of prologue + columns (1 per SM, 4 deep) / the first CTAs of the column grid (2 per SM, 2 deep).
struct ZeroTune { int n_cta, chunk, window; };
static ZeroTune zero_tune(size_t seg_bytes, bool own_grid) {
    ZeroTune t = {own_grid ? rp_num_sms() : 2 * rp_num_sms(), ZERO_CHUNK, own_grid ? 4 : 2};
    const size_t cps = (seg_bytes + t.chunk - 1) / t.chunk;
    if ((size_t)t.n_cta > cps) t.n_cta = (int)cps;
    return t;
}

// `ready`: launched programmatically behind k_zero_flags (no dependence on it); the last lift / index CTA raises READY
static int launch_prologue(const lss_problem *p, const lss_runplan_layout *L, void *workspace, bool index, const CalibPtrs &c, bool raw,
                           const float *dn, float *prob, float *ctx_t, float *prob_col, float *bev, size_t bev_bytes, bool ready,
                           cudaStream_t s, bool clear_rows = false) {
    const Dims d = make_dims(p);
    PrologueArgs a = {};
    a.rd = make_run_dims(p);
    char *w = (char *)workspace;
    if (w != nullptr) a.counters = (int32_t *)(w + L->off_counters);
    if (clear_rows) {                           // persistent output: no zero role, the index role clears what the last forward wrote
        a.clear_bev = index ? bev : nullptr;
    } else if (bev != nullptr && bev_bytes > 0) {      // two issuing CTAs per SM (one: 67.5 us per step at cfg 2 against 63)
        a.bev = bev; a.bev_bytes = bev_bytes;
        a.n_zero = (int)min((size_t)2 * rp_num_sms(), (bev_bytes + ZERO_CHUNK - 1) / ZERO_CHUNK);
    }
    if (ready) a.ready = (int32_t *)(w + L->off_ready);
    if (index) {
        const int runs_per_cta = RP_WARPS * RP_U * a.rd.RPW;
        if (raw) LSS_REQUIRE(raw_build_fits(p), LSS_ERR_UNSUPPORTED);
        a.n_index = (a.rd.R + runs_per_cta - 1) / runs_per_cta;
        a.c = c;
        a.prow = (int32_t *)(w + L->off_prow); a.head = (unsigned long long *)(w + L->off_head);
        a.sub = (int2 *)(w + L->off_sub); a.sub2 = (int2 *)(w + L->off_sub2); a.recs = (int4 *)(w + L->off_recs);
        a.qcount = (unsigned long long *)(w + L->off_qcount);
    }
    size_t smem = a.n_zero ? ZERO_CHUNK : 0;
    if (dn != nullptr) {
        a.n_lift = p->B * p->N * ((d.HW + LIFT_PX - 1) / LIFT_PX);
        a.dn = dn; a.prob = prob; a.ctx_t = ctx_t; a.prob_col = prob_col;
        smem = max(smem, (size_t)(d.D + d.C) * (LIFT_PX + 1) * sizeof(float));
    }
    const int grid = a.n_zero + a.n_index + a.n_lift;
    if (grid == 0) return LSS_OK;
    if (ready) LSS_REQUIRE(a.n_index + a.n_lift > 0, LSS_ERR_BAD_ARG);      // somebody has to raise READY
    if (smem > 227 * 1024) return LSS_ERR_UNSUPPORTED;
    auto kern = raw ? k_prologue<true> : k_prologue<false>;
    if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return LSS_ERR_CUDA;
    // programmatic launch only behind k_zero_flags of the same call: behind anybody else's kernel the grid keeps stream order
    if (lss_launch(kern, dim3(grid), dim3(RP_THREADS), smem, s, ready && !clear_rows, d, a) != cudaSuccess) return LSS_ERR_CUDA;
    LSS_CHECK_LAUNCH();
    return LSS_OK;
}

extern "C" int lss_runplan_build(const lss_problem *p, const lss_runplan_layout *L, void *workspace, const float *frustum,
                                 const float *post_trans, const float *M1, const float *M2, const float *trans,
                                 const float *rots, const float *intrins, const float *post_rots, void *stream) {
    return lss_liftsplat_prologue(p, L, workspace, frustum, post_trans, M1, M2, trans, rots, intrins, post_rots, nullptr, nullptr,
                                  nullptr, nullptr, nullptr, stream);
}

static int prologue_validate(const lss_problem *p, const lss_runplan_layout *L, void *workspace, const float *frustum,
                             const float *post_trans, const float *M1, const float *M2, const float *trans, const float *rots,
                             const float *intrins, const float *post_rots, const float *depthnet_out, float *prob, float *ctx_t,
                             float *prob_col, float *bev, bool ready) {
    int st = runplan_supported(p);
    if (st != LSS_OK) return st;
    const bool index = frustum != nullptr;
    const bool raw = M1 == nullptr || M2 == nullptr;
    if (index || ready) {
        LSS_REQUIRE(L && workspace, LSS_ERR_WORKSPACE);
        LSS_REQUIRE(L->n_points == (int64_t)p->B * p->N * p->D * p->fH * p->fW, LSS_ERR_WORKSPACE);
    }
    if (index) {
        LSS_REQUIRE(post_trans && trans, LSS_ERR_BAD_ARG);
        LSS_REQUIRE(!raw || (rots && intrins && post_rots), LSS_ERR_BAD_ARG);
        if (raw) LSS_REQUIRE(raw_build_fits(p), LSS_ERR_UNSUPPORTED);
    }
    if (depthnet_out != nullptr) LSS_REQUIRE(prob && ctx_t && prob_col, LSS_ERR_BAD_ARG);
    if (ready) LSS_REQUIRE(index || depthnet_out != nullptr, LSS_ERR_BAD_ARG);      // somebody has to raise READY
    if ((size_t)(p->D + p->C) * (LIFT_PX + 1) * sizeof(float) > 227 * 1024) return LSS_ERR_UNSUPPORTED;
    if (bev != nullptr) LSS_REQUIRE(lss_aligned(bev, 16), LSS_ERR_ALIGN);
    return LSS_OK;
}

static int prologue_checked(const lss_problem *p, const lss_runplan_layout *L, void *workspace, const float *frustum,
                            const float *post_trans, const float *M1, const float *M2, const float *trans,
                            const float *rots, const float *intrins, const float *post_rots, const float *depthnet_out,
                            float *prob, float *ctx_t, float *prob_col, float *bev, bool ready, void *stream, bool clear_rows = false) {
    int st = prologue_validate(p, L, workspace, frustum, post_trans, M1, M2, trans, rots, intrins, post_rots, depthnet_out, prob, ctx_t,
                               prob_col, bev, ready);
    if (st != LSS_OK) return st;
    const bool index = frustum != nullptr;
    const bool raw = M1 == nullptr || M2 == nullptr;
    const size_t bytes = (size_t)p->B * p->nz * p->C * p->nx * p->ny * 4;
    const CalibPtrs c{frustum, post_trans, M1, M2, trans, rots, intrins, post_rots};
    return launch_prologue(p, L, (L && workspace) ? workspace : nullptr, index, c, index && raw, depthnet_out, prob, ctx_t, prob_col, bev,
                           bytes, ready, (cudaStream_t)stream, clear_rows);
}

extern "C" int lss_liftsplat_prologue(const lss_problem *p, const lss_runplan_layout *L, void *workspace, const float *frustum,
                                      const float *post_trans, const float *M1, const float *M2, const float *trans,
                                      const float *rots, const float *intrins, const float *post_rots, const float *depthnet_out,
                                      float *prob, float *ctx_t, float *prob_col, float *bev, void *stream) {
    return prologue_checked(p, L, workspace, frustum, post_trans, M1, M2, trans, rots, intrins, post_rots, depthnet_out, prob, ctx_t,
                            prob_col, bev, false, stream);
}

static int launch_bev_zero(float *bev, size_t bytes, cudaStream_t s) {
    if (bytes % 16 != 0 || !lss_aligned(bev, 16))
        return cudaMemsetAsync(bev, 0, bytes, s) == cudaSuccess ? LSS_OK : LSS_ERR_CUDA;
    lss_problem dummy = {};
    dummy.B = dummy.N = dummy.D = dummy.fH = dummy.fW = 1; dummy.C = 32; dummy.nx = dummy.ny = dummy.nz = 1;
    dummy.dx[0] = dummy.dx[1] = dummy.dx[2] = 1.f;
    return launch_prologue(&dummy, nullptr, nullptr, false, CalibPtrs{}, false, nullptr, nullptr, nullptr, nullptr, bev, bytes, false, s);
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

static size_t fwd_columns_smem(const lss_problem *p, bool zero) {
    const size_t per = (size_t)p->D * p->fH;
    const size_t col = ((size_t)p->fH * p->C + 3 * per) * 4 + (per + 1) * 2 + 16;
    const size_t shr = (size_t)(GCL_THREADS / 32) * (4 * GCL_SHORT_CAP + (size_t)GCL_ROWS * p->C) * 4;
    const size_t lng = ((size_t)GCL_SORT_CAP + (size_t)GCL_NG * p->C) * 4;
    size_t total = col > shr ? col : shr;
    if (lng > total) total = lng;
    return zero && total < ZERO_CHUNK ? (size_t)ZERO_CHUNK : total;
}

template <int CPL>
static int launch_fwd_kernel(const Dims &d, FwdArgs a, size_t smem, bool pdl, cudaStream_t s) {
    if (smem > 48 * 1024 && cudaFuncSetAttribute(k_fwd_columns<CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return LSS_ERR_CUDA;
    // Where the shared-voxel CTAs sit in the grid.  If the column CTAs fit the GPU in one wave they come first and the shared-voxel
    // CTAs fill in as they leave (cfg 2: 53.2 against 54.0 us per step); if the columns need several waves anyway (big frusta:
    // fewer CTAs per SM), the shared-voxel CTAs -- few, independent of the columns, the longest-running -- go first (cfg 4: 439
    // against 496 us).
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fwd_columns<CPL>, GCL_THREADS, smem) != cudaSuccess) return LSS_ERR_CUDA;
    a.cons_first = a.n_keys > per_sm * rp_num_sms() ? 1 : 0;
    if (lss_launch(k_fwd_columns<CPL>, dim3(a.n_zero + a.n_keys + a.n_cons), dim3(GCL_THREADS), smem, s, pdl, d, a) != cudaSuccess) return LSS_ERR_CUDA;
    LSS_CHECK_LAUNCH();
    return LSS_OK;
}

// zero_target: 0 = pre-cleared; > 0 with in_kernel = zero CTAs inside the column grid; > 0 without = the CTAs of k_zero_flags
static int launch_fwd_grid(const lss_problem *p, const lss_runplan_layout *L, void *workspace, const float *prob_col,
                            const float *ctx_t, float *bev, int zero_target, bool in_kernel, bool wait_ready, cudaStream_t s) {
    const Dims d = make_dims(p);
    char *w = (char *)workspace;
    FwdArgs a = {};
    a.seg_bytes = (size_t)p->nx * p->ny * p->nz * p->C * 4;
    a.n_zero = in_kernel ? zero_target : 0;
    a.zero_target = zero_target;
    { const ZeroTune t = zero_tune(a.seg_bytes, false); a.z_chunk = t.chunk; a.z_window = t.window; }
    a.wait_ready = wait_ready ? 1 : 0;
    a.zero_done = (int32_t *)(w + L->off_zero_done); a.ready = (int32_t *)(w + L->off_ready);
    a.n_keys = p->B * p->N * p->fW; a.fWD = p->fW * p->D;
    a.mfH = ((1ull << 40) + (unsigned)p->fH - 1) / (unsigned)p->fH;
    a.prow = (const int32_t *)(w + L->off_prow); a.sub = (const int2 *)(w + L->off_sub); a.sub2 = (const int2 *)(w + L->off_sub2);
    a.head = (const unsigned long long *)(w + L->off_head);
    a.counters = (int32_t *)(w + L->off_counters); a.pool = (uint32_t *)(w + L->off_pool);
    a.recs = (const int4 *)(w + L->off_recs); a.n_rec_cap = (long long)L->n_rec_cap;
    a.qcount = (const unsigned long long *)(w + L->off_qcount);
    a.longs = (int4 *)(w + L->off_longs);
    a.prob_col = prob_col; a.ctx_t = ctx_t; a.bev = bev;
    a.n_cons = 4 * rp_num_sms();                          // one warp per shared voxel, 4 warps per CTA
    const size_t smem = fwd_columns_smem(p, a.n_zero > 0);
    if (smem > 227 * 1024) return LSS_ERR_UNSUPPORTED;
    // The grid is launched programmatically only behind the prologue of the same call (wait_ready): a grid that starts while
    // SMs are still full gets its first CTAs -- zero CTAs, in the one-launch form -- wherever room appears.
    if (p->C == 32) return launch_fwd_kernel<4>(d, a, smem, wait_ready, s);
    if (p->C == 64) return launch_fwd_kernel<8>(d, a, smem, wait_ready, s);
    return launch_fwd_kernel<16>(d, a, smem, wait_ready, s);
}

static int fwd_args_ok(const lss_problem *p, const lss_runplan_layout *L, const void *workspace, const float *prob_col,
                       const float *ctx_t, const float *bev) {
    int st = runplan_supported(p);
    if (st != LSS_OK) return st;
    LSS_REQUIRE(L && workspace, LSS_ERR_WORKSPACE);
    LSS_REQUIRE(prob_col && ctx_t && bev, LSS_ERR_BAD_ARG);
    LSS_REQUIRE(lss_aligned(ctx_t, 16) && lss_aligned(bev, 16), LSS_ERR_ALIGN);
    LSS_REQUIRE(L->n_points == (int64_t)p->B * p->N * p->D * p->fH * p->fW, LSS_ERR_WORKSPACE);
    return LSS_OK;
}

extern "C" int lss_liftsplat_fwd_cl(const lss_problem *p, const lss_runplan_layout *L, void *workspace,
                                    const float *prob_col, const float *ctx_t, float *bev, int zero_mode, void *stream) {
    int st = fwd_args_ok(p, L, workspace, prob_col, ctx_t, bev);
    if (st != LSS_OK) return st;
    LSS_REQUIRE(zero_mode == LSS_ZERO_ORDERED || zero_mode == LSS_ZERO_PRECLEARED, LSS_ERR_BAD_ARG);
    int n_zero = 0;
    if (zero_mode == LSS_ZERO_ORDERED) n_zero = zero_tune((size_t)p->nx * p->ny * p->nz * p->C * 4, false).n_cta;
    return launch_fwd_grid(p, L, workspace, prob_col, ctx_t, bev, n_zero, true, false, (cudaStream_t)stream);
}

extern "C" int lss_liftsplat_forward(const lss_problem *p, const lss_runplan_layout *L, void *workspace, const float *frustum,
                                     const float *post_trans, const float *M1, const float *M2, const float *trans,
                                     const float *rots, const float *intrins, const float *post_rots, const float *depthnet_out,
                                     float *prob, float *ctx_t, float *prob_col, float *bev, void *stream) {
    int st = fwd_args_ok(p, L, workspace, prob_col, ctx_t, bev);
    if (st != LSS_OK) return st;
    LSS_REQUIRE(depthnet_out != nullptr, LSS_ERR_BAD_ARG);
    st = prologue_validate(p, L, workspace, frustum, post_trans, M1, M2, trans, rots, intrins, post_rots, depthnet_out, prob, ctx_t,
                           prob_col, nullptr, true);      // before anything is launched: a zero-fill without its consumer
    if (st != LSS_OK) return st;                          // would leave the progress counters raised
    cudaStream_t s = (cudaStream_t)stream;
    const size_t seg_bytes = (size_t)p->nx * p->ny * p->nz * p->C * 4;
    const ZeroTune t = zero_tune(seg_bytes, true);
    k_zero_flags<<<t.n_cta, 32, ZERO_CHUNK, s>>>(bev, seg_bytes, p->B, (int32_t *)((char *)workspace + L->off_zero_done), t.chunk, t.window);
    LSS_CHECK_LAUNCH();
    st = prologue_checked(p, L, workspace, frustum, post_trans, M1, M2, trans, rots, intrins, post_rots, depthnet_out, prob, ctx_t,
                          prob_col, nullptr, true, stream);
    if (st != LSS_OK) return st;
    return launch_fwd_grid(p, L, workspace, prob_col, ctx_t, bev, t.n_cta, false, true, s);
}

// The same forward into a PERSISTENT output tensor: `bev` still holds what the previous lss_liftsplat_forward* call on this workspace
// wrote (or is all zero, with a workspace that has not been built since lss_runplan_reset).  Nothing but the rows that call wrote
// is non-zero, and the workspace still names them (the leaders of its sub-runs): the index role zeroes exactly those -- 10 MB
// instead of the tensor's 82 MB at cfg 2 -- before it overwrites the plan; without a rebuild (frustum == NULL) the rows about to
// be written are the rows that were written, and nothing is cleared at all.  No zero-fill grid, no waiting for zeros.
extern "C" int lss_liftsplat_forward_persistent(const lss_problem *p, const lss_runplan_layout *L, void *workspace, const float *frustum,
                                                const float *post_trans, const float *M1, const float *M2, const float *trans,
                                                const float *rots, const float *intrins, const float *post_rots,
                                                const float *depthnet_out, float *prob, float *ctx_t, float *prob_col, float *bev,
                                                void *stream) {
    int st = fwd_args_ok(p, L, workspace, prob_col, ctx_t, bev);
    if (st != LSS_OK) return st;
    LSS_REQUIRE(depthnet_out != nullptr, LSS_ERR_BAD_ARG);
    st = prologue_checked(p, L, workspace, frustum, post_trans, M1, M2, trans, rots, intrins, post_rots, depthnet_out, prob, ctx_t,
                          prob_col, bev, true, stream, true);
    if (st != LSS_OK) return st;
    return launch_fwd_grid(p, L, workspace, prob_col, ctx_t, bev, 0, false, true, (cudaStream_t)stream);
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
