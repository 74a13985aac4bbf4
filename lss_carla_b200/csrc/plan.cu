// Geometry, voxel indices and the per-batch plan (tile buckets sorted in the reference's rank order).
//
// Replaces, for the lift-splat path of shdragron/LSS-Carla:
//   LiftSplatShoot.get_geometry          src/models.py:170-190
//   the index half of voxel_pooling      src/models.py:212-231  (quantise, mask, rank, argsort)
//
// All per-point arithmetic uses the round-to-nearest intrinsics (__fmul_rn/__fadd_rn/__fsub_rn/
// __fdiv_rn), which nvcc never contracts into FMAs, so results are bit-identical to the reference's
// CPU evaluation (oracle/lss_oracle.py documents the measured association of the 3x3 products).
#include "common.cuh"
#include "geom.cuh"

// ------------------------------------------------------------------------------------------------
// kernels: calibration matrices, geometry, voxel index (+ tile histogram)
// ------------------------------------------------------------------------------------------------

__global__ void k_calib_matrices(int n_cams, const float *__restrict__ rots, const float *__restrict__ intrins,
                                 const float *__restrict__ post_rots, float *__restrict__ M1, float *__restrict__ M2) {
    const int cam = blockIdx.x * blockDim.x + threadIdx.x;
    if (cam >= n_cams) return;
    calib_matrices_of(rots, intrins, post_rots, cam, M1 + cam * 9, M2 + cam * 9);
}

__global__ void k_geometry(Dims d, CalibPtrs c, float *__restrict__ geom) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= d.n_points) return;
    const int cam = p / d.DHW, in_cam = p - cam * d.DHW;
    float g[3];
    ego_point(c, cam, in_cam, g);
    geom[(size_t)p * 3 + 0] = g[0];
    geom[(size_t)p * 3 + 1] = g[1];
    geom[(size_t)p * 3 + 2] = g[2];
}

// PPT points per thread (each round of 256 consecutive points is warp-aggregated on its own).  COUNT adds the
// per-tile histogram of kept points and, in the last CTA to finish, the exclusive scan tile_count -> tile_start
// (and leaves tile_count / cursor zeroed).  The plan build uses PPT = 2: the grid then fits the GPU in ONE wave (the
// kernel is a chain of latencies -- calibration loads, histogram atomics, fence, ticket -- so a second, nearly empty
// wave doubled its run time).
template <bool FROM_GEOM, bool COUNT, bool RAW = false, int PPT = 1, bool TAIL = true>
__global__ void __launch_bounds__(256)
k_voxel_index(Dims d, Tiling tl, const float *__restrict__ geom, CalibPtrs c, int32_t *__restrict__ vox,
              long long *__restrict__ idx, uint8_t *__restrict__ kept, long long *__restrict__ rank,
              int32_t *__restrict__ tile_count, int32_t *__restrict__ tile_start, int32_t *__restrict__ cursor,
              int32_t *__restrict__ sync, int32_t *__restrict__ counters, int32_t *__restrict__ key_count,
              int32_t *__restrict__ prow) {
    if (COUNT) {
        lss_pdl_trigger();                         // the scatter kernel may be scheduled while this grid drains
        if (!RAW && !FROM_GEOM) lss_pdl_wait();    // M1 / M2 may come from k_calib_matrices right before this launch
    }
    __shared__ float s_m[RAW ? LSS_RAW_CAMS : 1][18];
    const int cam0 = (int)((blockIdx.x * (PPT * 256u)) / (unsigned)d.DHW);
    if (RAW) {      // the calibration matrices of the few cameras this CTA touches, made on the fly (no extra launch)
        const int cam = cam0 + (int)threadIdx.x;
        if (threadIdx.x < LSS_RAW_CAMS && cam < d.B * d.N) calib_matrices_of(c.rots, c.intrins, c.post_rots, cam, s_m[threadIdx.x], s_m[threadIdx.x] + 9);
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int u = 0; u < PPT; ++u) {
        const int p = (blockIdx.x * PPT + u) * 256 + threadIdx.x;
        int v = -1;
        if (p < d.n_points) {
            const int b = p / d.P;
            float g[3];
            if (FROM_GEOM) {
                g[0] = __ldg(geom + (size_t)p * 3 + 0);
                g[1] = __ldg(geom + (size_t)p * 3 + 1);
                g[2] = __ldg(geom + (size_t)p * 3 + 2);
            } else {
                const int cam = p / d.DHW;
                if (RAW) ego_point(c, cam, p - cam * d.DHW, g, s_m[cam - cam0], s_m[cam - cam0] + 9);
                else ego_point(c, cam, p - cam * d.DHW, g);
            }
            long long ii[3];
            v = voxel_of_point(d, b, g, ii);
            if (vox) vox[p] = v;
            if (COUNT && v < 0)                      // kept points get their compact row from k_plan_sort
                prow[(size_t)b * d.P + lss_column_major(d, (unsigned)(p - b * d.P))] = -1;
            if (idx) { idx[(size_t)p * 3 + 0] = ii[0]; idx[(size_t)p * 3 + 1] = ii[1]; idx[(size_t)p * 3 + 2] = ii[2]; }
            if (kept) kept[p] = v >= 0;
            if (rank)   // models.py:226-229, int64
                rank[p] = v >= 0 ? ii[0] * ((long long)d.ny * d.nz * d.B) + ii[1] * ((long long)d.nz * d.B) + ii[2] * d.B + b : -1;
        }
        if (COUNT) {    // ---- per-tile histogram, warp-aggregated
            int tile = -1 - lane;   // unique negative key for dropped points: singleton groups
            if (v >= 0) {
                const int iy = v % d.ny;
                tile = (v / d.ny) * tl.nty + iy / tl.TY;
            }
            const unsigned peers = __match_any_sync(LSS_FULL_MASK, tile);
            if (v >= 0 && lane == __ffs(peers) - 1) atomicAdd(tile_count + tile, __popc(peers));
        }
    }
    if (!COUNT || !TAIL) return;                   // !TAIL: k_plan_scatter<PPT, true> scans the histogram itself

    // ---- last CTA scans the histogram
    __shared__ int s_last;
    __shared__ int s_warp[8];
    __shared__ int s_carry;
    __syncthreads();                               // the CTA's histogram atomics are ordered before thread 0's fence
    if (threadIdx.x == 0) {
        __threadfence();                           // (cumulative): release them device-wide, then take a ticket
        s_last = (atomicAdd(sync, 1) == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    for (int base = 0; base < tl.n_tiles; base += 1024) {
        const int i0 = base + threadIdx.x * 4;
        int a[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            a[k] = (i0 + k < tl.n_tiles) ? __ldcg(tile_count + i0 + k) : 0;
            if (i0 + k < tl.n_tiles) { tile_count[i0 + k] = 0; cursor[i0 + k] = 0; }
        }
        const int tsum = a[0] + a[1] + a[2] + a[3];
        int inc = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(LSS_FULL_MASK, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < warp; ++w) woff += s_warp[w];
        int run = s_carry + woff + inc - tsum;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (i0 + k < tl.n_tiles) tile_start[i0 + k] = run;
            run += a[k];
        }
        __syncthreads();
        if (threadIdx.x == 255) s_carry = run;
        __syncthreads();
    }
    for (int i = threadIdx.x; i < d.B * d.N * d.fW; i += blockDim.x) key_count[i] = 0;
    if (threadIdx.x == 0) { tile_start[tl.n_tiles] = s_carry; *sync = 0; counters[0] = 0; counters[1] = 0; counters[2] = 0; }
}

// Scatter kept points into their tile buckets: entries[tile_start[t] + k] = col << 20 | point-in-sample.
// PPT points per thread (256 consecutive points per warp-aggregation round): with PPT = 2 the grid is a single wave and
// both cursor atomics of a thread are in flight together (the kernel waits for atomic round trips, nothing else).
// SCAN: the exclusive scan tile_count -> tile_start is done HERE, redundantly by every CTA into shared memory (a few
// thousand counters from L2), instead of by the last CTA of k_voxel_index behind a fence and a ticket: that kernel then
// ends with its histogram atomics.  CTA 0 publishes tile_start and clears the per-build counters; tile_count / cursor
// are cleared by k_plan_sort (sorted plans) or a memset (unsorted plans).
template <int PPT, bool SCAN>
__global__ void __launch_bounds__(256)
k_plan_scatter(Dims d, Tiling tl, const int32_t *__restrict__ vox, int32_t *__restrict__ tile_start,
               int32_t *__restrict__ cursor, uint32_t *__restrict__ entries, const int32_t *__restrict__ tile_count,
               int32_t *__restrict__ key_count, int32_t *__restrict__ counters) {
    extern __shared__ int s_start[];               // SCAN: [n_tiles + 1]
    lss_pdl_trigger();
    lss_pdl_wait();                                // voxel ids and the histogram (or tile_start) come from k_voxel_index
    const int lane = threadIdx.x & 31;
    int v[PPT], tile[PPT], col[PPT], base[PPT], ts[PPT];
    unsigned peers[PPT];
#pragma unroll
    for (int u = 0; u < PPT; ++u) {                // requested before the scan: in flight while it runs
        const int p = (blockIdx.x * PPT + u) * 256 + threadIdx.x;
        v[u] = p < d.n_points ? __ldg(vox + p) : -1;
    }
    if (SCAN) {
        __shared__ int s_warp[8];
        __shared__ int s_carry;
        const int warp = threadIdx.x >> 5, nt = tl.n_tiles;
        if (threadIdx.x == 0) s_carry = 0;
        __syncthreads();
        for (int base0 = 0; base0 < nt; base0 += 2048) {  // 8 counters per thread and step
            const int i0 = base0 + threadIdx.x * 8;      // (tile_count is padded to 256 bytes, the padding is zero)
            int a[8];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int4 q = i0 + 4 * h < nt ? __ldcg(reinterpret_cast<const int4 *>(tile_count + i0 + 4 * h)) : make_int4(0, 0, 0, 0);
                a[4 * h] = q.x; a[4 * h + 1] = q.y; a[4 * h + 2] = q.z; a[4 * h + 3] = q.w;
            }
            int tsum = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) tsum += a[k];
            int inc = tsum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(LSS_FULL_MASK, inc, o);
                if (lane >= o) inc += t;
            }
            if (lane == 31) s_warp[warp] = inc;
            __syncthreads();
            int run = s_carry + inc - tsum;
            for (int w = 0; w < warp; ++w) run += s_warp[w];
#pragma unroll
            for (int k = 0; k < 8; ++k) {                // s_start has n_tiles + 1 slots: slot nt receives the total
                if (i0 + k <= nt) s_start[i0 + k] = run;
                run += a[k];
            }
            if (i0 + 8 == nt) s_start[nt] = run;
            __syncthreads();
            if (threadIdx.x == 255) s_carry = run;
            __syncthreads();
        }
        if (blockIdx.x == 0) {
            for (int i = threadIdx.x; i <= nt; i += 256) tile_start[i] = s_start[i];
            for (int i = threadIdx.x; i < d.B * d.N * d.fW; i += 256) key_count[i] = 0;
            if (threadIdx.x < 3) counters[threadIdx.x] = 0;
        }
    }
#pragma unroll
    for (int u = 0; u < PPT; ++u) {
        tile[u] = -1 - lane; col[u] = 0; ts[u] = 0;
        if (v[u] >= 0) {
            const int iy = v[u] % d.ny;
            const int ty = iy / tl.TY;
            tile[u] = (v[u] / d.ny) * tl.nty + ty;
            col[u] = iy - ty * tl.TY;
        }
        peers[u] = __match_any_sync(LSS_FULL_MASK, tile[u]);
        base[u] = 0;
        if (v[u] >= 0) {
            if (lane == __ffs(peers[u]) - 1) base[u] = atomicAdd(cursor + tile[u], __popc(peers[u]));
            ts[u] = SCAN ? s_start[tile[u]] : __ldg(tile_start + tile[u]);
        }
    }
#pragma unroll
    for (int u = 0; u < PPT; ++u) {
        const int p = (blockIdx.x * PPT + u) * 256 + threadIdx.x;
        const int b0 = __shfl_sync(LSS_FULL_MASK, base[u], __ffs(peers[u]) - 1);
        if (v[u] >= 0) {
            const int slot = ts[u] + b0 + __popc(peers[u] & ((1u << lane) - 1u));
            entries[slot] = ((uint32_t)col[u] << LSS_PIDX_BITS) | (uint32_t)(p % d.P);
        }
    }
}

#define LSS_SORT_SMEM_CAP 4096   // entries grouped in shared memory (16 KB); larger buckets sort in global memory
#define LSS_SORT_THREADS 256          // CTA size of k_plan_sort (128-thread CTAs measured no faster: the phases are latency chains)

// Exclusive scan of a[0..L) in shared memory by the whole CTA (L <= LSS_MAX_TILE_COLS); a[L] = total.
template <int NT>
__device__ __forceinline__ void block_scan_excl(int *a, int L, int *s_warp) {
    const int per = (L + NT - 1) / NT;
    const int lo = threadIdx.x * per, hi = min(L, lo + per);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int sum = 0;
    for (int i = lo; i < hi; ++i) sum += a[i];
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(LSS_FULL_MASK, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    int run = inc - sum;
    for (int w = 0; w < warp; ++w) run += s_warp[w];
    for (int i = lo; i < hi; ++i) { const int x = a[i]; a[i] = run; run += x; }
    if (threadIdx.x == NT - 1) a[L] = run;
    __syncthreads();
}

// Ordered enumeration by the whole CTA: calls emit(k, i) for every i in [0, L) with flag(i), k ascending
// in i starting at 0.  COUNT_ONLY skips the calls.  Returns the number of flagged items (uniform).
template <int NT, bool COUNT_ONLY, typename Flag, typename Emit>
__device__ __forceinline__ int block_enumerate(int L, int *s_warp, Flag flag, Emit emit) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int base = 0;
    for (int i0 = 0; i0 < L; i0 += NT) {
        const int i = i0 + threadIdx.x;
        const bool f = i < L && flag(i);
        const unsigned hb = __ballot_sync(LSS_FULL_MASK, f);
        if (lane == 0) s_warp[warp] = __popc(hb);
        __syncthreads();
        int off = base, total = 0;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) { const int c = s_warp[w]; if (w < warp) off += c; total += c; }
        if (!COUNT_ONLY && f) emit(off + __popc(hb & ((1u << lane) - 1u)), i);
        base += total;
        __syncthreads();
    }
    return base;
}

// One CTA per tile.  Counting sort by column (histogram, scan, unordered placement into shared memory),
// then every entry finds its rank inside its column segment by counting the smaller keys of that
// segment (segments are short: the points of ONE voxel) and is written to its final slot.  Keys are
// unique (col << 20 | point index), so the ranks of a segment are a permutation.
// Also emits the tables of the tile's ns non-empty voxels: the tile reserves ns consecutive compact rows
// [row0, row0+ns) with one atomic (the order of tiles in row space is irrelevant), and for its k-th
// non-empty voxel writes segs[s+k] = col << 20 | first entry within the bucket and appends the record
// {first entry (global), length, batch index, row0+k} to the bucket of the camera column (b, n, w) of the
// voxel's first point -- the forward gather walks these buckets, so that the few context rows of one
// camera column stay in L1 while all their voxels are summed.
// `kind`: 0 = every point of the voxel is an image row of the first point's (camera column, depth bin);
// 1 = all points lie in that camera column (several depths); 2 = points of other columns / cameras.
// Kinds 0 and 1 go to the bucket of the column.  If a kind-0 voxel holds exactly the fH image rows, in order,
// the record is PURE and its length field carries -(depth bin + 1): the gather needs neither its entries
// nor any decoding.  Kind 2 goes to the batch-wide queue of mixed voxels (counters[1]).
__device__ __forceinline__ void emit_voxel_record(const Dims &d, int b, uint32_t first_entry, int kind, int e0,
                                                  int len, int row, int4 *__restrict__ seg_recs,
                                                  int32_t *__restrict__ key_count, int4 *__restrict__ mixed_recs,
                                                  int32_t *__restrict__ counters, long long n_rows_cap) {
    if (len >= LSS_LONG_VOXEL) {                   // long voxels: their own queue, filled from the END of mixed_recs
        mixed_recs[n_rows_cap - 1 - atomicAdd(counters + 2, 1)] = make_int4(e0, len, b, row);
        return;
    }
    if (kind == 2) {
        mixed_recs[atomicAdd(counters + 1, 1)] = make_int4(e0, len, b, row);
        return;
    }
    const unsigned pidx = first_entry & LSS_PIDX_MASK;
    const unsigned cam = lss_div20(pidx, d.mDHW);
    const unsigned r = pidx - cam * d.DHW;
    const unsigned dd = lss_div20(r, d.mHW);
    const unsigned hw = r - dd * d.HW;
    const unsigned h = lss_div20(hw, d.mfW), w = hw - h * d.fW;
    const int key = (b * d.N + (int)cam) * d.fW + (int)w;
    const int slot = atomicAdd(key_count + key, 1);
    const bool pure = kind == 0 && len == d.fH && h == 0;
    seg_recs[(size_t)key * (d.D * d.fH) + slot] = make_int4(e0, pure ? -(int)(dd + 1) : len, b, row);
}

template <int NT>
__global__ void __launch_bounds__(NT)
k_plan_sort(Dims d, Tiling tl, const int32_t *__restrict__ tile_start, uint32_t *__restrict__ entries,
            uint32_t *__restrict__ segs, int32_t *__restrict__ tile_nseg, int32_t *__restrict__ tile_row0,
            int4 *__restrict__ seg_recs, int32_t *__restrict__ key_count, int4 *__restrict__ mixed_recs,
            int32_t *__restrict__ counters, int32_t *__restrict__ prow, long long n_rows_cap,
            int32_t *__restrict__ clear_count, int32_t *__restrict__ clear_cursor) {
    extern __shared__ int s_int[];                 // start[TY+1], cursor[TY], mixed[TY]
    __shared__ uint32_t s_grp[LSS_SORT_SMEM_CAP];
    __shared__ int s_warp[NT / 32];
    lss_pdl_trigger();                             // the forward gather may be scheduled while this grid drains (it waits at its top)
    lss_pdl_wait();                                // the buckets come from k_plan_scatter
    const int t = blockIdx.x;
    if (clear_count != nullptr && threadIdx.x == 0) { clear_count[t] = 0; clear_cursor[t] = 0; }   // scratch of the next build
    const int s = tile_start[t], n = tile_start[t + 1] - s;
    const int b = t / (tl.nty * d.nx * d.nz);
    uint32_t *g = entries + s;
    if (n == 0) { if (threadIdx.x == 0) { tile_nseg[t] = 0; tile_row0[t] = 0; } return; }
    // Compact rows: the tile's k-th non-empty voxel owns row s + k (s = the tile's first bucket slot; a voxel has at
    // least one entry, so the rows of different tiles never collide and stay below the number of kept points).  No
    // reservation, hence no atomic round trip in the middle of the CTA; the row space is sparse but tile-contiguous.
    const int row0 = s;
    auto publish = [&](int ns) { if (threadIdx.x == 0) { tile_nseg[t] = ns; tile_row0[t] = s; atomicAdd(counters, ns); } };
    if (n > LSS_SORT_SMEM_CAP) {                   // rare: correctness fallback, sorts in global memory (L2)
        bitonic_sort_block((volatile uint32_t *)g, n);
        __syncthreads();
        auto head = [&](int i) { return i == 0 || (g[i] >> LSS_PIDX_BITS) != (g[i - 1] >> LSS_PIDX_BITS); };
        const int ns = block_enumerate<NT, false>(n, s_warp, head, [&](int k, int i) {
            const uint32_t col = g[i] >> LSS_PIDX_BITS;
            int j = i + 1;                          // segment end: next head (long runs only occur here)
            while (j < n && (g[j] >> LSS_PIDX_BITS) == col) ++j;
            segs[s + k] = (col << LSS_PIDX_BITS) | (uint32_t)i;
            emit_voxel_record(d, b, g[i], 2, s + i, j - i, row0 + k, seg_recs, key_count, mixed_recs, counters, n_rows_cap);
            for (int q = i; q < j; ++q) prow[(size_t)b * d.P + lss_column_major(d, g[q] & LSS_PIDX_MASK)] = row0 + k;
        });
        publish(ns);
        return;
    }
    const int TY = tl.TY;
    int *start = s_int, *cursor = s_int + TY + 1, *mixed = cursor + TY;
    for (int i = threadIdx.x; i <= TY; i += NT) start[i] = 0;
    for (int i = threadIdx.x; i < TY; i += NT) { cursor[i] = 0; mixed[i] = 0; }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += NT) atomicAdd(start + (g[i] >> LSS_PIDX_BITS), 1);
    __syncthreads();
    block_scan_excl<NT>(start, TY, s_warp);
    for (int i = threadIdx.x; i < n; i += NT) {
        const uint32_t e = g[i];
        const int col = (int)(e >> LSS_PIDX_BITS);
        s_grp[start[col] + atomicAdd(cursor + col, 1)] = e;
    }
    __syncthreads();
    {   // rank inside the voxel, and in the same walk the voxel's first point (its smallest key): every entry is
        // classified against it -- same (camera column, depth) / same column / foreign points
        const unsigned span = (unsigned)d.fH * d.fW;
        for (int i = threadIdx.x; i < n; i += NT) {
            const uint32_t e = s_grp[i];
            const int col = (int)(e >> LSS_PIDX_BITS);
            const int a = start[col], bb = start[col + 1];
            int rank = 0;
            uint32_t first = e;
            for (int j = a; j < bb; ++j) { const uint32_t x = s_grp[j]; rank += x < e; first = min(first, x); }
            g[a + rank] = e;
            if (rank == 0) { cursor[col] = (int)e; continue; }       // the first point itself is kind 0 (cursor is free again)
            const unsigned p = e & LSS_PIDX_MASK, p0 = first & LSS_PIDX_MASK;
            const unsigned delta = p - p0;
            const bool col_aligned = delta - lss_div20(delta, d.mfW) * (unsigned)d.fW == 0u;   // delta < 2^20
            if (delta < span && col_aligned) continue;                                       // kind 0
            const bool same_col = col_aligned && lss_div20(p, d.mDHW) == lss_div20(p0, d.mDHW);
            atomicMax(mixed + col, same_col ? 1 : 2);
        }
    }
    __syncthreads();
    auto hit = [&](int c) { return start[c + 1] > start[c]; };
    const int ns = block_enumerate<NT, false>(TY, s_warp, hit, [&](int k, int c) {
        const int kind = mixed[c] & 3;
        mixed[c] = kind | (k << 2);               // the voxel's ordinal in the tile, for the per-point rows below
        segs[s + k] = ((uint32_t)c << LSS_PIDX_BITS) | (uint32_t)start[c];
        emit_voxel_record(d, b, (uint32_t)cursor[c], kind, s + start[c], start[c + 1] - start[c], row0 + k, seg_recs, key_count,
                          mixed_recs, counters, n_rows_cap);
    });
    publish(ns);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += NT) {
        const uint32_t e = s_grp[i];
        prow[(size_t)b * d.P + lss_column_major(d, e & LSS_PIDX_MASK)] = row0 + (mixed[e >> LSS_PIDX_BITS] >> 2);
    }
}

// ------------------------------------------------------------------------------------------------
// parity dump: the reference's sort permutation (models.py:226-231)
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ int rank_of_voxel(const Dims &d, int v) {
    const int iy = v % d.ny; int t = v / d.ny;
    const int ix = t % d.nx; t /= d.nx;
    const int iz = t % d.nz; const int b = t / d.nz;
    return ((ix * d.ny + iy) * d.nz + iz) * d.B + b;
}

__device__ __forceinline__ int voxel_of_entry(const Dims &d, const Tiling &tl, int tile, uint32_t e) {
    const int ty = tile % tl.nty;
    return (tile / tl.nty) * d.ny + ty * tl.TY + (int)(e >> LSS_PIDX_BITS);
}

__global__ void k_fill_i32(int32_t *a, int64_t n, int32_t val) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) a[i] = val;
}
__global__ void k_fill_i64(long long *a, int64_t n, long long val) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) a[i] = val;
}

// one CTA per tile: count points per reference rank
__global__ void k_ref_count(Dims d, Tiling tl, const int32_t *__restrict__ tile_start,
                            const uint32_t *__restrict__ entries, int32_t *__restrict__ rank_count) {
    const int t = blockIdx.x;
    const int s = tile_start[t], e = tile_start[t + 1];
    for (int i = s + threadIdx.x; i < e; i += blockDim.x)
        atomicAdd(rank_count + rank_of_voxel(d, voxel_of_entry(d, tl, t, entries[i])), 1);
}

__global__ void k_ref_place(Dims d, Tiling tl, const int32_t *__restrict__ tile_start,
                            const uint32_t *__restrict__ entries, const int32_t *__restrict__ rank_start,
                            long long *__restrict__ order) {
    const int t = blockIdx.x;
    const int s = tile_start[t], e = tile_start[t + 1];
    const int b = (t / tl.nty) / (d.nz * d.nx);
    for (int i = s + threadIdx.x; i < e; i += blockDim.x) {
        const uint32_t en = entries[i];
        int first = i;                                   // walk back to the first entry of this voxel
        while (first > s && (entries[first - 1] >> LSS_PIDX_BITS) == (en >> LSS_PIDX_BITS)) --first;
        const int r = rank_of_voxel(d, voxel_of_entry(d, tl, t, en));
        order[rank_start[r] + (i - first)] = (long long)b * d.P + (long long)(en & LSS_PIDX_MASK);
    }
}

// ------------------------------------------------------------------------------------------------
// host entry points
// ------------------------------------------------------------------------------------------------

static inline Tiling make_tiling(const lss_plan_layout *L) {
    Tiling t; t.TY = L->tile_cols; t.nty = L->tiles_per_row; t.n_tiles = L->n_tiles; return t;
}

extern "C" int lss_plan_layout_init(const lss_problem *p, int tile_cols, lss_plan_layout *out) {
    int st = lss_check_problem(p);
    if (st != LSS_OK) return st;
    LSS_REQUIRE(out != nullptr, LSS_ERR_BAD_ARG);
    if (tile_cols <= 0) tile_cols = p->ny <= 256 ? ((p->ny + 7) / 8) * 8 : 256;
    LSS_REQUIRE(tile_cols % 8 == 0 && tile_cols <= LSS_MAX_TILE_COLS, LSS_ERR_UNSUPPORTED);
    const Dims d = make_dims(p);
    out->tile_cols = tile_cols;
    out->tiles_per_row = (p->ny + tile_cols - 1) / tile_cols;
    const int64_t nt = (int64_t)p->B * p->nz * p->nx * out->tiles_per_row;
    LSS_REQUIRE(nt < ((int64_t)1 << 30), LSS_ERR_UNSUPPORTED);
    out->n_tiles = (int32_t)nt;
    out->n_points = d.n_points;
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    size_t off = 0;
    out->off_vox = off;        off += up((size_t)d.n_points * 4);
    out->off_entries = off;    off += up((size_t)d.n_points * 4);
    out->off_tile_start = off; off += up(((size_t)nt + 1) * 4);
    out->off_segs = off;       off += up((size_t)d.n_points * 4);
    out->off_tile_nseg = off;  off += up((size_t)nt * 4);
    out->off_tile_row0 = off;  off += up((size_t)nt * 4);
    const int64_t nvox = (int64_t)p->B * p->nx * p->ny * p->nz;
    (void)nvox;
    out->n_rows_cap = (int64_t)d.n_points;         // compact rows are indexed by bucket slot (k_plan_sort): below the number of kept points
    out->off_seg_recs = off;   off += up((size_t)d.n_points * 16);
    out->off_key_count = off;  off += up((size_t)p->B * p->N * p->fW * 4);
    out->off_mixed_recs = off; off += up((size_t)out->n_rows_cap * 16);
    out->off_prow = off;       off += up((size_t)d.n_points * 4);
    out->off_counters = off;   off += up(64 * 4);
    out->off_tile_count = off; off += up((size_t)nt * 4);
    out->off_cursor = off;     off += up((size_t)nt * 4);
    out->off_sync = off;       off += up(64 * 4);
    out->bytes = off;
    return LSS_OK;
}

extern "C" int lss_plan_reset(const lss_plan_layout *L, void *ws, void *stream) {
    LSS_REQUIRE(L && ws, LSS_ERR_WORKSPACE);
    cudaStream_t s = (cudaStream_t)stream;
    char *w = (char *)ws;
    // tile_count, cursor and sync are contiguous
    if (cudaMemsetAsync(w + L->off_tile_count, 0, L->bytes - L->off_tile_count, s) != cudaSuccess) return LSS_ERR_CUDA;
    return LSS_OK;
}

extern "C" int lss_calib_matrices(int32_t n_cams, const float *rots, const float *intrins, const float *post_rots,
                                  float *M1, float *M2, void *stream) {
    LSS_REQUIRE(n_cams > 0 && rots && intrins && post_rots && M1 && M2, LSS_ERR_BAD_ARG);
    k_calib_matrices<<<(n_cams + 63) / 64, 64, 0, (cudaStream_t)stream>>>(n_cams, rots, intrins, post_rots, M1, M2);
    LSS_CHECK_LAUNCH();
    return LSS_OK;
}

extern "C" int lss_geometry(const lss_problem *p, const float *frustum, const float *post_trans, const float *M1,
                            const float *M2, const float *trans, float *geom_out, void *stream) {
    int st = lss_check_problem(p);
    if (st != LSS_OK) return st;
    LSS_REQUIRE(frustum && post_trans && M1 && M2 && trans && geom_out, LSS_ERR_BAD_ARG);
    const Dims d = make_dims(p);
    CalibPtrs c{frustum, post_trans, M1, M2, trans, nullptr, nullptr, nullptr};
    k_geometry<<<(d.n_points + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d, c, geom_out);
    LSS_CHECK_LAUNCH();
    return LSS_OK;
}

extern "C" int lss_voxel_index(const lss_problem *p, const float *geom, const float *frustum, const float *post_trans,
                               const float *M1, const float *M2, const float *trans, int32_t *vox, int64_t *idx,
                               uint8_t *kept, int64_t *rank, void *stream) {
    int st = lss_check_problem(p);
    if (st != LSS_OK) return st;
    const bool from_geom = geom != nullptr;
    LSS_REQUIRE(from_geom || (frustum && post_trans && M1 && M2 && trans), LSS_ERR_BAD_ARG);
    const Dims d = make_dims(p);
    CalibPtrs c{frustum, post_trans, M1, M2, trans, nullptr, nullptr, nullptr};
    Tiling tl{8, 1, 1};
    const int grid = (d.n_points + 255) / 256;
    cudaStream_t s = (cudaStream_t)stream;
    if (from_geom)
        k_voxel_index<true, false><<<grid, 256, 0, s>>>(d, tl, geom, c, vox, (long long *)idx, kept, (long long *)rank,
                                                        nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    else
        k_voxel_index<false, false><<<grid, 256, 0, s>>>(d, tl, geom, c, vox, (long long *)idx, kept, (long long *)rank,
                                                         nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    LSS_CHECK_LAUNCH();
    return LSS_OK;
}

static int plan_build_impl(const lss_problem *p, const lss_plan_layout *L, void *workspace, const float *geom,
                           const CalibPtrs &c, bool raw, int sorted, cudaStream_t s) {
    const Dims d = make_dims(p);
    LSS_REQUIRE(L->n_points == d.n_points, LSS_ERR_WORKSPACE);
    const Tiling tl = make_tiling(L);
    char *w = (char *)workspace;
    int32_t *vox = (int32_t *)(w + L->off_vox);
    uint32_t *entries = (uint32_t *)(w + L->off_entries);
    int32_t *tile_start = (int32_t *)(w + L->off_tile_start);
    int32_t *tile_count = (int32_t *)(w + L->off_tile_count);
    int32_t *cursor = (int32_t *)(w + L->off_cursor);
    int32_t *sync = (int32_t *)(w + L->off_sync);
    int32_t *counters = (int32_t *)(w + L->off_counters);
    int32_t *key_count = (int32_t *)(w + L->off_key_count);
    int32_t *prow = (int32_t *)(w + L->off_prow);
    const int grid = (d.n_points + 255) / 256;
    // the histogram scan rides in the scatter kernel when its table fits shared memory
    const size_t scan_smem = ((size_t)tl.n_tiles + 1) * sizeof(int);
    // ... and is one 2048-counter step: every scatter CTA repeats the scan, which stops paying with many tiles AND many
    // CTAs (cfg 4, 6400 tiles x 3895 CTAs: plan 138.7 us with the scan in the scatter kernel, 122.6 us with the ticket path)
    const bool scan_in_scatter = tl.n_tiles <= 2048 && scan_smem <= 64 * 1024;
    const int grid_vi = (grid + 1) / 2;
#define VI_ARGS d, tl, geom, c, vox, nullptr, nullptr, nullptr, tile_count, tile_start, cursor, sync, counters, key_count, prow
    if (geom != nullptr) {
        if (scan_in_scatter) k_voxel_index<true, true, false, 2, false><<<grid_vi, 256, 0, s>>>(VI_ARGS);
        else k_voxel_index<true, true, false, 2, true><<<grid_vi, 256, 0, s>>>(VI_ARGS);
    } else if (raw) {
        if (scan_in_scatter) k_voxel_index<false, true, true, 2, false><<<grid_vi, 256, 0, s>>>(VI_ARGS);
        else k_voxel_index<false, true, true, 2, true><<<grid_vi, 256, 0, s>>>(VI_ARGS);
    } else if ((scan_in_scatter ? lss_launch(k_voxel_index<false, true, false, 2, false>, dim3(grid_vi), dim3(256), 0, s, true, VI_ARGS)
                                : lss_launch(k_voxel_index<false, true, false, 2, true>, dim3(grid_vi), dim3(256), 0, s, true, VI_ARGS)) != cudaSuccess) return LSS_ERR_CUDA;
#undef VI_ARGS
    LSS_CHECK_LAUNCH();
    if (scan_in_scatter) {
        if (scan_smem > 40 * 1024 &&       // (per device, cheap: no per-process cache)
            cudaFuncSetAttribute(k_plan_scatter<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024) != cudaSuccess) return LSS_ERR_CUDA;
        if (lss_launch(k_plan_scatter<2, true>, dim3(grid_vi), dim3(256), scan_smem, s, true, d, tl, vox, tile_start, cursor, entries,
                       tile_count, key_count, counters) != cudaSuccess) return LSS_ERR_CUDA;
        if (!sorted) {      // no sort kernel to clear the scratch counters of the next build: tile_count and cursor are contiguous
            if (cudaMemsetAsync(tile_count, 0, (size_t)((char *)sync - (char *)tile_count), s) != cudaSuccess) return LSS_ERR_CUDA;
        }
    } else if (lss_launch(k_plan_scatter<2, false>, dim3(grid_vi), dim3(256), 0, s, true, d, tl, vox, tile_start, cursor, entries,
                          tile_count, key_count, counters) != cudaSuccess) return LSS_ERR_CUDA;
    LSS_CHECK_LAUNCH();
    if (sorted) {
        const size_t sort_smem = (size_t)(3 * tl.TY + 1) * sizeof(int);
        if (sort_smem > 24 * 1024 &&
            cudaFuncSetAttribute(k_plan_sort<LSS_SORT_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_smem) != cudaSuccess)
            return LSS_ERR_CUDA;
        if (lss_launch(k_plan_sort<LSS_SORT_THREADS>, dim3(tl.n_tiles), dim3(LSS_SORT_THREADS), sort_smem, s, true,
                       d, tl, tile_start, entries, (uint32_t *)(w + L->off_segs), (int32_t *)(w + L->off_tile_nseg),
                       (int32_t *)(w + L->off_tile_row0), (int4 *)(w + L->off_seg_recs), key_count,
                       (int4 *)(w + L->off_mixed_recs), counters, prow, (long long)L->n_rows_cap,
                       scan_in_scatter ? tile_count : (int32_t *)nullptr, cursor) != cudaSuccess) return LSS_ERR_CUDA;
        LSS_CHECK_LAUNCH();
    }
    return LSS_OK;
}

extern "C" int lss_plan_build(const lss_problem *p, const lss_plan_layout *L, void *workspace, const float *geom,
                              const float *frustum, const float *post_trans, const float *M1, const float *M2,
                              const float *trans, int sorted, void *stream) {
    int st = lss_check_problem(p);
    if (st != LSS_OK) return st;
    LSS_REQUIRE(L && workspace, LSS_ERR_WORKSPACE);
    LSS_REQUIRE(geom != nullptr || (frustum && post_trans && M1 && M2 && trans), LSS_ERR_BAD_ARG);
    CalibPtrs c{frustum, post_trans, M1, M2, trans, nullptr, nullptr, nullptr};
    return plan_build_impl(p, L, workspace, geom, c, false, sorted, (cudaStream_t)stream);
}

extern "C" int lss_plan_build_raw(const lss_problem *p, const lss_plan_layout *L, void *workspace, const float *frustum,
                                  const float *rots, const float *trans, const float *intrins, const float *post_rots,
                                  const float *post_trans, int sorted, void *stream) {
    int st = lss_check_problem(p);
    if (st != LSS_OK) return st;
    LSS_REQUIRE(L && workspace, LSS_ERR_WORKSPACE);
    LSS_REQUIRE(frustum && rots && trans && intrins && post_rots && post_trans, LSS_ERR_BAD_ARG);
    const long long dhw = (long long)p->D * p->fH * p->fW;
    LSS_REQUIRE(512 / dhw + 2 <= LSS_RAW_CAMS, LSS_ERR_UNSUPPORTED);   // cameras one CTA (up to 512 points) may span
    CalibPtrs c{frustum, post_trans, nullptr, nullptr, trans, rots, intrins, post_rots};
    return plan_build_impl(p, L, workspace, nullptr, c, true, sorted, (cudaStream_t)stream);
}

extern "C" int lss_plan_reference_order(const lss_problem *p, const lss_plan_layout *L, const void *workspace,
                                        int32_t *scratch, int64_t *order_out, int32_t *n_kept_out, void *stream) {
    int st = lss_check_problem(p);
    if (st != LSS_OK) return st;
    LSS_REQUIRE(L && workspace, LSS_ERR_WORKSPACE);
    LSS_REQUIRE(scratch && order_out, LSS_ERR_BAD_ARG);
    const Dims d = make_dims(p);
    const Tiling tl = make_tiling(L);
    const char *w = (const char *)workspace;
    const uint32_t *entries = (const uint32_t *)(w + L->off_entries);
    const int32_t *tile_start = (const int32_t *)(w + L->off_tile_start);
    const int n_ranks = d.B * d.nx * d.ny * d.nz;
    cudaStream_t s = (cudaStream_t)stream;
    k_fill_i32<<<592, 256, 0, s>>>(scratch, (int64_t)n_ranks + 1, 0);
    k_fill_i64<<<592, 256, 0, s>>>((long long *)order_out, d.n_points, -1);
    k_ref_count<<<tl.n_tiles, 128, 0, s>>>(d, tl, tile_start, entries, scratch);
    k_scan_single<<<1, 1024, 0, s>>>(scratch, n_ranks, n_kept_out);
    k_ref_place<<<tl.n_tiles, 128, 0, s>>>(d, tl, tile_start, entries, scratch, (long long *)order_out);
    LSS_CHECK_LAUNCH();
    return LSS_OK;
}
