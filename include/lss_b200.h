/*
 * lss_b200.h  --  C ABI of the B200-native lift-splat library (liblss_b200.so, sm_100a only).
 *
 * Drop-in boundary for ONE path of shdragron/LSS-Carla: frustum geometry + lift + splat, forward and
 * backward (BASELINE.json north_star).  The reference has no FFI of its own -- the
 * path is Python calling stock ATen ops -- so every entry point below names the reference function
 * (file:line under the reference root) whose work it replaces; `INTEGRATION.md` shows the ctypes stub
 * and the `src/models.py` patch a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C: pointers, ints, floats, one POD struct; no torch / C++ types cross this boundary.
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`; the caller owns all memory,
 *     including workspaces (sizes from lss_plan_layout); nothing is allocated or freed inside.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it; no call synchronises
 *     the device or the stream, so every call is CUDA-graph capturable.
 *   - return value: LSS_OK (0) or a negative lss_status; no exceptions; re-entrant (distinct plans may be used
 *     from distinct streams concurrently); no global state apart from the process-wide options of
 *     lss_set_option, no environment variables.
 *   - there is NO CPU fallback: without a CUDA device the compute entry points return LSS_ERR_CUDA.
 *   - float tensors are IEEE binary32, C-contiguous in the stated shape unless strides are passed.
 *
 * Point / voxel numbering
 *   point  p = (((b*N + n)*D + d)*fH + h)*fW + w          "flat (b,n,d,h,w) index", models.py:205-213
 *   voxel  v = ((b*nz + iz)*nx + ix)*ny + iy               dense id used by the library (-1 = dropped)
 *   rank     = ix*(ny*nz*B) + iy*(nz*B) + iz*B + b         the reference's sort key, models.py:226-229
 *   BEV element (b, iz*C + c, ix, iy), shape [B, nz*C, nx, ny]                      models.py:240-244
 */
#ifndef LSS_B200_H
#define LSS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LSS_B200_VERSION 100 /* 0.1.0 */

typedef enum lss_status {
    LSS_OK = 0,
    LSS_ERR_BAD_ARG = -1,     /* null pointer, non-positive dimension, inconsistent sizes            */
    LSS_ERR_ALIGN = -2,       /* a pointer is not aligned as documented (16 B for float rows)        */
    LSS_ERR_UNSUPPORTED = -3, /* dimensions outside the compiled limits (see lss_limits)             */
    LSS_ERR_CUDA = -4,        /* a CUDA runtime call failed (launch error, no device, ...)           */
    LSS_ERR_WORKSPACE = -5    /* workspace pointer null / too small for this problem                 */
} lss_status;

/* Problem description.  Mirrors what LiftSplatShoot.__init__ derives from grid_conf / data_aug_conf
 * (models.py:134-150) plus the batch shape.  `lo[k] = bx[k] - dx[k]/2` must be computed in float32 by
 * the host exactly as the reference's tensor expression does (models.py:212). */
typedef struct lss_problem {
    int32_t B, N, D, fH, fW, C; /* batch, cameras, depth bins, feature map, context channels        */
    int32_t nx, ny, nz;         /* voxel counts (gen_dx_bx, tools.py:174-179)                        */
    float dx[3];                /* voxel size  x, y, z                                               */
    float lo[3];                /* lower corner of voxel 0 = bx - dx/2                               */
} lss_problem;

/* BEV memory layouts.  Both describe the same logical tensor [B, nz*C, nx, ny]. */
enum { LSS_LAYOUT_NCHW = 0,          /* contiguous, y fastest (what the reference returns)           */
       LSS_LAYOUT_CHANNELS_LAST = 1  /* torch.channels_last strides: physical [B, nx, ny, nz*C]      */ };

/* Splat accumulation modes (north_star: "atomic" and "deterministic sort-by-rank segmented sum"). */
enum { LSS_SPLAT_SORTED = 0,      /* per voxel, ascending flat point index, sequential fp32 adds      */
       LSS_SPLAT_SMEM_ATOMIC = 1, /* tile-owner CTAs, shared-memory atomics, order not fixed          */
       LSS_SPLAT_RED_GLOBAL = 2   /* memset + red.global.add.f32 straight into the BEV tensor         */ };

/* ---------------------------------------------------------------------------------------------- */
/* Introspection                                                                                   */
/* ---------------------------------------------------------------------------------------------- */

int lss_version(void);
const char *lss_status_string(int status);

/* Process-wide options (the library's only global state).  LSS_OPT_PDL: launch the kernel chains with programmatic
 * dependent launch (default 1); 0 = plain stream order (compute-sanitizer runs, A/B measurements). */
enum { LSS_OPT_PDL = 0, LSS_OPT_COUNT = 1 };
int lss_set_option(int option, int value);
int lss_get_option(int option);

/* Compiled limits: max per-sample points (N*D*fH*fW), max tile width, max D for the fused backward. */
typedef struct lss_limits { int32_t max_points_per_sample, max_tile_cols, max_depth_bins, max_channels; } lss_limits;
void lss_get_limits(lss_limits *out);

/* ---------------------------------------------------------------------------------------------- */
/* Plan: device-side index structures shared by forward and backward of one batch                  */
/* ---------------------------------------------------------------------------------------------- */

/* Byte offsets of the arrays inside the caller-allocated plan workspace (all 256-B aligned). */
typedef struct lss_plan_layout {
    int32_t tile_cols;      /* TY: BEV columns (iy) owned by one CTA tile, multiple of 8             */
    int32_t tiles_per_row;  /* ceil(ny / TY)                                                         */
    int32_t n_tiles;        /* B * nz * nx * tiles_per_row                                           */
    int64_t n_points;       /* B*N*D*fH*fW                                                           */
    size_t off_vox;         /* int32 [n_points]   dense voxel id or -1                               */
    size_t off_entries;     /* uint32[n_points]   bucketed (col << 20 | point-in-sample), per tile   */
    size_t off_tile_start;  /* int32 [n_tiles+1]  exclusive prefix of per-tile kept-point counts     */
    size_t off_segs;        /* uint32[n_points]   sorted plans: per tile, (col << 20 | first entry   */
                            /*                    within the bucket) of every non-empty voxel        */
    size_t off_tile_nseg;   /* int32 [n_tiles]    sorted plans: non-empty voxels per tile            */
    size_t off_tile_row0;   /* int32 [n_tiles]    sorted plans: first compact row of the tile (its   */
                            /*                    k-th non-empty voxel owns row tile_row0 + k)       */
    size_t off_seg_recs;    /* int32 [n_points,4] sorted plans: voxel records {first entry (global), */
                            /*                    length (or -(depth bin+1) if the voxel is exactly  */
                            /*                    the fH image rows of one column and depth),        */
                            /*                    batch index, compact row}, bucketed by the         */
                            /*                    camera column (b, n, w) of the voxel's first point:*/
                            /*                    bucket k owns slots [k*D*fH, (k+1)*D*fH)           */
    size_t off_key_count;   /* int32 [B*N*fW]     sorted plans: records in every bucket              */
    size_t off_mixed_recs;  /* int32 [n_rows_cap,4] sorted plans: records of the voxels that hold    */
                            /*                    points of several camera columns (not bucketed)    */
    size_t off_prow;        /* int32 [B,N,fW,D,fH] sorted plans: compact row of the point's voxel, -1 */
                            /*                    for dropped points, camera-column major (backward) */
    size_t off_counters;    /* int32 [64]         [0] = non-empty voxels of the batch (rows in use), */
                            /*                    [1] = records in mixed_recs, [2] = records of long */
                            /*                    voxels (>= 64 points), stored from the END of      */
                            /*                    mixed_recs downwards                               */
    int64_t n_rows_cap;     /* n_points: capacity (rows) of the voxel_sums / grad_rows workspaces     */
    size_t off_tile_count;  /* int32 [n_tiles]    scratch, all-zero between calls                    */
    size_t off_cursor;      /* int32 [n_tiles]    scratch                                            */
    size_t off_sync;        /* int32 [64]         scratch counters, all-zero between calls           */
    size_t bytes;           /* total workspace size                                                  */
} lss_plan_layout;

/* Fill `out` for problem `p`.  tile_cols <= 0 selects the default. Host-only, no CUDA calls. */
int lss_plan_layout_init(const lss_problem *p, int tile_cols, lss_plan_layout *out);

/* Zero the scratch counters of a freshly allocated (or possibly dirty) workspace. */
int lss_plan_reset(const lss_plan_layout *L, void *workspace, void *stream);

/* ---------------------------------------------------------------------------------------------- */
/* Geometry and voxel indices                                                                      */
/* ---------------------------------------------------------------------------------------------- */

/* Device-side replacement of the two host round trips in get_geometry (models.py:180,186):
 *   M1 = inverse(post_rots), M2 = rots @ inverse(intrins), closed-form adjugate inverse, fp32.
 * NOT bit-identical to the LAPACK inverse the reference calls on the CPU; the bit-exact contract of
 * this library starts at (M1, M2), which the Python layer by default prepares with the reference's
 * own torch calls.  rots/intrins/post_rots/M1/M2: f32[B*N,3,3]. */
int lss_calib_matrices(int32_t n_cams, const float *rots, const float *intrins, const float *post_rots,
                       float *M1, float *M2, void *stream);

/* LiftSplatShoot.get_geometry (models.py:170-190) given the prepared matrices.
 *   frustum f32[D,fH,fW,3] (the model's `frustum` parameter, models.py:157-168)
 *   post_trans, trans f32[B*N,3]; M1, M2 f32[B*N,3,3]
 *   geom_out f32[B,N,D,fH,fW,3]
 * Bit-exact w.r.t. the reference on CPU: un-fused fp32, rows as (a0*v0 + a1*v1) + a2*v2. */
int lss_geometry(const lss_problem *p, const float *frustum, const float *post_trans, const float *M1,
                 const float *M2, const float *trans, float *geom_out, void *stream);

/* Quantise + mask + (optionally) dump what voxel_pooling derives per point (models.py:212-229).
 * Either `geom` (f32[n_points,3], the `geom_feats` argument of voxel_pooling) is given, or it is null
 * and the calibration set (frustum, post_trans, M1, M2, trans) is used to evaluate the geometry in
 * registers (fused path; the geometry tensor is never written).
 * Outputs (each may be null): vox int32[n_points]; idx int64[n_points,3] (ix,iy,iz, truncation toward
 * zero, INT64_MIN for non-finite); kept uint8[n_points]; rank int64[n_points] (-1 where dropped). */
int lss_voxel_index(const lss_problem *p, const float *geom, const float *frustum,
                    const float *post_trans, const float *M1, const float *M2, const float *trans,
                    int32_t *vox, int64_t *idx, uint8_t *kept, int64_t *rank, void *stream);

/* Build the plan for one batch: voxel ids (as lss_voxel_index, from `geom` or from calibration),
 * per-tile buckets of kept points, each bucket sorted by (column, flat point index) -- i.e. the
 * reference's stable `ranks.argsort()` order inside every voxel (models.py:230).
 * Replaces models.py:212-231.  `sorted` = 0 skips the in-bucket sort (enough for the atomic modes and
 * for the backward pass). */
int lss_plan_build(const lss_problem *p, const lss_plan_layout *L, void *workspace, const float *geom,
                   const float *frustum, const float *post_trans, const float *M1, const float *M2,
                   const float *trans, int sorted, void *stream);

/* Same plan straight from the raw calibration of LiftSplatShoot.forward (models.py:256): the closed-form 3x3
 * inverses of lss_calib_matrices are evaluated inside the voxel-index kernel (identical bits), which saves
 * a launch per batch.  rots, intrins, post_rots f32[B*N,3,3]; trans, post_trans f32[B*N,3].
 * LSS_ERR_UNSUPPORTED if a camera has fewer than 43 frustum points (use lss_calib_matrices + lss_plan_build). */
int lss_plan_build_raw(const lss_problem *p, const lss_plan_layout *L, void *workspace, const float *frustum,
                       const float *rots, const float *trans, const float *intrins, const float *post_rots,
                       const float *post_trans, int sorted, void *stream);

/* Parity dump: the reference's sort permutation.  order_out int64[n_points] receives, for
 * i < n_kept, the flat point index of the i-th element of `x[kept][sorts]` (models.py:222-231), and -1
 * for i >= n_kept; n_kept_out int32[1].  Needs a plan built with sorted=1.  scratch int32[n_ranks+1]
 * with n_ranks = B*nx*ny*nz. */
int lss_plan_reference_order(const lss_problem *p, const lss_plan_layout *L, const void *workspace,
                             int32_t *scratch, int64_t *order_out, int32_t *n_kept_out, void *stream);

/* ---------------------------------------------------------------------------------------------- */
/* Lift + splat, fused level (depthnet output in, BEV out; the B*N*D*fH*fW*C tensor never exists)  */
/* ---------------------------------------------------------------------------------------------- */

/* CamEncode.get_depth_dist + the operand layout of get_depth_feat (models.py:49-61):
 *   depthnet_out f32[B*N, D+C, fH, fW]  ->  prob f32[B*N, D, fH, fW] = softmax over D,
 *                                            ctx_t f32[B*N, fH*fW, C] = context, pixel-major;
 *   optional (may be null) prob_col f32[B*N, fW, D, fH]: the same weights, camera-column major -- the block
 *   the GROUP forward gather stages per column; pass it on to lss_splat_fwd. */
int lss_lift_prepare(const lss_problem *p, const float *depthnet_out, float *prob, float *ctx_t,
                     float *prob_col, void *stream);
/* The same for a bfloat16 depthnet output (autocast training; the reference's AMP-less loop has no counterpart):
 * values are widened on load, everything downstream is the float32 path bit for bit -- the result equals
 * lss_lift_prepare on the widened tensor.  Against float32 inputs that were rounded to bfloat16 the BEV differs by the
 * input rounding only (stated tolerance of the bf16 path: rtol 3e-2 / atol 5e-3, tests/test_cuda_parity.py). */
int lss_lift_prepare_bf16(const lss_problem *p, const void *depthnet_out_bf16, float *prob, float *ctx_t,
                          float *prob_col, void *stream);

/* Kernel variants of the tile-owner forward in SORTED mode (bit-identical results). */
enum { LSS_VARIANT_AUTO = 0,   /* GROUP when the shape and workspace allow it, else WARP                          */
       LSS_VARIANT_WARP = 1,   /* ONE tile-owner kernel: a warp per 32-entry chunk, lane = channel; any C         */
       LSS_VARIANT_GROUP = 2,  /* TWO kernels: gather by 8-lane groups (lane = C/8 channels, operands staged per   */
                               /* camera column) into compact per-voxel rows, then a streaming tile-owner store;  */
                               /* C in {32,64,128}, fused level only, needs the `voxel_sums` workspace            */
       LSS_VARIANT_GROUP_GATHER = 3, /* measurement aid: only the gather kernel of GROUP (fills `voxel_sums`)   */
       LSS_VARIANT_GROUP_STORE = 4   /* measurement aid: only the store kernel of GROUP (reads `voxel_sums`)    */ };

/* Zero a BEV tensor (what `torch.zeros` does at models.py:240).  Only the RED_GLOBAL mode needs a zeroed
 * grid; callers may issue the clear early / on another stream and pass precleared=1 to lss_splat_fwd. */
int lss_bev_clear(const lss_problem *p, float *bev, void *stream);

/* voxel_pooling of the lifted features (models.py:59 outer product + :204-246), without materialising
 * them:  bev[b, iz*C+c, ix, iy] = sum_{p in voxel} prob[p] * ctx_t[pixel(p), c].
 * Modes SORTED / SMEM_ATOMIC write every BEV element exactly once (zeros included): no memset, no global
 * atomics.  `precleared` != 0 promises that `bev` is already all-zero (only RED_GLOBAL looks at it).
 * `prob_col`: optional column-major weights from lss_lift_prepare (null: staged from `prob`).
 * `b0, b1`: sample range [b0, b1) to process ((0, 0) = all).  Samples are independent, so a caller can issue the
 * GROUP variant in parts on two streams and let the store of one part overlap the gather of the next (the part
 * that starts at sample 0 also sums the voxels shared between camera columns of ALL samples: issue it first).
 * `voxel_sums`: caller workspace f32[n_rows_cap, C] (n_rows_cap = n_points) for the GROUP variant (may be null:
 * the WARP variant is used).  bev f32[B, nz*C, nx, ny] in `layout`. */
int lss_splat_fwd(const lss_problem *p, const lss_plan_layout *L, const void *workspace,
                  const float *prob, const float *ctx_t, const float *prob_col, float *voxel_sums, float *bev,
                  int mode, int layout, int variant, int precleared, int b0, int b1, void *stream);

/* Backward of lift+splat to the depthnet output (replaces QuickCumsum.backward tools.py:212-219 and the
 * autograd backward of models.py:58-59,:199-200,:240-244):
 *   grad_depthnet f32[B*N, D+C, fH, fW]  (first D channels: logits through the softmax; last C: context)
 * `grad_rows` is a caller workspace f32[max(B*nz*nx*ny, n_rows_cap), C]: the gradient rows of the non-empty voxels are
 * gathered into it, channel-contiguous (compact row order for sorted plans, voxel order otherwise); it may
 * be null only for channels_last gradients.  `plan_sorted` != 0 promises that the plan was built with
 * sorted=1, which (together with `prob_col` from lss_lift_prepare) enables the compact-row kernels
 * (C in {32,64,128}); `prob_col` may be null.
 * `stage`: 0 = everything, 1 = only the gradient-row gather, 2 = only the pixel gather (compact-row kernels);
 * `b0, b1`: sample range as in lss_splat_fwd -- the DRAM-bound row gather of one part overlaps the pixel gather of
 * the previous one when they are issued on two streams. */
int lss_splat_bwd(const lss_problem *p, const lss_plan_layout *L, const void *workspace,
                  const float *grad_bev, int layout, const float *prob, const float *ctx_t, const float *prob_col,
                  float *grad_rows, float *grad_depthnet, int plan_sorted, int stage, int b0, int b1, void *stream);

/* ---------------------------------------------------------------------------------------------- */
/* Run plan: the channels_last fast path of the fused level (default of bench.py and of the API)   */
/* ---------------------------------------------------------------------------------------------- */

/* With the BEV in channels_last (physical [B, nx, ny, nz*C]; what bevencode.conv1, models.py:97-98,:118, is fed
 * through cuDNN's NHWC kernels) every voxel is ONE contiguous C-float row, so the forward needs no tile-owner
 * store and the backward no gradient transposition.  The plan shrinks accordingly: points are handled as RUNS --
 * the fH image rows of one (camera, column, depth bin), which almost always share a voxel -- and nothing is sorted:
 *   plan build      per point: geometry + voxel row (models.py:179-188, :212-221); per sub-run (the points of a run
 *                   that share a voxel): one push on the voxel's list (a 64-bit exchange on head[voxel], tagged with
 *                   the build's epoch, so that nothing has to be cleared between builds)
 *   forward         the zero-fill of the tensor (bulk-copy engine, sample by sample, with a progress counter per
 *                   sample) runs next to the camera-column CTAs, which classify their sub-runs (alone on the voxel's
 *                   list = EXCLUSIVE: fH staged products in image-row order; otherwise the voxel is queued and a warp
 *                   at the end of the grid walks its list and sums its few points sorted by flat index), wait for
 *                   their sample's zeros and write every non-empty voxel row ONCE -- the same per-voxel sequential
 *                   float32 sum in ascending flat (b,n,d,h,w) order as LSS_SPLAT_SORTED, hence the same bits.  The
 *                   forward only reads the plan: a plan may be kept and used again while the calibration repeats
 *   backward        the pixel-owner gather reads gradient rows straight from the channels_last gradient
 * voxel "row" r = ((b*nx + ix)*ny + iy)*nz + iz; its C floats start at element r*C of the channels_last tensor. */
typedef struct lss_runplan_layout {
    int64_t n_points;       /* B*N*D*fH*fW                                                                   */
    int64_t n_runs;         /* B*N*fW*D                                                                      */
    int64_t n_voxels;       /* B*nx*ny*nz                                                                    */
    size_t off_prow;        /* int32 [B,N,fW,D,fH]  voxel row of the point or -1, camera-column major        */
    size_t off_sub;         /* int32 [n_points,2]   at the first point of every sub-run: {the sub-run pushed on the   */
                            /*                      voxel's list before it (point index + 1, 0 = none), row mask:     */
                            /*                      bit j = image row h+j of the same run belongs to it}; {0,0} elsewhere */
    size_t off_sub2;        /* int32 [n_points,2]   per point {context row = pixel (bn*fH + h)*fW + w, flat index in the sample */
                            /*                      ((n*D + d)*fH + h)*fW + w}: what the sums of shared voxels need        */
    size_t off_pool;        /* uint32[n_points]     forward scratch: point keys of voxels beyond the shared-memory sort */
    int64_t n_rec_cap;      /* capacity (records) of recs                                                    */
    size_t off_recs;        /* int32 [n_rec_cap,4]  queue of the voxels shared by several sub-runs, filled by the build: every    */
                            /*                      sub-run that found another one on its voxel's list queued {voxel row,         */
                            /*                      previous sub-run, itself (both: point index + 1), its row mask}               */
    size_t off_longs;       /* int32 [n_points/64+2,4] forward scratch: overflow list of voxels with >= 64 points                 */
    size_t off_counters;    /* int32 [64]           [0] epoch of the last build; [6] voxels shared by several sub-runs   */
                            /*                      and [7] voxels with >= 64 points, as summed by the last forward;     */
                            /*                      the rest: scratch, zero between launches                             */
    size_t off_qcount;      /* uint64[32] (128 bytes apart) (epoch << 32) | records of sub-queue s; record k of sub-queue s */
                            /*                      sits in recs[k*32 + s]                                               */
    size_t off_zero_done;   /* int32 [B,32]         forward scratch: zero-fill progress per sample ([b][0]; one 128-byte  */
                            /*                      line each), zero between launches                                    */
    size_t off_ready;       /* int32 [32,32]        forward scratch: 32 copies ([i][0]) of the "plan + lift operands     */
                            /*                      complete" flag of lss_liftsplat_forward, zero between launches       */
    size_t off_head;        /* uint64[n_voxels]     (build epoch << 32) | (point index + 1) of the last sub-run pushed on */
                            /*                      the voxel's list; entries of older epochs are stale and read as empty */
    size_t bytes;
} lss_runplan_layout;

/* Fill `out` for problem `p`.  LSS_ERR_UNSUPPORTED if fH > 32 (a run must fit a warp) or C is not 32 / 64 / 128:
 * use the tile plan (lss_plan_build + lss_splat_fwd) for such shapes.  Host-only. */
int lss_runplan_layout_init(const lss_problem *p, lss_runplan_layout *out);
/* Zero the counters and list heads of a freshly allocated workspace (build epochs keep them consistent afterwards;
 * the forward tells builds apart by 31 bits of the epoch: reset once more before 2^31 builds into the same workspace). */
int lss_runplan_reset(const lss_runplan_layout *L, void *workspace, void *stream);
/* Build the run plan of one batch from calibration (replaces models.py:170-190 + :212-231).  Either the prepared
 * matrices M1 = inverse(post_rots), M2 = rots @ inverse(intrins) are given (bit-exact w.r.t. the reference's host
 * inverses) and rots / intrins / post_rots may be null, or M1 and M2 are null and the closed-form inverses of
 * lss_calib_matrices are evaluated inside the kernel from the raw calibration. */
/* LSS_OK if lss_runplan_build / lss_liftsplat_prologue take the raw calibration (M1 == M2 == null) for this problem,
 * LSS_ERR_UNSUPPORTED if the cameras are so small that an index CTA would span too many of them: prepare M1 / M2 with
 * lss_calib_matrices then.  Host-only. */
int lss_runplan_raw_supported(const lss_problem *p);
int lss_runplan_build(const lss_problem *p, const lss_runplan_layout *L, void *workspace, const float *frustum,
                      const float *post_trans, const float *M1, const float *M2, const float *trans,
                      const float *rots, const float *intrins, const float *post_rots, void *stream);
/* Fused prologue, ONE launch: the run plan (frustum == null: off, e.g. a kept plan; arguments as lss_runplan_build), the
 * lift operands (depthnet_out == null: off; outputs as lss_lift_prepare, prob_col required) and, optionally, a zero-fill of
 * `bev` (null: off) as independent CTA roles of one grid.  Afterwards: lss_liftsplat_fwd_cl(..., LSS_ZERO_PRECLEARED) if `bev`
 * was given here, LSS_ZERO_ORDERED otherwise.  (The pieces of lss_liftsplat_forward as separate, stream-ordered calls.) */
int lss_liftsplat_prologue(const lss_problem *p, const lss_runplan_layout *L, void *workspace, const float *frustum,
                           const float *post_trans, const float *M1, const float *M2, const float *trans,
                           const float *rots, const float *intrins, const float *post_rots, const float *depthnet_out,
                           float *prob, float *ctx_t, float *prob_col, float *bev, void *stream);
/* torch.zeros of models.py:240 through the bulk-copy engine (cp.async.bulk shared -> global): one thread per CTA
 * issues the stores, so the kernel leaves the SMs to whatever runs next to it.
 * `part` of `n_parts`: zero only that slice of the tensor; (0, 1) = everything. */
int lss_bev_zero(const lss_problem *p, float *bev, int part, int n_parts, void *stream);
/* Forward from an existing plan and existing lift operands, ONE launch (k_fwd_columns): bev (channels_last) receives the
 * sum of every non-empty voxel and zeros elsewhere.  prob_col f32[B*N, fW, D, fH], ctx_t f32[B*N, fH*fW, C] from lss_lift_prepare /
 * lss_liftsplat_prologue.  `zero_mode`:
 *   LSS_ZERO_ORDERED     the kernel zero-fills `bev` itself: its first CTAs stream the zeros sample by sample, the camera-column
 *                        CTAs wait for their sample's progress counter before they write
 *   LSS_ZERO_PRECLEARED  `bev` is already all-zero (lss_bev_zero, the prologue's zero role)
 * The workspace is scratch for the duration of the launch (progress counters, pool): one forward per plan at a time. */
enum { LSS_ZERO_ORDERED = 0, LSS_ZERO_PRECLEARED = 1 };
int lss_liftsplat_fwd_cl(const lss_problem *p, const lss_runplan_layout *L, void *workspace,
                         const float *prob_col, const float *ctx_t, float *bev, int zero_mode, void *stream);
/* The whole forward of a step (replaces models.py:170-190, :49-61, :192-246 for one batch) as THREE launches that run
 * side by side:
 *   k_zero_flags   zero-fill of `bev` by the bulk-copy engine, sample by sample, one progress counter per sample; two
 *                  one-warp CTAs per SM, launched first so that they are spread evenly; lets its successors start at once
 *   k_prologue     lift operands of `depthnet_out` || plan build from the calibration (frustum == null: the plan in the
 *                  workspace is kept); its last CTA raises READY
 *   k_fwd_columns  launched programmatically: one CTA per camera column polls READY (not the completion of the grids
 *                  before it), classifies its sub-runs, sums the exclusive voxels from staged operands, waits for its
 *                  sample's zeros and writes the rows; the voxels shared by several sub-runs are queued and summed, one
 *                  warp each, by the CTAs at the end of the same grid
 * so that the only bandwidth-bound piece (G bytes of zeros) runs from the first microsecond of the step, and the latency
 * chains next to it.  Arguments as lss_liftsplat_prologue (depthnet_out, prob, ctx_t, prob_col, bev required). */
int lss_liftsplat_forward(const lss_problem *p, const lss_runplan_layout *L, void *workspace, const float *frustum,
                          const float *post_trans, const float *M1, const float *M2, const float *trans,
                          const float *rots, const float *intrins, const float *post_rots, const float *depthnet_out,
                          float *prob, float *ctx_t, float *prob_col, float *bev, void *stream);
/* The same forward into a PERSISTENT output tensor (no counterpart in the reference, whose voxel_pooling allocates a fresh
 * torch.zeros tensor per call, models.py:240): `bev` must still hold what the previous lss_liftsplat_forward /
 * lss_liftsplat_forward_persistent call on this workspace wrote -- or be all zero, with a workspace that has not been built since
 * lss_runplan_reset.  Only the rows that call wrote can be non-zero, and the workspace still names them: the plan build zeroes
 * exactly those (10 MB instead of the tensor's 82 MB at the benchmark configuration) before it overwrites the plan; with
 * frustum == NULL (plan kept) the rows to be written are the rows that were written and nothing is cleared.  There is no
 * zero-fill grid and nobody waits for zeros: two launches.  The result is bit-identical to lss_liftsplat_forward's.
 * The caller owns the pairing of workspace and tensor; a tensor that anybody else wrote to must go through
 * lss_liftsplat_forward (or lss_bev_zero + lss_runplan_reset) once. */
int lss_liftsplat_forward_persistent(const lss_problem *p, const lss_runplan_layout *L, void *workspace, const float *frustum,
                                     const float *post_trans, const float *M1, const float *M2, const float *trans,
                                     const float *rots, const float *intrins, const float *post_rots,
                                     const float *depthnet_out, float *prob, float *ctx_t, float *prob_col, float *bev,
                                     void *stream);
/* Backward to the depthnet output from a channels_last BEV gradient (tools.py:212-219 + autograd of models.py:58-59). */
int lss_liftsplat_bwd_cl(const lss_problem *p, const lss_runplan_layout *L, const void *workspace,
                         const float *grad_bev, const float *prob_col, const float *ctx_t, float *grad_depthnet,
                         void *stream);

/* ---------------------------------------------------------------------------------------------- */
/* Operator level: LiftSplatShoot.voxel_pooling(geom_feats, x) with a materialised x                */
/* ---------------------------------------------------------------------------------------------- */

/* x f32 with logical shape [B,N,D,fH,fW,C] and element strides xs[6] (it usually arrives as a permuted
 * view, models.py:199-200).  bev as above.  Replaces models.py:204-246 given a plan built from `geom`. */
int lss_voxel_pooling_fwd(const lss_problem *p, const lss_plan_layout *L, const void *workspace,
                          const float *x, const int64_t *xs_host, float *bev, int mode, int layout,
                          int variant, int precleared, void *stream);

/* grad_x f32[n_points, C] contiguous: row p = grad_bev[b, iz*C:(iz+1)*C, ix, iy] of p's voxel, or 0. */
int lss_voxel_pooling_bwd(const lss_problem *p, const lss_plan_layout *L, const void *workspace,
                          const float *grad_bev, int layout, float *grad_rows, float *grad_x,
                          void *stream);

/* ---------------------------------------------------------------------------------------------- */
/* Operator level: tools.QuickCumsum / tools.cumsum_trick (tools.py:182-219)                        */
/* ---------------------------------------------------------------------------------------------- */

/* Pass 1: run structure of an already sorted rank vector.  run_id int32[n] = index of the run each
 * element belongs to; n_runs int32[1] (device).  scratch int32[lss_quickcumsum_scratch_elems(n)]. */
size_t lss_quickcumsum_scratch_elems(int64_t n);
int lss_quickcumsum_runs(int64_t n, const int64_t *ranks, int32_t *run_id, int32_t *n_runs,
                         int32_t *scratch, void *stream);
/* Pass 2 (after the caller read n_runs to size the outputs, as the reference's boolean indexing does,
 * tools.py:200):  sums f32[n_runs, C] = per-run sequential sum of x rows (x f32[n, C], row stride
 * `x_row_stride` elements); geom_out int64[n_runs, 4] = geom_feats row of each run's LAST element.
 * `scratch` is the buffer lss_quickcumsum_runs filled (it holds the run offsets). */
int lss_quickcumsum_fwd(int64_t n, int32_t C, const float *x, int64_t x_row_stride,
                        const int64_t *geom_feats, const int32_t *scratch, int32_t n_runs, float *sums,
                        int64_t *geom_out, void *stream);
/* Backward = gather (tools.py:212-219): grad_x[i, :] = grad_sums[run_id[i], :]. */
int lss_quickcumsum_bwd(int64_t n, int32_t C, const float *grad_sums, const int32_t *run_id,
                        float *grad_x, void *stream);

/* ---------------------------------------------------------------------------------------------- */
/* Host-buffer pipeline helpers (used by lss_carla_b200.api.StepPipeline; the reference has no        */
/* counterpart: its loader hands CUDA tensors to LiftSplatShoot.forward, train_simbev.py:232-239)     */
/* ---------------------------------------------------------------------------------------------- */

/* Events without timing, for stream-to-stream ordering.  lss_pipe_event_create returns NULL on failure. */
void *lss_pipe_event_create(void);
int lss_pipe_event_destroy(void *event);
int lss_pipe_event_synchronize(void *event);
/* One pipeline stage, enqueued with a single call: `stream` waits for wait_a and wait_b (each may be NULL), runs
 * n_copies (<= 4) cudaMemcpyAsync(dst[i], src[i], bytes[i], default kind: pinned host <-> device), then records
 * `record` (may be NULL).  Returns LSS_OK or LSS_ERR_CUDA. */
int lss_pipe_stage(void *stream, void *wait_a, void *wait_b, int32_t n_copies, void *const *dst,
                   const void *const *src, const size_t *bytes, void *record);
/* One whole pipelined step, enqueued with a single call (the host side of api.StepPipeline.run):
 *   copy_in_stream:  wait ev_compute (the previous run of this step has read its inputs), copy in_bytes from pinned in_host to
 *                    in_dev, record ev_in;
 *   compute_stream:  wait ev_in and ev_done (the previous results have left the device), launch `graph_exec` (a cudaGraphExec_t
 *                    holding the step's kernels, e.g. torch.cuda.CUDAGraph.raw_cuda_graph_exec()), record ev_compute;
 *   copy_out_stream: wait ev_compute, copy out_bytes from out_dev to pinned out_host, record ev_done.
 * All three events must have been recorded once before the first call.  Returns LSS_OK, LSS_ERR_BAD_ARG or LSS_ERR_CUDA. */
int lss_pipe_step(void *copy_in_stream, void *compute_stream, void *copy_out_stream, void *graph_exec,
                  void *ev_in, void *ev_compute, void *ev_done, void *in_dev, const void *in_host, size_t in_bytes,
                  void *out_host, const void *out_dev, size_t out_bytes);

#ifdef __cplusplus
}
#endif
#endif /* LSS_B200_H */
