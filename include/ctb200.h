/* ctb200.h -- C ABI of libctb200.so: the B200 (sm_100a) Splat / Slice hot path of Cloud Transformers.
 *
 * Drop-in boundary for the reference's L1 operators (all citations relative to the reference repo):
 *   layers/cloud_transform.py:62-121  DifferentiablePositions.forward  (+ layers/utils.py:100-186)
 *   layers/cloud_transform.py:124-180 Splat.forward   (torch_scatter.scatter_max call at :171-173)
 *   layers/cloud_transform.py:183-227 Slice.forward   (torch.gather call at :216-218)
 * and the autograd backward of each.
 *
 * Conventions
 *   - Plain C: raw DEVICE pointers, sizes, and a `void* stream` (a cudaStream_t; NULL = legacy default
 *     stream).  No torch types.  The library allocates nothing that outlives a call, keeps no state
 *     between calls, never synchronises the device and only enqueues work on the given stream, so it
 *     is safe to call from PyTorch's autograd thread and alongside DDP's side streams.
 *   - All tensors are dense, row-major (C-contiguous) with the reference's shapes:
 *       keys      f32 [B, H*dim, N]      head-major then axis (cloud_transform.py:89)
 *       lc        f32 [B, H, S, N]       S = 2^dim corner weights ("local_coordinate")
 *       idx       i64 [B, H, S, N]       flattened cell index of each corner ("flattened_index")
 *       features  f32 [B, H*F, N]
 *       pad       f32 [B, N] or NULL     "pts_padding" (cloud_transform.py:158-159, :224-225)
 *       grid      f32 [B, H*F, C]        C = prod(size); NCHW / NCDHW exactly as the reference returns
 *       arg       i32 [B, H*F, C]        winner e = s*N + n of every cell, -1 where nothing beat the 0 floor
 *   - Every entry returns 0 on success or a negative ctb_status; it never throws and never exits.
 *   - Corner order: bit0 of s is +1 on grid axis 0 (the slowest axis), bit1 axis 1, bit2 axis 2
 *     (layers/utils.py:103-110, :161-164); flat index = x*W1*W2 + y*W2 + z (cloud_transform.py:113-119).
 */
#ifndef CTB200_H
#define CTB200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTB_VERSION 200 /* major*100 + minor; 2.0: plan argument of ctb_slice_fwd_keys, EMD, SyncBN, Chamfer entries */

typedef enum ctb_status {
  CTB_OK = 0,
  CTB_ERR_INVALID_ARGUMENT = -1, /* NULL pointer, non-positive size, dim not in {2,3}, size < 2 ... */
  CTB_ERR_UNSUPPORTED = -2,      /* shape outside what the kernels support (e.g. S*N >= 2^31) */
  CTB_ERR_CUDA = -3,             /* a CUDA runtime call failed; see ctb_last_cuda_error() */
  CTB_ERR_WORKSPACE = -4         /* workspace missing or too small */
} ctb_status;

/* reduce operator of Splat: MAX is the reference (scatter_max onto a zero grid => implicit 0 floor,
 * cloud_transform.py:164-173); SUM is the scatter-add variant named by the north star. */
typedef enum ctb_reduce { CTB_REDUCE_MAX = 0, CTB_REDUCE_SUM = 1 } ctb_reduce;

/* algorithm selector of the fused entries:
 *   ATOMIC        point-stationary kernels on the NCHW grid in L2: red.global atomics for the scatters,
 *                 direct loads for the gathers.  Any shape.  Splat-max is reproducible (max is order
 *                 independent, the arg winner is resolved by a min-e pass); float sums are not.
 *   TILE          CTA-owned shared-memory tiles of the grid: scatters accumulate with native
 *                 shared-memory atomics and store every cell once (no zero-fill, no L2 atomics), gathers
 *                 receive the slab by TMA bulk copy.  Splat-max is reproducible; the sums (Slice backward
 *                 grad_grid, Splat-sum) are accumulated as exact fixed-point integers (|err| <= 2^-24 |sum|
 *                 + N 2^-39 max|v|), so they are bit-identical from run to run and under any permutation of
 *                 the points; only inputs with Inf / NaN fall back to float atomics.  The fast path.
 *   DETERMINISTIC as TILE, but the two scatters are REQUIRED to be the output-stationary kernels over the plan
 *                 (see ctb_plan_build): every cell is reduced by one owner in ascending e = s*N + n with plain
 *                 fp32 adds and stores -- no atomics anywhere, bit-identical from run to run and equal to a
 *                 sequential loop over the entries.  TILE uses the same kernels whenever it is given a plan. */
typedef enum ctb_mode { CTB_MODE_ATOMIC = 0, CTB_MODE_DETERMINISTIC = 1, CTB_MODE_TILE = 2 } ctb_mode;

/* element type of the GRID tensors (z, convolved grid, grad_grid, grad_z) of the fused entries.  BF16 is the storage
 * mode of the north star: grids live in HBM as bf16, every product / max / sum is still computed in fp32 on chip
 * (tolerance rel 1e-2 instead of 1e-5).  Keys, features, per-point outputs and arg are unaffected. */
typedef enum ctb_dtype { CTB_DTYPE_F32 = 0, CTB_DTYPE_BF16 = 1 } ctb_dtype;

/* geometry of one call.  size[2] is ignored when dim == 2. */
typedef struct ctb_shape {
  int32_t B;       /* batch of clouds */
  int32_t H;       /* heads */
  int32_t F;       /* feature channels per head (ignored by the positions entries) */
  int32_t N;       /* points per cloud */
  int32_t dim;     /* 2 or 3 */
  int32_t size[3]; /* grid extent per axis, each >= 2 (tensor_size, cloud_transform.py:41-46) */
  int32_t grid_dtype; /* ctb_dtype of the grid tensors; BF16 only with CTB_MODE_TILE fused entries */
} ctb_shape;

int ctb_version(void);
const char* ctb_strerror(int status);
/* cudaError_t of the last failing CUDA call made by this library on the calling thread (0 if none). */
int ctb_last_cuda_error(void);

/* ---- A1 / A7: DifferentiablePositions (cloud_transform.py:72-121, utils.py:100-186) ------------- */
/* keys -> lc, idx.  Bit-exact replay of clamp(+-(1-1e-7)) -> +1 -> *(W-1)/2 -> floor -> corner products. */
int ctb_positions_fwd(const float* keys, float* lc, int64_t* idx, const ctb_shape* shape, void* stream);
/* grad_lc -> grad_keys: identity through GradientBalancing (cloud_transform.py:21-23, no (W-1)/2
 * factor), zero where the key was clamped (:91). */
int ctb_positions_bwd(const float* keys, const float* grad_lc, float* grad_keys, const ctb_shape* shape,
                      void* stream);

/* ---- reference-API operators: lc / idx are tensors handed in by the caller ---------------------- */
/* A2+A3  Splat.forward (cloud_transform.py:131-180): z[b,h,f,c] = max(0, max_{(s,n): idx=c} feat*pad*lc)
 * (or the plain sum for CTB_REDUCE_SUM).  z and arg are fully overwritten; arg may be NULL for SUM. */
int ctb_splat_fwd(const float* lc, const int64_t* idx, const float* features, const float* pad, float* z,
                  int32_t* arg, const ctb_shape* shape, int reduce, void* stream);
/* A6  Splat backward (ScatterMax::backward + MulBackward): grad_z -> grad_features, grad_lc. */
int ctb_splat_bwd(const float* lc, const int64_t* idx, const float* features, const float* pad,
                  const float* grad_z, const int32_t* arg, float* grad_features, float* grad_lc,
                  const ctb_shape* shape, int reduce, void* stream);
/* A4  Slice.forward (cloud_transform.py:190-227): out[b,h,f,n] = pad * sum_s lc * grid[b,h,f,idx]. */
int ctb_slice_fwd(const float* lc, const int64_t* idx, const float* grid, const float* pad, float* out,
                  const ctb_shape* shape, void* stream);
/* A5  Slice backward (gather backward = scatter-add): grad_out -> grad_grid (fully overwritten), grad_lc. */
int ctb_slice_bwd(const float* lc, const int64_t* idx, const float* grid, const float* pad,
                  const float* grad_out, float* grad_grid, float* grad_lc, const ctb_shape* shape,
                  void* stream);

/* ---- fused operators: positions are recomputed from keys inside every kernel -------------------- */
/* Same results as the reference-API operators fed with ctb_positions_fwd's outputs (identical device
 * code computes the positions), but lc / idx are never materialised and the backward entries emit
 * grad_keys directly (A7 folded in).
 *
 * mode: see ctb_mode.  CTB_MODE_DETERMINISTIC needs the `plan` built by ctb_plan_build from the same
 *        keys for the two scatters (ctb_splat_fwd_keys, grad_grid of ctb_slice_bwd_keys).  TILE and
 *        DETERMINISTIC return CTB_ERR_UNSUPPORTED for shapes the tile kernels do not cover (ask
 *        ctb_mode_supported first) -- there is no silent fallback inside the library. */
typedef enum ctb_op {
  CTB_OP_SPLAT_FWD = 0,
  CTB_OP_SPLAT_BWD = 1,
  CTB_OP_SLICE_FWD = 2,
  CTB_OP_SLICE_BWD = 3
} ctb_op;
/* 1 if `op` can run in `mode` on this shape (reduce only matters for the Splat ops). */
int ctb_mode_supported(const ctb_shape* shape, int op, int reduce, int mode);

/* The plan: the points of every (b, h) unit sorted (stably) by their base cell, built once per set of keys and
 * shared by the operators that see those keys (an MHCT block feeds the same positions to Splat and Slice,
 * layers/multihead_ct.py:99-107).  With a plan the two scatters (ctb_splat_fwd_keys, grad_grid of
 * ctb_slice_bwd_keys) run output-stationary: every grid cell reduces the points of its 2^dim neighbouring bins in
 * registers, in ascending e = s*N + n -- no atomics, fixed summation order.
 * ctb_plan_used: 1 if the kernels of `mode` use a plan on this shape (then the caller should build one and pass it;
 * DETERMINISTIC requires it, TILE falls back to its shared-memory tile scatters when plan == NULL). */
int ctb_plan_used(const ctb_shape* shape, int mode);
/* 1 if `op` (CTB_OP_SPLAT_FWD with `reduce`, CTB_OP_SLICE_FWD, or CTB_OP_SLICE_BWD) reads the plan in `mode` on this shape.  Lets a
 * caller build the plan off the critical path when only the backward needs it (16^3 x F16: the forward max keeps the
 * tile scatter, the grad_grid sum of Slice backward is plan-based). */
int ctb_op_uses_plan(const ctb_shape* shape, int op, int reduce, int mode);
/* bytes of the plan for this shape, 0 if the shape cannot be planned. */
size_t ctb_plan_bytes(const ctb_shape* shape);
int ctb_plan_build(const float* keys, void* plan, size_t plan_bytes, const ctb_shape* shape, void* stream);

/* grid tensors (z, grad_z, grid, grad_grid) are void*: f32 or bf16 according to shape->grid_dtype */
int ctb_splat_fwd_keys(const float* keys, const float* features, const float* pad, void* z, int32_t* arg,
                       const ctb_shape* shape, int reduce, int mode, const void* plan, void* stream);
int ctb_splat_bwd_keys(const float* keys, const float* features, const float* pad, const void* grad_z,
                       const int32_t* arg, float* grad_features, float* grad_keys, const ctb_shape* shape,
                       int reduce, int mode, void* stream);
/* plan: may be NULL; where ctb_op_uses_plan(shape, CTB_OP_SLICE_FWD, ..) is 1, a plan lets the gather walk the points
 * in cell-sorted order (same results bit for bit, fewer shared-memory bank conflicts). */
int ctb_slice_fwd_keys(const float* keys, const void* grid, const float* pad, float* out,
                       const ctb_shape* shape, int mode, const void* plan, void* stream);
int ctb_slice_bwd_keys(const float* keys, const void* grid, const float* pad, const float* grad_out,
                       void* grad_grid, float* grad_keys, const ctb_shape* shape, int mode,
                       const void* plan, void* stream);

/* ---- A8: per-head projection + tanh in front of DifferentiablePositions ------------------------------ */
/* keys[b, h*dim + j, n] = tanh( ((pcd[b,:,n] + res_scale*keys_res[b,h,:,n] + shift[h]) . rot[h])_j * scales[h][j] )
 * Replaces VolTransformer / PlaneTransformer.forward + torch.tanh (layers/utils.py:25-34, :53-61;
 * layers/multihead_ct.py:93-97; multihead_ct_adain.py:112-115 where res_scale is the learnable `scale`).
 *   pcd f32 [B,3,N], keys_res f32 [B,H,3,N] or NULL, shift f32 [H,3], rot f32 [H,3,3] (= so3_exponential_map(log_R),
 *   row-vector convention q_j = sum_c p_c rot[h][c][j]), scales f32 [H,dim] or NULL, keys f32 [B,H*dim,N]. */
int ctb_project_fwd(const float* pcd, const float* keys_res, float res_scale, const float* shift, const float* rot,
                    const float* scales, float* keys, const ctb_shape* shape, void* stream);
/* Same, and additionally accumulates into key_stats (f64 [2], device memory, caller zeroes; may be NULL) the sum and
 * the sum of squares of the PRE-tanh keys: the mean / variance every MHCT block logs (layers/multihead_ct.py:109-113,
 * multihead_ct_adain.py:127-131, multihead_ct_pool.py:76-80) without materialising the pre-tanh tensor. */
int ctb_project_fwd_stats(const float* pcd, const float* keys_res, float res_scale, const float* shift, const float* rot,
                          const float* scales, float* keys, double* key_stats, const ctb_shape* shape, void* stream);
/* Backward of ctb_project_fwd.  grad_pcd f32 [B,3,N] (summed over heads), grad_keys_res f32 [B,H,3,N] or NULL,
 * param_acc f32 [H,16] ACCUMULATED into (caller zeroes): cols 0-2 d shift, 3-11 d rot (row major c,j; only j < dim
 * are written), 12-14 d scales, 15 d res_scale (per head; sum over heads for the scalar). */
int ctb_project_bwd(const float* pcd, const float* keys_res, float res_scale, const float* shift, const float* rot,
                    const float* scales, const float* keys, const float* grad_keys, float* grad_pcd,
                    float* grad_keys_res, float* param_acc, float* workspace, size_t workspace_bytes,
                    const ctb_shape* shape, void* stream);
/* bytes of the caller-provided scratch of ctb_project_bwd (per-CTA partial parameter gradients, reduced in a
 * fixed order so the result is deterministic). */
size_t ctb_project_bwd_workspace_bytes(const ctb_shape* shape);

/* A9  occupancy statistic of MultiHead blocks (layers/multihead_ct.py:104-105): count of |z| > 1e-9
 * accumulated into *count (u64, device memory, caller zeroes it). */
int ctb_count_occupied(const float* z, uint64_t n_elements, unsigned long long* count, void* stream);

/* ---- N3: Chamfer nearest-neighbour distances (the completion loss, train_inpainter.py:190) ------------------------ */
/* Replaces chamfer_extension/chamfer.cu:136-153 chamfer_cuda_forward (NmDistanceKernel :12-134) as bound by
 * chamfer_extension/dist_chamfer.py:10-38:  dist1[b,j] = min_k |xyz1[b,j] - xyz2[b,k]|^2, idx1 = the FIRST minimising k
 * (ascending k, strict <), and the same with the clouds swapped.  xyz1 f32 [B,n,3], xyz2 f32 [B,m,3] (contiguous),
 * dist f32, idx i32.  workspace: ctb_chamfer_workspace_bytes(B, n, m) bytes of device scratch. */
size_t ctb_chamfer_workspace_bytes(int B, int n, int m);
int ctb_chamfer_fwd(const float* xyz1, const float* xyz2, float* dist1, float* dist2, int32_t* idx1, int32_t* idx2,
                    void* workspace, size_t workspace_bytes, int B, int n, int m, void* stream);
/* Replaces chamfer_cuda_backward (chamfer.cu:176-195, NmDistanceGradKernel :155-174): grad_xyz1 / grad_xyz2 are fully
 * overwritten (zeroed, then accumulated with atomics like the reference: g = 2 grad_dist, +g (p1 - p2) on the query,
 * -g (p1 - p2) on its nearest neighbour). */
int ctb_chamfer_bwd(const float* xyz1, const float* xyz2, const float* grad_dist1, const float* grad_dist2,
                    const int32_t* idx1, const int32_t* idx2, float* grad_xyz1, float* grad_xyz2, int B, int n, int m,
                    void* stream);

/* ---- N2: SyncBatchNorm with the cross-GPU exchange fused into its kernels (train_classification.py:107-109) -------- */
/* Replaces emd_linear/emd_cuda.cu:223-275 emd_cuda_forward as bound by emd_linear/emd_module.py:30-63: the auction
 * approximation of the Earth Mover's Distance between two clouds of n points each, xyz f32 [B, n, 3] in [0, 1]^3.
 * dist f32 [B, n] = squared distance of every point of xyz1 to its assigned point of xyz2, assignment i32 [B, n]
 * (not guaranteed to be a bijection: sources still unassigned in the last iteration take the target they bid on).
 * One launch for all `iters` iterations.  Up to 4096 points the reference's eleven work tensors live in shared memory
 * (ctb_emd_workspace_bytes = 0, workspace may be NULL); larger clouds (the inpainting decoder: 16384) need
 * ctb_emd_workspace_bytes(B, n) bytes of device scratch, uninitialised.  n <= ctb_emd_max_points() (32768), any n (the
 * reference needs multiples of 1024), any B. */
int ctb_emd_max_points(void);
size_t ctb_emd_workspace_bytes(int B, int n);
int ctb_emd_fwd(const float* xyz1, const float* xyz2, float* dist, int32_t* assignment, void* workspace,
                size_t workspace_bytes, int B, int n, float eps, int iters, void* stream);
/* Replaces emd_cuda_backward (emd_cuda.cu:277-316): grad_xyz1 f32 [B, n, 3], fully written; xyz2 gets no gradient. */
int ctb_emd_bwd(const float* xyz1, const float* xyz2, const float* grad_dist, const int32_t* assignment, float* grad_xyz1,
                int B, int n, void* stream);

/* Replaces torch.nn.SyncBatchNorm's statistics kernel -> NCCL all_gather / all_reduce -> combine -> elementwise chain
 * (106 layers = 212 small collectives per step in model_zoo/scanobject/classifier.py) by two kernels per direction whose
 * exchange is a one-shot all-gather over NVLink peer memory: the statistics kernel stores its per-channel partial sums
 * straight into the exchange block of every rank and publishes an epoch flag; the elementwise kernel waits for the flags
 * of all ranks and sums the partials in rank order (identical bits on every rank).
 *   exchange block of one layer and direction, on EVERY rank in peer-mapped (symmetric) memory, zero-initialised:
 *       data f32 [2][world][2*C] followed anywhere by flag u32 [2][world]
 *   peer_data / peer_flag: DEVICE arrays of `world` pointers to that block's data / flag part on each rank;
 *   epoch u32 [1], done u32 [2], scratch (ctb_syncbn_scratch_bytes(C) bytes, 8-byte aligned): ordinary device memory of
 *   this rank, zero-initialised, owned by this layer + direction.
 * Equal B and L on every rank.  x, y, grad_y, grad_x: f32 [B, C, L] contiguous. */
typedef struct ctb_bn_exchange {
  void* const* peer_data;
  void* const* peer_flag;
  uint32_t* epoch;
  uint32_t* done;
  void* scratch;
  int32_t rank;
  int32_t world;
} ctb_bn_exchange;
uint64_t ctb_syncbn_scratch_bytes(int C);
/* training-mode forward: y, save_mean / save_invstd f32 [C] (for the backward), running statistics updated in place
 * (may be NULL) with momentum and the unbiased variance, like nn.SyncBatchNorm. */
int ctb_syncbn_fwd(const float* x, const float* weight, const float* bias, float* y, float* save_mean, float* save_invstd,
                   float* running_mean, float* running_var, const ctb_bn_exchange* exchange, int B, int C, int L, float eps,
                   float momentum, void* stream);
/* backward: grad_x; grad_weight / grad_bias f32 [C] are this rank's LOCAL sums (DDP reduces parameter gradients). */
int ctb_syncbn_bwd(const float* x, const float* grad_y, const float* weight, const float* save_mean, const float* save_invstd,
                   float* grad_x, float* grad_weight, float* grad_bias, const ctb_bn_exchange* exchange, int B, int C, int L,
                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CTB200_H */
