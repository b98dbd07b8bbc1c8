/*
 * devo_b200.h -- C ABI of libdevo_b200.so: the B200 (sm_100a) implementation of the
 * DEVO update-operator hot path (sparse patch correlation, SE3/Sim3 Lie-group ops,
 * projective transform, Gauss-Newton bundle adjustment).
 *
 * This is the drop-in boundary.  Every entry point takes plain device pointers, sizes
 * and a CUDA stream (as void*); no torch types.  The functions correspond one-to-one
 * to what the reference binds through pybind11 (file:line of the reference interface
 * each one replaces is cited).  Tensors are dense row-major ("contiguous") unless a
 * stride is given.  All functions return 0 on success, a negative DEVO_E* code on an
 * argument / capacity error, or a positive cudaError_t; devo_last_error() gives text.
 * All work is enqueued on `stream`; nothing synchronises with the host.
 */
#ifndef DEVO_B200_H
#define DEVO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DEVO_B200_ABI_VERSION 1

/* element types (the reference dispatches half/float/double: correlation_kernel.cu:211,
 * float/double: lietorch/include/dispatch.h:41-42) */
enum { DEVO_F16 = 0, DEVO_BF16 = 1, DEVO_F32 = 2, DEVO_F64 = 3 };
/* Lie groups (devo/lietorch/groups.py:236,252,268,290; dispatch.h:16-31) */
enum { DEVO_SO3 = 1, DEVO_RXSO3 = 2, DEVO_SE3 = 3, DEVO_SIM3 = 4 };
/* error codes */
enum { DEVO_OK = 0, DEVO_EINVAL = -1, DEVO_ECAPACITY = -2, DEVO_EWORKSPACE = -3, DEVO_EUNSUPPORTED = -4 };

int         devo_abi_version(void);
const char* devo_last_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
uint64_t    devo_launch_count(void);
/* device-to-device copy of nbytes done by a KERNEL on `stream` (16-byte vectors when dst, src and nbytes allow).  For the
 * small state copies inside a captured step (what `tensor.copy_()` does in the reference's loop, devo/devo.py:232-239,
 * 523-527): a cudaMemcpyAsync / cudaMemsetAsync there becomes a copy-engine node of the graph, which costs ~4 us of
 * engine hand-over and queues behind whatever host upload is in flight on that engine (tools/e2e_probe.py). */
int         devo_copy_bytes(void* dst, const void* src, size_t nbytes, void* stream);

/* ------------------------------------------------------------------ altcorr (cuda_corr) */
/* cuda_corr.forward  (devo/altcorr/correlation.cpp:57, correlation_kernel.cu:82-136,193-233)
 * fmap1 [B,Np,C,P,P], fmap2 [B,Nf,C,H,W] (dtype), coords [B,E,2,P,P] f32, ii/jj i64[E].
 * out: [B,E,2r+1 (x-off),2r+1 (y-off),P,P] dtype, contiguous -- the value of the
 * reference's (permuted) return tensor, with the bilinear blend fused in. */
int devo_corr_forward(const void* fmap1, const void* fmap2, const float* coords,
                      const int64_t* ii, const int64_t* jj, void* out, int dtype,
                      int B, int Np, int Nf, int C, int H, int W, int E, int P, int radius,
                      void* stream);
/* cuda_corr.backward (correlation.cpp:58, correlation_kernel.cu:139-190,236-286)
 * grad: [B,E,2r+1,2r+1,P,P] f32 contiguous (x-off,y-off order, as returned by forward).
 * fmap1_grad/fmap2_grad (dtype) are fully overwritten (zeroed inside, then accumulated). */
int devo_corr_backward(const void* fmap1, const void* fmap2, const float* coords,
                       const int64_t* ii, const int64_t* jj, const float* grad,
                       void* fmap1_grad, void* fmap2_grad, int dtype,
                       int B, int Np, int Nf, int C, int H, int W, int E, int P, int radius,
                       void* stream);
/* The same backward on PIXEL-MAJOR float32 volumes (training shape: B = 1, P = 3, radius 3, C = 64 or 128): one warp per
 * bounding-box pixel, one coalesced load and one 16-byte vector reduction per lane instead of C scalar atomics per pixel
 * (csrc/corr_bwd_pm.cu).  fmap1 planar [Np,C,3,3]; fmap2_pm [Nf,H,W,C]; grad [E,7,7,3,3] as above; the two gradients are
 * ACCUMULATED into fmap1_grad_pm [Np,9,C] and fmap2_grad_pm [Nf,H,W,C] (the caller zeroes them and converts the layouts). */
int devo_corr_backward_pm(const float* fmap1, const float* fmap2_pm, const float* coords, const int64_t* ii,
                          const int64_t* jj, const float* grad, float* fmap1_grad_pm, float* fmap2_grad_pm,
                          int Np, int Nf, int C, int H, int W, int E, void* stream);
/* cuda_corr.patchify_forward (correlation.cpp:60, correlation_kernel.cu:16-47,288-307)
 * net [B,C,H,W], coords [B,M,2] f32 -> patches [B,M,C,2r+2,2r+2] (fully written, OOB = 0). */
int devo_patchify_forward(const void* net, const float* coords, void* patches, int dtype,
                          int B, int C, int H, int W, int M, int radius, void* stream);
/* cuda_corr.patchify_backward (correlation.cpp:61, correlation_kernel.cu:49-80,309-333)
 * net_grad [B,C,H,W] is fully overwritten. */
int devo_patchify_backward(const void* patch_grad, const float* coords, void* net_grad, int dtype,
                           int B, int C, int H, int W, int M, int radius, void* stream);

/* The tail of Patchifier.forward (devo/enet.py:179-191) as one launch: for patch centres coords [N,M,2] (x, y at feature
 * resolution) the bilinear-mode altcorr.patchify of fmap [N,C,H,W] at radius P/2 (-> gmap_planar [N*M,C,P,P] and / or
 * gmap_pm [N*M,P*P,C], the layout devo_corr_lookup_fused reads), of imap [N,D,H,W] at radius 0 (-> imap_out [N*M,D]) and of
 * the (x, y, disps) grid of coords_grid_with_index at radius P/2 (-> patches [N*M,3,P,P] float; disps [N,H,W] or NULL =
 * ones).  Blend in float32 like the reference (float32 offsets x half windows), outputs rounded to `dtype`.  Any of
 * fmap / imap / patches may be NULL. */
int devo_patch_gather(const void* fmap, const void* imap, const float* disps, const float* coords, void* gmap_planar,
                      void* gmap_pm, void* imap_out, float* patches, int dtype, int N, int C, int D, int H, int W, int M,
                      int P, void* stream);

/* --- B200-native fast path of the lookup: pixel-major (channels-last) pyramid + fused
 * multi-level lookup.  Same arithmetic as devo_corr_forward over every level; replaces
 * pyramidify + 2x altcorr.corr + torch.stack (devo/devo.py:210-217, devo/enet.py:203-216,
 * devo/utils.py:70-79). */
/* average-pool `pool`x`pool` (stride pool) of planar fmap [N,C,H,W] and write it
 * pixel-major [N,H/pool,W/pool,C] (f16 or bf16). */
int devo_pyramid_pack(const void* fmap_planar, void* out_pixel_major, int dtype,
                      int N, int C, int H, int W, int pool, void* stream);
/* both levels of a [1, pool] pyramid from ONE read of the frames: out_level1 [N,H,W,C] and out_pooled [N,H/pool,W/pool,C];
 * bit-identical to two devo_pyramid_pack calls.  pool in {2,4,8}; C % 8 == 0, W % 8 == 0, H and W multiples of pool. */
int devo_pyramid_pack2(const void* fmap_planar, void* out_level1, void* out_pooled, int dtype, int N, int C, int H, int W,
                       int pool, void* stream);
/* repack gmap [Np,C,P,P] -> [Np,P*P,C] */
int devo_gmap_pack(const void* gmap_planar, void* out, int dtype, int Np, int C, int PP, void* stream);
#define DEVO_MAX_LEVELS 4
typedef struct {
  int n_levels;
  const void* level[DEVO_MAX_LEVELS];   /* pixel-major [Nf,H_l,W_l,C] */
  int H[DEVO_MAX_LEVELS], W[DEVO_MAX_LEVELS];
  float scale[DEVO_MAX_LEVELS];         /* level-1 coords are divided by this per level (1, 4, ...) */
} devo_pyramid_t;
/* out [E, 49*P*P*n_levels]: index (((xo*7+yo)*P+i0)*P+j0)*L+l  == torch.stack(corrs,-1).view(1,E,-1)
 * gmap_pm: [Np,P*P,C] pixel-major; coords [E,2,P,P] f32 at level-1 resolution; radius 3, P 3, C%64==0.
 * ii indexes gmap (already reduced modulo the ring buffer by the caller), jj indexes frames. */
int devo_corr_lookup_fused(const void* gmap_pm, const devo_pyramid_t* pyr, const float* coords,
                           const int64_t* ii, const int64_t* jj, void* out, int dtype,
                           int Np, int Nf, int C, int E, void* stream);

/* same with an explicit output row stride ld_out >= 49*P*P*n_levels (elements): lets the caller keep the
 * rows padded to a GEMM-friendly K (882 -> 896) for the first Linear of the corr MLP (enet.py:60). */
int devo_corr_lookup_fused_ld(const void* gmap_pm, const devo_pyramid_t* pyr, const float* coords,
                              const int64_t* ii, const int64_t* jj, void* out, int ld_out, int dtype,
                              int Np, int Nf, int C, int E, void* stream);

/* float32 features on the same tensor-core path (training runs altcorr in float32): split precision.  Every feature is
 * stored as two halves, a = hi + 2^-11 lo (22 significant bits; devo_pyramid_pack_split / devo_gmap_pack_split, float
 * planar in, two pixel-major half buffers out, pooling in float), and the lookup runs three passes of the half kernel
 * (<hi,hi> + 2^-11 (<hi,lo> + <lo,hi>)) with float accumulation into a FLOAT output of the layout above.  Agrees with
 * the float32 kernel (devo_corr_forward) to ~1e-6 relative; the reference kernel it replaces: correlation_kernel.cu:82-136
 * instantiated for float. */
int devo_pyramid_pack_split(const float* fmap_planar, void* out_hi, void* out_lo, int N, int C, int H, int W, int pool,
                            void* stream);
int devo_gmap_pack_split(const float* gmap_planar, void* out_hi, void* out_lo, int Np, int C, int PP, void* stream);
int devo_corr_lookup_fused_split(const void* gmap_hi, const void* gmap_lo, const devo_pyramid_t* pyr_hi,
                                 const devo_pyramid_t* pyr_lo, const float* coords, const int64_t* ii, const int64_t* jj,
                                 float* out, int ld_out, int Np, int Nf, int C, int E, void* stream);

/* ------------------------------------------------------------------ fastba (cuda_ba) */
/* Edge-graph analysis shared by neighbors / BA / segment softmax: edges sorted by
 * (ka, kb, edge index).  All outputs are device arrays; any may be NULL.
 *   perm   i32[E]   edge index at each sorted position
 *   gid    i32[E]   dense id (rank of ka among the sorted unique ka) of each EDGE  (= torch.unique inverse)
 *   gstart i32[E+1] start of each group in `perm`; gstart[ngroups] = E
 *   gkey   i64[E]   unique ka values (first ngroups valid)                          (= torch.unique values)
 *   ngroups i32[1]
 *   ix,jx  i64[E]   previous / next edge of the same ka in kb order, -1 at the ends (= cuda_ba.neighbors)
 * max_ka / max_kb: exclusive upper bounds if known, else -1. */
size_t devo_graph_plan_workspace(int E);
int devo_graph_plan(const int64_t* ka, const int64_t* kb, int E, int64_t max_ka, int64_t max_kb,
                    int32_t* perm, int32_t* gid, int32_t* gstart, int64_t* gkey, int32_t* ngroups,
                    int64_t* ix, int64_t* jx, void* workspace, size_t workspace_bytes, void* stream);
/* cuda_ba.neighbors (devo/fastba/ba.cpp:104-149,154) -- on the GPU, no host round trip */
int devo_neighbors(const int64_t* ii, const int64_t* jj, int64_t* ix, int64_t* jx, int E,
                   void* workspace, size_t workspace_bytes, void* stream);
/* cuda_ba.forward (devo/fastba/ba.cpp:153, ba_cuda.cu:422-540): `iterations` Gauss-Newton
 * steps, IN PLACE on poses [n_poses,7] and patches [n_patches,3,P,P] (f32).
 * status (device i32[1]): 0, or iteration+1 at which the Schur system was not positive
 * definite (the reference raises there: later iterations are skipped), or a DEVO_E* code. */
size_t devo_ba_workspace(int E, int n_free_poses);
int devo_ba_forward(float* poses, float* patches, const float* intrinsics, const float* target,
                    const float* weight, const float* lmbda,
                    const int64_t* ii, const int64_t* jj, const int64_t* kk,
                    int E, int n_poses, int n_patches, int P, int t0, int t1, int iterations,
                    void* workspace, size_t workspace_bytes, int32_t* status, void* stream);
/* same, reusing a devo_graph_plan(kk, jj) the caller already holds (perm/gstart/gkey/ngroups): the update
 * operator needs that plan anyway for neighbors and the patch-wise SoftAgg, so one sort serves all three. */
int devo_ba_forward_planned(float* poses, float* patches, const float* intrinsics, const float* target,
                            const float* weight, const float* lmbda,
                            const int64_t* ii, const int64_t* jj, const int64_t* kk,
                            int E, int n_poses, int n_patches, int P, int t0, int t1, int iterations,
                            const int32_t* perm, const int32_t* gstart, const int64_t* gkey, const int32_t* ngroups,
                            void* workspace, size_t workspace_bytes, int32_t* status, void* stream);

/* The same call with its housekeeping off the critical path (used by the engine; no reference counterpart).
 * devo_ba_prepare: the two clears of a call (status word, 128-byte ticket area of the workspace; one small kernel), on any stream that
 * is made to precede the BA.  devo_ba_forward_prepared: after a prepare for this workspace / status, launches nothing
 * but the iterations (+ the final depth update); the caller vouches that the plan arrays are older than the kernel
 * preceding the call in `stream`, so they are read ahead of the programmatic dependent-launch wait from the first
 * iteration on; `status_or` (may be NULL): device word the call's status is OR-ed into by the last launch. */
int devo_ba_prepare(void* workspace, size_t workspace_bytes, int E, int n_free_poses, int32_t* status, void* stream);
int devo_ba_forward_prepared(float* poses, float* patches, const float* intrinsics, const float* target,
                             const float* weight, const float* lmbda, const int64_t* ii, const int64_t* jj,
                             const int64_t* kk, int E, int n_poses, int n_patches, int P, int t0, int t1,
                             int iterations, const int32_t* perm, const int32_t* gstart, const int64_t* gkey,
                             const int32_t* ngroups, void* workspace, size_t workspace_bytes, int32_t* status,
                             int32_t* status_or, void* stream);
/* Edge-sharded form of cuda_ba.forward for ONE frame graph split over several GPUs by owning patch (the reference has
 * no such path: train.py:90-93 runs one sequence per DDP rank; BASELINE.json north_star / SURVEY 8e ask for it).
 * Per Gauss-Newton iteration, on every rank:
 *   devo_ba_sharded_accumulate  local edges -> sys_out, fp64 [devo_ba_system_doubles(t1-t0)]: this rank's partial of the
 *                               reduced system [S|y] (upper triangle, undamped) + a status word
 *   all-reduce(sum) of sys_out over the ranks  (the single collective; host side, NCCL)
 *   devo_ba_sharded_solve       damping, LDL^T, pose retraction -- identical on every rank (poses are replicated)
 * flags: bit0 = apply the previous iteration's depth update to the local patches first, bit1 = accumulate,
 * bit2 = first call for this edge list (re-plan, reset status).  Finish with flags=1 (depth update only).
 * Same workspace (devo_ba_workspace(E, t1-t0)) and E for all calls of one BA. */
size_t devo_ba_system_doubles(int n_free_poses);
int devo_ba_sharded_accumulate(float* poses, float* patches, const float* intrinsics, const float* target,
                               const float* weight, const float* lmbda,
                               const int64_t* ii, const int64_t* jj, const int64_t* kk,
                               int E, int n_poses, int n_patches, int P, int t0, int t1, int itr, int flags,
                               double* sys_out, void* workspace, size_t workspace_bytes, int32_t* status, void* stream);
int devo_ba_sharded_solve(float* poses, const double* sys, int E, int n_poses, int t0, int t1, int itr,
                          void* workspace, size_t workspace_bytes, int32_t* status, void* stream);
/* The same solve with the all-reduce fused in over NVLink peer memory (no NCCL call): every rank's partial lives in a
 * symmetric, peer-mapped buffer laid out as [2][devo_ba_system_doubles()] doubles + 2 u64 epoch flags (double-buffered by
 * the parity of `epoch`, which the caller increments once per Gauss-Newton iteration, starting at 1, in lockstep on all
 * ranks).  devo_ba_sharded_accumulate must have written sys_out = own buffer + (epoch & 1) * nsys.  peer_ptrs_dev: device
 * array of `world` u64 device pointers to the ranks' buffers (e.g. torch symmetric memory `buffer_ptrs_dev`). */
int devo_ba_sharded_solve_peer(float* poses, const void* peer_ptrs_dev, int world, int rank, uint64_t epoch, int E,
                               int n_poses, int t0, int t1, int itr, void* workspace, size_t workspace_bytes,
                               int32_t* status, void* stream);
/* cuda_ba.reproject (devo/fastba/ba.cpp:155, ba_cuda.cu:368-418,543-575) -> coords [E,2,P,P] f32 */
int devo_reproject(const float* poses, const float* patches, const float* intrinsics,
                   const int64_t* ii, const int64_t* jj, const int64_t* kk, float* coords,
                   int E, int P, void* stream);

/* ------------------------------------------------------------------ projective_ops */
/* Fused forward of devo/projective_ops.py:53-105 `transform` for SE3 f32:
 * Gij = Gj * Gi^-1, X1 = Gij * iproj(patch), proj with 1/clamp(Z,0.1).
 * coords_out [E,P,P,2] if layout==0 (what transform returns), [E,2,P,P] if layout==1
 * (what devo.py:223 / enet.py:342 permute it to).  valid_out [E] (Z_centre>0.2) and
 * Ji,Jj [E,2,6], Jz [E,2,1] may be NULL.  tonly zeroes the rotation (:63-64). */
int devo_transform_forward(const float* poses, const float* patches, const float* intrinsics,
                           const int64_t* ii, const int64_t* jj, const int64_t* kk,
                           float* coords_out, float* valid_out, float* Ji, float* Jj, float* Jz,
                           int E, int P, int layout, int tonly, void* stream);
/* backward of devo_transform_forward for training (projective_ops.transform under autograd, called 6-8x per training
 * iteration, enet.py:341,363-372, and inside ba.BA): g_coords [E,P,P,2] (layout 0) / [E,2,P,P] (layout 1), g_Ji / g_Jj
 * [E,2,6], g_Jz [E,2] -- any may be NULL -- are accumulated (atomicAdd; the caller zeroes) into grad_poses [n_poses,7]
 * (lietorch convention: gradient w.r.t. a LEFT tangent perturbation in slots 0..5, slot 6 untouched) and grad_patches
 * [n_patches,3,P,P]. */
int devo_transform_backward(const float* poses, const float* patches, const float* intrinsics,
                            const int64_t* ii, const int64_t* jj, const int64_t* kk,
                            const float* g_coords, const float* g_Ji, const float* g_Jj, const float* g_Jz,
                            float* grad_poses, float* grad_patches, int E, int P, int layout, int tonly, void* stream);

/* ------------------------------------------------------------------ lietorch_backends */
/* The 19 entry points of devo/lietorch/src/lietorch.cpp:286-316.  `n` = batch (rows).
 * X,Y,Z group elements [n,N]; a,b tangents [n,K]; p,q points [n,3] or [n,4]; gradients
 * w.r.t. group elements are [n,N] with the K-vector in the first K slots, rest 0
 * (lietorch_gpu.cu:41-42,120-123).  dtype in {DEVO_F32, DEVO_F64}. */
int devo_lie_expm(int group, int dtype, const void* a, void* X, int64_t n, void* stream);
int devo_lie_expm_backward(int group, int dtype, const void* grad, const void* a, void* da, int64_t n, void* stream);
int devo_lie_logm(int group, int dtype, const void* X, void* a, int64_t n, void* stream);
int devo_lie_logm_backward(int group, int dtype, const void* grad, const void* X, void* dX, int64_t n, void* stream);
int devo_lie_inv(int group, int dtype, const void* X, void* Y, int64_t n, void* stream);
int devo_lie_inv_backward(int group, int dtype, const void* grad, const void* X, void* dX, int64_t n, void* stream);
int devo_lie_mul(int group, int dtype, const void* X, const void* Y, void* Z, int64_t n, void* stream);
int devo_lie_mul_backward(int group, int dtype, const void* grad, const void* X, const void* Y, void* dX, void* dY, int64_t n, void* stream);
int devo_lie_adj(int group, int dtype, const void* X, const void* a, void* b, int64_t n, void* stream);
int devo_lie_adj_backward(int group, int dtype, const void* grad, const void* X, const void* a, void* dX, void* da, int64_t n, void* stream);
int devo_lie_adjT(int group, int dtype, const void* X, const void* a, void* b, int64_t n, void* stream);
int devo_lie_adjT_backward(int group, int dtype, const void* grad, const void* X, const void* a, void* dX, void* da, int64_t n, void* stream);
int devo_lie_act(int group, int dtype, const void* X, const void* p, void* q, int64_t n, void* stream);
int devo_lie_act_backward(int group, int dtype, const void* grad, const void* X, const void* p, void* dX, void* dp, int64_t n, void* stream);
int devo_lie_act4(int group, int dtype, const void* X, const void* p, void* q, int64_t n, void* stream);
int devo_lie_act4_backward(int group, int dtype, const void* grad, const void* X, const void* p, void* dX, void* dp, int64_t n, void* stream);
int devo_lie_as_matrix(int group, int dtype, const void* X, void* T4x4, int64_t n, void* stream);
int devo_lie_projector(int group, int dtype, const void* X, void* PNxN, int64_t n, void* stream);
int devo_lie_jinv(int group, int dtype, const void* X, const void* a, void* b, int64_t n, void* stream);

/* ------------------------------------------------------------------ segment ops (torch_scatter role) */
/* scatter_softmax / scatter_sum over dim 1 of x [n_rows, dim] with group ids gid i32[n_rows]
 * (devo/blocks.py:40-48).  softmax_out [n_rows,dim]; sum_out [n_groups,dim] (overwritten). */
int devo_segment_softmax_sum(const void* g, const void* f, const int32_t* perm, const int32_t* gstart,
                             const int32_t* ngroups, int max_groups, void* y_out, int dtype,
                             int n_rows, int dim, void* stream);

/* ------------------------------------------------------------------ GRU glue */
/* Fused element-wise chains between the cuBLAS Linear layers of `Update` (devo/enet.py:80-99,
 * devo/blocks.py:15-29).  dtype = the autocast type (DEVO_F16 / DEVO_BF16); rounding points follow
 * torch.autocast's dtype flow.  rows x dim row-major.
 * layernorm modes: 0: LN(T(T(a+b)+c)) -> out32 ; 1: LN(x32) -> out32 (+ out16 copy) ; 2: T(relu(LN(a))) -> out16 */
int devo_glue_layernorm(int mode, int dtype, const void* a, const void* b, const void* c, const float* x32,
                        const float* gamma, const float* beta, float eps, float* out32, void* out16,
                        int rows, int dim, void* stream);
/* out16[e,:] = idx[e] >= 0 ? T(x32[idx[e],:]) : 0   (mask * net[:, ix], enet.py:87-91) */
int devo_glue_gather_mask_cast(int dtype, const float* x32, const int64_t* idx, void* out16, int rows, int dim, void* stream);
/* net32[e,:] += y16[gid ? gid[e] : e, :] ; optional T copy of the sum in out16 */
int devo_glue_residual_add(int dtype, float* net32, const void* y16, const int32_t* gid, void* out16, int rows, int dim, void* stream);
/* out32 = x32 + T(T(sigmoid(gate_pre)) * res)   (GatedResidual, blocks.py:15-29) */
int devo_glue_gated_residual(int dtype, const float* x32, const void* gate_pre, const void* res, float* out32, int64_t total, void* stream);
/* out16 = T(relu ? max(x32,0) : x32) */
int devo_glue_relu_cast(int dtype, const float* x32, void* out16, int64_t total, int relu, void* stream);
/* both heads of Update in one pass (enet.py:68-77): h = T(relu(x32)); delta[rows,2] = T(W[0:2].h + b[0:2]);
 * weight[rows,2] = T(sigmoid(T(W[2:4].h + b[2:4]))); W16 [4,dim], b16 [4] */
int devo_glue_heads(int dtype, const float* x32, const void* W16, const void* b16, void* delta, void* weight,
                    int rows, int dim, void* stream);

/* ------------------------------------------------------------------ fused update operator */
/* The whole of Update.forward (devo/enet.py:80-99: corr MLP, norm, neighbour convolutions c1/c2, the two SoftAgg
 * aggregations, the gated-residual GRU and both heads) as fused tcgen05 kernels: a pair of CTAs (cta_group::2 MMAs,
 * 64 whole rows per CTA) keeps a tile of 128 edges resident (activations in shared memory, accumulators in TMEM)
 * through a whole chain of Linear layers, weights streamed by TMA.  Inference only, autocast rounding points.
 * The recurrent hidden state is FLOAT32 in a TILE layout [ceil(E/128)*2 tiles][96 float4 groups][64 rows][4]
 * (devo_gru_state_floats(E) floats; convert / gather rows with devo_gru_state_gather): the reference's state is float32
 * from the second update on (GatedResidual returns float32 under autocast; devo.py:232-233 concatenates half zeros).
 * Stacked layer order of W [18*384,384] and bias rows 1..18: corr[2], corr[5], c1[0], c1[2], c2[0], c2[2],
 * agg_kk.g, agg_kk.f, agg_kk.h, agg_ij.g, agg_ij.f, agg_ij.h, gru[1].gate[0], gru[1].res[0], gru[1].res[2],
 * gru[3].gate[0], gru[3].res[0], gru[3].res[2].  LayerNorm rows: corr[3], norm, gru[0], gru[2]. */
typedef struct {
  const void* W;          /* [18*384, 384] (out, in) row-major, autocast dtype */
  const void* W0;         /* corr[0] weight [384, corr_ld], input dim zero-padded to corr_ld */
  const void* bias;       /* [19, 384]: row 0 = corr[0], rows 1..18 = the stacked layers */
  const float* ln_gamma;  /* [4, 384] f32 */
  const float* ln_beta;   /* [4, 384] f32 */
  float ln_eps;
  const void* head_W;     /* [4, 384]: d.x d.y w.x w.y */
  const void* head_b;     /* [4] */
} devo_gru_weights_t;
typedef struct {
  int E, dim, corr_ld;
  const void* corr16;     /* [E, corr_ld] correlation features (rows zero-padded) */
  float* state32;         /* hidden state, float32, tile layout: read (unless net16 is given) and overwritten in place */
  const void* net16;      /* optional [E, 384] row-major hidden state in the autocast dtype (the first update of a sequence:
                             half dtype flow); NULL => the input is state32 */
  const void* imap16;     /* [n_patches, 384] context features; inp = imap16[kk] */
  const int64_t* kk;      /* [E] */
  const int64_t* ix;      /* [E] neighbours from devo_graph_plan(kk, jj): previous / next edge or -1 */
  const int64_t* jx;
  const int32_t* perm_kk; const int32_t* gstart_kk; const int32_t* ngroups_kk; const int32_t* gid_kk; int max_groups_kk;
  const int32_t* perm_ij; const int32_t* gstart_ij; const int32_t* ngroups_ij; const int32_t* gid_ij; int max_groups_ij;
  void* net16_out;        /* optional [E, 384] row-major copy of the new hidden state in the autocast dtype, or NULL */
  void* delta;            /* [E, 2] */
  void* weight;           /* [E, 2] */
  /* optional fused BA inputs (devo.py:326-331): target = coords[:, :, 1, 1] + float(delta), weight as f32 */
  const float* coords;    /* [E, 2, 3, 3] reprojected patch coordinates, or NULL */
  float* target32;        /* [E, 2] or NULL */
  float* weight32;        /* [E, 2] or NULL */
  int tile_local;         /* 1: the caller promises that ix[e] and jx[e] are -1 or lie in e's own 64-edge tile (e / 64) for every
                             edge -- true for a patch-major edge list whose patches do not straddle a multiple of 64, e.g. the
                             all-pairs graph of enet.py:300-301 with 8 frames.  The neighbour gathers of c1 / c2 then happen
                             inside the CTA's shared memory and the first three programs run as one launch (the kernel
                             traps if the promise is broken).  0: always correct. */
} devo_gru_io_t;
size_t devo_gru_workspace(int E, int max_groups);
size_t devo_gru_state_floats(int E);
/* dst row e = idx ? (idx[e] >= 0 ? src row idx[e] : 0) : src row e ; layouts: 0 = row-major [rows,384], 1 = tile layout.
 * Packs / unpacks the hidden state and implements net[:, ~m] / torch.cat([net, zeros]) of devo.py:225-239 on the device. */
int devo_gru_state_gather(const float* src, int src_layout, int src_rows, const int64_t* idx, float* dst, int dst_layout,
                          int dst_rows, void* stream);
int devo_gru_update(const devo_gru_weights_t* weights, const devo_gru_io_t* io, int dtype, void* workspace,
                    size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ frame front */
/* Voxel-grid normalisation (utils/voxel_utils.py:6-52; devo/devo.py:419-452): x, y [groups][n_per_group] f32 (y may alias x).
 * mode 0 "std": standardise the NON-ZERO entries of each group with their own mean / std (zeros stay zero), applied only if
 * every group has a non-zero entry (voxel_utils.py:18); mode 1 "rescale": positives / max, negatives / -min (1e-5 when a
 * polarity is absent).  stats_out (optional) [groups][5] = nnz, sum, sumsq, max, min.  Deterministic; two launches. */
size_t devo_voxel_workspace(int groups);
int devo_voxel_normalize(const float* x, float* y, long long n_per_group, int groups, int mode, float* stats_out,
                         void* workspace, size_t workspace_bytes, void* stream);
/* Event stream -> voxel grid (utils/event_utils.py:180-231): xs, ys f32, ts f64 (sorted), ps f32 (0 or -1: negative) of
 * n_events events are ACCUMULATED into grid [bins][H][W] f32 (the caller zeroes) with trilinear weights. */
int devo_events_to_voxel(const float* xs, const float* ys, const double* ts, const float* ps, long long n_events,
                         float* grid, int bins, int H, int W, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEVO_B200_H */
