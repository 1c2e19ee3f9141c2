/*
 * tpugan_b200.h — C ABI of libtpugan_b200.so
 *
 * B200-native (sm_100a) point-neighbourhood kernels behind the call surfaces
 * that TPU-GAN (zijieli-Jlee/Temporal-Pointcloud-Upsampling-GAN) imports from
 * its un-vendored native dependencies.  Every entry point below names the
 * reference call site (file:line under the reference tree) whose native op it
 * replaces.
 *
 * Conventions
 *   - plain pointers + sizes, no torch types.  All pointers are DEVICE pointers
 *     unless a parameter is documented as a host scalar.
 *   - the caller owns every buffer (inputs, outputs, workspace); the library
 *     never allocates or frees device memory.  No data state survives a call; the
 *     only process-global state is the two scheduling hints of tpg_set_option()
 *     (they never change a result) and the launch counter.
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*);
 *     no call synchronises the device or reads anything back to the host.
 *   - return value: TPG_OK (0) or a negative TPG_E* code; the message of the
 *     last failure on the calling thread is returned by tpg_last_error().
 *   - tensors are dense row-major ("contiguous") float32 / int32 / int64.
 *   - `lengths*` pointers may be NULL (all clouds full length).
 *
 * Canonical numerics
 *   squared distance  d2 = sum_d (a_d - b_d)^2, accumulated sequentially for
 *   d = 0..D-1 in fp32 with SEPARATE multiply and add (no FMA contraction);
 *   neighbour order is (d2, index) lexicographic ascending, so ties go to the
 *   lowest index.
 */
#ifndef TPUGAN_B200_H_
#define TPUGAN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TPG_ABI_VERSION 9

typedef void* tpg_stream_t; /* cudaStream_t */

enum {
  TPG_OK = 0,
  TPG_EINVAL = -1,       /* bad shape / argument                         */
  TPG_EUNSUPPORTED = -2, /* argument combination outside the built range */
  TPG_ECUDA = -3,        /* a CUDA runtime call or launch failed         */
  TPG_EWORKSPACE = -4    /* workspace pointer NULL or too small          */
};

/* reduce ops of tpg_group_reduce_* */
enum { TPG_REDUCE_MAX = 0, TPG_REDUCE_SUM = 1, TPG_REDUCE_MIN = 2 };

/* which Chamfer directions to evaluate */
enum { TPG_CHAMFER_FWD = 1, TPG_CHAMFER_REV = 2, TPG_CHAMFER_BOTH = 3 };

/* ---- library ---------------------------------------------------------- */
int tpg_abi_version(void);
const char* tpg_last_error(void);
/* number of kernel launches issued by this library on the calling process
 * since load (bench.py's "gpu_launches" claim is read from here). */
uint64_t tpg_launch_count(void);
/* Scheduling hints; never change a result.  Returns TPG_OK or TPG_EINVAL (unknown name / value).
 *   "fps.sms_per_cloud"  1 | 2 | 4 | 8 (default 8; initial value from env TPG_FPS_CLUSTER):
 *       size of the thread-block cluster that shares one cloud of 2049..65536 points in
 *       tpg_fps_f32 / tpg_fps_sampling_f32.  8 gives the shortest call; 1 holds 8x fewer SMs
 *       for a ~1.4x longer call — the better choice when calls overlap with other work
 *       (multi-stream capture of a train step).
 *   "fps.exclusive_sm"   0 | 1 (default 0; env TPG_FPS_EXCLUSIVE): a one-CTA-per-cloud FPS launch requests
 *       200 KB of shared memory so that no smem-using CTA of a concurrent kernel shares its SM — the
 *       latency-bound rounds then keep their solo speed (2.6x slower when co-resident). */
int tpg_set_option(const char* name, long value);

/* ---- K1/K2: k nearest neighbours ---------------------------------------
 * replaces pytorch3d.ops.knn_points(p1, p2, lengths1, lengths2, K,
 *   return_sorted=True) — reference call sites gcn_lib/pointnet/gcn.py:16,38,
 *   discriminator.py:15,33, gcn_lib/interpolation.py:47,
 *   gcn_lib/graph_utils.py:71.
 * p1 [B,P1,D], p2 [B,P2,D] -> dists [B,P1,K] (squared), idx [B,P1,K] int64.
 * Slots beyond min(K, lengths2[b]) and rows beyond lengths1[b] hold 0 / 0.
 * 1 <= D <= 256, 1 <= K <= 1024.
 * Feature-space searches (D = 32/64, K <= 24, P2 >= 1024, hit-mask workspace B*P1*P2/8 bytes <= 1 GiB) run on the
 * tensor cores (tcgen05: centred operands split into two bf16 terms, three-product bf16 contraction with fp32
 * accumulation as candidate search + exact fp32 re-rank on the original rows) and 3-D searches over clouds of
 * >= 2048 points (K <= 32) walk a uniform grid; both return results identical to the
 * brute-force path and need tpg_knn_workspace_bytes() bytes of workspace; for every
 * other shape that function returns 0 and workspace may be NULL.            */
size_t tpg_knn_workspace_bytes(int B, int P1, int P2, int D, int K);
/* Diagnostics of the tensor-core path: byte offset, inside the workspace of a finished call, of an
 * int32 counter = number of queries of that call that were answered by the exact SIMT fallback
 * (candidate margin not provably complete).  Results are identical either way. */
size_t tpg_knn_fallback_count_offset(int B);
int tpg_knn_f32(const float* p1, const float* p2, const int64_t* lengths1,
                const int64_t* lengths2, int B, int P1, int P2, int D, int K,
                float* dists, int64_t* idx, void* workspace, size_t workspace_bytes,
                tpg_stream_t stream);

/* Memoised search (exact): the reference's IDGCNLayer runs three knn_points calls on bit-identical
 * feature maps (gcn_lib/pointnet/gcn.py:258-265: K = 9, 20, 20) — each on a fresh .contiguous() copy,
 * so only a content comparison can tell.  Everything stays on the device and capturable:
 *   tpg_bytes_equal_and   flag[0] is cleared when the two buffers differ (caller sets it to 1 first;
 *                         several comparisons may share one flag)
 *   tpg_knn_cond_f32      tpg_knn_f32, except that on the tensor-core path every kernel of the call
 *                         returns at once when *skip_flag != 0 (skip_flag may be NULL; other paths
 *                         ignore it and compute)
 *   tpg_knn_take_prefix   when *flag != 0: dists/idx[r, :K] = cached_dists/idx[r, :K] of a cached
 *                         [rows, Kc] result, Kc >= K (the K nearest are a prefix of the Kc nearest
 *                         in the canonical (d2, index) order)
 * tpugan_b200.functional.knn_memo drives them; it never keeps anything across train steps.       */
int tpg_bytes_equal_and(const void* a, const void* b, size_t nbytes, int32_t* flag, tpg_stream_t stream);
int tpg_knn_cond_f32(const float* p1, const float* p2, const int64_t* lengths1,
                     const int64_t* lengths2, int B, int P1, int P2, int D, int K,
                     float* dists, int64_t* idx, void* workspace, size_t workspace_bytes,
                     const int32_t* skip_flag, tpg_stream_t stream);
int tpg_knn_take_prefix(const int32_t* flag, const float* cached_dists, const int64_t* cached_idx, int Kc,
                        float* dists, int64_t* idx, int K, long long rows, tpg_stream_t stream);

/* ---- K3: fixed-radius nearest neighbours --------------------------------
 * replaces frnn.frnn_grid_points(points1, points2, lengths1, lengths2, K, r,
 *   return_sorted=True) — discriminator.py:27, loss.py:105,142,229,256,261,
 *   gcn_lib/interpolation.py:20,33, gcn_lib/graph_utils.py:46,
 *   gcn_lib/pointnet/gcn.py:30.
 * The K nearest points with d2 < r*r (strict), ordered by (d2, idx); unused
 * slots hold dist -1 / idx -1.  D = 3 (2 also accepted).
 * r_per_cloud: device [B] or NULL, in which case the host scalar `r` is used.
 * workspace: tpg_frnn_workspace_bytes() bytes (uniform grid: bounding box, cell
 * size and counting sort are computed on the device, no host synchronisation). */
size_t tpg_frnn_workspace_bytes(int B, int P1, int P2, int D, int K);
int tpg_frnn_f32(const float* p1, const float* p2, const int64_t* lengths1,
                 const int64_t* lengths2, int B, int P1, int P2, int D, int K,
                 float r, const float* r_per_cloud, float* dists, int64_t* idx,
                 void* workspace, size_t workspace_bytes, tpg_stream_t stream);

/* ---- K4: ball query ------------------------------------------------------
 * replaces pointnet2_ops.pointnet2_utils.ball_query(radius, nsample, xyz,
 *   new_xyz) used inside QueryAndGroup — discriminator.py:190.
 * xyz [B,N,3], new_xyz [B,M,3] -> idx [B,M,nsample] int32: the `nsample`
 * LOWEST indices with d2 < radius^2, slots beyond the hit count repeat the
 * first hit, no hit -> all 0.  Clouds of >= 16384 points are searched through the
 * uniform grid (workspace: tpg_ball_query_workspace_bytes(), else 0 / NULL).    */
size_t tpg_ball_query_workspace_bytes(int B, int N, int M, int nsample);
int tpg_ball_query_f32(const float* xyz, const float* new_xyz, int B, int N,
                       int M, float radius, int nsample, int32_t* idx,
                       void* workspace, size_t workspace_bytes,
                       tpg_stream_t stream);

/* ---- K5: farthest point sampling -----------------------------------------
 * (a) replaces pointnet2_utils.furthest_point_sample(xyz, npoint) —
 *     discriminator.py:114.  Starts at index 0, skips points with
 *     x^2+y^2+z^2 <= 1e-3, ties -> lowest index.  xyz [B,N,3] -> idx
 *     [B,npoint] int32.
 * (b) replaces sampling.farthest_point_sampling(pts, k, initial_idx) —
 *     sampling.py:50-106 (numba loop sampling.py:36-44).  pts [B,N,D] (D<=3),
 *     start [B] int64 -> idx [B,k] int64 and, when dist_rows != NULL, the
 *     [B,k,N] rows of squared distances the reference returns.             */
size_t tpg_fps_workspace_bytes(int B, int N); /* 0 for N <= 65536 */
int tpg_fps_f32(const float* xyz, int B, int N, int npoint, int32_t* idx,
                void* workspace, size_t workspace_bytes, tpg_stream_t stream);
int tpg_fps_start_f32(const float* pts, int B, int N, int D, int k,
                      const int64_t* start, int64_t* idx, float* dist_rows,
                      void* workspace, size_t workspace_bytes,
                      tpg_stream_t stream);

/* ---- K6: grouping / gather ------------------------------------------------
 * replaces pointnet2_utils.grouping_operation(features, idx) —
 *   gcn_lib/pointnet/gcn.py:207,261, discriminator.py:270,273 — and
 *   pointnet2_utils.gather_operation(features, idx) — discriminator.py:132
 *   (the k == 1 case).
 * f [B,C,N], idx [B,M,k] int32 -> out [B,C,M,k] = f[b,c,idx[b,m,j]].
 * center != NULL ([B,C,M]) fuses the "- centre" of gcn.py:209 /
 * discriminator.py:271: out = f[b,c,idx] - center[b,c,m].
 * Backward: grad_f [B,C,N] = sum over (m,j) with idx == n of grad_out, in
 * ascending (m,j) order, through an inverse index (CSR) built once per idx
 * tensor by tpg_inverse_index_build and shared by every grouping call that
 * reuses that idx.
 *   seg_offsets [B,N+1] int32, seg_items [B,L] int32 (L = M*k; item = m*k+j).
 *   Contract of tpg_group_bwd_f32: the items of a segment are ASCENDING (what
 *   tpg_inverse_index_build produces) and lie in [0,L); the shared-memory-staged
 *   kernel walks them with a cursor through chunks of the grad_out rows, so an
 *   unsorted list would drop contributions. */
int tpg_group_fwd_f32(const float* f, const int32_t* idx, const float* center,
                      int B, int C, int N, int M, int k, float* out,
                      tpg_stream_t stream);
/* Rows that do not fit shared memory (N > 24576 source points, C >= 16): with tpg_group_fwd_workspace_bytes() bytes
 * of workspace the features are first copied point-major and gathered as contiguous channel rows
 * (same result; 3-4x faster than 4-byte gathers from L2).  workspace NULL / 0 -> tpg_group_fwd_f32. */
size_t tpg_group_fwd_workspace_bytes(int B, int C, int N, int M, int k);
int tpg_group_fwd_ws_f32(const float* f, const int32_t* idx, const float* center,
                         int B, int C, int N, int M, int k, float* out, void* workspace,
                         size_t workspace_bytes, tpg_stream_t stream);
size_t tpg_inverse_index_workspace_bytes(int B, int N, int L);
int tpg_inverse_index_build(const int32_t* idx, int B, int N, int L,
                            int32_t* seg_offsets, int32_t* seg_items,
                            void* workspace, size_t workspace_bytes,
                            tpg_stream_t stream);
int tpg_group_bwd_f32(const float* grad_out, const int32_t* seg_offsets,
                      const int32_t* seg_items, int B, int C, int N, int L,
                      float* grad_f, tpg_stream_t stream);

/* Long rows (L = M*k > 65536 positions, N <= 8192 source points): the shared-memory-staged backward keeps
 * 16-bit row positions, so the row is processed in S segments of L/S positions, each continuing the
 * sequential sums of the one before (same summation order as one pass: results stay bit-identical).
 *   tpg_group_bwd_segments   S for this shape (0 or 1: use tpg_group_bwd_f32 with the plain inverse index)
 *   tpg_group_bwd_segmented_f32  seg_offsets [B*S, N+1], seg_items [B*S, L/S] = the inverse index of idx
 *                            viewed as [B*S, L/S] (tpg_inverse_index_build(idx, B*S, N, L/S, ...)).        */
int tpg_group_bwd_segments(int B, int C, int N, int L);
int tpg_group_bwd_segmented_f32(const float* grad_out, const int32_t* seg_offsets,
                                const int32_t* seg_items, int B, int C, int N, int L, int S,
                                float* grad_f, tpg_stream_t stream);

/* ---- K7: fused gather + reduce over the k neighbours ---------------------
 * replaces grouping_operation followed by torch.max(dim=-1) —
 *   gcn_lib/pointnet/gcn.py:261-263 — without materialising [B,C,M,k].
 * out [B,C,M]; arg [B,C,M] int32 (slot j of the first extremum; may be NULL,
 * ignored for SUM).  Backward routes grad_out[b,c,m] to f[b,c,idx[b,m,arg]]
 * (MAX/MIN) or to every neighbour (SUM) through the same CSR.              */
int tpg_group_reduce_fwd_f32(const float* f, const int32_t* idx, int B, int C,
                             int N, int M, int k, int op, float* out,
                             int32_t* arg, tpg_stream_t stream);
/* Rows that do not fit shared memory (N > 12800, C >= 8): with tpg_group_reduce_workspace_bytes() bytes of
 * workspace the features are copied point-major and every neighbour is one coalesced channel-row read.
 * `w` != NULL computes the weighted sum of three_interpolate (k = 3) instead of `op`.  Same results. */
size_t tpg_group_reduce_workspace_bytes(int B, int C, int N);
int tpg_group_reduce_fwd_ws_f32(const float* f, const int32_t* idx, const float* w, int B, int C,
                                int N, int M, int k, int op, float* out, int32_t* arg,
                                void* workspace, size_t workspace_bytes, tpg_stream_t stream);
int tpg_group_reduce_bwd_f32(const float* grad_out, const int32_t* arg,
                             const int32_t* seg_offsets,
                             const int32_t* seg_items, int B, int C, int N,
                             int M, int k, int op, float* grad_f,
                             tpg_stream_t stream);

/* ---- K11: conv-input assembly around the grouping ---------------------------
 * replaces, in one pass and without the per-tensor [B,C_p,M,k] intermediates and the torch.cat copy,
 *   pointnet2_utils.QueryAndGroup.forward — discriminator.py:190:
 *       cat([xyz[idx] - new_xyz, features[idx]], dim=1)
 *   FlowEmbedding.forward — discriminator.py:270-277:
 *       cat([pos2[idx] - pos1, feat2[idx], feat1.view(B,-1,N,1).repeat(1,1,1,k)], dim=1)
 * out [B, sum_p C_p, M, k]; part p fills the channel range after the parts before it:
 *   TPG_PART_GATHER     out[b,c,m,j] = src[b,c,idx[b,m,j]] (- center[b,c,m] if center != NULL); src [B,C,N]
 *   TPG_PART_BROADCAST  out[b,c,m,j] = src[b,c,m];  src [B,C,M], center must be NULL, N ignored
 * `parts` is a HOST array (1..TPG_ASSEMBLE_MAX_PARTS entries) of device pointers; idx [B,M,k] int32.
 * Values are exact copies / single fp32 subtractions: identical to the unfused composition.           */
#define TPG_ASSEMBLE_MAX_PARTS 4
enum { TPG_PART_GATHER = 0, TPG_PART_BROADCAST = 1 };
typedef struct tpg_assemble_part {
  const float* src;
  const float* center;
  int C;
  int N;
  int mode;
} tpg_assemble_part;
int tpg_group_assemble_f32(const tpg_assemble_part* parts, int nparts, const int32_t* idx,
                           int B, int M, int k, float* out, tpg_stream_t stream);

/* ---- K12: EdgeConv pre-activation after the algebraic restructure -----------
 * replaces the k-expanded front half of EdgeConv.forward — gcn_lib/pointnet/gcn.py:206-211:
 *       feat = grouping_operation(feat, idx); edge = feat - center
 *       feat = node_affine(feat) + edge_affine(edge)          (1x1 conv + LeakyReLU each, no norm layer)
 * With the two 1x1 convolutions applied per NODE — p = LeakyReLU(W_n f + b_n), q = W_e f + b_e, both [B,C,N] —
 *       out[b,c,m,j] = p[b,c,i] + LeakyReLU(q[b,c,i] - center[b,c,m]),  i = idx[b,m,j],  center = q_m - b_e
 * (W(f_j - f_i) = W f_j - W f_i: k times fewer convolution columns, no [B,C,N,k] intermediates; rounding
 * differs from the reference's order of operations by a few ulp, tests bound it by 1e-5 relative).
 * Backward: tpg_edge_affine_bwd_f32 writes g2 = grad_out * LeakyReLU'(q[i] - center) ([B,C,M,k]) and, if
 * grad_center != NULL, grad_center[b,c,m] = -sum_j g2 (sequential in j); grad_p and grad_q are the ordinary
 * grouping backwards of grad_out and g2 (tpg_group_bwd_f32 over the shared inverse index).            */
int tpg_edge_affine_fwd_f32(const float* p, const float* q, const float* center, const int32_t* idx,
                            float slope, int B, int C, int N, int M, int k, float* out,
                            tpg_stream_t stream);
int tpg_edge_affine_bwd_f32(const float* grad_out, const float* q, const float* center,
                            const int32_t* idx, float slope, int B, int C, int N, int M, int k,
                            float* g2, float* grad_center, tpg_stream_t stream);

/* ---- K8: three_nn / three_interpolate -------------------------------------
 * replaces pointnet2_utils.three_nn(unknown, known) and
 *   three_interpolate(features, idx, weight) (north_star surface; no call
 *   site in the reference tree).
 * unknown [B,n,3], known [B,m,3] -> dist [B,n,3] = sqrt(d2), idx [B,n,3] int32.
 * f [B,c,m], idx, w [B,n,3] -> out [B,c,n] = sum_k w_k f[idx_k].           */
size_t tpg_three_nn_workspace_bytes(int B, int n, int m); /* > 0: grid search (m >= 2048) */
int tpg_three_nn_f32(const float* unknown, const float* known, int B, int n,
                     int m, float* dist, int32_t* idx, void* workspace,
                     size_t workspace_bytes, tpg_stream_t stream);
int tpg_three_interpolate_fwd_f32(const float* f, const int32_t* idx,
                                  const float* w, int B, int c, int m, int n,
                                  float* out, tpg_stream_t stream);
int tpg_three_interpolate_bwd_f32(const float* grad_out, const float* w,
                                  const int32_t* seg_offsets,
                                  const int32_t* seg_items, int B, int c, int m,
                                  int n, float* grad_f, tpg_stream_t stream);

/* ---- K9: Chamfer distance ---------------------------------------------------
 * replaces chamferdist.ChamferDistance.forward (two knn_points(K=1) searches +
 *   reductions + the pytorch3d knn backward) — loss.py:125,176,224,280.
 * src [B,P1,D], tgt [B,P2,D].  directions: TPG_CHAMFER_*.
 * fwd writes per-point nearest squared distance / index for each evaluated
 * direction (d_src,i_src: src->tgt [B,P1]; d_tgt,i_tgt: tgt->src [B,P2]) and
 * the per-cloud sums sum_src [B], sum_tgt [B] (fixed summation order).
 * bwd: grad_src/grad_tgt (either may be NULL) from per-cloud upstream
 * gradients g_src [B], g_tgt [B] (d loss / d sum_src[b], d sum_tgt[b]):
 *   grad_src[i] = 2 g_src (s_i - t_nn(i)) + sum_{j: nn(j)=i} 2 g_tgt (s_i - t_j)
 * the scattered part summed in ascending j through a CSR over i_tgt.
 * workspace: tpg_chamfer_fwd_workspace_bytes() (uniform-grid search for 3-D clouds
 * of >= 2048 points; 0 otherwise) / tpg_chamfer_bwd_workspace_bytes().       */
size_t tpg_chamfer_fwd_workspace_bytes(int B, int P1, int P2, int D);
int tpg_chamfer_fwd_f32(const float* src, const float* tgt,
                        const int64_t* lengths_src, const int64_t* lengths_tgt,
                        int B, int P1, int P2, int D, int directions,
                        float* d_src, int32_t* i_src, float* d_tgt,
                        int32_t* i_tgt, float* sum_src, float* sum_tgt,
                        void* workspace, size_t workspace_bytes,
                        tpg_stream_t stream);
size_t tpg_chamfer_bwd_workspace_bytes(int B, int P1, int P2);
int tpg_chamfer_bwd_f32(const float* src, const float* tgt,
                        const int64_t* lengths_src, const int64_t* lengths_tgt,
                        const int32_t* i_src, const int32_t* i_tgt,
                        const float* g_src, const float* g_tgt, int B, int P1,
                        int P2, int D, int directions, float* grad_src,
                        float* grad_tgt, void* workspace,
                        size_t workspace_bytes, tpg_stream_t stream);

/* ---- K10: SPH cubic-kernel field interpolation ----------------------------
 * replaces gcn_lib.cubic_interpolation(query_pos, field, pos, cutoff) —
 *   gcn_lib/interpolation.py:103-123 (graph :16-80, l2dist :11-14, kernel
 *   :92-100) — batched over the S samples that train_step_final.py:54-65
 *   loops over in Python.
 * query [S,Q,3], field [S,P,F], pos [S,P,3] -> out [S,Q,F].  F <= 16.       */
size_t tpg_cubic_interp_workspace_bytes(int S, int Q, int P);
int tpg_cubic_interp_f32(const float* query, const float* field,
                         const float* pos, int S, int Q, int P, int F,
                         float cutoff, float* out, void* workspace,
                         size_t workspace_bytes, tpg_stream_t stream);

/* ---- point-major gathers ---------------------------------------------------
 * replaces pytorch3d.ops.knn_gather(x, idx) (north_star surface) and the
 *   advanced-indexing row gather index_points(points, idx) —
 *   discriminator.py:43-60, loss.py:10-27.
 * x [B,N,U], idx [B,L] int64 -> out [B,L,U] = x[b, idx[b,l], :].  An index of
 * -1 (index_points on FRNN output, loss.py:273-275) wraps to row N-1 exactly
 * as Python negative indexing does.                                        */
int tpg_gather_rows_f32(const float* x, const int64_t* idx, int B, int N, int U,
                        int L, float* out, tpg_stream_t stream);
/* backward of tpg_gather_rows_f32 (= pytorch3d knn_gather backward): grad_x [B,N,U] = sum of
 * grad_out[b,l,:] over the positions l with idx[b,l] == n, in ascending l (deterministic, no
 * atomics), through the inverse index (tpg_inverse_index_build) of the int32 copy of idx with
 * negative entries clamped to 0; `idx` (the original int64 tensor, may be NULL) lets the kernel skip
 * positions whose index is negative (FRNN padding).                                          */
int tpg_gather_rows_bwd_f32(const float* grad_out, const int64_t* idx,
                            const int32_t* seg_offsets, const int32_t* seg_items, int B,
                            int N, int U, int L, float* grad_x, tpg_stream_t stream);

/* ---- backward of the distances of tpg_knn_f32 / tpg_frnn_f32 -----------------------------------
 * replaces the autograd backward of pytorch3d.ops.knn_points / frnn.frnn_grid_points (never
 * taken on the reference's train step: no caller consumes `dists` — gcn_lib/pointnet/gcn.py:91,258,
 * discriminator.py:33 — but part of the drop-in surface, and what chamferdist is built on).
 * grad_p1[b,i,:] = sum_k 2 g[b,i,k] (p1[b,i] - p2[b,idx[b,i,k]]); grad_p2 = the scattered negative,
 * summed in ascending (i,k) through the inverse index of idx (keys clamped to >= 0, L = P1*K).
 * Slots k >= min(K, lengths2[b]), rows i >= lengths1[b] and negative indices carry no gradient.
 * Either gradient pointer may be NULL.                                                        */
int tpg_knn_bwd_f32(const float* p1, const float* p2, const int64_t* idx,
                    const float* grad_dists, const int64_t* lengths1, const int64_t* lengths2,
                    const int32_t* seg_offsets, const int32_t* seg_items, int B, int P1, int P2,
                    int D, int K, float* grad_p1, float* grad_p2, tpg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TPUGAN_B200_H_ */
