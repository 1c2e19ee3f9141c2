// internal.cuh — cross-file internals of libtpugan_b200.so (not part of the ABI).
#pragma once
#include <atomic>
#include "common.cuh"

namespace tpg {

enum { OUT_KNN = 0, OUT_FRNN = 1, OUT_THREE = 2 };

struct KnnArgs {
  const float* p1;
  const float* p2;
  const int64_t* len1;
  const int64_t* len2;
  int B, P1, P2, D, K;
  float r;
  const float* r_per_cloud;
  int use_radius;
  float* dists;
  void* idx;
  int out_mode;
  const int32_t* skip = nullptr;  // device flag: nonzero -> the tensor-core path returns without writing (memoised search)
};

// knn.cu
int knn_dispatch(const KnnArgs& a, cudaStream_t st);

// knn_feat.cu — tcgen05 path for D in {32,64}, K <= 24 (exact results, tensor-core candidate search)
bool knn_feat_eligible(const KnnArgs& a);
size_t knn_feat_workspace_bytes(int B, int P1, int P2, int D);
size_t knn_feat_fallback_count_offset(int B);
int knn_feat_dispatch(const KnnArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t st);

// grid.cu — uniform-grid search for 3-D clouds of >= 2048 points, K <= 32 (results identical to knn_dispatch)
bool grid_eligible(int D, int P2, int K);
size_t grid_workspace_bytes(int B, int P);
int grid_knn_dispatch(const KnnArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t st);
int grid_ball_query(const float* xyz, const float* new_xyz, int B, int N, int M, float radius, int nsample,
                    int32_t* idx, void* workspace, size_t workspace_bytes, cudaStream_t st);
size_t grid_chamfer_workspace_bytes(int B, int P1, int P2);
int grid_chamfer_nn(const float* src, const float* tgt, const int64_t* ls, const int64_t* lt, int B, int P1, int P2,
                    int directions, float* d_src, int32_t* i_src, float* d_tgt, int32_t* i_tgt, void* workspace,
                    size_t workspace_bytes, int* handled, cudaStream_t st);

// group.cu — inverse index (CSR) of an int32 index tensor idx [B,L] with keys in
// [0,N): seg_offsets [B,N+1], seg_items [B,L] (ascending positions per key).
// item_len (device [B] int64 or null) limits the positions of cloud b to [0,len).
size_t csr_workspace_bytes(int B, int N, int L);
int build_csr(const int32_t* idx, const int64_t* item_len, int B, int N, int L, int32_t* seg_offsets,
              int32_t* seg_items, void* workspace, size_t workspace_bytes, cudaStream_t st);

// group_bwd_staged.cu: grouping backward that streams grad_out rows through shared memory (TMA bulk copies)
bool group_bwd_staged_eligible(const float* go, const int32_t* items, int B, int C, int N, int L);
int group_bwd_staged(const float* go, const int32_t* off, const int32_t* items, int B, int C, int N, int L, float* gf,
                     int force_tcg, cudaStream_t st);
int group_bwd_staged_segments(int B, int C, int N, int L);
int group_bwd_staged_segmented(const float* go, const int32_t* off, const int32_t* items, int B, int C, int N, int L,
                               int S, float* gf, cudaStream_t st);
// fps.cu: CTAs (SMs) per cloud for 2048 < N <= 65536 (tpg_set_option "fps.sms_per_cloud")
std::atomic<int>& fps_cluster_option();
std::atomic<int>& fps_exclusive_option();
}  // namespace tpg
