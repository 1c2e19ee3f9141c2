// cubic_interp.cu — K10: SPH cubic-kernel field interpolation, batched.
//
// Replaces gcn_lib.cubic_interpolation (gcn_lib/interpolation.py:103-123) and the
// per-frame x per-sample Python loop around it (train_step_final.py:54-65):
//   1. FRNN(K=32, r=cutoff) neighbour lists of every query            (:19-42)
//      (the second FRNN over the compacted candidates returns the same lists —
//      compaction preserves index order — so it is not repeated);
//   2. in-range mask of candidates + "some query has no neighbour" flag per sample
//      (:26-31, :44);
//   3. one warp per query: lane k owns neighbour slot k, evaluates the reference's
//      expanded-form distance (:11-14) and piecewise cubic weight (:92-100) and the
//      warp reduces sum(w f) and sum(w); when the flag is set, queries with fewer
//      than 32 neighbours add the 4 nearest in-range candidates as extra edges
//      (:46-60, duplicates of in-range neighbours, zero weight outside the cutoff);
//      out = sum(w f) / (sum(w) + 1e-6)                               (:119-122).
#include "common.cuh"
#include "internal.cuh"

namespace tpg {

constexpr int CI_K = 32;
constexpr int CI_THREADS = 256;

__global__ void cubic_mark_kernel(const int64_t* __restrict__ ni, int S, int Q, int P, unsigned char* mark,
                                  int* any_empty) {
  const long long total = (long long)S * Q * CI_K;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int s = (int)(e / ((long long)Q * CI_K));
    const int64_t j = ni[e];
    if (j >= 0) mark[(size_t)s * P + j] = 1;
    else if ((e % CI_K) == 0) any_empty[s] = 1;
  }
}

// reference l2dist(pos_src, pos_dst): sum_c (s^2 + q^2 - (2 q) s), clamp, sqrt
__device__ __forceinline__ float l2dist_ref(const float* s, float qx, float qy, float qz) {
  const float q[3] = {qx, qy, qz};
  float acc = 0.0f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float a = __fmul_rn(s[c], s[c]);
    const float bq = __fmul_rn(q[c], q[c]);
    const float sum = __fadd_rn(a, bq);
    const float e = __fmul_rn(__fmul_rn(2.0f, q[c]), s[c]);
    acc = __fadd_rn(acc, __fsub_rn(sum, e));
  }
  if (acc < 1e-8f) acc = 0.0f;
  return __fsqrt_rn(acc);
}

__device__ __forceinline__ float bicubic_ref(float r, float cutoff, float coeff) {
  const float q = __fdiv_rn(r, cutoff);
  float ker = 0.0f;
  if (q >= 0.0f && q <= 0.5f) {
    const float q2 = __fmul_rn(q, q);
    const float q3 = __fmul_rn(q2, q);
    ker = __fadd_rn(__fmul_rn(6.0f, __fsub_rn(q3, q2)), 1.0f);
  } else if (q > 0.5f && q <= 1.0f) {
    const float u = __fsub_rn(1.0f, q);
    ker = __fmul_rn(2.0f, __fmul_rn(__fmul_rn(u, u), u));
  }
  return __fmul_rn(ker, coeff);
}

__global__ void __launch_bounds__(CI_THREADS) cubic_interp_kernel(
    const float* __restrict__ query, const float* __restrict__ field, const float* __restrict__ pos,
    const int64_t* __restrict__ ni, const unsigned char* __restrict__ mark, const int* __restrict__ any_empty,
    int S, int Q, int P, int F, float cutoff, float coeff, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long wq = (long long)blockIdx.x * (CI_THREADS / 32) + (threadIdx.x >> 5);
  if (wq >= (long long)S * Q) return;
  const int s = (int)(wq / Q);
  const float* qp = query + (size_t)wq * 3;
  const float qx = qp[0], qy = qp[1], qz = qp[2];
  const float* pp = pos + (size_t)s * P * 3;
  const float* ff = field + (size_t)s * P * F;

  const int64_t j = ni[(size_t)wq * CI_K + lane];
  const bool valid = j >= 0;
  const int cnt = __popc(__ballot_sync(FULL, valid));
  float w = 0.0f;
  float acc[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c] = 0.0f;
  if (valid) {
    w = bicubic_ref(l2dist_ref(pp + (size_t)j * 3, qx, qy, qz), cutoff, coeff);
#pragma unroll
    for (int c = 0; c < 16; ++c)
      if (c < F) acc[c] = __fmul_rn(ff[(size_t)j * F + c], w);
  }

  if (any_empty[s] && cnt < CI_K) {
    // 4 nearest among in-range candidates, canonical (d2, idx) order
    WarpList lst;
    lst.init();
    float tau = __int_as_float(0x7f800000);
    int first = -1;
    for (int j0 = 0; j0 < P; j0 += 32) {
      const int jj = j0 + lane;
      const bool in = jj < P && mark[(size_t)s * P + jj];
      float d = 0.0f;
      if (in) d = sqdist3(qx, qy, qz, pp[(size_t)jj * 3], pp[(size_t)jj * 3 + 1], pp[(size_t)jj * 3 + 2]);
      const unsigned min_ = __ballot_sync(FULL, in);
      if (first < 0 && min_) first = j0 + __ffs(min_) - 1;
      unsigned m = __ballot_sync(FULL, in && d < tau);
      while (m) {
        const int l = __ffs(m) - 1;
        m &= m - 1;
        const float dc = __shfl_sync(FULL, d, l);
        if (dc < tau) {
          lst.insert_tail(dc, j0 + l, lane);
          tau = lst.kth(4);
        }
      }
    }
    if (lane < 4) {
      const int pj = lst.i >= 0 ? lst.i : first;
      if (pj >= 0) {
        const float wp = bicubic_ref(l2dist_ref(pp + (size_t)pj * 3, qx, qy, qz), cutoff, coeff);
        w = __fadd_rn(w, wp);
#pragma unroll
        for (int c = 0; c < 16; ++c)
          if (c < F) acc[c] = __fadd_rn(acc[c], __fmul_rn(ff[(size_t)pj * F + c], wp));
      }
    }
  }

#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    w += __shfl_xor_sync(FULL, w, o);
#pragma unroll
    for (int c = 0; c < 16; ++c)
      if (c < F) acc[c] += __shfl_xor_sync(FULL, acc[c], o);
  }
  const float denom = __fadd_rn(w, 1e-6f);
#pragma unroll
  for (int c = 0; c < 16; ++c)
    if (c < F && lane == c) out[(size_t)wq * F + c] = __fdiv_rn(acc[c], denom);
}

struct CubicWs {
  float* nd;
  int64_t* ni;
  unsigned char* mark;
  int* any_empty;
  char* grid;
  size_t grid_bytes;
  size_t total;
};

static CubicWs carve(void* base, int S, int Q, int P) {
  CubicWs w;
  char* p = reinterpret_cast<char*>(base);
  size_t o = 0;
  w.ni = reinterpret_cast<int64_t*>(p + o); o += align_up(sizeof(int64_t) * (size_t)S * Q * CI_K, 256);
  w.nd = reinterpret_cast<float*>(p + o);   o += align_up(sizeof(float) * (size_t)S * Q * CI_K, 256);
  w.mark = reinterpret_cast<unsigned char*>(p + o); o += align_up((size_t)S * P, 256);
  w.any_empty = reinterpret_cast<int*>(p + o); o += align_up(sizeof(int) * (size_t)S, 256);
  w.grid = p + o;
  w.grid_bytes = grid_eligible(3, P, CI_K) ? grid_workspace_bytes(S, P) : 0;
  o += align_up(w.grid_bytes, 256);
  w.total = o;
  return w;
}

}  // namespace tpg

using namespace tpg;

TPG_API size_t tpg_cubic_interp_workspace_bytes(int S, int Q, int P) { return carve(nullptr, S, Q, P).total; }

TPG_API int tpg_cubic_interp_f32(const float* query, const float* field, const float* pos, int S, int Q, int P,
                                 int F, float cutoff, float* out, void* workspace, size_t workspace_bytes,
                                 tpg_stream_t stream) {
  TPG_REQUIRE(S >= 0 && Q >= 0 && P >= 0, TPG_EINVAL, "cubic_interp: negative size");
  TPG_REQUIRE(F >= 1 && F <= 16, TPG_EUNSUPPORTED, "cubic_interp: F=%d outside [1,16]", F);
  TPG_REQUIRE(S <= 65535, TPG_EUNSUPPORTED, "cubic_interp: S > 65535");
  if (S == 0 || Q == 0) return TPG_OK;
  TPG_REQUIRE(P >= 1, TPG_EINVAL, "cubic_interp: empty candidate cloud");
  TPG_REQUIRE(query && field && pos && out, TPG_EINVAL, "cubic_interp: null pointer");
  TPG_REQUIRE(workspace && workspace_bytes >= tpg_cubic_interp_workspace_bytes(S, Q, P), TPG_EWORKSPACE,
              "cubic_interp: workspace too small");
  cudaStream_t st = as_stream(stream);
  CubicWs w = carve(workspace, S, Q, P);
  KnnArgs ka{query, pos, nullptr, nullptr, S, Q, P, 3, CI_K, cutoff, nullptr, 1, w.nd, w.ni, OUT_FRNN};
  int rc = w.grid_bytes ? grid_knn_dispatch(ka, w.grid, w.grid_bytes, st) : knn_dispatch(ka, st);
  if (rc) return rc;
  TPG_CUDA(cudaMemsetAsync(w.mark, 0, (size_t)S * P, st));
  TPG_CUDA(cudaMemsetAsync(w.any_empty, 0, sizeof(int) * (size_t)S, st));
  const long long total = (long long)S * Q * CI_K;
  const unsigned blocks = (unsigned)min((total + 255) / 256, (long long)num_sms() * 32);
  cubic_mark_kernel<<<blocks, 256, 0, st>>>(w.ni, S, Q, P, w.mark, w.any_empty);
  TPG_CHECK_LAUNCH("cubic_mark_kernel");
  const float coeff = (float)(8.0 / (3.14159265358979323846 * (double)cutoff * (double)cutoff * (double)cutoff));
  const long long warps = (long long)S * Q;
  cubic_interp_kernel<<<(unsigned)((warps + 7) / 8), CI_THREADS, 0, st>>>(query, field, pos, w.ni, w.mark,
                                                                          w.any_empty, S, Q, P, F, cutoff, coeff, out);
  TPG_CHECK_LAUNCH("cubic_interp_kernel");
  return TPG_OK;
}
