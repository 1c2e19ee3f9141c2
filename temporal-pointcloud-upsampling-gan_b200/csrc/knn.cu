// knn.cu — K1: exact brute-force k-nearest / fixed-radius neighbours for any D.
//
// Replaces pytorch3d knn_points (gcn_lib/pointnet/gcn.py:16,38; discriminator.py:15,33;
// gcn_lib/interpolation.py:47), frnn.frnn_grid_points in its brute-force form
// (discriminator.py:27; loss.py:256,261) and pointnet2_ops three_nn.
//
// Design (sm_100a, SIMT fp32 — the distance must be the exact sequential, unfused
// sum so indices are bit-exact):
//   * a CTA owns QB = 8 warps x QPW queries of one cloud and streams that cloud's
//     candidates through shared memory in tiles of TJ points;
//   * the tile is stored chunk-major, [D/4][TJ+1] float4, so a warp reading 32
//     consecutive candidates issues conflict-free LDS.128; the +1 float4 of padding
//     makes the transposing tile fill conflict-free too;
//   * lanes own CANDIDATES, not queries: each lane evaluates CPL candidates against
//     the warp's QPW queries (query chunks are warp-uniform broadcast LDS.128), so
//     the admission test against the current k-th distance is one ballot per query
//     and the branchy part (insertion) is warp-uniform — no SIMT divergence;
//   * the running top-K of a query is a WarpList: one (d, idx) per lane, rank ==
//     lane; an insertion is ballot + popc + shfl_up.  Candidates arrive in ascending
//     index order, so inserting "after every element with d <= dc" realises the
//     canonical (d2, idx) order with strict '<' admission;
//   * K > 32 runs extra passes, each collecting the next 32 keys greater than the
//     last key of the previous pass.
#include "common.cuh"
#include "internal.cuh"

namespace tpg {

constexpr int KNN_THREADS = 256;
constexpr int KNN_WARPS = KNN_THREADS / 32;
constexpr int KNN_TJ = 128;  // candidates per shared-memory tile

template <int DV, int QPW, int CPL>
__global__ void __launch_bounds__(KNN_THREADS) knn_warp_kernel(KnnArgs a) {
  extern __shared__ float4 smem4[];
  const int dv = DV > 0 ? DV : (a.D + 3) >> 2;
  const int dp = dv * 4;
  constexpr int TJS = KNN_TJ + 1;
  constexpr int QB = KNN_WARPS * QPW;
  float4* cand = smem4;             // [dv][TJS]
  float4* qs = smem4 + dv * TJS;    // [QB][dv]

  const int b = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * QB;
  const int n1 = a.len1 ? min((int)a.len1[b], a.P1) : a.P1;
  const int n2 = a.len2 ? min((int)a.len2[b], a.P2) : a.P2;
  const float INF = __int_as_float(0x7f800000);
  float r2 = INF;
  if (a.use_radius) {
    float r = a.r_per_cloud ? a.r_per_cloud[b] : a.r;
    r2 = __fmul_rn(r, r);
  }

  // queries of this CTA, zero padded to dp floats ((0-0)^2 adds exactly +0)
  {
    float* qsf = reinterpret_cast<float*>(qs);
    const float* src = a.p1 + (size_t)b * a.P1 * a.D;
    for (int e = tid; e < QB * dp; e += KNN_THREADS) {
      int q = e / dp, d = e - q * dp;
      int qi = q0 + q;
      qsf[e] = (qi < n1 && d < a.D) ? src[(size_t)qi * a.D + d] : 0.0f;
    }
  }
  const bool warp_active = (q0 + warp * QPW) < n1;
  const float* p2b = a.p2 + (size_t)b * a.P2 * a.D;

  float lo_d[QPW];
  int lo_i[QPW];
#pragma unroll
  for (int q = 0; q < QPW; ++q) { lo_d[q] = -1.0f; lo_i[q] = -1; }

  const int npass = (a.K + 31) >> 5;
  for (int pass = 0; pass < npass; ++pass) {
    const int Kp = min(32, a.K - 32 * pass);
    WarpList lst[QPW];
    float tau[QPW];
#pragma unroll
    for (int q = 0; q < QPW; ++q) { lst[q].init(); tau[q] = INF; }

    for (int t0 = 0; t0 < n2; t0 += KNN_TJ) {
      __syncthreads();  // previous tile fully consumed (first trip: queries visible)
      for (int e = tid; e < KNN_TJ * dp; e += KNN_THREADS) {
        int j = e / dp, d = e - j * dp;
        int gj = t0 + j;
        float v = (gj < n2 && d < a.D) ? p2b[(size_t)gj * a.D + d] : 0.0f;
        reinterpret_cast<float*>(&cand[(d >> 2) * TJS + j])[d & 3] = v;
      }
      __syncthreads();
      if (!warp_active) continue;
      const int tile_n = min(KNN_TJ, n2 - t0);
#pragma unroll 1
      for (int s = 0; s < tile_n; s += 32 * CPL) {
        float acc[CPL][QPW];
#pragma unroll
        for (int cp = 0; cp < CPL; ++cp)
#pragma unroll
          for (int q = 0; q < QPW; ++q) acc[cp][q] = 0.0f;
        // DV == 1 (D <= 4) is fully unrolled; wider points loop over float4 chunks with a
        // small unroll so the accumulators, not hoisted loads, own the registers.
#pragma unroll(DV == 1 ? 1 : 2)
        for (int c = 0; c < dv; ++c) {
          float4 cv[CPL];
#pragma unroll
          for (int cp = 0; cp < CPL; ++cp) cv[cp] = cand[c * TJS + s + cp * 32 + lane];
#pragma unroll
          for (int q = 0; q < QPW; ++q) {
            const float4 qv = qs[(warp * QPW + q) * dv + c];
#pragma unroll
            for (int cp = 0; cp < CPL; ++cp) {
              float t = acc[cp][q];
              t = sq_acc(t, qv.x, cv[cp].x);
              t = sq_acc(t, qv.y, cv[cp].y);
              t = sq_acc(t, qv.z, cv[cp].z);
              t = sq_acc(t, qv.w, cv[cp].w);
              acc[cp][q] = t;
            }
          }
        }
#pragma unroll
        for (int cp = 0; cp < CPL; ++cp) {
          const int jbase = t0 + s + cp * 32;  // index of lane 0's candidate
          const int gj = jbase + lane;
          const bool valid = (s + cp * 32 + lane) < tile_n;
#pragma unroll
          for (int q = 0; q < QPW; ++q) {
            const float dq = acc[cp][q];
            bool ok = valid && dq < tau[q] && dq < r2;
            if (pass > 0) ok = ok && (dq > lo_d[q] || (dq == lo_d[q] && gj > lo_i[q]));
            unsigned m = __ballot_sync(FULL, ok);
            while (m) {
              const int l = __ffs(m) - 1;
              m &= m - 1;
              const float dc = __shfl_sync(FULL, dq, l);
              if (dc < tau[q]) {
                lst[q].insert_tail(dc, jbase + l, lane);
                tau[q] = lst[q].kth(Kp);
              }
            }
          }
        }
      }
    }

    // write this pass's slots
#pragma unroll
    for (int q = 0; q < QPW; ++q) {
      const int qi = q0 + warp * QPW + q;
      if (qi < a.P1 && lane < Kp) {
        const size_t o = ((size_t)b * a.P1 + qi) * a.K + pass * 32 + lane;
        const bool found = (qi < n1) && lst[q].i >= 0;
        if (a.out_mode == OUT_THREE) {
          a.dists[o] = found ? sqrtf(lst[q].d) : 0.0f;
          reinterpret_cast<int32_t*>(a.idx)[o] = found ? lst[q].i : 0;
        } else {
          const float padd = a.out_mode == OUT_FRNN ? -1.0f : 0.0f;
          const int64_t padi = a.out_mode == OUT_FRNN ? -1 : 0;
          a.dists[o] = found ? lst[q].d : padd;
          reinterpret_cast<int64_t*>(a.idx)[o] = found ? (int64_t)lst[q].i : padi;
        }
      }
      // lower bound of the next pass = last key of this one (inf when not full)
      lo_d[q] = __shfl_sync(FULL, lst[q].d, 31);
      lo_i[q] = __shfl_sync(FULL, lst[q].i, 31);
    }
  }
}

template <int DV, int QPW, int CPL>
static int launch_knn(const KnnArgs& a, cudaStream_t st) {
  const int dv = DV > 0 ? DV : (a.D + 3) / 4;
  const size_t smem = sizeof(float4) * ((size_t)dv * (KNN_TJ + 1) + (size_t)KNN_WARPS * QPW * dv);
  auto kern = knn_warp_kernel<DV, QPW, CPL>;
  if (smem > 48 * 1024) {
    TPG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  dim3 grid(ceil_div(a.P1, KNN_WARPS * QPW), a.B);
  kern<<<grid, KNN_THREADS, smem, st>>>(a);
  TPG_CHECK_LAUNCH("knn_warp_kernel");
  return TPG_OK;
}

int knn_dispatch(const KnnArgs& a, cudaStream_t st) {
  if (a.B == 0 || a.P1 == 0 || a.K == 0) return TPG_OK;
  const int dv = (a.D + 3) / 4;
  // queries per warp: 4 amortises the candidate loads best, but the insertions of a warp's
  // queries are serial, so small problems (few CTAs) spread over more warps instead
  const long long queries = (long long)a.B * a.P1;
  const int two_waves = 2 * num_sms();
  const int qpw = queries / (KNN_WARPS * 4) >= two_waves ? 4 : (queries / (KNN_WARPS * 2) >= two_waves ? 2 : 1);
  if (dv == 1) {
    if (qpw == 4) return launch_knn<1, 4, 2>(a, st);
    if (qpw == 2) return launch_knn<1, 2, 2>(a, st);
    return launch_knn<1, 1, 2>(a, st);
  }
  if (qpw == 4) return launch_knn<0, 4, 2>(a, st);
  if (qpw == 2) return launch_knn<0, 2, 2>(a, st);
  return launch_knn<0, 1, 2>(a, st);
}

}  // namespace tpg

using namespace tpg;

TPG_API size_t tpg_knn_workspace_bytes(int B, int P1, int P2, int D, int K) {
  KnnArgs a{reinterpret_cast<const float*>(16), reinterpret_cast<const float*>(16), nullptr, nullptr, B, P1, P2, D, K,
            0.f, nullptr, 0, nullptr, nullptr, OUT_KNN};
  if (knn_feat_eligible(a)) return knn_feat_workspace_bytes(B, P1, P2, D);
  return grid_eligible(D, P2, K) ? grid_workspace_bytes(B, P2) : 0;
}

TPG_API size_t tpg_knn_fallback_count_offset(int B) { return knn_feat_fallback_count_offset(B); }

namespace tpg {
// flag[0] &= (a == b) bytewise; nbytes % 16 == 0, 16-byte aligned
__global__ void bytes_equal_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, size_t n16,
                                   int32_t* __restrict__ flag) {
  bool diff = false;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n16; e += (size_t)gridDim.x * blockDim.x) {
    const uint4 x = __ldg(a + e), y = __ldg(b + e);
    diff |= (x.x != y.x) | (x.y != y.y) | (x.z != y.z) | (x.w != y.w);
  }
  if (diff) *flag = 0;
}
// flag != 0: out[r, :K] = cached[r, :Kc][:K]
__global__ void knn_take_prefix_kernel(const int32_t* __restrict__ flag, const float* __restrict__ cd,
                                       const int64_t* __restrict__ ci, int Kc, float* __restrict__ od,
                                       int64_t* __restrict__ oi, int K, long long rows) {
  if (*flag == 0) return;
  const long long total = rows * K;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / K;
    const int k = (int)(e - r * K);
    od[e] = cd[r * Kc + k];
    oi[e] = ci[r * Kc + k];
  }
}
}  // namespace tpg

TPG_API int tpg_bytes_equal_and(const void* a, const void* b, size_t nbytes, int32_t* flag, tpg_stream_t stream) {
  TPG_REQUIRE(flag, TPG_EINVAL, "bytes_equal: null flag");
  if (nbytes == 0) return TPG_OK;
  TPG_REQUIRE(a && b, TPG_EINVAL, "bytes_equal: null pointer");
  TPG_REQUIRE(nbytes % 16 == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0,
              TPG_EUNSUPPORTED, "bytes_equal: needs 16-byte aligned buffers of a multiple of 16 bytes");
  const size_t n16 = nbytes / 16;
  const unsigned blocks = (unsigned)min((n16 + 255) / 256, (size_t)num_sms() * 8);
  bytes_equal_kernel<<<blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<const uint4*>(a),
                                                           reinterpret_cast<const uint4*>(b), n16, flag);
  TPG_CHECK_LAUNCH("bytes_equal_kernel");
  return TPG_OK;
}

TPG_API int tpg_knn_take_prefix(const int32_t* flag, const float* cached_dists, const int64_t* cached_idx, int Kc,
                                float* dists, int64_t* idx, int K, long long rows, tpg_stream_t stream) {
  TPG_REQUIRE(K >= 1 && Kc >= K && rows >= 0, TPG_EINVAL, "knn_take_prefix: bad sizes K=%d Kc=%d", K, Kc);
  if (rows == 0) return TPG_OK;
  TPG_REQUIRE(flag && cached_dists && cached_idx && dists && idx, TPG_EINVAL, "knn_take_prefix: null pointer");
  const long long total = rows * K;
  const unsigned blocks = (unsigned)min((total + 255) / 256, (long long)num_sms() * 8);
  knn_take_prefix_kernel<<<blocks, 256, 0, as_stream(stream)>>>(flag, cached_dists, cached_idx, Kc, dists, idx, K, rows);
  TPG_CHECK_LAUNCH("knn_take_prefix_kernel");
  return TPG_OK;
}

TPG_API int tpg_knn_cond_f32(const float* p1, const float* p2, const int64_t* lengths1, const int64_t* lengths2, int B,
                             int P1, int P2, int D, int K, float* dists, int64_t* idx, void* workspace,
                             size_t workspace_bytes, const int32_t* skip_flag, tpg_stream_t stream);

TPG_API int tpg_knn_f32(const float* p1, const float* p2, const int64_t* lengths1,
                        const int64_t* lengths2, int B, int P1, int P2, int D, int K,
                        float* dists, int64_t* idx, void* workspace, size_t workspace_bytes,
                        tpg_stream_t stream) {
  return tpg_knn_cond_f32(p1, p2, lengths1, lengths2, B, P1, P2, D, K, dists, idx, workspace, workspace_bytes, nullptr,
                          stream);
}

TPG_API int tpg_knn_cond_f32(const float* p1, const float* p2, const int64_t* lengths1, const int64_t* lengths2, int B,
                             int P1, int P2, int D, int K, float* dists, int64_t* idx, void* workspace,
                             size_t workspace_bytes, const int32_t* skip_flag, tpg_stream_t stream) {
  TPG_REQUIRE(B >= 0 && P1 >= 0 && P2 >= 0, TPG_EINVAL, "knn: negative size");
  TPG_REQUIRE(D >= 1 && D <= 256, TPG_EUNSUPPORTED, "knn: D=%d outside [1,256]", D);
  TPG_REQUIRE(K >= 1 && K <= 1024, TPG_EUNSUPPORTED, "knn: K=%d outside [1,1024]", K);
  TPG_REQUIRE(B <= 65535, TPG_EUNSUPPORTED, "knn: B=%d > 65535", B);
  if (B == 0 || P1 == 0) return TPG_OK;
  TPG_REQUIRE(p1 && (p2 || P2 == 0) && dists && idx, TPG_EINVAL, "knn: null pointer");
  KnnArgs a{p1, p2, lengths1, lengths2, B, P1, P2, D, K, 0.f, nullptr, 0, dists, idx, OUT_KNN};
  a.skip = skip_flag;  // honoured by the tensor-core path only; the other paths simply compute
  if (P2 > 0 && knn_feat_eligible(a)) return knn_feat_dispatch(a, workspace, workspace_bytes, as_stream(stream));
  if (grid_eligible(D, P2, K)) return grid_knn_dispatch(a, workspace, workspace_bytes, as_stream(stream));
  return knn_dispatch(a, as_stream(stream));
}

TPG_API size_t tpg_frnn_workspace_bytes(int B, int P1, int P2, int D, int K) {
  (void)P1;
  return grid_eligible(D, P2, K) ? grid_workspace_bytes(B, P2) : 256;
}

TPG_API int tpg_frnn_f32(const float* p1, const float* p2, const int64_t* lengths1,
                         const int64_t* lengths2, int B, int P1, int P2, int D, int K, float r,
                         const float* r_per_cloud, float* dists, int64_t* idx, void* workspace,
                         size_t workspace_bytes, tpg_stream_t stream) {
  TPG_REQUIRE(B >= 0 && P1 >= 0 && P2 >= 0, TPG_EINVAL, "frnn: negative size");
  TPG_REQUIRE(D == 2 || D == 3, TPG_EUNSUPPORTED, "frnn: D=%d (only 2 or 3, like upstream)", D);
  TPG_REQUIRE(K >= 1 && K <= 1024, TPG_EUNSUPPORTED, "frnn: K=%d outside [1,1024]", K);
  TPG_REQUIRE(B <= 65535, TPG_EUNSUPPORTED, "frnn: B=%d > 65535", B);
  if (B == 0 || P1 == 0) return TPG_OK;
  TPG_REQUIRE(p1 && (p2 || P2 == 0) && dists && idx, TPG_EINVAL, "frnn: null pointer");
  KnnArgs a{p1, p2, lengths1, lengths2, B, P1, P2, D, K, r, r_per_cloud, 1, dists, idx, OUT_FRNN};
  if (grid_eligible(D, P2, K)) return grid_knn_dispatch(a, workspace, workspace_bytes, as_stream(stream));
  return knn_dispatch(a, as_stream(stream));
}

TPG_API size_t tpg_three_nn_workspace_bytes(int B, int n, int m) {
  (void)n;
  return grid_eligible(3, m, 3) ? grid_workspace_bytes(B, m) : 0;
}

TPG_API int tpg_three_nn_f32(const float* unknown, const float* known, int B, int n, int m,
                             float* dist, int32_t* idx, void* workspace, size_t workspace_bytes,
                             tpg_stream_t stream) {
  TPG_REQUIRE(B >= 0 && n >= 0 && m >= 0, TPG_EINVAL, "three_nn: negative size");
  TPG_REQUIRE(B <= 65535, TPG_EUNSUPPORTED, "three_nn: B=%d > 65535", B);
  if (B == 0 || n == 0) return TPG_OK;
  TPG_REQUIRE(unknown && (known || m == 0) && dist && idx, TPG_EINVAL, "three_nn: null pointer");
  KnnArgs a{unknown, known, nullptr, nullptr, B, n, m, 3, 3, 0.f, nullptr, 0, dist, idx, OUT_THREE};
  if (grid_eligible(3, m, 3)) return grid_knn_dispatch(a, workspace, workspace_bytes, as_stream(stream));
  return knn_dispatch(a, as_stream(stream));
}
