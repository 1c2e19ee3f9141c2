/*
 * tpg_oracle.c — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the neighbourhood primitives that TPU-GAN calls
 * through pytorch3d / frnn / pointnet2_ops / chamferdist / dgl.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this; the product (libtpugan_b200.so) never does.
 *
 * PARITY STATUS: "parity unpinned" against upstream binaries — the reference
 * tree vendors none of the native packages, has no tests and no golden vectors
 * (SURVEY.md §4, §8c).  What pins this file instead:
 *   - the reference's own call sites (cited per function below);
 *   - live runs of the reference's importable Python (sampling.py, index_points,
 *     l2dist, bicubic_kernel, cubic_interpolation over a functional dgl stub) —
 *     tests/golden/make_golden.py, fixtures under tests/golden/;
 *   - independent float64 NumPy brute force and scipy.spatial.cKDTree
 *     (tests/test_oracle.py).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math [-fopenmp] -shared -fPIC
 * (-ffp-contract=off is REQUIRED: the canonical distance is a sequential fp32
 * sum with separate multiply and add).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
ORC_API void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* canonical squared distance: ((d0*d0 + d1*d1) + d2*d2) + ...  — pytorch3d's
 * CPU loop (knn_cpu.cpp, upstream) as called from gcn_lib/pointnet/gcn.py:16. */
static inline float sqdist(const float* a, const float* b, int D) {
  float acc = 0.0f;
  for (int d = 0; d < D; ++d) {
    float diff = a[d] - b[d];
    float sq = diff * diff;
    acc = acc + sq;
  }
  return acc;
}

/* ------------------------------------------------------------------------
 * kNN / FRNN  (reference: gcn_lib/pointnet/gcn.py:13-45, discriminator.py:13-40,
 * loss.py:253-263, gcn_lib/interpolation.py:19-42)
 *   r2 == NULL : knn_points semantics, pad (0, 0)
 *   r2 != NULL : frnn_grid_points semantics, keep d2 < r2[b], pad (-1, -1)
 * selection order: (d2, index) ascending; strict '<' admission.
 * ---------------------------------------------------------------------- */
ORC_API void orc_knn(const float* p1, const float* p2, const int64_t* len1,
                     const int64_t* len2, int B, int P1, int P2, int D, int K,
                     const float* r2, float* dists, int64_t* idx) {
  const float padd = r2 ? -1.0f : 0.0f;
  const int64_t padi = r2 ? -1 : 0;
#pragma omp parallel for collapse(2) schedule(dynamic, 64)
  for (int b = 0; b < B; ++b) {
    for (int i = 0; i < P1; ++i) {
      float* od = dists + ((size_t)b * P1 + i) * K;
      int64_t* oi = idx + ((size_t)b * P1 + i) * K;
      for (int s = 0; s < K; ++s) {
        od[s] = padd;
        oi[s] = padi;
      }
      int n1 = len1 ? (int)len1[b] : P1;
      int n2 = len2 ? (int)len2[b] : P2;
      if (i >= n1) continue; /* rows beyond lengths1 keep the pad value */
      const float* q = p1 + ((size_t)b * P1 + i) * D;
      int cnt = 0;
      for (int j = 0; j < n2; ++j) {
        float d = sqdist(q, p2 + ((size_t)b * P2 + j) * D, D);
        if (r2 && !(d < r2[b])) continue;
        if (cnt == K && !(d < od[K - 1])) continue;
        /* insert after every element with dist <= d (later index loses ties) */
        int pos = cnt < K ? cnt : K - 1;
        while (pos > 0 && od[pos - 1] > d) {
          od[pos] = od[pos - 1];
          oi[pos] = oi[pos - 1];
          --pos;
        }
        od[pos] = d;
        oi[pos] = j;
        if (cnt < K) ++cnt;
      }
      for (int s = cnt; s < K; ++s) {
        od[s] = padd;
        oi[s] = padi;
      }
    }
  }
}

/* ------------------------------------------------------------------------
 * ball_query (pointnet2_ops query_ball_point kernel semantics, upstream; used by
 * QueryAndGroup at discriminator.py:190).  d2 = dx*dx + dy*dy + dz*dz.
 * ---------------------------------------------------------------------- */
ORC_API void orc_ball_query(const float* xyz, const float* new_xyz, int B, int N,
                            int M, float radius, int nsample, int32_t* idx) {
  const float r2 = radius * radius;
#pragma omp parallel for collapse(2) schedule(dynamic, 64)
  for (int b = 0; b < B; ++b) {
    for (int m = 0; m < M; ++m) {
      int32_t* o = idx + ((size_t)b * M + m) * nsample;
      for (int s = 0; s < nsample; ++s) o[s] = 0;
      const float* c = new_xyz + ((size_t)b * M + m) * 3;
      int cnt = 0;
      for (int j = 0; j < N && cnt < nsample; ++j) {
        float d = sqdist(c, xyz + ((size_t)b * N + j) * 3, 3);
        if (d < r2) {
          if (cnt == 0)
            for (int s = 0; s < nsample; ++s) o[s] = j;
          o[cnt++] = j;
        }
      }
    }
  }
}

/* ------------------------------------------------------------------------
 * FPS, pointnet2 mode (furthest_point_sampling_kernel semantics, upstream;
 * discriminator.py:114): start 0, min-dist init 1e10, points with
 * x*x+y*y+z*z <= 1e-3 are skipped, argmax ties -> lowest index, best init -1
 * with index 0.
 * ---------------------------------------------------------------------- */
ORC_API void orc_fps(const float* xyz, int B, int N, int npoint, int32_t* idx) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < B; ++b) {
    const float* p = xyz + (size_t)b * N * 3;
    int32_t* o = idx + (size_t)b * npoint;
    float* temp = (float*)malloc(sizeof(float) * (size_t)(N > 0 ? N : 1));
    for (int j = 0; j < N; ++j) temp[j] = 1e10f;
    int old = 0;
    if (npoint > 0) o[0] = 0;
    for (int s = 1; s < npoint; ++s) {
      int besti = 0;
      float best = -1.0f;
      const float* a = p + (size_t)old * 3;
      for (int j = 0; j < N; ++j) {
        const float* c = p + (size_t)j * 3;
        float mag = (c[0] * c[0] + c[1] * c[1]) + c[2] * c[2];
        if (mag <= 1e-3f) continue;
        float d = sqdist(c, a, 3);
        float d2 = d < temp[j] ? d : temp[j];
        temp[j] = d2;
        if (d2 > best) {
          best = d2;
          besti = j;
        }
      }
      old = besti;
      o[s] = besti;
    }
    free(temp);
  }
}

/* ------------------------------------------------------------------------
 * FPS, sampling.py mode (sampling.py:36-44, 50-106): explicit start, no skip,
 * np.argmax first-max, int64 indices, optional [k,N] distance rows.
 * ---------------------------------------------------------------------- */
ORC_API void orc_fps_start(const float* pts, int B, int N, int D, int k,
                           const int64_t* start, int64_t* idx, float* rows) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < B; ++b) {
    const float* p = pts + (size_t)b * N * D;
    int64_t* o = idx + (size_t)b * k;
    float* mind = (float*)malloc(sizeof(float) * (size_t)(N > 0 ? N : 1));
    int64_t cur = start[b];
    for (int s = 0; s < k; ++s) {
      if (s > 0) {
        int64_t besti = 0;
        float best = mind[0];
        for (int j = 1; j < N; ++j)
          if (mind[j] > best) {
            best = mind[j];
            besti = j;
          }
        cur = besti;
      }
      o[s] = cur;
      const float* a = p + (size_t)cur * D;
      for (int j = 0; j < N; ++j) {
        float d = sqdist(a, p + (size_t)j * D, D);
        if (rows) rows[((size_t)b * k + s) * N + j] = d;
        if (s == 0)
          mind[j] = d;
        else
          mind[j] = mind[j] < d ? mind[j] : d;
      }
    }
    free(mind);
  }
}

/* ------------------------------------------------------------------------
 * grouping / gather (group_points_kernel, gather_points_kernel upstream;
 * gcn_lib/pointnet/gcn.py:207,261; discriminator.py:132,270,273)
 * ---------------------------------------------------------------------- */
ORC_API void orc_group_fwd(const float* f, const int32_t* idx, const float* center,
                           int B, int C, int N, int M, int k, float* out) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c) {
      const float* row = f + ((size_t)b * C + c) * N;
      float* o = out + ((size_t)b * C + c) * M * k;
      const int32_t* ib = idx + (size_t)b * M * k;
      for (int m = 0; m < M; ++m) {
        float ctr = center ? center[((size_t)b * C + c) * M + m] : 0.0f;
        for (int j = 0; j < k; ++j) {
          float v = row[ib[(size_t)m * k + j]];
          o[(size_t)m * k + j] = center ? v - ctr : v;
        }
      }
    }
}

/* backward: sum in ascending (m,j) order */
ORC_API void orc_group_bwd(const float* grad_out, const int32_t* idx, int B, int C,
                           int N, int M, int k, float* grad_f) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c) {
      float* g = grad_f + ((size_t)b * C + c) * N;
      for (int n = 0; n < N; ++n) g[n] = 0.0f;
      const float* go = grad_out + ((size_t)b * C + c) * M * k;
      const int32_t* ib = idx + (size_t)b * M * k;
      for (size_t l = 0; l < (size_t)M * k; ++l) g[ib[l]] = g[ib[l]] + go[l];
    }
}

/* fused gather + reduce over k (gcn.py:261-263: grouping then torch.max, first
 * maximum wins like torch.max on CPU) */
ORC_API void orc_group_reduce_fwd(const float* f, const int32_t* idx, int B, int C,
                                  int N, int M, int k, int op, float* out,
                                  int32_t* arg) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c) {
      const float* row = f + ((size_t)b * C + c) * N;
      const int32_t* ib = idx + (size_t)b * M * k;
      for (int m = 0; m < M; ++m) {
        float acc = row[ib[(size_t)m * k]];
        int32_t a = 0;
        for (int j = 1; j < k; ++j) {
          float v = row[ib[(size_t)m * k + j]];
          if (op == 0) {
            if (v > acc) { acc = v; a = j; }
          } else if (op == 2) {
            if (v < acc) { acc = v; a = j; }
          } else {
            acc = acc + v;
          }
        }
        out[((size_t)b * C + c) * M + m] = acc;
        if (arg) arg[((size_t)b * C + c) * M + m] = a;
      }
    }
}

ORC_API void orc_group_reduce_bwd(const float* grad_out, const int32_t* idx,
                                  const int32_t* arg, int B, int C, int N, int M,
                                  int k, int op, float* grad_f) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c) {
      float* g = grad_f + ((size_t)b * C + c) * N;
      for (int n = 0; n < N; ++n) g[n] = 0.0f;
      const int32_t* ib = idx + (size_t)b * M * k;
      for (int m = 0; m < M; ++m) {
        float go = grad_out[((size_t)b * C + c) * M + m];
        if (op == 1) {
          for (int j = 0; j < k; ++j) {
            int n = ib[(size_t)m * k + j];
            g[n] = g[n] + go;
          }
        } else {
          int n = ib[(size_t)m * k + arg[((size_t)b * C + c) * M + m]];
          g[n] = g[n] + go;
        }
      }
    }
}

/* ------------------------------------------------------------------------
 * three_nn / three_interpolate (pointnet2_ops, upstream; north_star surface)
 * Stated deviation: with fewer than 3 known points (m < 3) the unused slots are
 * dist 0 / idx 0 here (and in the CUDA kernel), where upstream's kernel leaves
 * its 1e40 initial value (-> sqrt = inf) / idx 0.  Upstream never calls it that
 * way (feature propagation needs >= 3 known points); finite padding keeps the
 * 1 / (dist + eps) weights of the usual caller finite.
 * ---------------------------------------------------------------------- */
ORC_API void orc_three_nn(const float* unknown, const float* known, int B, int n,
                          int m, float* dist, int32_t* idx) {
#pragma omp parallel for collapse(2) schedule(dynamic, 64)
  for (int b = 0; b < B; ++b)
    for (int i = 0; i < n; ++i) {
      float bd[3] = {0.f, 0.f, 0.f};
      int32_t bi[3] = {0, 0, 0};
      int cnt = 0;
      const float* q = unknown + ((size_t)b * n + i) * 3;
      for (int j = 0; j < m; ++j) {
        float d = sqdist(q, known + ((size_t)b * m + j) * 3, 3);
        if (cnt == 3 && !(d < bd[2])) continue;
        int pos = cnt < 3 ? cnt : 2;
        while (pos > 0 && bd[pos - 1] > d) {
          bd[pos] = bd[pos - 1];
          bi[pos] = bi[pos - 1];
          --pos;
        }
        bd[pos] = d;
        bi[pos] = j;
        if (cnt < 3) ++cnt;
      }
      for (int s = 0; s < 3; ++s) {
        dist[((size_t)b * n + i) * 3 + s] = sqrtf(bd[s]);
        idx[((size_t)b * n + i) * 3 + s] = bi[s];
      }
    }
}

ORC_API void orc_three_interpolate_fwd(const float* f, const int32_t* idx,
                                       const float* w, int B, int c, int m, int n,
                                       float* out) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b)
    for (int ch = 0; ch < c; ++ch) {
      const float* row = f + ((size_t)b * c + ch) * m;
      for (int i = 0; i < n; ++i) {
        const int32_t* ii = idx + ((size_t)b * n + i) * 3;
        const float* ww = w + ((size_t)b * n + i) * 3;
        float acc = ww[0] * row[ii[0]];
        acc = acc + ww[1] * row[ii[1]];
        acc = acc + ww[2] * row[ii[2]];
        out[((size_t)b * c + ch) * n + i] = acc;
      }
    }
}

ORC_API void orc_three_interpolate_bwd(const float* grad_out, const int32_t* idx,
                                       const float* w, int B, int c, int m, int n,
                                       float* grad_f) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b)
    for (int ch = 0; ch < c; ++ch) {
      float* g = grad_f + ((size_t)b * c + ch) * m;
      for (int j = 0; j < m; ++j) g[j] = 0.0f;
      for (int i = 0; i < n; ++i)
        for (int s = 0; s < 3; ++s) {
          size_t l = ((size_t)b * n + i) * 3 + s;
          g[idx[l]] = g[idx[l]] + grad_out[((size_t)b * c + ch) * n + i] * w[l];
        }
    }
}

/* ------------------------------------------------------------------------
 * Chamfer (chamferdist.ChamferDistance = 2x knn_points(K=1) + sums, upstream;
 * loss.py:176-181).  Per-point nearest (d2, idx) and per-cloud sums
 * (accumulated in double, rounded once).
 * ---------------------------------------------------------------------- */
static void nn1(const float* a, const float* bpts, int Pa, int Pb, int D, float* d,
                int32_t* ix) {
  for (int i = 0; i < Pa; ++i) {
    float best = 0.0f;
    int32_t bi = 0;
    for (int j = 0; j < Pb; ++j) {
      float v = sqdist(a + (size_t)i * D, bpts + (size_t)j * D, D);
      if (j == 0 || v < best) {
        best = v;
        bi = j;
      }
    }
    d[i] = best;
    ix[i] = bi;
  }
}

ORC_API void orc_chamfer_fwd(const float* src, const float* tgt, const int64_t* ls,
                             const int64_t* lt, int B, int P1, int P2, int D,
                             int directions, float* d_src, int32_t* i_src,
                             float* d_tgt, int32_t* i_tgt, float* sum_src,
                             float* sum_tgt) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < B; ++b) {
    int n1 = ls ? (int)ls[b] : P1, n2 = lt ? (int)lt[b] : P2;
    const float* s = src + (size_t)b * P1 * D;
    const float* t = tgt + (size_t)b * P2 * D;
    if (directions & 1) {
      for (int i = 0; i < P1; ++i) { d_src[(size_t)b * P1 + i] = 0.f; i_src[(size_t)b * P1 + i] = 0; }
      nn1(s, t, n1, n2, D, d_src + (size_t)b * P1, i_src + (size_t)b * P1);
      double acc = 0.0; /* double: the oracle value is the correctly rounded sum */
      for (int i = 0; i < n1; ++i) acc = acc + (double)d_src[(size_t)b * P1 + i];
      sum_src[b] = (float)acc;
    }
    if (directions & 2) {
      for (int i = 0; i < P2; ++i) { d_tgt[(size_t)b * P2 + i] = 0.f; i_tgt[(size_t)b * P2 + i] = 0; }
      nn1(t, s, n2, n1, D, d_tgt + (size_t)b * P2, i_tgt + (size_t)b * P2);
      double acc = 0.0;
      for (int i = 0; i < n2; ++i) acc = acc + (double)d_tgt[(size_t)b * P2 + i];
      sum_tgt[b] = (float)acc;
    }
  }
}

/* grad of sum_src[b]*g_src[b] + sum_tgt[b]*g_tgt[b] w.r.t. src and tgt
 * (pytorch3d KNearestNeighborBackward semantics, upstream: 2*g*(p1 - p2[idx]);
 * scattered contributions summed in ascending source-point order). */
ORC_API void orc_chamfer_bwd(const float* src, const float* tgt, const int64_t* ls,
                             const int64_t* lt, const int32_t* i_src,
                             const int32_t* i_tgt, const float* g_src,
                             const float* g_tgt, int B, int P1, int P2, int D,
                             int directions, float* grad_src, float* grad_tgt) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < B; ++b) {
    int n1 = ls ? (int)ls[b] : P1, n2 = lt ? (int)lt[b] : P2;
    const float* s = src + (size_t)b * P1 * D;
    const float* t = tgt + (size_t)b * P2 * D;
    float* gs = grad_src ? grad_src + (size_t)b * P1 * D : NULL;
    float* gt = grad_tgt ? grad_tgt + (size_t)b * P2 * D : NULL;
    if (gs) memset(gs, 0, sizeof(float) * (size_t)P1 * D);
    if (gt) memset(gt, 0, sizeof(float) * (size_t)P2 * D);
    if ((directions & 1) && n2 > 0) {
      for (int i = 0; i < n1; ++i) {
        int j = i_src[(size_t)b * P1 + i];
        for (int d = 0; d < D; ++d) {
          float v = 2.0f * g_src[b] * (s[(size_t)i * D + d] - t[(size_t)j * D + d]);
          if (gs) gs[(size_t)i * D + d] = gs[(size_t)i * D + d] + v;
          if (gt) gt[(size_t)j * D + d] = gt[(size_t)j * D + d] - v;
        }
      }
    }
    if ((directions & 2) && n1 > 0) {
      for (int j = 0; j < n2; ++j) {
        int i = i_tgt[(size_t)b * P2 + j];
        for (int d = 0; d < D; ++d) {
          float v = 2.0f * g_tgt[b] * (t[(size_t)j * D + d] - s[(size_t)i * D + d]);
          if (gt) gt[(size_t)j * D + d] = gt[(size_t)j * D + d] + v;
          if (gs) gs[(size_t)i * D + d] = gs[(size_t)i * D + d] - v;
        }
      }
    }
  }
}

/* ------------------------------------------------------------------------
 * cubic_interpolation (gcn_lib/interpolation.py:103-123) for one sample.
 *  (i)   FRNN(K=32, r=cutoff) query->pos; in_range = set of valid idx (:19-31)
 *  (ii)  second FRNN on the compacted candidates == same lists, indices
 *        remapped order-preservingly (:33-42), so (d2, idx) order is unchanged
 *  (iii) iff some query has no valid neighbour (:44) every query with a -1 slot
 *        (:46) gets 4 extra kNN edges over the in-range candidates (:47-60)
 *  (iv)  r = sqrt(clamp(sum_c (s_c^2 + q_c^2 - (2 q_c) s_c)))  (:11-14)
 *  (v)   w = 8/(pi c^3) * W(r/c)  (:92-100)
 *  (vi)  out = sum w f / (sum w + 1e-6)  (:119-122), edges summed in graph
 *        order: FRNN slots ascending, then the pad edges.
 * ---------------------------------------------------------------------- */
static inline float l2dist_ref(const float* s, const float* q) {
  float acc = 0.0f;
  for (int c = 0; c < 3; ++c) {
    float a = s[c] * s[c];
    float bq = q[c] * q[c];
    float sum = a + bq;
    float two = 2.0f * q[c];
    float e = two * s[c];
    float t = sum - e;
    acc = acc + t; /* torch.sum over 3 comps == ((t0+t1)+t2) */
  }
  if (acc < 1e-8f) acc = 0.0f;
  return sqrtf(acc);
}

static inline float bicubic_ref(float r, float cutoff, float coeff) {
  float q = r / cutoff;
  float ker = 0.0f;
  if (q >= 0.0f && q <= 0.5f) {
    float q3 = (q * q) * q;
    float q2 = q * q;
    ker = 6.0f * (q3 - q2) + 1.0f;
  } else if (q > 0.5f && q <= 1.0f) {
    float u = 1.0f - q;
    ker = 2.0f * ((u * u) * u);
  }
  return ker * coeff;
}

/* element-wise exports of the two helpers above, pinned against the reference's own
 * l2dist / bicubic_kernel run live (tests/golden/interp_kernels.npz). */
ORC_API void orc_l2dist(const float* src, const float* dst, int n, float* out) {
  for (int i = 0; i < n; ++i) out[i] = l2dist_ref(src + (size_t)i * 3, dst + (size_t)i * 3);
}
ORC_API void orc_bicubic(const float* r, int n, float cutoff, float* out) {
  const float coeff = (float)(8.0 / (3.14159265358979323846 * (double)cutoff *
                                     (double)cutoff * (double)cutoff));
  for (int i = 0; i < n; ++i) out[i] = bicubic_ref(r[i], cutoff, coeff);
}

ORC_API void orc_cubic_interp(const float* query, const float* field,
                              const float* pos, int S, int Q, int P, int F,
                              float cutoff, float* out) {
  const int K = 32;
  /* coeff is computed in Python double then applied to an fp32 tensor */
  const float coeff = (float)(8.0 / (3.14159265358979323846 * (double)cutoff *
                                     (double)cutoff * (double)cutoff));
  for (int s = 0; s < S; ++s) {
    const float* qp = query + (size_t)s * Q * 3;
    const float* pp = pos + (size_t)s * P * 3;
    const float* ff = field + (size_t)s * P * F;
    float* oo = out + (size_t)s * Q * F;
    float* nd = (float*)malloc(sizeof(float) * (size_t)Q * K);
    int64_t* ni = (int64_t*)malloc(sizeof(int64_t) * (size_t)Q * K);
    unsigned char* inr = (unsigned char*)calloc((size_t)(P > 0 ? P : 1), 1);
    float r2 = cutoff * cutoff;
    orc_knn(qp, pp, NULL, NULL, 1, Q, P, 3, K, &r2, nd, ni);
    int any_empty = 0;
    for (int i = 0; i < Q; ++i) {
      if (ni[(size_t)i * K] < 0) any_empty = 1;
      for (int k = 0; k < K; ++k)
        if (ni[(size_t)i * K + k] >= 0) inr[ni[(size_t)i * K + k]] = 1;
    }
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < Q; ++i) {
      float acc[16];
      for (int c = 0; c < F; ++c) acc[c] = 0.0f;
      float wsum = 0.0f;
      int cnt = 0;
      for (int k = 0; k < K; ++k) {
        int64_t j = ni[(size_t)i * K + k];
        if (j < 0) break;
        ++cnt;
        float w = bicubic_ref(l2dist_ref(pp + (size_t)j * 3, qp + (size_t)i * 3), cutoff, coeff);
        for (int c = 0; c < F; ++c) acc[c] = acc[c] + ff[(size_t)j * F + c] * w;
        wsum = wsum + w;
      }
      if (any_empty && cnt < K) {
        /* 4 nearest among in-range candidates, (d2, idx) order, pad idx = first
         * in-range candidate (compacted index 0) like knn_points' zero pad */
        float bd[4];
        int64_t bi[4];
        int c4 = 0;
        int64_t first = -1;
        for (int j = 0; j < P; ++j) {
          if (!inr[j]) continue;
          if (first < 0) first = j;
          float d = sqdist(qp + (size_t)i * 3, pp + (size_t)j * 3, 3);
          if (c4 == 4 && !(d < bd[3])) continue;
          int p = c4 < 4 ? c4 : 3;
          while (p > 0 && bd[p - 1] > d) { bd[p] = bd[p - 1]; bi[p] = bi[p - 1]; --p; }
          bd[p] = d; bi[p] = j;
          if (c4 < 4) ++c4;
        }
        for (int k = c4; k < 4; ++k) bi[k] = first;
        for (int k = 0; k < 4; ++k) {
          int64_t j = bi[k];
          if (j < 0) continue;
          float w = bicubic_ref(l2dist_ref(pp + (size_t)j * 3, qp + (size_t)i * 3), cutoff, coeff);
          for (int c = 0; c < F; ++c) acc[c] = acc[c] + ff[(size_t)j * F + c] * w;
          wsum = wsum + w;
        }
      }
      for (int c = 0; c < F; ++c) oo[(size_t)i * F + c] = acc[c] / (wsum + 1e-6f);
    }
    free(nd);
    free(ni);
    free(inr);
  }
}

/* row gather: knn_gather / index_points (discriminator.py:43-60, loss.py:10-27);
 * negative indices wrap like Python indexing (loss.py:273-275 relies on -1). */
ORC_API void orc_gather_rows(const float* x, const int64_t* idx, int B, int N, int U,
                             int L, float* out) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b)
    for (int l = 0; l < L; ++l) {
      int64_t j = idx[(size_t)b * L + l];
      if (j < 0) j += N;
      memcpy(out + ((size_t)b * L + l) * U, x + ((size_t)b * N + j) * U, sizeof(float) * (size_t)U);
    }
}
