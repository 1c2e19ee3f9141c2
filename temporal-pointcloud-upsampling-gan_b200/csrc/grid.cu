// grid.cu — K3: uniform-grid neighbour search for 3-D clouds (kNN, fixed-radius, Chamfer NN).
//
// Replaces the search inside frnn.frnn_grid_points (discriminator.py:27; loss.py:256,261;
// gcn_lib/interpolation.py:20,33), pytorch3d knn_points on positions (gcn_lib/pointnet/gcn.py:16)
// and the two K=1 searches of chamferdist (loss.py:176-181) for clouds of >= 2048 points, where
// brute force is bound by the fp32 pipe (67 M pairs per 8192-point cloud) and a grid makes the
// search proportional to the neighbourhood size.  Results are IDENTICAL to the brute-force
// kernels: the same canonical distance expression is evaluated for every visited candidate and
// candidates are ranked by the full (d2, index) key, which is independent of traversal order.
//
// Build (per call, no host synchronisation, grid parameters computed on the device):
//   setup  one CTA per cloud: bounding box -> origin / cell size / dims (<= 32 per axis), zeroes
//          the cell counters.  Radius searches use cell = r (27-cell block covers the ball);
//          kNN / NN use ~4 points per cell and grow the block until the K-th distance is
//          provably covered.
//   count  cell histogram (integer atomics: deterministic counts), scan, fill: counting sort of
//          (x, y, z, original index) into float4 records — one coalesced 16-byte load per
//          candidate in the search.
// Search:
//   K >= 2  one warp per query: the rows of the cell block are contiguous record ranges; the
//           ranges are flattened with a warp prefix sum so all 32 lanes evaluate candidates,
//           admission by ballot, insertion into the register-resident sorted list by key.
//   K == 1  one thread per query (no list), used by Chamfer and K=1 radius searches.
#include "common.cuh"
#include "internal.cuh"

namespace tpg {

constexpr int GRID_AXIS = 32;
constexpr int GRID_CELLS = GRID_AXIS * GRID_AXIS * GRID_AXIS;  // counters allocated per cloud

struct GridParams {
  float ox, oy, oz, h, inv_h, slack;
  int gx, gy, gz, n;
};

struct GridRef {
  const GridParams* prm;   // [B]
  const int* cell_start;   // [B][GRID_CELLS + 1]
  const float4* rec;       // [B][P] (x, y, z, bits(index)) sorted by cell
  int P;
};

__device__ __forceinline__ int cell_of(float v, float o, float inv_h, int g) {
  const int c = (int)floorf(__fmul_rn(__fsub_rn(v, o), inv_h));
  return min(max(c, 0), g - 1);
}

// ---- build --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) grid_setup_kernel(const float* __restrict__ p, const int64_t* __restrict__ len,
                                                          int P, int K, int use_radius, float r,
                                                          const float* __restrict__ r_per_cloud,
                                                          GridParams* __restrict__ prm, int* __restrict__ counts) {
  __shared__ float red[6][32];
  __shared__ int ncell_s;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = len ? min((int)len[b], P) : P;
  const float* pb = p + (size_t)b * P * 3;
  const float INF = __int_as_float(0x7f800000);
  float lo[3] = {INF, INF, INF}, hi[3] = {-INF, -INF, -INF};
  for (int i = tid; i < n; i += 1024) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = pb[(size_t)i * 3 + c];
      lo[c] = fminf(lo[c], v);
      hi[c] = fmaxf(hi[c], v);
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[c] = fminf(lo[c], __shfl_xor_sync(FULL, lo[c], o));
      hi[c] = fmaxf(hi[c], __shfl_xor_sync(FULL, hi[c], o));
    }
    if (lane == 0) { red[c][warp] = lo[c]; red[3 + c][warp] = hi[c]; }
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      lo[c] = red[c][lane];
      hi[c] = red[3 + c][lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        lo[c] = fminf(lo[c], __shfl_xor_sync(FULL, lo[c], o));
        hi[c] = fmaxf(hi[c], __shfl_xor_sync(FULL, hi[c], o));
      }
    }
    if (lane == 0) {
      GridParams g;
      if (n == 0) { lo[0] = lo[1] = lo[2] = 0.f; hi[0] = hi[1] = hi[2] = 0.f; }
      const float ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
      const float emax = fmaxf(ex, fmaxf(ey, ez));
      const float amax = fmaxf(fmaxf(fabsf(lo[0]), fabsf(hi[0])),
                               fmaxf(fmaxf(fabsf(lo[1]), fabsf(hi[1])), fmaxf(fabsf(lo[2]), fabsf(hi[2]))));
      float h;
      if (use_radius) {
        const float rr = r_per_cloud ? r_per_cloud[b] : r;
        h = rr * 1.0001f;
      } else {
        const float e0 = fmaxf(emax * 1e-3f, 1e-20f);
        const float vol = fmaxf(ex, e0) * fmaxf(ey, e0) * fmaxf(ez, e0);
        // the K-th neighbour sits at ~ (3K / (4 pi density))^(1/3); cells 1.3x that wide make the
        // 27-cell block cover it almost always: 0.55 K points per cell (at least 2)
        const float target = fmaxf(2.0f, 0.55f * (float)K);
        h = cbrtf(vol * target / (float)max(n, 1));
      }
      h = fmaxf(h, emax / (float)GRID_AXIS * 1.0001f);
      if (!(h > 0.f) || !(h < INF)) h = 1.0f;
      g.ox = lo[0]; g.oy = lo[1]; g.oz = lo[2];
      g.h = h;
      g.inv_h = 1.0f / h;
      g.gx = min(GRID_AXIS, (int)floorf(ex * g.inv_h) + 1);
      g.gy = min(GRID_AXIS, (int)floorf(ey * g.inv_h) + 1);
      g.gz = min(GRID_AXIS, (int)floorf(ez * g.inv_h) + 1);
      g.n = n;
      // slack of every coverage bound: rounding of (v - o) * inv_h at a cell face
      g.slack = 4e-6f * (amax + emax + h) + 1e-5f * h;
      prm[b] = g;
      ncell_s = g.gx * g.gy * g.gz;
    }
  }
  __syncthreads();
  int* cb = counts + (size_t)b * GRID_CELLS;
  const int ncell = ncell_s;
  for (int i = tid; i < ncell; i += 1024) cb[i] = 0;
}

__global__ void grid_count_kernel(const float* __restrict__ p, int B, int P, const GridParams* __restrict__ prm,
                                  int* __restrict__ counts, int* __restrict__ cellid) {
  const long long total = (long long)B * P;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(e / P), i = (int)(e - (long long)b * P);
    const GridParams g = prm[b];
    if (i >= g.n) continue;
    const float* q = p + (size_t)e * 3;
    const int c = (cell_of(q[2], g.oz, g.inv_h, g.gz) * g.gy + cell_of(q[1], g.oy, g.inv_h, g.gy)) * g.gx +
                  cell_of(q[0], g.ox, g.inv_h, g.gx);
    cellid[e] = c;
    atomicAdd(counts + (size_t)b * GRID_CELLS + c, 1);
  }
}

// exclusive scan of the used cell counters of a cloud -> cell_start[0..ncell]; counters become cursors
__global__ void __launch_bounds__(1024) grid_scan_kernel(const GridParams* __restrict__ prm, int* __restrict__ counts,
                                                         int* __restrict__ cell_start) {
  __shared__ int wsum[32];
  __shared__ int carry_s;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ncell = prm[b].gx * prm[b].gy * prm[b].gz;
  int* cb = counts + (size_t)b * GRID_CELLS;
  int* cs = cell_start + (size_t)b * (GRID_CELLS + 1);
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < ncell; base += 1024) {
    const int e = base + tid;
    const int v = e < ncell ? cb[e] : 0;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(FULL, inc, d);
      if (lane >= d) inc += t;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      int w = wsum[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(FULL, w, d);
        if (lane >= d) w += t;
      }
      wsum[lane] = w;
    }
    __syncthreads();
    const int excl = carry_s + (warp > 0 ? wsum[warp - 1] : 0) + inc - v;
    if (e < ncell) { cs[e] = excl; cb[e] = excl; }
    __syncthreads();
    if (tid == 1023) carry_s = excl + v;
    __syncthreads();
  }
  if (tid == 0) cs[ncell] = carry_s;
}

__global__ void grid_fill_kernel(const float* __restrict__ p, int B, int P, const GridParams* __restrict__ prm,
                                 int* __restrict__ cursor, const int* __restrict__ cellid, float4* __restrict__ rec) {
  const long long total = (long long)B * P;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(e / P), i = (int)(e - (long long)b * P);
    if (i >= prm[b].n) continue;
    const float* q = p + (size_t)e * 3;
    const int pos = atomicAdd(cursor + (size_t)b * GRID_CELLS + cellid[e], 1);
    rec[(size_t)b * P + pos] = make_float4(q[0], q[1], q[2], __int_as_float(i));
  }
}

// ---- block of cells around a query + the distance it provably covers -----------------------------
struct Block {
  int x0, x1, y0, y1, z0, z1;
  float cover;  // every point outside the block is farther than this (inf: block == grid)
};

__device__ __forceinline__ Block make_block(const GridParams& g, float qx, float qy, float qz, int R) {
  const float INF = __int_as_float(0x7f800000);
  const int cx = cell_of(qx, g.ox, g.inv_h, g.gx), cy = cell_of(qy, g.oy, g.inv_h, g.gy),
            cz = cell_of(qz, g.oz, g.inv_h, g.gz);
  Block k;
  k.x0 = max(cx - R, 0); k.x1 = min(cx + R, g.gx - 1);
  k.y0 = max(cy - R, 0); k.y1 = min(cy + R, g.gy - 1);
  k.z0 = max(cz - R, 0); k.z1 = min(cz + R, g.gz - 1);
  float c = INF;
  // a face that is not on the grid boundary limits the covered distance
  if (k.x0 > 0) c = fminf(c, qx - (g.ox + (float)k.x0 * g.h));
  if (k.x1 < g.gx - 1) c = fminf(c, (g.ox + (float)(k.x1 + 1) * g.h) - qx);
  if (k.y0 > 0) c = fminf(c, qy - (g.oy + (float)k.y0 * g.h));
  if (k.y1 < g.gy - 1) c = fminf(c, (g.oy + (float)(k.y1 + 1) * g.h) - qy);
  if (k.z0 > 0) c = fminf(c, qz - (g.oz + (float)k.z0 * g.h));
  if (k.z1 < g.gz - 1) c = fminf(c, (g.oz + (float)(k.z1 + 1) * g.h) - qz);
  k.cover = c == INF ? INF : fmaxf(c - g.slack, 0.0f) * 0.99999f;
  return k;
}

// ---- K >= 2: one warp per query --------------------------------------------------------------------
struct GridKnnArgs {
  const float* p1;
  const float4* qrec;  // != null: queries are visited in the cell order of their own grid (self search)
  const int64_t* len1;
  int B, P1, K, use_radius;
  float r;
  const float* r_per_cloud;
  GridRef g;
  float* dists;
  void* idx;
  int out_mode;  // OUT_KNN / OUT_FRNN / OUT_THREE, or OUT_BALL: the K LOWEST INDICES inside the radius (int32,
                 // remaining slots repeat the first hit, no hit -> 0): pointnet2 ball_query semantics
};

constexpr int OUT_BALL = 100;


// Upper end of the 16-bit radix bucket holding the k-th smallest (1-based) of the NV keys per lane of a warp
// (keys of absent values = 0xffffffff): a valid upper bound of the k-th smallest, < 1 % above it.
template <int NV>
__device__ __forceinline__ unsigned grid_radix_bound16(const unsigned (&uk)[NV], int k) {
  unsigned prefix = 0u, himask = 0u;
#pragma unroll 1
  for (int bit = 31; bit >= 16; --bit) {
    const unsigned bmask = 1u << bit;
    unsigned local = 0;
#pragma unroll
    for (int v = 0; v < NV; ++v) local += ((uk[v] & (himask | bmask)) == prefix) ? 1u : 0u;
    const int c = (int)__reduce_add_sync(FULL, local);
    if (k > c) { prefix |= bmask; k -= c; }
    himask |= bmask;
  }
  return prefix | 0xffffu;
}

constexpr int GK_PER = 12;  // candidates per lane the select-based fast path of grid_knn_kernel keeps in registers

template <bool FAST>  // FAST: plain kNN with K <= 24 -- adds the select-based first block (80 registers instead of 64)
__global__ void __launch_bounds__(256) grid_knn_kernel(GridKnnArgs a) {
  __shared__ float2 sel_s[FAST ? 8 : 1][2][32];  // per warp: selected candidates, then the sorted result
  const int lane = threadIdx.x & 31;
  long long wq = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (wq >= (long long)a.B * a.P1) return;
  const int b = (int)(wq / a.P1);
  int qi = (int)(wq - (long long)b * a.P1);
  const int n1 = a.len1 ? min((int)a.len1[b], a.P1) : a.P1;
  const float INF = __int_as_float(0x7f800000);
  const int K = a.K;
  WarpList L;
  L.init();
  float qx = 0.f, qy = 0.f, qz = 0.f;
  if (qi < n1) {
    if (a.qrec) {  // neighbouring warps handle neighbouring points: shared candidate cells, L1/L2 hits
      const float4 v = __ldg(a.qrec + (size_t)b * a.P1 + qi);
      qx = v.x; qy = v.y; qz = v.z;
      qi = __float_as_int(v.w);
      wq = (long long)b * a.P1 + qi;
    } else {
      const float* q = a.p1 + (size_t)wq * 3;
      qx = q[0]; qy = q[1]; qz = q[2];
    }
  }
  if (qi < n1) {
    const GridParams g = a.g.prm[b];
    const int* cs = a.g.cell_start + (size_t)b * (GRID_CELLS + 1);
    const float4* rec = a.g.rec + (size_t)b * a.g.P;
    float r2 = INF;
    if (a.use_radius) {
      const float rr = a.r_per_cloud ? a.r_per_cloud[b] : a.r;
      r2 = __fmul_rn(rr, rr);
    }
    for (int R = 1;; R *= 2) {
      const Block k = make_block(g, qx, qy, qz, R);
      L.init();
      float tau_d = INF;
      int tau_i = 0x7fffffff;
      const int ny = k.y1 - k.y0 + 1, nrows = ny * (k.z1 - k.z0 + 1);
      // ---- fast path (first block, <= 32 * GK_PER candidates, K <= 32): no sorted-list insertions.  All candidate
      // distances stay in registers; a 16-bit radix select bounds the K-th smallest; the few candidates under the
      // bound are ranked by (d, idx) with shuffles.  Same result as the insertion path (canonical order).
      if (FAST && R == 1 && nrows <= 32) {
        int start = 0, cnt = 0;
        if (lane < nrows) {
          const int z = k.z0 + lane / ny, y = k.y0 + lane % ny;
          const int c0 = (z * g.gy + y) * g.gx;
          start = cs[c0 + k.x0];
          cnt = cs[c0 + k.x1 + 1] - start;
        }
        int inc = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int t = __shfl_up_sync(FULL, inc, d);
          if (lane >= d) inc += t;
        }
        const int total = __shfl_sync(FULL, inc, 31);
        if (total <= 32 * GK_PER) {
          float dv[GK_PER];
          int iv[GK_PER];
          int nvalid = 0;
#pragma unroll
          for (int s2 = 0; s2 < GK_PER; ++s2) {
            dv[s2] = INF;
            iv[s2] = 0x7fffffff;
            if (s2 * 32 < total) {  // (warp-uniform)
              const int t = s2 * 32 + lane;
              int rsel = 0;
              for (int rr = 0; rr < nrows - 1; ++rr) rsel += (t >= __shfl_sync(FULL, inc, rr)) ? 1 : 0;
              const int rstart = __shfl_sync(FULL, start, rsel);
              const int rinc = __shfl_sync(FULL, inc, rsel);
              const int rcnt = __shfl_sync(FULL, cnt, rsel);
              if (t < total) {
                const float4 v = __ldg(rec + rstart + (t - (rinc - rcnt)));
                const float d = sqdist3(qx, qy, qz, v.x, v.y, v.z);
                if (d < r2) { dv[s2] = d; iv[s2] = __float_as_int(v.w); ++nvalid; }
              }
            }
          }
          nvalid = (int)__reduce_add_sync(FULL, (unsigned)nvalid);
          unsigned bound = 0x7f7fffffu;  // every finite distance
          if (nvalid > K) {
            unsigned uk[GK_PER];
#pragma unroll
            for (int s2 = 0; s2 < GK_PER; ++s2) uk[s2] = iv[s2] == 0x7fffffff ? 0xffffffffu : __float_as_uint(dv[s2]);
            bound = grid_radix_bound16(uk, K);
          }
          // compact the candidates under the bound (all ties at the K-th distance included) into <= 32 slots
          const unsigned lt = (1u << lane) - 1u;
          const int wl = threadIdx.x >> 5;
          int nsel = 0;
#pragma unroll
          for (int s2 = 0; s2 < GK_PER; ++s2) {
            const bool sel = iv[s2] != 0x7fffffff && __float_as_uint(dv[s2]) <= bound;
            const unsigned m = __ballot_sync(FULL, sel);
            const int pos = nsel + __popc(m & lt);
            if (sel && pos < 32) sel_s[wl][0][pos] = make_float2(dv[s2], __int_as_float(iv[s2]));
            nsel += __popc(m);
          }
          if (nsel <= 32) {
            __syncwarp();
            float dc = INF;
            int ic = 0x7fffffff;
            if (lane < nsel) { const float2 v = sel_s[wl][0][lane]; dc = v.x; ic = __float_as_int(v.y); }
            int rank = 0;
            for (int m2 = 0; m2 < nsel; ++m2) {
              const float od = __shfl_sync(FULL, dc, m2);
              const int oi = __shfl_sync(FULL, ic, m2);
              rank += (od < dc || (od == dc && oi < ic)) ? 1 : 0;
            }
            if (lane < nsel) sel_s[wl][1][rank] = make_float2(dc, __int_as_float(ic));
            __syncwarp();
            const int nres = min(nsel, K);
            if (lane < nres) { const float2 v = sel_s[wl][1][lane]; L.d = v.x; L.i = __float_as_int(v.y); }
            __syncwarp();
            tau_d = nsel >= K ? __shfl_sync(FULL, L.d, K - 1) : INF;
            if (a.use_radius || k.cover == INF) break;
            if (tau_d < INF && tau_d < __fmul_rn(k.cover, k.cover)) break;
            continue;  // not covered yet: next block size through the insertion path
          }
          L.init();
        }
      }
      for (int row0 = 0; row0 < nrows; row0 += 32) {
        // rows of the block: contiguous record ranges; flatten them with a prefix sum
        const int row = row0 + lane;
        int start = 0, cnt = 0;
        if (row < nrows) {
          const int z = k.z0 + row / ny, y = k.y0 + row % ny;
          const int c0 = (z * g.gy + y) * g.gx;
          start = cs[c0 + k.x0];
          cnt = cs[c0 + k.x1 + 1] - start;
        }
        int inc = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int t = __shfl_up_sync(FULL, inc, d);
          if (lane >= d) inc += t;
        }
        const int total = __shfl_sync(FULL, inc, 31);
        const int nr = min(32, nrows - row0);
        for (int t0 = 0; t0 < total; t0 += 32) {
          const int t = t0 + lane;
          int rsel = 0;
          for (int rr = 0; rr < nr - 1; ++rr) rsel += (t >= __shfl_sync(FULL, inc, rr)) ? 1 : 0;
          const int rstart = __shfl_sync(FULL, start, rsel);
          const int rinc = __shfl_sync(FULL, inc, rsel);
          const int rcnt = __shfl_sync(FULL, cnt, rsel);
          float d = INF;
          int ci = 0x7fffffff;
          if (t < total) {
            const float4 v = __ldg(rec + rstart + (t - (rinc - rcnt)));
            // ball query: the centre is the FIRST operand upstream (new_xyz - xyz); (a-b)^2 == (b-a)^2 exactly
            d = sqdist3(qx, qy, qz, v.x, v.y, v.z);
            ci = __float_as_int(v.w);
            if (a.out_mode == OUT_BALL) d = d < r2 ? 0.0f : __int_as_float(0x7f800000);  // rank by index only
          }
          unsigned m = __ballot_sync(FULL, t < total && d < r2 && (d < tau_d || (d == tau_d && ci < tau_i)));
          while (m) {
            const int l = __ffs(m) - 1;
            m &= m - 1;
            const float dc = __shfl_sync(FULL, d, l);
            const int ic = __shfl_sync(FULL, ci, l);
            if (dc < tau_d || (dc == tau_d && ic < tau_i)) {
              L.insert_key(dc, ic, lane);
              tau_d = __shfl_sync(FULL, L.d, K - 1);
              tau_i = __shfl_sync(FULL, L.i, K - 1);
              if (tau_i < 0) tau_i = 0x7fffffff;
            }
          }
        }
      }
      // radius search: the 27-cell block (cell >= r) covers the ball; kNN: stop once the K-th
      // distance is provably inside the covered region
      if (a.use_radius || k.cover == INF) break;
      if (tau_d < INF && tau_d < __fmul_rn(k.cover, k.cover)) break;
    }
  }
  if (a.out_mode == OUT_BALL) {
    if (qi < a.P1 && lane < K) {
      const int first = __shfl_sync(FULL, L.i, 0);
      reinterpret_cast<int32_t*>(a.idx)[(size_t)wq * K + lane] = L.i >= 0 ? L.i : (first >= 0 ? first : 0);
    }
    return;
  }
  if (qi < a.P1 && lane < K) {
    const size_t o = (size_t)wq * K + lane;
    const bool found = qi < n1 && L.i >= 0;
    if (a.out_mode == OUT_THREE) {
      a.dists[o] = found ? sqrtf(L.d) : 0.0f;
      reinterpret_cast<int32_t*>(a.idx)[o] = found ? L.i : 0;
    } else {
      const float padd = a.out_mode == OUT_FRNN ? -1.0f : 0.0f;
      const int64_t padi = a.out_mode == OUT_FRNN ? -1 : 0;
      a.dists[o] = found ? L.d : padd;
      reinterpret_cast<int64_t*>(a.idx)[o] = found ? (int64_t)L.i : padi;
    }
  }
}

// ---- K == 1: one thread per query --------------------------------------------------------------------
struct GridNn1Args {
  const float* q;          // [B,Pq,3]
  const float4* qrec;      // != null: queries visited in the cell order of their own grid
  const int64_t* qlen;
  int B, Pq, use_radius;
  float r;
  const float* r_per_cloud;
  GridRef g;
  float* d_out;            // [B,Pq]
  int32_t* i32_out;        // chamfer mode: int32 [B,Pq]
  int64_t* i64_out;        // knn / frnn mode: int64 [B,Pq]
  int out_mode;            // OUT_KNN / OUT_FRNN pads; i32_out != null -> chamfer (0 / 0)
};

__global__ void __launch_bounds__(128) grid_nn1_kernel(GridNn1Args a) {
  long long e = (long long)blockIdx.x * 128 + threadIdx.x;
  if (e >= (long long)a.B * a.Pq) return;
  const int b = (int)(e / a.Pq);
  int qi = (int)(e - (long long)b * a.Pq);
  const int nq = a.qlen ? min((int)a.qlen[b], a.Pq) : a.Pq;
  const float INF = __int_as_float(0x7f800000);
  float best = INF;
  int bi = -1;
  float qx = 0.f, qy = 0.f, qz = 0.f;
  if (qi < nq) {
    if (a.qrec) {
      const float4 v = __ldg(a.qrec + (size_t)b * a.Pq + qi);
      qx = v.x; qy = v.y; qz = v.z;
      qi = __float_as_int(v.w);
      e = (long long)b * a.Pq + qi;
    } else {
      const float* q = a.q + (size_t)e * 3;
      qx = q[0]; qy = q[1]; qz = q[2];
    }
  }
  if (qi < nq) {
    const GridParams g = a.g.prm[b];
    const int* cs = a.g.cell_start + (size_t)b * (GRID_CELLS + 1);
    const float4* rec = a.g.rec + (size_t)b * a.g.P;
    float r2 = INF;
    if (a.use_radius) {
      const float rr = a.r_per_cloud ? a.r_per_cloud[b] : a.r;
      r2 = __fmul_rn(rr, rr);
    }
    for (int R = 1;; R *= 2) {
      const Block k = make_block(g, qx, qy, qz, R);
      best = INF;
      bi = -1;
      for (int z = k.z0; z <= k.z1; ++z)
        for (int y = k.y0; y <= k.y1; ++y) {
          const int c0 = (z * g.gy + y) * g.gx;
          const int s0 = cs[c0 + k.x0], s1 = cs[c0 + k.x1 + 1];
          for (int s = s0; s < s1; ++s) {
            const float4 v = __ldg(rec + s);
            const float d = sqdist3(qx, qy, qz, v.x, v.y, v.z);
            const int ci = __float_as_int(v.w);
            if (d < r2 && (d < best || (d == best && ci < bi))) { best = d; bi = ci; }
          }
        }
      if (a.use_radius || k.cover == INF) break;
      if (bi >= 0 && best < __fmul_rn(k.cover, k.cover)) break;
    }
  }
  const bool found = qi < nq && bi >= 0;
  if (a.i32_out) {
    a.d_out[e] = found ? best : 0.0f;
    a.i32_out[e] = found ? bi : 0;
  } else {
    const float padd = a.out_mode == OUT_FRNN ? -1.0f : 0.0f;
    const int64_t padi = a.out_mode == OUT_FRNN ? -1 : 0;
    a.d_out[e] = found ? best : padd;
    a.i64_out[e] = found ? (int64_t)bi : padi;
  }
}

// ---- host side ---------------------------------------------------------------------------------------
struct GridWs {
  GridParams* prm;
  int* counts;
  int* cell_start;
  int* cellid;
  float4* rec;
  size_t total;
};

static GridWs grid_carve(void* base, int B, int P) {
  GridWs w;
  char* p = reinterpret_cast<char*>(base);
  size_t o = 0;
  w.prm = reinterpret_cast<GridParams*>(p + o);  o += align_up(sizeof(GridParams) * (size_t)B, 256);
  w.counts = reinterpret_cast<int*>(p + o);      o += align_up(sizeof(int) * (size_t)B * GRID_CELLS, 256);
  w.cell_start = reinterpret_cast<int*>(p + o);  o += align_up(sizeof(int) * (size_t)B * (GRID_CELLS + 1), 256);
  w.cellid = reinterpret_cast<int*>(p + o);      o += align_up(sizeof(int) * (size_t)B * P, 256);
  w.rec = reinterpret_cast<float4*>(p + o);      o += align_up(sizeof(float4) * (size_t)B * P, 256);
  w.total = o;
  return w;
}

size_t grid_workspace_bytes(int B, int P) { return grid_carve(nullptr, B, P).total; }

bool grid_eligible(int D, int P2, int K) { return D == 3 && P2 >= 2048 && K >= 1 && K <= 32; }

// builds the grid of cloud set p [B,P,3] inside `workspace` and returns a reference to it
static int grid_build(const float* p, const int64_t* len, int B, int P, int K, int use_radius, float r,
                      const float* r_per_cloud, void* workspace, GridRef* out, cudaStream_t st) {
  GridWs w = grid_carve(workspace, B, P);
  grid_setup_kernel<<<B, 1024, 0, st>>>(p, len, P, K, use_radius, r, r_per_cloud, w.prm, w.counts);
  TPG_CHECK_LAUNCH("grid_setup_kernel");
  const long long total = (long long)B * P;
  const unsigned blocks = (unsigned)min((total + 255) / 256, (long long)num_sms() * 16);
  grid_count_kernel<<<blocks, 256, 0, st>>>(p, B, P, w.prm, w.counts, w.cellid);
  TPG_CHECK_LAUNCH("grid_count_kernel");
  grid_scan_kernel<<<B, 1024, 0, st>>>(w.prm, w.counts, w.cell_start);
  TPG_CHECK_LAUNCH("grid_scan_kernel");
  grid_fill_kernel<<<blocks, 256, 0, st>>>(p, B, P, w.prm, w.counts, w.cellid, w.rec);
  TPG_CHECK_LAUNCH("grid_fill_kernel");
  out->prm = w.prm;
  out->cell_start = w.cell_start;
  out->rec = w.rec;
  out->P = P;
  return TPG_OK;
}

int grid_ball_query(const float* xyz, const float* new_xyz, int B, int N, int M, float radius, int nsample,
                    int32_t* idx, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  TPG_REQUIRE(workspace && workspace_bytes >= grid_workspace_bytes(B, N), TPG_EWORKSPACE,
              "ball_query grid search: workspace too small (need %zu bytes)", grid_workspace_bytes(B, N));
  GridRef g;
  int rc = grid_build(xyz, nullptr, B, N, nsample, 1, radius, nullptr, workspace, &g, st);
  if (rc) return rc;
  GridKnnArgs k{new_xyz, nullptr, nullptr, B, M, nsample, 1, radius, nullptr, g, nullptr, idx, OUT_BALL};
  grid_knn_kernel<false><<<(unsigned)(((long long)B * M + 7) / 8), 256, 0, st>>>(k);
  TPG_CHECK_LAUNCH("grid_knn_kernel");
  return TPG_OK;
}

int grid_knn_dispatch(const KnnArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  TPG_REQUIRE(workspace && workspace_bytes >= grid_workspace_bytes(a.B, a.P2), TPG_EWORKSPACE,
              "grid search: workspace too small (need %zu bytes)", grid_workspace_bytes(a.B, a.P2));
  TPG_REQUIRE(a.B <= 65535, TPG_EUNSUPPORTED, "grid search: B > 65535");
  GridRef g;
  int rc = grid_build(a.p2, a.len2, a.B, a.P2, a.K, a.use_radius, a.r, a.r_per_cloud, workspace, &g, st);
  if (rc) return rc;
  const long long queries = (long long)a.B * a.P1;
  // self search: visit the queries in the cell order of the grid that was just built
  const float4* qrec = (a.p1 == a.p2 && a.P1 == a.P2 && a.len1 == a.len2) ? g.rec : nullptr;
  if (a.K == 1 && a.out_mode != OUT_THREE) {
    GridNn1Args n{a.p1, qrec, a.len1, a.B, a.P1, a.use_radius, a.r, a.r_per_cloud, g, a.dists, nullptr,
                  reinterpret_cast<int64_t*>(a.idx), a.out_mode};
    grid_nn1_kernel<<<(unsigned)((queries + 127) / 128), 128, 0, st>>>(n);
    TPG_CHECK_LAUNCH("grid_nn1_kernel");
    return TPG_OK;
  }
  GridKnnArgs k{a.p1, qrec, a.len1, a.B, a.P1, a.K, a.use_radius, a.r, a.r_per_cloud, g, a.dists, a.idx, a.out_mode};
  // plain kNN with K <= 24 (<= 32 * GK_PER candidates in the first block at the grid's cell size): select-based kernel
  // (three_nn, K = 3, stays on the insertion kernel: measured 140 vs 175 us at 8 x 8192 x 2048 — the select's fixed
  // cost does not pay for three neighbours)
  if (a.out_mode == OUT_KNN && !a.use_radius && a.K <= 24)
    grid_knn_kernel<true><<<(unsigned)((queries + 7) / 8), 256, 0, st>>>(k);
  else
    grid_knn_kernel<false><<<(unsigned)((queries + 7) / 8), 256, 0, st>>>(k);
  TPG_CHECK_LAUNCH("grid_knn_kernel");
  return TPG_OK;
}

// Chamfer: both directions.  Clouds of >= 2048 points get a grid; a direction whose candidate cloud
// has a grid searches it, visiting its queries in the cell order of the query cloud's own grid when
// that exists.  Returns which directions were handled (bit 0: src->tgt, bit 1: tgt->src).
size_t grid_chamfer_workspace_bytes(int B, int P1, int P2) {
  return (grid_eligible(3, P1, 1) ? grid_workspace_bytes(B, P1) : 0) + (grid_eligible(3, P2, 1) ? grid_workspace_bytes(B, P2) : 0);
}

int grid_chamfer_nn(const float* src, const float* tgt, const int64_t* ls, const int64_t* lt, int B, int P1, int P2,
                    int directions, float* d_src, int32_t* i_src, float* d_tgt, int32_t* i_tgt, void* workspace,
                    size_t workspace_bytes, int* handled, cudaStream_t st) {
  *handled = 0;
  const bool gs = grid_eligible(3, P1, 1), gt = grid_eligible(3, P2, 1);
  if (!gs && !gt) return TPG_OK;
  TPG_REQUIRE(workspace && workspace_bytes >= grid_chamfer_workspace_bytes(B, P1, P2), TPG_EWORKSPACE,
              "chamfer grid search: workspace too small");
  char* ws = reinterpret_cast<char*>(workspace);
  GridRef rs{}, rt{};
  int rc;
  if (gs) {
    if ((rc = grid_build(src, ls, B, P1, 1, 0, 0.f, nullptr, ws, &rs, st))) return rc;
    ws += grid_workspace_bytes(B, P1);
  }
  if (gt && (rc = grid_build(tgt, lt, B, P2, 1, 0, 0.f, nullptr, ws, &rt, st))) return rc;
  if ((directions & TPG_CHAMFER_FWD) && gt && P1 > 0) {
    GridNn1Args n{src, gs ? rs.rec : nullptr, ls, B, P1, 0, 0.f, nullptr, rt, d_src, i_src, nullptr, OUT_KNN};
    grid_nn1_kernel<<<(unsigned)(((long long)B * P1 + 127) / 128), 128, 0, st>>>(n);
    TPG_CHECK_LAUNCH("grid_nn1_kernel");
    *handled |= TPG_CHAMFER_FWD;
  }
  if ((directions & TPG_CHAMFER_REV) && gs && P2 > 0) {
    GridNn1Args n{tgt, gt ? rt.rec : nullptr, lt, B, P2, 0, 0.f, nullptr, rs, d_tgt, i_tgt, nullptr, OUT_KNN};
    grid_nn1_kernel<<<(unsigned)(((long long)B * P2 + 127) / 128), 128, 0, st>>>(n);
    TPG_CHECK_LAUNCH("grid_nn1_kernel");
    *handled |= TPG_CHAMFER_REV;
  }
  return TPG_OK;
}

}  // namespace tpg
