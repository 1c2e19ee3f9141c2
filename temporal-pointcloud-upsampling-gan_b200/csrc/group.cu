// group.cu — K6/K7/K8b: grouping / gather forward, atomic-free backward through an
// inverse index, fused gather+reduce, three_interpolate.
//
// Replaces pointnet2_ops grouping_operation (gcn_lib/pointnet/gcn.py:207,261;
// discriminator.py:270,273), gather_operation (discriminator.py:132), the
// grouping+max pair of gcn.py:261-263, and three_interpolate.
//
// Forward design (HBM-bound: the [B,C,M,k] output dominates the traffic):
//   a CTA owns (cloud b, a tile of TC channels, a range of flat (m,j) positions).
//   The TC feature rows f[b,c,:] are staged in shared memory once (coalesced), then
//   every thread turns one int4 of indices into TC float4 stores: the random
//   accesses hit shared memory (4-byte bank-granular) instead of pulling a 32-byte
//   sector per 4 useful bytes through L1/L2, and the output is written with fully
//   coalesced 16-byte stores.  Rows that do not fit in shared memory fall back to
//   read-only global gathers.
// Backward design: scatter-add is inverted into a gather.  tpg_inverse_index_build
//   turns idx into a CSR (for every source point n: the ascending list of flat
//   positions that read it); grad_f[b,c,n] is then a private sequential sum —
//   no atomics, deterministic, and the CSR is reused by every op sharing the idx.
#include <cooperative_groups.h>

#include "common.cuh"
#include "internal.cuh"

namespace tpg {

constexpr int GRP_THREADS = 256;
constexpr size_t GRP_SMEM_MAX = 96 * 1024;

struct GroupFwdArgs {
  const float* f;        // [B,C,N]
  const int32_t* idx;    // [B,L]
  const float* center;   // [B,C,M] or null
  int B, C, N, M, k, L;
  int TC, LT, rows_vec;
  float* out;            // [B,C,L]
};

template <bool SMEM, bool VEC4>
__global__ void __launch_bounds__(GRP_THREADS) group_fwd_kernel(GroupFwdArgs a) {
  extern __shared__ float rows_s[];
  const int b = blockIdx.z, c0 = blockIdx.y * a.TC, tid = threadIdx.x;
  const int tc = min(a.TC, a.C - c0);
  const float* fb = a.f + ((size_t)b * a.C + c0) * a.N;
  if (SMEM) {
    const int total = tc * a.N;
    if (a.rows_vec) {
      const float4* src = reinterpret_cast<const float4*>(fb);
      float4* dst = reinterpret_cast<float4*>(rows_s);
      for (int e = tid; e < (total >> 2); e += GRP_THREADS) dst[e] = __ldg(src + e);
    } else {
      for (int e = tid; e < total; e += GRP_THREADS) rows_s[e] = __ldg(fb + e);
    }
    __syncthreads();
  }
  const float* rows = SMEM ? rows_s : fb;
  const int l0 = blockIdx.x * a.LT;
  const int l1 = min(a.L, l0 + a.LT);
  const int32_t* ib = a.idx + (size_t)b * a.L;
  float* ob = a.out + ((size_t)b * a.C + c0) * a.L;
  const float* cb = a.center ? a.center + ((size_t)b * a.C + c0) * a.M : nullptr;
  if (VEC4) {
    const int4* ib4 = reinterpret_cast<const int4*>(ib);
    for (int g = (l0 >> 2) + tid; g < (l1 >> 2); g += GRP_THREADS) {
      const int4 ii = __ldg(ib4 + g);
      int m0 = 0, m1 = 0, m2 = 0, m3 = 0;
      if (cb) { const int l = g << 2; m0 = l / a.k; m1 = (l + 1) / a.k; m2 = (l + 2) / a.k; m3 = (l + 3) / a.k; }
#pragma unroll 4
      for (int c = 0; c < tc; ++c) {
        const float* r = rows + (size_t)c * a.N;
        float4 v;
        if (SMEM) { v.x = r[ii.x]; v.y = r[ii.y]; v.z = r[ii.z]; v.w = r[ii.w]; }
        else { v.x = __ldg(r + ii.x); v.y = __ldg(r + ii.y); v.z = __ldg(r + ii.z); v.w = __ldg(r + ii.w); }
        if (cb) {
          const float* cc = cb + (size_t)c * a.M;
          v.x = __fsub_rn(v.x, __ldg(cc + m0)); v.y = __fsub_rn(v.y, __ldg(cc + m1));
          v.z = __fsub_rn(v.z, __ldg(cc + m2)); v.w = __fsub_rn(v.w, __ldg(cc + m3));
        }
        reinterpret_cast<float4*>(ob + (size_t)c * a.L)[g] = v;
      }
    }
  } else {
    for (int l = l0 + tid; l < l1; l += GRP_THREADS) {
      const int i = __ldg(ib + l);
      const int m = cb ? l / a.k : 0;
      for (int c = 0; c < tc; ++c) {
        float v = SMEM ? rows[(size_t)c * a.N + i] : __ldg(rows + (size_t)c * a.N + i);
        if (cb) v = __fsub_rn(v, __ldg(cb + (size_t)c * a.M + m));
        ob[(size_t)c * a.L + l] = v;
      }
    }
  }
}

// ---- grouping forward for rows that do not fit shared memory (N > 24576) ---------------------------------
// Gathering 4-byte values from a 256 KB row pulls a 32-byte L2 sector per element (13-16 % of the HBM peak).
// Here the features are first transposed to point-major [B, N, C] (one tiled pass over C*N values, tiny next to
// the k-times larger output); a gather then reads the CONTIGUOUS channel row of a neighbour (128 bytes per warp
// request, every byte used) into a shared-memory tile [CT][LT] and the tile is written out along l with
// coalesced 16-byte stores.
constexpr int GT_LT = 128;   // positions per tile
constexpr int GT_CT = 64;    // channels per tile

__global__ void __launch_bounds__(256) transpose_cn_kernel(const float* __restrict__ f, int C, int N, float* __restrict__ ft) {
  __shared__ float t[32][33];
  const int b = blockIdx.z, n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, n = n0 + tx;
    t[r][tx] = (c < C && n < N) ? __ldg(f + ((size_t)b * C + c) * N + n) : 0.0f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int n = n0 + r, c = c0 + tx;
    if (n < N && c < C) ft[((size_t)b * N + n) * C + c] = t[tx][r];
  }
}

struct GroupFwdTArgs {
  const float* ft;       // [B,N,C] point-major copy
  const int32_t* idx;    // [B,L]
  const float* center;   // [B,C,M] or null
  int B, C, N, M, k, L;
  float* out;            // [B,C,L]
};

__global__ void __launch_bounds__(256) group_fwd_pointmajor_kernel(GroupFwdTArgs a) {
  __shared__ float tile[GT_LT][GT_CT + 1];  // position-major: conflict-free both ways (stride 65 words)
  const int b = blockIdx.z, c0 = blockIdx.y * GT_CT, l0 = blockIdx.x * GT_LT;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ct = min(GT_CT, a.C - c0), lt = min(GT_LT, a.L - l0);
  const int32_t* ib = a.idx + (size_t)b * a.L + l0;
  const float* fb = a.ft + (size_t)b * a.N * a.C + c0;
  // gather: one warp per position, lanes over channels (2 x 128 B for 64 channels), 4 positions in flight
  for (int p0 = warp * 16; p0 < warp * 16 + 16; p0 += 4) {
    int ii[4];
    float v[4][2];
#pragma unroll
    for (int u = 0; u < 4; ++u) ii[u] = (p0 + u < lt) ? __ldg(ib + p0 + u) : 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float* r = fb + (size_t)ii[u] * a.C;
      v[u][0] = lane < ct ? __ldg(r + lane) : 0.0f;
      v[u][1] = lane + 32 < ct ? __ldg(r + lane + 32) : 0.0f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      tile[p0 + u][lane] = v[u][0];
      tile[p0 + u][lane + 32] = v[u][1];
    }
  }
  __syncthreads();
  // write out: one row of lt floats per channel, coalesced along l (a warp stores 128 contiguous bytes)
  float* ob = a.out + ((size_t)b * a.C + c0) * a.L + l0;
  for (int e = threadIdx.x; e < ct * GT_LT; e += 256) {
    const int c = e / GT_LT, l = e - c * GT_LT;
    if (l < lt) {
      float val = tile[l][c];
      if (a.center) val = __fsub_rn(val, __ldg(a.center + ((size_t)b * a.C + c0 + c) * a.M + (l0 + l) / a.k));
      ob[(size_t)c * a.L + l] = val;
    }
  }
}

// ---- fused gather + reduce over k (also three_interpolate when w != null) --------
struct ReduceFwdArgs {
  const float* f;      // [B,C,N]
  const int32_t* idx;  // [B,M,k]
  const float* w;      // [B,M,k] or null (weighted sum)
  int B, C, N, M, k, op, TC, MT;
  float* out;          // [B,C,M]
  int32_t* arg;        // [B,C,M] or null
};

// Fused gather + reduce, shared-memory version (K7 v2).  A CTA owns (cloud b, TC channels, a range of output
// points).  The TC feature rows sit in shared memory CHANNEL-QUAD INTERLEAVED — [TC/4][N][4] — so one LDS.128
// fetches four channels of a neighbour: random point indices then cost ~2.2 bank-conflict cycles per 8-lane phase
// for 128 useful words instead of ~3.5 per 32 words with scalar gathers.  Every thread reduces ONE output point over
// its k neighbours for all TC channels; its neighbour list is read straight from global memory with 16-byte loads
// (a thread's list is contiguous, neighbouring threads' lists share cache lines), so each index is read once and
// serves TC gathers; outputs are written coalesced along m.  (The first version read the indices with scalar loads
// at stride 4k bytes across the warp, served at most 8 channels per pass and gathered 4 bytes per LDS.)
constexpr int RT_THREADS = 512;
constexpr size_t RT_SMEM_MAX = 200 * 1024;

template <int TC>  // 4, 8 or 16 channels per tile (rows beyond C are zero-filled)
__global__ void __launch_bounds__(RT_THREADS, 1) group_reduce_fwd_smem_kernel(ReduceFwdArgs a) {
  extern __shared__ __align__(16) float rows_s[];  // [TC/4][N][4]
  constexpr int Q = TC / 4;
  const int b = blockIdx.z, c0 = blockIdx.y * TC, tid = threadIdx.x;
  const int tc = min(TC, a.C - c0);
  {
    // four coalesced row loads (one per channel of the quad) -> one conflict-free STS.128 per point
    const float* fb = a.f + ((size_t)b * a.C + c0) * a.N;
    float4* dst = reinterpret_cast<float4*>(rows_s);
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const float* r0 = fb + (size_t)(4 * q) * a.N;
      const bool h0 = 4 * q < tc, h1 = 4 * q + 1 < tc, h2 = 4 * q + 2 < tc, h3 = 4 * q + 3 < tc;
      for (int n = tid; n < a.N; n += RT_THREADS) {
        float4 v;
        v.x = h0 ? __ldg(r0 + n) : 0.0f;
        v.y = h1 ? __ldg(r0 + a.N + n) : 0.0f;
        v.z = h2 ? __ldg(r0 + 2 * (size_t)a.N + n) : 0.0f;
        v.w = h3 ? __ldg(r0 + 3 * (size_t)a.N + n) : 0.0f;
        dst[(size_t)q * a.N + n] = v;
      }
    }
  }
  __syncthreads();
  const float4* rows4 = reinterpret_cast<const float4*>(rows_s);
  const int m_lo = blockIdx.x * a.MT, m_hi = min(a.M, m_lo + a.MT);
  const bool vec = ((a.k & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.idx) & 15) == 0);
  for (int m = m_lo + tid; m < m_hi; m += RT_THREADS) {
    const int32_t* im = a.idx + ((size_t)b * a.M + m) * a.k;
    const float* wm = a.w ? a.w + ((size_t)b * a.M + m) * a.k : nullptr;
    float acc[TC];
    int best[TC];
    auto take = [&](int j, int i) {
      float4 v[Q];
#pragma unroll
      for (int q = 0; q < Q; ++q) v[q] = rows4[(size_t)q * a.N + i];
      const float* vf = reinterpret_cast<const float*>(v);
      if (j == 0) {
        const float w0 = wm ? __ldg(wm) : 1.0f;
#pragma unroll
        for (int c = 0; c < TC; ++c) { acc[c] = wm ? __fmul_rn(w0, vf[c]) : vf[c]; best[c] = 0; }
      } else if (wm) {
        const float wj = __ldg(wm + j);
#pragma unroll
        for (int c = 0; c < TC; ++c) acc[c] = __fadd_rn(acc[c], __fmul_rn(wj, vf[c]));
      } else if (a.op == TPG_REDUCE_MAX) {
#pragma unroll
        for (int c = 0; c < TC; ++c) if (vf[c] > acc[c]) { acc[c] = vf[c]; best[c] = j; }
      } else if (a.op == TPG_REDUCE_MIN) {
#pragma unroll
        for (int c = 0; c < TC; ++c) if (vf[c] < acc[c]) { acc[c] = vf[c]; best[c] = j; }
      } else {
#pragma unroll
        for (int c = 0; c < TC; ++c) acc[c] = __fadd_rn(acc[c], vf[c]);
      }
    };
    if (vec) {
      const int4* im4 = reinterpret_cast<const int4*>(im);
      for (int j4 = 0; j4 < (a.k >> 2); ++j4) {
        const int4 ii = __ldg(im4 + j4);
        take(4 * j4, ii.x); take(4 * j4 + 1, ii.y); take(4 * j4 + 2, ii.z); take(4 * j4 + 3, ii.w);
      }
    } else {
      for (int j = 0; j < a.k; ++j) take(j, __ldg(im + j));
    }
#pragma unroll
    for (int c = 0; c < TC; ++c) {
      if (c < tc) {
        const size_t o = ((size_t)b * a.C + c0 + c) * a.M + m;
        a.out[o] = acc[c];
        if (a.arg) a.arg[o] = best[c];
      }
    }
  }
}

// Fused gather + reduce for rows that do not fit shared memory (N > 12800): point-major features [B,N,C] (the
// transposed copy of transpose_cn_kernel); one warp per output point, lanes over channels, so every neighbour costs one
// or two coalesced 128-byte row reads; results of 32 points are staged in a [64][33] tile and written along m.
struct ReduceFwdTArgs {
  const float* ft;     // [B,N,C]
  const int32_t* idx;  // [B,M,k]
  const float* w;      // [B,M,k] or null
  int B, C, N, M, k, op;
  float* out;          // [B,C,M]
  int32_t* arg;        // [B,C,M] or null
};

__global__ void __launch_bounds__(256) group_reduce_pointmajor_kernel(ReduceFwdTArgs a) {
  __shared__ float tile_v[64][33];
  __shared__ int tile_a[64][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 64, m0 = blockIdx.x * 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ct = min(64, a.C - c0), mt = min(32, a.M - m0);
  const float* fb = a.ft + (size_t)b * a.N * a.C + c0;
  const bool h0 = lane < ct, h1 = lane + 32 < ct;
  for (int p = warp * 4; p < warp * 4 + 4; ++p) {
    if (p >= mt) break;  // (warp-uniform)
    const int32_t* im = a.idx + ((size_t)b * a.M + m0 + p) * a.k;
    const float* wm = a.w ? a.w + ((size_t)b * a.M + m0 + p) * a.k : nullptr;
    float acc0 = 0.f, acc1 = 0.f;
    int b0 = 0, b1 = 0;
    for (int j0 = 0; j0 < a.k; j0 += 8) {
      float v0[8], v1[8], wj[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = j0 + u;
        v0[u] = v1[u] = 0.f; wj[u] = 1.f;
        if (j < a.k) {
          const float* r = fb + (size_t)__ldg(im + j) * a.C;
          if (h0) v0[u] = __ldg(r + lane);
          if (h1) v1[u] = __ldg(r + lane + 32);
          if (wm) wj[u] = __ldg(wm + j);
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = j0 + u;
        if (j >= a.k) break;
        if (wm) {
          const float t0 = __fmul_rn(wj[u], v0[u]), t1 = __fmul_rn(wj[u], v1[u]);
          acc0 = j == 0 ? t0 : __fadd_rn(acc0, t0);
          acc1 = j == 0 ? t1 : __fadd_rn(acc1, t1);
        } else if (j == 0) {
          acc0 = v0[u]; acc1 = v1[u];
        } else if (a.op == TPG_REDUCE_MAX) {
          if (v0[u] > acc0) { acc0 = v0[u]; b0 = j; }
          if (v1[u] > acc1) { acc1 = v1[u]; b1 = j; }
        } else if (a.op == TPG_REDUCE_MIN) {
          if (v0[u] < acc0) { acc0 = v0[u]; b0 = j; }
          if (v1[u] < acc1) { acc1 = v1[u]; b1 = j; }
        } else {
          acc0 = __fadd_rn(acc0, v0[u]); acc1 = __fadd_rn(acc1, v1[u]);
        }
      }
    }
    tile_v[lane][p] = acc0; tile_v[lane + 32][p] = acc1;
    tile_a[lane][p] = b0; tile_a[lane + 32][p] = b1;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < ct * 32; e += 256) {
    const int c = e >> 5, p = e & 31;
    if (p < mt) {
      const size_t o = ((size_t)b * a.C + c0 + c) * a.M + m0 + p;
      a.out[o] = tile_v[c][p];
      if (a.arg) a.arg[o] = tile_a[c][p];
    }
  }
}

template <bool SMEM>
__global__ void __launch_bounds__(GRP_THREADS) group_reduce_fwd_kernel(ReduceFwdArgs a) {
  extern __shared__ float rows_s[];
  const int b = blockIdx.z, c0 = blockIdx.y * a.TC, tid = threadIdx.x;
  const int tc = min(a.TC, a.C - c0);
  const float* fb = a.f + ((size_t)b * a.C + c0) * a.N;
  if (SMEM) {
    const int total = tc * a.N;
    for (int e = tid; e < total; e += GRP_THREADS) rows_s[e] = __ldg(fb + e);
    __syncthreads();
  }
  const float* rows = SMEM ? rows_s : fb;
  const int m0 = blockIdx.x * a.MT, m1 = min(a.M, m0 + a.MT);
  constexpr int TCMAX = 8;
  for (int m = m0 + tid; m < m1; m += GRP_THREADS) {
    const int32_t* im = a.idx + ((size_t)b * a.M + m) * a.k;
    const float* wm = a.w ? a.w + ((size_t)b * a.M + m) * a.k : nullptr;
    // neighbour-major: every index (and weight) is loaded once and serves all the channels of the tile
    float acc[TCMAX];
    int best[TCMAX];
    {
      const int i0 = __ldg(im);
      const float w0 = wm ? __ldg(wm) : 1.0f;
#pragma unroll
      for (int c = 0; c < TCMAX; ++c) {
        float v = 0.0f;
        if (c < tc) v = SMEM ? rows[(size_t)c * a.N + i0] : __ldg(rows + (size_t)c * a.N + i0);
        acc[c] = wm ? __fmul_rn(w0, v) : v;
        best[c] = 0;
      }
    }
    for (int j = 1; j < a.k; ++j) {
      const int i = __ldg(im + j);
      const float wj = wm ? __ldg(wm + j) : 1.0f;
#pragma unroll
      for (int c = 0; c < TCMAX; ++c) {
        if (c < tc) {
          const float v = SMEM ? rows[(size_t)c * a.N + i] : __ldg(rows + (size_t)c * a.N + i);
          if (wm) acc[c] = __fadd_rn(acc[c], __fmul_rn(wj, v));
          else if (a.op == TPG_REDUCE_MAX) { if (v > acc[c]) { acc[c] = v; best[c] = j; } }
          else if (a.op == TPG_REDUCE_MIN) { if (v < acc[c]) { acc[c] = v; best[c] = j; } }
          else acc[c] = __fadd_rn(acc[c], v);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < TCMAX; ++c) {
      if (c < tc) {
        const size_t o = ((size_t)b * a.C + c0 + c) * a.M + m;
        a.out[o] = acc[c];
        if (a.arg) a.arg[o] = best[c];
      }
    }
  }
}

// ---- inverse index (CSR) ------------------------------------------------------------
__global__ void csr_count_kernel(const int32_t* __restrict__ idx, const int64_t* __restrict__ item_len,
                                 int B, int N, int L, int32_t* off) {
  const long long total = (long long)B * L;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(e / L);
    if (item_len && (e - (long long)b * L) >= item_len[b]) continue;
    const int key = idx[e];
    if (key >= 0 && key < N) atomicAdd(off + (size_t)b * (N + 1) + key + 1, 1);
  }
}

// in-place inclusive scan of off[b][0..N] (off[b][0] == 0), one CTA per cloud; also
// seeds the fill cursors with the segment starts.
__global__ void __launch_bounds__(1024) csr_scan_kernel(int32_t* off, int32_t* cursor, int N) {
  __shared__ int warp_sums[32];
  __shared__ int carry_s;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int32_t* o = off + (size_t)b * (N + 1);
  int32_t* cur = cursor + (size_t)b * N;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < N + 1; base += 1024) {
    const int e = base + tid;
    int v = e < N + 1 ? o[e] : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int t = __shfl_up_sync(FULL, v, d);
      if (lane >= d) v += t;
    }
    if (lane == 31) warp_sums[warp] = v;
    __syncthreads();
    if (warp == 0) {
      int w = warp_sums[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(FULL, w, d);
        if (lane >= d) w += t;
      }
      warp_sums[lane] = w;
    }
    __syncthreads();
    const int carry = carry_s;
    v += carry + (warp > 0 ? warp_sums[warp - 1] : 0);
    if (e < N + 1) {
      o[e] = v;
      if (e < N) cur[e] = v;  // start of segment e+1 ... fixed below
    }
    __syncthreads();
    if (tid == 1023) carry_s = v;
    __syncthreads();
  }
  // cursor[n] must be the START of segment n == off[n]; the loop stored off[n] at
  // cur[n] already (inclusive scan value at position n is the sum of counts < n
  // because counts live at key+1).
}

__global__ void csr_fill_kernel(const int32_t* __restrict__ idx, const int64_t* __restrict__ item_len,
                                int B, int N, int L, int32_t* cursor, int32_t* tmp_items) {
  const long long total = (long long)B * L;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(e / L);
    const int l = (int)(e - (long long)b * L);
    if (item_len && l >= item_len[b]) continue;
    const int key = idx[e];
    if (key >= 0 && key < N) {
      const int pos = atomicAdd(cursor + (size_t)b * N + key, 1);
      tmp_items[(size_t)b * L + pos] = l;
    }
  }
}

// one warp per segment: rank-sort the (distinct) positions ascending, out of place.
__global__ void csr_sort_kernel(const int32_t* __restrict__ off, const int32_t* __restrict__ tmp_items,
                                int B, int N, int L, int32_t* __restrict__ items) {
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= (long long)B * N) return;
  const int b = (int)(w / N), n = (int)(w - (long long)b * N);
  const int s0 = off[(size_t)b * (N + 1) + n], s1 = off[(size_t)b * (N + 1) + n + 1];
  const int s = s1 - s0;
  if (s == 0) return;
  const int32_t* src = tmp_items + (size_t)b * L + s0;
  int32_t* dst = items + (size_t)b * L + s0;
  if (s <= 32) {
    const int x = lane < s ? src[lane] : 0x7fffffff;
    int rank = 0;
    for (int t = 0; t < s; ++t) rank += (__shfl_sync(FULL, x, t) < x) ? 1 : 0;
    if (lane < s) dst[rank] = x;
  } else {
    for (int e = lane; e < s; e += 32) {
      const int x = src[e];
      int rank = 0;
      for (int t = 0; t < s; ++t) rank += (__ldg(src + t) < x) ? 1 : 0;
      dst[rank] = x;
    }
  }
}

// ---- inverse index, stable single-kernel build (one CTA per cloud) ---------------------------
// Positions are split into W contiguous chunks, one per warp.  A warp walks its chunk 32
// positions at a time; MATCH.ANY groups the lanes that hold the same key, so one lane per key
// updates the warp-private counter cnt[w][key] — no atomics anywhere.  After an exclusive scan
// over (key, chunk) the second walk hands out slots in ascending position order directly
// (rank inside a MATCH group = lanes below with the same key), so no sort pass is needed and
// the result is deterministic.  CT = uint16 when a chunk cannot overflow it, else uint32.
// lanes of the warp that hold the same key.  MATCH.ANY costs ~10 cycles per distinct value, so it
// only runs when a duplicate exists among the 32 keys; duplicates are detected through a per-warp
// tag array: every lane writes its id to tag[key] and a lane that reads back another id lost.
__device__ __forceinline__ unsigned same_key_mask(unsigned char* tag, int key, bool valid, int lane, int N) {
  if (valid) tag[key] = (unsigned char)lane;
  __syncwarp();
  const bool lost = valid && tag[key] != (unsigned char)lane;
  if (__ballot_sync(FULL, lost) == 0u) return 1u << lane;
  return __match_any_sync(FULL, valid ? key : N + lane);
}

template <typename CT>
__global__ void __launch_bounds__(1024) csr_stable_kernel(const int32_t* __restrict__ idx,
                                                          const int64_t* __restrict__ item_len, int N, int L,
                                                          int32_t* __restrict__ off_g, int32_t* __restrict__ items_g) {
  extern __shared__ __align__(16) unsigned char csr_smem[];
  __shared__ int warp_tot[32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, T = blockDim.x, W = T >> 5;
  int32_t* off_s = reinterpret_cast<int32_t*>(csr_smem);                 // [N + 1]
  CT* cnt = reinterpret_cast<CT*>(csr_smem + sizeof(int32_t) * (size_t)((N + 1 + 3) & ~3));  // [W][N]
  const int Lb = item_len ? (int)min((long long)item_len[b], (long long)L) : L;
  const int32_t* ib = idx + (size_t)b * L;
  int32_t* itb = items_g + (size_t)b * L;
  int32_t* ob = off_g + (size_t)b * (N + 1);
  for (int e = tid; e < W * N; e += T) cnt[e] = 0;
  __syncthreads();
  const int CL = ((Lb + W - 1) / W + 127) & ~127;
  const int l_lo = warp * CL, l_hi = min(Lb, l_lo + CL);
  CT* mycnt = cnt + (size_t)warp * N;
  unsigned char* mytag = reinterpret_cast<unsigned char*>(cnt + (size_t)W * N) + (size_t)warp * N;  // [W][N]
  const unsigned lt = (1u << lane) - 1u;
  for (int l0 = l_lo; l0 < l_hi; l0 += 128) {  // 4 independent loads in flight per lane
    int keys[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int l = l0 + u * 32 + lane;
      keys[u] = l < l_hi ? __ldg(ib + l) : -1;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int key = keys[u];
      const bool valid = key >= 0 && key < N;
      const unsigned m = same_key_mask(mytag, key, valid, lane, N);
      if (valid && (m & lt) == 0) mycnt[key] = (CT)(mycnt[key] + __popc(m));
      __syncwarp();
    }
  }
  __syncthreads();
  // exclusive scan over keys of the per-key totals; cnt[w][key] becomes the prefix over chunks
  const int KPT = (N + T - 1) / T;
  const int k0 = tid * KPT, k1 = min(N, k0 + KPT);
  int mine = 0;
  for (int key = k0; key < k1; ++key) {
    int tot = 0;
    for (int w = 0; w < W; ++w) {
      const int c = cnt[(size_t)w * N + key];
      cnt[(size_t)w * N + key] = (CT)tot;
      tot += c;
    }
    off_s[key] = tot;  // per-key total for now
    mine += tot;
  }
  int v = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(FULL, v, d);
    if (lane >= d) v += t;
  }
  if (lane == 31) warp_tot[warp] = v;
  __syncthreads();
  if (warp == 0) {
    int w = lane < W ? warp_tot[lane] : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(FULL, w, d);
      if (lane >= d) w += t;
    }
    warp_tot[lane] = w;
  }
  __syncthreads();
  int run = v - mine + (warp > 0 ? warp_tot[warp - 1] : 0);  // exclusive prefix of this thread's first key
  for (int key = k0; key < k1; ++key) {
    const int tot = off_s[key];
    off_s[key] = run;
    ob[key] = run;
    run += tot;
  }
  if (k1 == N && k0 < N) ob[N] = run;
  if (N == 0 && tid == 0) ob[0] = 0;
  __syncthreads();
  for (int l0 = l_lo; l0 < l_hi; l0 += 128) {
    int keys[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int l = l0 + u * 32 + lane;
      keys[u] = l < l_hi ? __ldg(ib + l) : -1;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int key = keys[u];
      const int l = l0 + u * 32 + lane;
      const bool valid = key >= 0 && key < N;
      const unsigned m = same_key_mask(mytag, key, valid, lane, N);
      int base = 0;
      if (valid) base = off_s[key] + (int)mycnt[key];
      __syncwarp();
      if (valid) {
        itb[base + __popc(m & lt)] = l;
        if ((m & lt) == 0) mycnt[key] = (CT)(mycnt[key] + __popc(m));
      }
      __syncwarp();
    }
  }
}

// ---- inverse index, stable build over a thread-block cluster (Q CTAs per cloud) ---------------
// Same scheme as csr_stable_kernel, with the positions of a cloud cut into Q * 32 contiguous chunks
// (CTA r of the cluster owns chunks r*32 .. r*32+31, one per warp), so a warp walks L / (32 Q)
// positions instead of L / 32.  The only cross-CTA step is the per-key prefix over CTAs: every CTA
// publishes its per-key totals in shared memory, and after one cluster barrier each CTA reads its
// peers' totals through DSMEM.  Deterministic, stable, no atomics.
template <typename CT>
__global__ void __launch_bounds__(1024, 1) csr_cluster_kernel(const int32_t* __restrict__ idx,
                                                              const int64_t* __restrict__ item_len, int N, int L, int Q,
                                                              int32_t* __restrict__ off_g,
                                                              int32_t* __restrict__ items_g) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) unsigned char csr_smem[];
  __shared__ int warp_tot[32];
  constexpr int W = 32, T = 1024;
  const int r = (int)cluster.block_rank(), b = blockIdx.x / Q;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Np = (N + 3) & ~3;
  int32_t* off_s = reinterpret_cast<int32_t*>(csr_smem);        // [Np]  slot of the first item of (this CTA, key)
  int32_t* tot_s = off_s + Np;                                   // [Np]  per-key count of this CTA (read by peers)
  CT* cnt = reinterpret_cast<CT*>(tot_s + Np);                   // [W][N]
  unsigned char* tag = reinterpret_cast<unsigned char*>(cnt + (size_t)W * N);  // [W][N]
  const int Lb = item_len ? (int)min((long long)item_len[b], (long long)L) : L;
  const int32_t* ib = idx + (size_t)b * L;
  int32_t* itb = items_g + (size_t)b * L;
  int32_t* ob = off_g + (size_t)b * (N + 1);
  {
    uint32_t* z = reinterpret_cast<uint32_t*>(cnt);
    const int words = (int)(((size_t)W * N * sizeof(CT) + 3) / 4);
    for (int e = tid; e < words; e += T) z[e] = 0u;
  }
  __syncthreads();
  const int CL = (((Lb + Q * W - 1) / (Q * W)) + 31) & ~31;
  const int l_lo = min(Lb, (r * W + warp) * CL), l_hi = min(Lb, l_lo + CL);
  CT* mycnt = cnt + (size_t)warp * N;
  unsigned char* mytag = tag + (size_t)warp * N;
  const unsigned lt = (1u << lane) - 1u;
  for (int l0 = l_lo; l0 < l_hi; l0 += 128) {
    int keys[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int l = l0 + u * 32 + lane;
      keys[u] = l < l_hi ? __ldg(ib + l) : -1;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (l0 + u * 32 >= l_hi) break;  // (warp-uniform)
      const int key = keys[u];
      const bool valid = key >= 0 && key < N;
      const unsigned m = same_key_mask(mytag, key, valid, lane, N);
      if (valid && (m & lt) == 0) mycnt[key] = (CT)(mycnt[key] + __popc(m));
      __syncwarp();
    }
  }
  __syncthreads();
  // per key: exclusive prefix over this CTA's warps, CTA total published for the peers
  const int KPT = (N + T - 1) / T;  // <= 3
  const int k0 = tid * KPT, k1 = min(N, k0 + KPT);
  for (int key = k0; key < k1; ++key) {
    int tot = 0;
#pragma unroll 8
    for (int w = 0; w < W; ++w) {
      const int c = cnt[(size_t)w * N + key];
      cnt[(size_t)w * N + key] = (CT)tot;
      tot += c;
    }
    tot_s[key] = tot;
  }
  cluster.sync();
  int ktot[3], kbase[3], mine = 0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    ktot[i] = kbase[i] = 0;
    const int key = k0 + i;
    if (i < KPT && key < k1) {
      for (int q = 0; q < Q; ++q) {
        const int t = cluster.map_shared_rank(tot_s, q)[key];
        ktot[i] += t;
        if (q < r) kbase[i] += t;
      }
      mine += ktot[i];
    }
  }
  cluster.barrier_arrive();  // this CTA no longer reads its peers' shared memory
  int v = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(FULL, v, d);
    if (lane >= d) v += t;
  }
  if (lane == 31) warp_tot[warp] = v;
  __syncthreads();
  if (warp == 0) {
    int w = warp_tot[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(FULL, w, d);
      if (lane >= d) w += t;
    }
    warp_tot[lane] = w;
  }
  __syncthreads();
  int run = v - mine + (warp > 0 ? warp_tot[warp - 1] : 0);  // exclusive prefix of this thread's first key
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int key = k0 + i;
    if (i < KPT && key < k1) {
      off_s[key] = run + kbase[i];
      if (r == 0) ob[key] = run;
      run += ktot[i];
    }
  }
  if (r == 0 && k1 == N && k0 < N) ob[N] = run;
  __syncthreads();
  for (int l0 = l_lo; l0 < l_hi; l0 += 128) {
    int keys[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int l = l0 + u * 32 + lane;
      keys[u] = l < l_hi ? __ldg(ib + l) : -1;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (l0 + u * 32 >= l_hi) break;
      const int key = keys[u];
      const int l = l0 + u * 32 + lane;
      const bool valid = key >= 0 && key < N;
      const unsigned m = same_key_mask(mytag, key, valid, lane, N);
      int base = 0;
      if (valid) base = off_s[key] + (int)mycnt[key];
      __syncwarp();
      if (valid) {
        itb[base + __popc(m & lt)] = l;
        if ((m & lt) == 0) mycnt[key] = (CT)(mycnt[key] + __popc(m));
      }
      __syncwarp();
    }
  }
  cluster.barrier_wait();  // no CTA may exit while a peer still reads its totals
}

// cluster size for the cluster build (0 = not eligible)
static int csr_cluster_size(int N, int L, bool* wide, size_t* smem) {
  if (N < 1 || N > 2048 || L < 2048) return 0;
  int Q = 8;
  while (Q > 1 && L < 1024 * Q) Q >>= 1;
  const int CL = (((L + Q * 32 - 1) / (Q * 32)) + 31) & ~31;
  // cnt[w][key] ends up holding the prefix over the CTA's 32 warps: up to 32 * CL
  *wide = 32LL * CL > 65535;
  *smem = sizeof(int32_t) * 2 * (size_t)((N + 3) & ~3) + (size_t)32 * N * ((*wide ? 4 : 2) + 1);
  return *smem <= 220 * 1024 ? Q : 0;
}

template <typename CT>
static int launch_csr_cluster(const int32_t* idx, const int64_t* item_len, int B, int N, int L, int Q, size_t smem,
                              int32_t* off, int32_t* items, cudaStream_t st) {
  auto kern = csr_cluster_kernel<CT>;
  // per call: the attribute is per device, and a process may drive several
  TPG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(B * Q));
  cfg.blockDim = dim3(1024);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)Q;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  TPG_CUDA(cudaLaunchKernelEx(&cfg, kern, idx, item_len, N, L, Q, off, items));
  return TPG_OK;
}

// warps per cloud for the stable build (0 = does not fit: use the count/scan/fill/sort path)
static int csr_stable_warps(int N, int L, bool* wide) {
  *wide = false;
  for (int W = 32; W >= 4; W >>= 1) {
    const long long chunk = ((long long)(L + W - 1) / W + 127) & ~127LL;
    // the kernel overwrites a warp's per-key counter with the exclusive prefix over ALL warps, which can reach the
    // whole row (one key repeated L times): 16-bit counters only when that total fits
    const bool w32 = (long long)W * chunk > 65535;
    const size_t bytes = sizeof(int32_t) * (size_t)((N + 1 + 3) & ~3) + (size_t)W * N * (w32 ? 5 : 3);
    if (bytes <= 220 * 1024) { *wide = w32; return W; }
  }
  return 0;
}

// ---- backward: per-source-point sequential sums over the CSR ---------------------
enum { BWD_GROUP = 0, BWD_ARG = 1, BWD_SUM = 2, BWD_WEIGHTED = 3 };

struct BwdArgs {
  const float* go;        // GROUP: [B,C,L]; others: [B,C,M]
  const int32_t* arg;     // BWD_ARG: [B,C,M]
  const float* w;         // BWD_WEIGHTED: [B,L]
  const int32_t* off;     // [B,N+1]
  const int32_t* items;   // [B,L]
  int B, C, N, M, k, L, mode;
  float* gf;              // [B,C,N]
};

__global__ void __launch_bounds__(GRP_THREADS) group_bwd_kernel(BwdArgs a) {
  const int b = blockIdx.z, c = blockIdx.y;
  const int n = blockIdx.x * GRP_THREADS + threadIdx.x;
  if (n >= a.N) return;
  const int s0 = a.off[(size_t)b * (a.N + 1) + n], s1 = a.off[(size_t)b * (a.N + 1) + n + 1];
  const int32_t* it = a.items + (size_t)b * a.L;
  float acc = 0.0f;
  if (a.mode == BWD_GROUP) {
    const float* go = a.go + ((size_t)b * a.C + c) * a.L;
    for (int p = s0; p < s1; ++p) acc = __fadd_rn(acc, __ldg(go + it[p]));
  } else {
    const float* go = a.go + ((size_t)b * a.C + c) * a.M;
    for (int p = s0; p < s1; ++p) {
      const int l = it[p];
      const int m = l / a.k;
      if (a.mode == BWD_ARG) {
        if (a.arg[((size_t)b * a.C + c) * a.M + m] == l - m * a.k) acc = __fadd_rn(acc, __ldg(go + m));
      } else if (a.mode == BWD_SUM) {
        acc = __fadd_rn(acc, __ldg(go + m));
      } else {
        acc = __fadd_rn(acc, __fmul_rn(__ldg(go + m), __ldg(a.w + (size_t)b * a.L + l)));
      }
    }
  }
  a.gf[((size_t)b * a.C + c) * a.N + n] = acc;
}

static void pick_tiles(int B, int C, int N, int L, int& TC, int& LT, bool& smem) {
  const int target = 4 * num_sms();
  const int max_tc = (int)(GRP_SMEM_MAX / ((size_t)N * sizeof(float)));
  smem = max_tc >= 1;
  TC = 1;
  const int base_lt = max(1, L / max(1, 4 * N));  // l-tiles that keep row loads <= 25% of stores
  for (int t = 8; t >= 1; t >>= 1) {
    if (t > C && t > 1) continue;
    if (smem && t > max_tc) continue;
    TC = t;
    if ((long long)B * ceil_div(C, t) * base_lt >= target) break;
  }
  int lt = max(base_lt, ceil_div(target, B * ceil_div(C, TC)));
  lt = min(lt, max(1, L / 1024));
  LT = ceil_div(ceil_div(L, lt), 1024) * 1024;
}

}  // namespace tpg

using namespace tpg;

TPG_API size_t tpg_group_fwd_workspace_bytes(int B, int C, int N, int M, int k) {
  (void)M; (void)k;
  // rows that do not fit shared memory: point-major copy of the features
  return ((size_t)N * sizeof(float) > GRP_SMEM_MAX && C >= 16) ? sizeof(float) * (size_t)B * (size_t)C * (size_t)N : 0;
}

TPG_API int tpg_group_fwd_ws_f32(const float* f, const int32_t* idx, const float* center, int B, int C, int N, int M,
                                 int k, float* out, void* workspace, size_t workspace_bytes, tpg_stream_t stream) {
  const size_t need = tpg_group_fwd_workspace_bytes(B, C, N, M, k);
  if (need == 0 || !workspace || workspace_bytes < need) return tpg_group_fwd_f32(f, idx, center, B, C, N, M, k, out, stream);
  TPG_REQUIRE(B >= 0 && C >= 0 && N >= 1 && M >= 0 && k >= 0, TPG_EINVAL, "group_fwd: bad size");
  const long long L64 = (long long)M * k;
  TPG_REQUIRE(L64 < (1LL << 31), TPG_EUNSUPPORTED, "group_fwd: M*k too large");
  if (B == 0 || C == 0 || L64 == 0) return TPG_OK;
  TPG_REQUIRE(f && idx && out, TPG_EINVAL, "group_fwd: null pointer");
  TPG_REQUIRE(B <= 65535 && ceil_div(C, 32) <= 65535, TPG_EUNSUPPORTED, "group_fwd: B or C too large");
  cudaStream_t st = as_stream(stream);
  float* ft = reinterpret_cast<float*>(workspace);
  transpose_cn_kernel<<<dim3(ceil_div(N, 32), ceil_div(C, 32), B), 256, 0, st>>>(f, C, N, ft);
  TPG_CHECK_LAUNCH("transpose_cn_kernel");
  GroupFwdTArgs a{ft, idx, center, B, C, N, M, k, (int)L64, out};
  group_fwd_pointmajor_kernel<<<dim3(ceil_div((int)L64, GT_LT), ceil_div(C, GT_CT), B), 256, 0, st>>>(a);
  TPG_CHECK_LAUNCH("group_fwd_pointmajor_kernel");
  return TPG_OK;
}

TPG_API int tpg_group_fwd_f32(const float* f, const int32_t* idx, const float* center, int B, int C, int N,
                              int M, int k, float* out, tpg_stream_t stream) {
  TPG_REQUIRE(B >= 0 && C >= 0 && N >= 0 && M >= 0 && k >= 0, TPG_EINVAL, "group_fwd: negative size");
  const long long L64 = (long long)M * k;
  TPG_REQUIRE(L64 < (1LL << 31), TPG_EUNSUPPORTED, "group_fwd: M*k too large");
  if (B == 0 || C == 0 || L64 == 0) return TPG_OK;
  TPG_REQUIRE(N >= 1, TPG_EINVAL, "group_fwd: empty source cloud");
  TPG_REQUIRE(f && idx && out, TPG_EINVAL, "group_fwd: null pointer");
  TPG_REQUIRE(B <= 65535, TPG_EUNSUPPORTED, "group_fwd: B > 65535");
  GroupFwdArgs a{f, idx, center, B, C, N, M, k, (int)L64, 1, 1024, 0, out};
  a.rows_vec = ((N & 3) == 0) && ((reinterpret_cast<uintptr_t>(f) & 15) == 0);
  bool smem;
  pick_tiles(B, C, N, a.L, a.TC, a.LT, smem);
  TPG_REQUIRE(ceil_div(C, a.TC) <= 65535, TPG_EUNSUPPORTED, "group_fwd: C too large");
  const bool vec4 = (a.L & 3) == 0 && ((reinterpret_cast<uintptr_t>(idx) & 15) == 0) &&
                    ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  dim3 grid(ceil_div(a.L, a.LT), ceil_div(C, a.TC), B);
  const size_t sm = smem ? (size_t)a.TC * N * sizeof(float) : 0;
  cudaStream_t st = as_stream(stream);
#define LAUNCH_GF(S, V)                                                                         \
  do {                                                                                          \
    auto kern = group_fwd_kernel<S, V>;                                                         \
    if (sm > 48 * 1024)                                                                         \
      TPG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
    kern<<<grid, GRP_THREADS, sm, st>>>(a);                                                     \
  } while (0)
  if (smem) { if (vec4) LAUNCH_GF(true, true); else LAUNCH_GF(true, false); }
  else { if (vec4) LAUNCH_GF(false, true); else LAUNCH_GF(false, false); }
#undef LAUNCH_GF
  TPG_CHECK_LAUNCH("group_fwd_kernel");
  return TPG_OK;
}

template <int TC>
static int launch_reduce_smem(ReduceFwdArgs a, size_t sm, cudaStream_t st) {
  auto kern = group_reduce_fwd_smem_kernel<TC>;
  a.TC = TC;
  const int target = 2 * num_sms();
  int mt = max(1, ceil_div(target, a.B * ceil_div(a.C, TC)));
  mt = min(mt, max(1, a.M / RT_THREADS));  // every range re-stages the TC rows: keep ranges >= one tile
  a.MT = ceil_div(ceil_div(a.M, mt), RT_THREADS) * RT_THREADS;
  dim3 grid(ceil_div(a.M, a.MT), ceil_div(a.C, TC), a.B);
  if (sm > 48 * 1024) TPG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  kern<<<grid, RT_THREADS, sm, st>>>(a);
  TPG_CHECK_LAUNCH("group_reduce_fwd_smem_kernel");
  return TPG_OK;
}

static int launch_reduce_fwd(ReduceFwdArgs a, cudaStream_t st) {
  // shared-memory version: the widest channel tile (16 / 8 / 4) whose interleaved rows fit
  for (int TC = 16; TC >= 4; TC >>= 1) {
    if (TC > 4 && TC >= 2 * a.C) continue;  // do not stage (zero) rows that do not exist
    const size_t sm = (size_t)TC * a.N * sizeof(float);
    if (sm > RT_SMEM_MAX || ceil_div(a.C, TC) > 65535) continue;
    if (TC == 16) return launch_reduce_smem<16>(a, sm, st);
    if (TC == 8) return launch_reduce_smem<8>(a, sm, st);
    return launch_reduce_smem<4>(a, sm, st);
  }
  // rows too long for shared memory: read-only global gathers
  const int target = 2 * num_sms();
  int TC = 8;
  while (TC > 1 && TC > a.C) TC >>= 1;
  a.TC = TC;
  int mt = max(1, ceil_div(target, a.B * ceil_div(a.C, TC)));
  mt = min(mt, max(1, a.M / GRP_THREADS));
  a.MT = ceil_div(ceil_div(a.M, mt), GRP_THREADS) * GRP_THREADS;
  dim3 grid(ceil_div(a.M, a.MT), ceil_div(a.C, TC), a.B);
  group_reduce_fwd_kernel<false><<<grid, GRP_THREADS, 0, st>>>(a);
  TPG_CHECK_LAUNCH("group_reduce_fwd_kernel");
  return TPG_OK;
}

// long rows: the widest shared-memory tile (4 channels) does not fit -> point-major gathers
static bool reduce_wants_pointmajor(int C, int N) { return (size_t)4 * N * sizeof(float) > RT_SMEM_MAX && C >= 8; }

static int launch_reduce_pointmajor(const float* f, const int32_t* idx, const float* w, int B, int C, int N, int M, int k,
                                    int op, float* out, int32_t* arg, void* workspace, cudaStream_t st) {
  TPG_REQUIRE(B <= 65535 && ceil_div(C, 32) <= 65535, TPG_EUNSUPPORTED, "group_reduce: B or C too large");
  float* ft = reinterpret_cast<float*>(workspace);
  transpose_cn_kernel<<<dim3(ceil_div(N, 32), ceil_div(C, 32), B), 256, 0, st>>>(f, C, N, ft);
  TPG_CHECK_LAUNCH("transpose_cn_kernel");
  ReduceFwdTArgs a{ft, idx, w, B, C, N, M, k, op, out, arg};
  group_reduce_pointmajor_kernel<<<dim3(ceil_div(M, 32), ceil_div(C, 64), B), 256, 0, st>>>(a);
  TPG_CHECK_LAUNCH("group_reduce_pointmajor_kernel");
  return TPG_OK;
}

TPG_API size_t tpg_group_reduce_workspace_bytes(int B, int C, int N) {
  return reduce_wants_pointmajor(C, N) ? sizeof(float) * (size_t)B * (size_t)C * (size_t)N : 0;
}

TPG_API int tpg_group_reduce_fwd_ws_f32(const float* f, const int32_t* idx, const float* w, int B, int C, int N, int M,
                                        int k, int op, float* out, int32_t* arg, void* workspace,
                                        size_t workspace_bytes, tpg_stream_t stream) {
  TPG_REQUIRE(B >= 0 && C >= 0 && N >= 1 && M >= 0 && k >= 1, TPG_EINVAL, "group_reduce_fwd: bad size");
  TPG_REQUIRE(op >= 0 && op <= 2, TPG_EINVAL, "group_reduce_fwd: bad op %d", op);
  if (B == 0 || C == 0 || M == 0) return TPG_OK;
  TPG_REQUIRE(f && idx && out, TPG_EINVAL, "group_reduce_fwd: null pointer");
  const size_t need = tpg_group_reduce_workspace_bytes(B, C, N);
  if (need && workspace && workspace_bytes >= need)
    return launch_reduce_pointmajor(f, idx, w, B, C, N, M, k, w ? TPG_REDUCE_SUM : op, out, (w || op == TPG_REDUCE_SUM) ? nullptr : arg,
                                    workspace, as_stream(stream));
  if (w) return tpg_three_interpolate_fwd_f32(f, idx, w, B, C, N, M, out, stream);
  return tpg_group_reduce_fwd_f32(f, idx, B, C, N, M, k, op, out, arg, stream);
}

TPG_API int tpg_group_reduce_fwd_f32(const float* f, const int32_t* idx, int B, int C, int N, int M, int k,
                                     int op, float* out, int32_t* arg, tpg_stream_t stream) {
  TPG_REQUIRE(B >= 0 && C >= 0 && N >= 0 && M >= 0, TPG_EINVAL, "group_reduce_fwd: negative size");
  TPG_REQUIRE(k >= 1, TPG_EINVAL, "group_reduce_fwd: k must be >= 1");
  TPG_REQUIRE(op >= 0 && op <= 2, TPG_EINVAL, "group_reduce_fwd: bad op %d", op);
  if (B == 0 || C == 0 || M == 0) return TPG_OK;
  TPG_REQUIRE(N >= 1 && f && idx && out, TPG_EINVAL, "group_reduce_fwd: null pointer / empty cloud");
  TPG_REQUIRE(B <= 65535 && C <= 65535 * 8, TPG_EUNSUPPORTED, "group_reduce_fwd: B or C too large");
  ReduceFwdArgs a{f, idx, nullptr, B, C, N, M, k, op, 1, GRP_THREADS, out, op == TPG_REDUCE_SUM ? nullptr : arg};
  return launch_reduce_fwd(a, as_stream(stream));
}

TPG_API int tpg_three_interpolate_fwd_f32(const float* f, const int32_t* idx, const float* w, int B, int c,
                                          int m, int n, float* out, tpg_stream_t stream) {
  TPG_REQUIRE(B >= 0 && c >= 0 && m >= 0 && n >= 0, TPG_EINVAL, "three_interpolate_fwd: negative size");
  if (B == 0 || c == 0 || n == 0) return TPG_OK;
  TPG_REQUIRE(m >= 1 && f && idx && w && out, TPG_EINVAL, "three_interpolate_fwd: null pointer / empty cloud");
  TPG_REQUIRE(B <= 65535, TPG_EUNSUPPORTED, "three_interpolate_fwd: B too large");
  ReduceFwdArgs a{f, idx, w, B, c, m, n, 3, TPG_REDUCE_SUM, 1, GRP_THREADS, out, nullptr};
  return launch_reduce_fwd(a, as_stream(stream));
}

namespace tpg {
size_t csr_workspace_bytes(int B, int N, int L) {
  return align_up(sizeof(int32_t) * (size_t)B * (size_t)N, 256) + align_up(sizeof(int32_t) * (size_t)B * (size_t)L, 256);
}

int build_csr(const int32_t* idx, const int64_t* item_len, int B, int N, int L, int32_t* seg_offsets,
              int32_t* seg_items, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  TPG_REQUIRE(B >= 0 && N >= 0 && L >= 0, TPG_EINVAL, "inverse_index: negative size");
  if (B == 0) return TPG_OK;
  TPG_REQUIRE(seg_offsets, TPG_EINVAL, "inverse_index: null seg_offsets");
  if (N == 0 || L == 0) {
    TPG_CUDA(cudaMemsetAsync(seg_offsets, 0, sizeof(int32_t) * (size_t)B * (N + 1), st));
    return TPG_OK;
  }
  TPG_REQUIRE(idx && seg_items, TPG_EINVAL, "inverse_index: null pointer");
  TPG_REQUIRE(workspace && workspace_bytes >= csr_workspace_bytes(B, N, L), TPG_EWORKSPACE,
              "inverse_index: workspace too small");
  {
    bool wide = false;
    size_t sm = 0;
    const int Q = csr_cluster_size(N, L, &wide, &sm);
    if (Q >= 1)
      return wide ? launch_csr_cluster<uint32_t>(idx, item_len, B, N, L, Q, sm, seg_offsets, seg_items, st)
                  : launch_csr_cluster<uint16_t>(idx, item_len, B, N, L, Q, sm, seg_offsets, seg_items, st);
  }
  {
    bool wide = false;
    const int W = csr_stable_warps(N, L, &wide);
    if (W == 32) {  // fewer warps per cloud: the count / scan / fill / sort path over all SMs is faster
      const size_t sm = sizeof(int32_t) * (size_t)((N + 1 + 3) & ~3) + (size_t)W * N * (wide ? 5 : 3);
      if (wide) {
        auto kern = csr_stable_kernel<uint32_t>;
        if (sm > 48 * 1024) TPG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        kern<<<B, 32 * W, sm, st>>>(idx, item_len, N, L, seg_offsets, seg_items);
      } else {
        auto kern = csr_stable_kernel<uint16_t>;
        if (sm > 48 * 1024) TPG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        kern<<<B, 32 * W, sm, st>>>(idx, item_len, N, L, seg_offsets, seg_items);
      }
      TPG_CHECK_LAUNCH("csr_stable_kernel");
      return TPG_OK;
    }
  }
  TPG_CUDA(cudaMemsetAsync(seg_offsets, 0, sizeof(int32_t) * (size_t)B * (N + 1), st));
  int32_t* cursor = reinterpret_cast<int32_t*>(workspace);
  int32_t* tmp = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(workspace) +
                                            align_up(sizeof(int32_t) * (size_t)B * (size_t)N, 256));
  const long long total = (long long)B * L;
  const int threads = 256;
  const unsigned blocks = (unsigned)min((total + threads - 1) / threads, (long long)num_sms() * 32);
  csr_count_kernel<<<blocks, threads, 0, st>>>(idx, item_len, B, N, L, seg_offsets);
  TPG_CHECK_LAUNCH("csr_count_kernel");
  csr_scan_kernel<<<B, 1024, 0, st>>>(seg_offsets, cursor, N);
  TPG_CHECK_LAUNCH("csr_scan_kernel");
  csr_fill_kernel<<<blocks, threads, 0, st>>>(idx, item_len, B, N, L, cursor, tmp);
  TPG_CHECK_LAUNCH("csr_fill_kernel");
  const long long segs = (long long)B * N;
  csr_sort_kernel<<<(unsigned)((segs + 7) / 8), 256, 0, st>>>(seg_offsets, tmp, B, N, L, seg_items);
  TPG_CHECK_LAUNCH("csr_sort_kernel");
  return TPG_OK;
}
}  // namespace tpg

TPG_API size_t tpg_inverse_index_workspace_bytes(int B, int N, int L) { return tpg::csr_workspace_bytes(B, N, L); }

TPG_API int tpg_inverse_index_build(const int32_t* idx, int B, int N, int L, int32_t* seg_offsets,
                                    int32_t* seg_items, void* workspace, size_t workspace_bytes,
                                    tpg_stream_t stream) {
  return tpg::build_csr(idx, nullptr, B, N, L, seg_offsets, seg_items, workspace, workspace_bytes, as_stream(stream));
}

static int launch_bwd(const BwdArgs& a, cudaStream_t st) {
  TPG_REQUIRE(a.C <= 65535 && a.B <= 65535, TPG_EUNSUPPORTED, "group_bwd: B or C too large");
  dim3 grid(ceil_div(a.N, GRP_THREADS), a.C, a.B);
  group_bwd_kernel<<<grid, GRP_THREADS, 0, st>>>(a);
  TPG_CHECK_LAUNCH("group_bwd_kernel");
  return TPG_OK;
}

TPG_API int tpg_group_bwd_f32(const float* grad_out, const int32_t* seg_offsets, const int32_t* seg_items,
                              int B, int C, int N, int L, float* grad_f, tpg_stream_t stream) {
  TPG_REQUIRE(B >= 0 && C >= 0 && N >= 0 && L >= 0, TPG_EINVAL, "group_bwd: negative size");
  if (B == 0 || C == 0 || N == 0) return TPG_OK;
  TPG_REQUIRE(grad_f && seg_offsets && (L == 0 || (grad_out && seg_items)), TPG_EINVAL, "group_bwd: null pointer");
  if (L > 0 && tpg::group_bwd_staged_eligible(grad_out, seg_items, B, C, N, L))
    return tpg::group_bwd_staged(grad_out, seg_offsets, seg_items, B, C, N, L, grad_f, 0, as_stream(stream));
  BwdArgs a{grad_out, nullptr, nullptr, seg_offsets, seg_items, B, C, N, 0, 1, L, BWD_GROUP, grad_f};
  return launch_bwd(a, as_stream(stream));
}

TPG_API int tpg_group_bwd_segments(int B, int C, int N, int L) { return tpg::group_bwd_staged_segments(B, C, N, L); }

TPG_API int tpg_group_bwd_segmented_f32(const float* grad_out, const int32_t* seg_offsets, const int32_t* seg_items,
                                        int B, int C, int N, int L, int S, float* grad_f, tpg_stream_t stream) {
  TPG_REQUIRE(B >= 0 && C >= 0 && N >= 1 && L >= 1 && S >= 1, TPG_EINVAL, "group_bwd_segmented: bad size");
  if (B == 0 || C == 0) return TPG_OK;
  TPG_REQUIRE(grad_f && seg_offsets && grad_out && seg_items, TPG_EINVAL, "group_bwd_segmented: null pointer");
  return tpg::group_bwd_staged_segmented(grad_out, seg_offsets, seg_items, B, C, N, L, S, grad_f, as_stream(stream));
}

// tuning hook (tools/bench_group_bwd.py; not part of the ABI): variant 0 = plain CSR gather, 1/2/4 = staged kernel
// with that many channels per thread, -1 = the dispatch of tpg_group_bwd_f32
TPG_API int tpg_debug_group_bwd_variant(const float* grad_out, const int32_t* seg_offsets, const int32_t* seg_items,
                                        int B, int C, int N, int L, float* grad_f, int variant, tpg_stream_t stream) {
  if (variant < 0) return tpg_group_bwd_f32(grad_out, seg_offsets, seg_items, B, C, N, L, grad_f, stream);
  if (variant == 0) {
    BwdArgs a{grad_out, nullptr, nullptr, seg_offsets, seg_items, B, C, N, 0, 1, L, BWD_GROUP, grad_f};
    return launch_bwd(a, as_stream(stream));
  }
  TPG_REQUIRE(tpg::group_bwd_staged_eligible(grad_out, seg_items, B, C, N, L), TPG_EUNSUPPORTED, "group_bwd: shape not eligible for the staged kernel");
  return tpg::group_bwd_staged(grad_out, seg_offsets, seg_items, B, C, N, L, grad_f, variant, as_stream(stream));
}

TPG_API int tpg_group_reduce_bwd_f32(const float* grad_out, const int32_t* arg, const int32_t* seg_offsets,
                                     const int32_t* seg_items, int B, int C, int N, int M, int k, int op,
                                     float* grad_f, tpg_stream_t stream) {
  TPG_REQUIRE(B >= 0 && C >= 0 && N >= 0 && M >= 0 && k >= 1, TPG_EINVAL, "group_reduce_bwd: bad size");
  TPG_REQUIRE(op >= 0 && op <= 2, TPG_EINVAL, "group_reduce_bwd: bad op %d", op);
  if (B == 0 || C == 0 || N == 0) return TPG_OK;
  TPG_REQUIRE(grad_f && seg_offsets, TPG_EINVAL, "group_reduce_bwd: null pointer");
  TPG_REQUIRE(op == TPG_REDUCE_SUM || arg, TPG_EINVAL, "group_reduce_bwd: max/min need arg");
  BwdArgs a{grad_out, arg, nullptr, seg_offsets, seg_items, B, C, N, M, k, M * k,
            op == TPG_REDUCE_SUM ? BWD_SUM : BWD_ARG, grad_f};
  return launch_bwd(a, as_stream(stream));
}

TPG_API int tpg_three_interpolate_bwd_f32(const float* grad_out, const float* w, const int32_t* seg_offsets,
                                          const int32_t* seg_items, int B, int c, int m, int n, float* grad_f,
                                          tpg_stream_t stream) {
  TPG_REQUIRE(B >= 0 && c >= 0 && m >= 0 && n >= 0, TPG_EINVAL, "three_interpolate_bwd: negative size");
  if (B == 0 || c == 0 || m == 0) return TPG_OK;
  TPG_REQUIRE(grad_f && seg_offsets && w, TPG_EINVAL, "three_interpolate_bwd: null pointer");
  BwdArgs a{grad_out, nullptr, w, seg_offsets, seg_items, B, c, m, n, 3, n * 3, BWD_WEIGHTED, grad_f};
  return launch_bwd(a, as_stream(stream));
}
