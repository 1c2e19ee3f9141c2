// fps.cu — K5: farthest point sampling, two modes.
//   mode A  pointnet2_utils.furthest_point_sample (discriminator.py:114): start 0,
//           running min-dist init 1e10, points with x^2+y^2+z^2 <= 1e-3 never
//           updated nor selected, int32 out.
//   mode B  sampling.farthest_point_sampling (sampling.py:36-44, 50-106): explicit
//           start, no skip, np.argmax (first maximum), int64 out, optional [k,N]
//           rows of squared distances.
// Both: argmax ties -> lowest index; d2 = ((dx*dx)+dy*dy)+dz*dz unfused.
//
// Design: FPS is a chain of `npoint` dependent arg-max rounds, so the kernel is
// latency-bound, not bandwidth-bound.  One CTA owns one cloud; every point lives in
// REGISTERS of its thread (coordinates + running min-dist, PPT points per thread),
// a round is: PPT distance updates per thread -> two REDUX instructions per warp
// (max of the distance bits, then min index among the maxima) -> one 32-entry
// shared-memory exchange with a single __syncthreads (double-buffered) -> two more
// REDUX -> the winner's coordinates come back through L1 (the cloud was loaded
// through L1 by this CTA and stays resident).
// Clouds larger than 8192 points use the same round structure with the running
// min-dist in a caller-provided global workspace.
#include "common.cuh"

namespace tpg {

struct FpsArgs {
  const float* pts;  // [B,N,D]
  int B, N, D, npoint;
  const int64_t* start;  // mode B
  void* out;             // int32 [B,npoint] (A) or int64 (B)
  float* rows;           // mode B optional [B,npoint,N]
  float* temp_ws;        // [B,N] for the large-cloud kernel
};

__device__ __forceinline__ void load_point(const float* p, int D, float& x, float& y, float& z) {
  x = p[0];
  y = D > 1 ? p[1] : 0.0f;
  z = D > 2 ? p[2] : 0.0f;
}

// one round's cross-thread arg-max.  Returns the winning index (or 0 if no
// candidate at all, which is what upstream's "best=-1, besti=0" start yields).
__device__ __forceinline__ int block_argmax(unsigned bits, unsigned j, bool has, unsigned (*s_bits)[32],
                                            unsigned (*s_j)[32], int buf, int nwarps) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned m = __reduce_max_sync(FULL, has ? bits : 0u);
  unsigned cj = (has && bits == m) ? j : 0xffffffffu;
  unsigned jm = __reduce_min_sync(FULL, cj);
  if (lane == 0) { s_bits[buf][warp] = m; s_j[buf][warp] = jm; }
  __syncthreads();
  unsigned wb = lane < nwarps ? s_bits[buf][lane] : 0u;
  unsigned wj = lane < nwarps ? s_j[buf][lane] : 0xffffffffu;
  // a warp without candidates reports (0, 0xffffffff) and loses to any real entry
  unsigned m2 = __reduce_max_sync(FULL, wj != 0xffffffffu ? wb : 0u);
  unsigned cj2 = (wj != 0xffffffffu && wb == m2) ? wj : 0xffffffffu;
  unsigned j2 = __reduce_min_sync(FULL, cj2);
  return j2 == 0xffffffffu ? 0 : (int)j2;
}

template <int PPT, bool MODEB>
__global__ void __launch_bounds__(1024) fps_reg_kernel(FpsArgs a) {
  __shared__ unsigned s_bits[2][32];
  __shared__ unsigned s_j[2][32];
  const int b = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
  const int nwarps = T >> 5;
  const float* p = a.pts + (size_t)b * a.N * a.D;
  float x[PPT], y[PPT], z[PPT], t[PPT];
  bool live[PPT];
#pragma unroll
  for (int s = 0; s < PPT; ++s) {
    const int j = tid + s * T;
    live[s] = j < a.N;
    x[s] = y[s] = z[s] = 0.0f;
    if (live[s]) load_point(p + (size_t)j * a.D, a.D, x[s], y[s], z[s]);
    if (MODEB) {
      t[s] = __int_as_float(0x7f800000);
    } else {
      t[s] = 1e10f;
      float mag = __fadd_rn(__fadd_rn(__fmul_rn(x[s], x[s]), __fmul_rn(y[s], y[s])), __fmul_rn(z[s], z[s]));
      live[s] = live[s] && !(mag <= 1e-3f);
    }
  }
  int cur = MODEB ? (int)a.start[b] : 0;
  for (int it = 0; it < a.npoint; ++it) {
    if (tid == 0) {
      if (MODEB) reinterpret_cast<int64_t*>(a.out)[(size_t)b * a.npoint + it] = cur;
      else reinterpret_cast<int32_t*>(a.out)[(size_t)b * a.npoint + it] = cur;
    }
    if (it == a.npoint - 1 && !(MODEB && a.rows)) break;
    float cx, cy, cz;
    load_point(p + (size_t)cur * a.D, a.D, cx, cy, cz);
    unsigned bb = 0u, bj = 0xffffffffu;
    bool has = false;
#pragma unroll
    for (int s = 0; s < PPT; ++s) {
      const int j = tid + s * T;
      if (live[s]) {
        float d = sqdist3(x[s], y[s], z[s], cx, cy, cz);
        if (MODEB && a.rows) a.rows[((size_t)b * a.npoint + it) * a.N + j] = d;
        float v = fminf(d, t[s]);
        t[s] = v;
        unsigned vb = __float_as_uint(v);
        if (!has || vb > bb) { bb = vb; bj = (unsigned)j; has = true; }
      }
    }
    cur = block_argmax(bb, bj, has, s_bits, s_j, it & 1, nwarps);
  }
}

// large clouds: running min-dist in global memory, coordinates re-read (L2-resident)
template <bool MODEB>
__global__ void __launch_bounds__(1024) fps_mem_kernel(FpsArgs a) {
  __shared__ unsigned s_bits[2][32];
  __shared__ unsigned s_j[2][32];
  const int b = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
  const int nwarps = T >> 5;
  const float* p = a.pts + (size_t)b * a.N * a.D;
  float* temp = a.temp_ws + (size_t)b * a.N;
  for (int j = tid; j < a.N; j += T) temp[j] = MODEB ? __int_as_float(0x7f800000) : 1e10f;
  int cur = MODEB ? (int)a.start[b] : 0;
  for (int it = 0; it < a.npoint; ++it) {
    if (tid == 0) {
      if (MODEB) reinterpret_cast<int64_t*>(a.out)[(size_t)b * a.npoint + it] = cur;
      else reinterpret_cast<int32_t*>(a.out)[(size_t)b * a.npoint + it] = cur;
    }
    if (it == a.npoint - 1 && !(MODEB && a.rows)) break;
    float cx, cy, cz;
    load_point(p + (size_t)cur * a.D, a.D, cx, cy, cz);
    unsigned bb = 0u, bj = 0xffffffffu;
    bool has = false;
    for (int j = tid; j < a.N; j += T) {
      float x, y, z;
      load_point(p + (size_t)j * a.D, a.D, x, y, z);
      if (!MODEB) {
        float mag = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
        if (mag <= 1e-3f) continue;
      }
      float d = sqdist3(x, y, z, cx, cy, cz);
      if (MODEB && a.rows) a.rows[((size_t)b * a.npoint + it) * a.N + j] = d;
      float v = fminf(d, temp[j]);
      temp[j] = v;
      unsigned vb = __float_as_uint(v);
      if (!has || vb > bb) { bb = vb; bj = (unsigned)j; has = true; }
    }
    cur = block_argmax(bb, bj, has, s_bits, s_j, it & 1, nwarps);
  }
}

template <bool MODEB>
static int fps_dispatch(const FpsArgs& a, cudaStream_t st) {
  if (a.B == 0 || a.npoint == 0) return TPG_OK;
  const int N = a.N;
  if (N <= 8192) {
    int T, ppt;
    if (N <= 1024) { T = max(32, (N + 31) / 32 * 32); ppt = 1; }
    else { T = 1024; ppt = N <= 2048 ? 2 : (N <= 4096 ? 4 : 8); }
    switch (ppt) {
      case 1: fps_reg_kernel<1, MODEB><<<a.B, T, 0, st>>>(a); break;
      case 2: fps_reg_kernel<2, MODEB><<<a.B, T, 0, st>>>(a); break;
      case 4: fps_reg_kernel<4, MODEB><<<a.B, T, 0, st>>>(a); break;
      default: fps_reg_kernel<8, MODEB><<<a.B, T, 0, st>>>(a); break;
    }
    TPG_CHECK_LAUNCH("fps_reg_kernel");
  } else {
    TPG_REQUIRE(a.temp_ws != nullptr, TPG_EWORKSPACE, "fps: N=%d > 8192 needs a [B,N] float workspace", N);
    fps_mem_kernel<MODEB><<<a.B, 1024, 0, st>>>(a);
    TPG_CHECK_LAUNCH("fps_mem_kernel");
  }
  return TPG_OK;
}

}  // namespace tpg

using namespace tpg;

TPG_API size_t tpg_fps_workspace_bytes(int B, int N) {
  return N > 8192 ? sizeof(float) * (size_t)B * (size_t)N : 0;
}

TPG_API int tpg_fps_f32(const float* xyz, int B, int N, int npoint, int32_t* idx, void* workspace,
                        size_t workspace_bytes, tpg_stream_t stream) {
  TPG_REQUIRE(B >= 0 && N >= 1 && npoint >= 0, TPG_EINVAL, "fps: bad size B=%d N=%d npoint=%d", B, N, npoint);
  if (B == 0 || npoint == 0) return TPG_OK;
  TPG_REQUIRE(xyz && idx, TPG_EINVAL, "fps: null pointer");
  TPG_REQUIRE(workspace_bytes >= tpg_fps_workspace_bytes(B, N), TPG_EWORKSPACE, "fps: workspace too small");
  FpsArgs a{xyz, B, N, 3, npoint, nullptr, idx, nullptr, reinterpret_cast<float*>(workspace)};
  return fps_dispatch<false>(a, as_stream(stream));
}

TPG_API int tpg_fps_start_f32(const float* pts, int B, int N, int D, int k, const int64_t* start,
                              int64_t* idx, float* dist_rows, void* workspace, size_t workspace_bytes,
                              tpg_stream_t stream) {
  TPG_REQUIRE(B >= 0 && N >= 1 && k >= 0, TPG_EINVAL, "fps_start: bad size");
  TPG_REQUIRE(D >= 1 && D <= 3, TPG_EUNSUPPORTED, "fps_start: D=%d outside [1,3]", D);
  if (B == 0 || k == 0) return TPG_OK;
  TPG_REQUIRE(pts && start && idx, TPG_EINVAL, "fps_start: null pointer");
  TPG_REQUIRE(workspace_bytes >= tpg_fps_workspace_bytes(B, N), TPG_EWORKSPACE, "fps_start: workspace too small");
  FpsArgs a{pts, B, N, D, k, start, idx, dist_rows, reinterpret_cast<float*>(workspace)};
  return fps_dispatch<true>(a, as_stream(stream));
}
