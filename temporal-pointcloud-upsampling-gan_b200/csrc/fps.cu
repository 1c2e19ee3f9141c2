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
// Clouds of more than 2048 points are split over a thread-block cluster of 4 (<= 8192 points) or 8 CTAs;
// the local winners are exchanged ONE-WAY through DSMEM (st.async completing bytes on the receiver's
// mbarrier: no cluster barrier in the round); beyond 65536 points the running min-dist moves to a
// caller-provided global workspace.
#include <atomic>
#include <cstdlib>

#include "common.cuh"

#include <cooperative_groups.h>

namespace tpg {

struct FpsArgs {
  const float* pts;  // [B,N,D]
  int B, N, D, npoint;
  const int64_t* start;  // mode B
  void* out;             // int32 [B,npoint] (A) or int64 (B)
  float* rows;           // mode B optional [B,npoint,N]
  float* temp_ws;        // [B,N] for the large-cloud kernel
  int flags;             // tuning hook (tools/bench_fps.py), unused by the shipped kernels
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

// Register-resident FPS.  One CTA (CL == 1) or one thread-block cluster of CL CTAs (CL == 8,
// one SM each) owns a cloud; a CTA keeps its slice of the cloud in registers, PPT points per
// thread, and uses as FEW warps as the slice allows: the per-round cost of a warp is
// ~14 instructions per point plus ~30 of fixed reduction work, so 8 points per thread keeps
// the round latency-bound (~300 cycles) instead of issue-bound.
// Cluster rounds: CTA-local arg-max as above, then warp 0 pushes {key, index, xyz} of the local
// winner into every peer's shared memory (DSMEM), one cluster barrier, and every thread reduces
// the CL slots — the winner's coordinates arrive with the key, so no global access is on the
// critical path (cluster.sync invalidates L1D anyway).
struct __align__(16) FpsSlot {
  unsigned long long key;  // (bits(min-dist) + 1) << 32 | ~index; 0 == no candidate
  float x, y, z;
  unsigned pad;
};

__device__ __forceinline__ unsigned long long fps_pack(unsigned key, unsigned j) {
  return ((unsigned long long)key << 32) | (unsigned long long)(0xffffffffu - j);
}

// ---- one-way cluster exchange (replaces one cluster.sync() per round) -------------------------------------
// Every CTA pushes {key, xyz} of its local winner into every peer's slot with st.async: the store itself
// completes transaction bytes on the RECEIVER's mbarrier, so a round costs one DSMEM store latency plus an
// mbarrier wake-up instead of store + full cluster barrier (arrive.release / wait.acquire, which also
// flushes L1D).  Slots and barriers are double-buffered by round parity; a sender can only be two rounds
// ahead of a receiver after having received that receiver's message of the round in between, which the
// receiver sent after reading the slots being overwritten — so no slot is overwritten while still in use.
__device__ __forceinline__ uint32_t fps_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t fps_mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void fps_st_async_v4(uint32_t raddr, uint32_t rbar, unsigned a0, unsigned a1, unsigned a2, unsigned a3) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];\n" ::"r"(raddr),
               "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(rbar)
               : "memory");
}
__device__ __forceinline__ void fps_st_async_v2(uint32_t raddr, uint32_t rbar, unsigned a0, unsigned a1) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];\n" ::"r"(raddr), "r"(a0),
               "r"(a1), "r"(rbar)
               : "memory");
}
__device__ __forceinline__ void fps_mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  for (uint32_t spin = 0; spin < (1u << 26); ++spin) {  // bounded: a protocol bug must trap, not hang the GPU
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
  }
  __trap();
}

template <int PPT, int CL, bool MODEB, bool ONEWAY = false>
__global__ void __launch_bounds__(1024) fps_reg_kernel(FpsArgs a) {
  __shared__ unsigned long long s_key[2][32];
  __shared__ FpsSlot xchg[2][CL];
  __shared__ __align__(8) unsigned long long xbar[2];
  extern __shared__ float pts_s[];  // this CTA's slice of the cloud, xyz interleaved (winner lookup)
  const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31;
  const int nwarps = T >> 5;
  const int b = blockIdx.x / CL, rank = blockIdx.x % CL;
  const float* p = a.pts + (size_t)b * a.N * a.D;
  const int ppc = CL == 1 ? a.N : ((a.N + CL - 1) / CL + 31) & ~31;  // points per CTA
  const int lo = rank * ppc, hi = min(a.N, lo + ppc);
  float x[PPT], y[PPT], z[PPT], t[PPT];
  bool live[PPT];
#pragma unroll
  for (int s = 0; s < PPT; ++s) {
    const int jl = tid + s * T, j = lo + jl;
    live[s] = j < hi;
    x[s] = y[s] = z[s] = 0.0f;
    if (live[s]) {
      load_point(p + (size_t)j * a.D, a.D, x[s], y[s], z[s]);
      pts_s[3 * jl] = x[s]; pts_s[3 * jl + 1] = y[s]; pts_s[3 * jl + 2] = z[s];
    }
    if (MODEB) {
      t[s] = __int_as_float(0x7f800000);
    } else {
      t[s] = 1e10f;
      float mag = __fadd_rn(__fadd_rn(__fmul_rn(x[s], x[s]), __fmul_rn(y[s], y[s])), __fmul_rn(z[s], z[s]));
      live[s] = live[s] && !(mag <= 1e-3f);
    }
  }
  int cur = MODEB ? (int)a.start[b] : 0;
  float cx, cy, cz;
  load_point(p + (size_t)cur * a.D, a.D, cx, cy, cz);
  if (CL > 1 && ONEWAY) {
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(fps_smem_u32(&xbar[0])) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(fps_smem_u32(&xbar[1])) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    cooperative_groups::this_cluster().sync();  // every peer's barriers exist before the first remote complete_tx
  }
  __syncthreads();
  for (int it = 0; it < a.npoint; ++it) {
    if (tid == 0 && rank == 0) {
      if (MODEB) reinterpret_cast<int64_t*>(a.out)[(size_t)b * a.npoint + it] = cur;
      else reinterpret_cast<int32_t*>(a.out)[(size_t)b * a.npoint + it] = cur;
    }
    if (it == a.npoint - 1 && !(MODEB && a.rows)) break;
    // branch-free update (the PPT points of a thread must overlap in the pipeline: a branchy
    // version serialises them at ~135 cycles per point).  key = bits(min-dist) + 1 for live
    // points, 0 for skipped ones; strict '>' keeps the lower index on ties.
    unsigned bb = 0u, bj = 0xffffffffu;
#pragma unroll
    for (int s = 0; s < PPT; ++s) {
      const int j = lo + tid + s * T;
      const float d = sqdist3(x[s], y[s], z[s], cx, cy, cz);
      if (MODEB && a.rows && live[s]) a.rows[((size_t)b * a.npoint + it) * a.N + j] = d;
      const float v = fminf(d, t[s]);
      t[s] = v;
      const unsigned key = live[s] ? __float_as_uint(v) + 1u : 0u;
      const bool take = key > bb;
      bb = take ? key : bb;
      bj = take ? (unsigned)j : bj;
    }
    // warp arg-max: (max key, lowest index)
    const int warp = tid >> 5, buf = it & 1;
    const unsigned m = __reduce_max_sync(FULL, bb);
    const unsigned jm = __reduce_min_sync(FULL, (bb == m && m != 0u) ? bj : 0xffffffffu);
    if (lane == 0) s_key[buf][warp] = m == 0u ? 0ull : fps_pack(m, jm);
    __syncthreads();
    // CTA arg-max over the warp entries: REDUX on the two halves of the packed key (measured faster
    // than a linear scan of broadcast LDS.64 even for 4 warps)
    unsigned long long best;
    {
      const unsigned long long e = lane < nwarps ? s_key[buf][lane] : 0ull;
      const unsigned hi32 = __reduce_max_sync(FULL, (unsigned)(e >> 32));
      const unsigned lo32 = __reduce_max_sync(FULL, (unsigned)(e >> 32) == hi32 ? (unsigned)e : 0u);
      best = ((unsigned long long)hi32 << 32) | lo32;
    }
    if (CL == 1) {
      cur = best == 0ull ? 0 : (int)(0xffffffffu - (unsigned)best);  // upstream starts at (best=-1, besti=0)
      cx = pts_s[3 * cur]; cy = pts_s[3 * cur + 1]; cz = pts_s[3 * cur + 2];
    } else {
      namespace cg = cooperative_groups;
      cg::cluster_group cluster = cg::this_cluster();
      if (warp == 0 && lane < CL) {
        FpsSlot v;
        v.key = best; v.x = v.y = v.z = 0.0f; v.pad = 0u;
        if (best != 0ull) {
          const int jl = (int)(0xffffffffu - (unsigned)best) - lo;
          v.x = pts_s[3 * jl]; v.y = pts_s[3 * jl + 1]; v.z = pts_s[3 * jl + 2];
        }
        if (ONEWAY) {
          const uint32_t rslot = fps_mapa(fps_smem_u32(&xchg[buf][rank]), (uint32_t)lane);
          const uint32_t rbar = fps_mapa(fps_smem_u32(&xbar[buf]), (uint32_t)lane);
          if (lane == 0)  // this CTA's own barrier: one arrival + the bytes of all CL senders (24 each)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(fps_smem_u32(&xbar[buf])), "r"(CL * 24) : "memory");
          fps_st_async_v4(rslot, rbar, (unsigned)v.key, (unsigned)(v.key >> 32), __float_as_uint(v.x), __float_as_uint(v.y));
          fps_st_async_v2(rslot + 16, rbar, __float_as_uint(v.z), 0u);
        } else {
          FpsSlot* dst = cluster.map_shared_rank(&xchg[buf][rank], lane);
          reinterpret_cast<uint4*>(dst)[0] = make_uint4((unsigned)v.key, (unsigned)(v.key >> 32), __float_as_uint(v.x), __float_as_uint(v.y));
          reinterpret_cast<float*>(dst)[4] = v.z;
        }
      }
      if (ONEWAY) fps_mbar_wait_cluster(fps_smem_u32(&xbar[buf]), (uint32_t)((it >> 1) & 1));
      else cluster.sync();
      unsigned long long g = xchg[buf][0].key;
      int src = 0;
#pragma unroll
      for (int r = 1; r < CL; ++r) { const unsigned long long o = xchg[buf][r].key; if (o > g) { g = o; src = r; } }
      if (g == 0ull) {
        cur = 0;
        load_point(p, a.D, cx, cy, cz);
      } else {
        cur = (int)(0xffffffffu - (unsigned)g);
        cx = xchg[buf][src].x; cy = xchg[buf][src].y; cz = xchg[buf][src].z;
      }
    }
  }
  if (CL > 1) cooperative_groups::this_cluster().sync();  // no CTA may exit while peers still write its smem
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

std::atomic<int>& fps_exclusive_option();

template <int PPT, int CL, bool MODEB, bool ONEWAY = (CL > 1)>
static int fps_launch(const FpsArgs& a, int threads, cudaStream_t st) {
  auto kern = fps_reg_kernel<PPT, CL, MODEB, ONEWAY>;
  const int ppc = CL == 1 ? a.N : (ceil_div(a.N, CL) + 31) & ~31;
  size_t smem = sizeof(float) * 3 * (size_t)ppc;
  // option "fps.exclusive_sm": a single-CTA-per-cloud launch asks for (nearly) all shared memory of its SM, so no
  // smem-using CTA of a concurrent kernel lands beside it and steals issue slots from the latency-bound rounds
  if (CL == 1 && fps_exclusive_option().load(std::memory_order_relaxed)) smem = max(smem, (size_t)200 * 1024);
  if (smem > 40 * 1024) TPG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  // static smem counts too
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(a.B * CL));
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CL > 1 ? 1 : 0;
  TPG_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
  TPG_CHECK_LAUNCH("fps_reg_kernel");
  return TPG_OK;
}

constexpr int FPS_CL = 8;         // CTAs (SMs) per cloud for N > FPS_SINGLE_MAX
constexpr int FPS_SINGLE_MAX = 2048;
constexpr int FPS_REG_MAX = 65536;  // 8 CTAs x 1024 threads x 8 points

std::atomic<int>& fps_cluster_option() {
  static std::atomic<int> v{[] {
    const char* e = getenv("TPG_FPS_CLUSTER");
    const int c = e ? atoi(e) : FPS_CL;
    return (c == 1 || c == 2 || c == 4 || c == 8) ? c : FPS_CL;
  }()};
  return v;
}

std::atomic<int>& fps_exclusive_option() {
  static std::atomic<int> v{[] { const char* e = getenv("TPG_FPS_EXCLUSIVE"); return e && e[0] == '1' ? 1 : 0; }()};
  return v;
}

template <bool MODEB>
static int fps_dispatch(const FpsArgs& a, cudaStream_t st) {
  if (a.B == 0 || a.npoint == 0) return TPG_OK;
  const int N = a.N;
  if (N <= FPS_SINGLE_MAX) {
    // one CTA; measured on B200: 2 points per thread up to 1024 points, 4 up to 2048
    if (N <= 1024) return fps_launch<2, 1, MODEB>(a, max(32, (ceil_div(N, 2) + 31) & ~31), st);
    return fps_launch<4, 1, MODEB>(a, (ceil_div(N, 4) + 31) & ~31, st);
  }
  // option "fps.sms_per_cloud" = 1|2|4: fewer SMs per cloud.  A round gets slower, but a call that overlaps with
  // other work (multi-stream replay) then holds 8/16/32 SMs instead of 64 for its whole duration.
  const int cl_env = fps_cluster_option().load(std::memory_order_relaxed);
  if (cl_env < 4 && N <= 8192 * cl_env && (cl_env == 1 || cl_env == 2)) {
    const int ppc = (ceil_div(N, cl_env) + 31) & ~31;
    const int t8 = (ceil_div(ppc, 8) + 31) & ~31;
    if (cl_env == 1) return fps_launch<8, 1, MODEB>(a, t8, st);
    return fps_launch<8, 2, MODEB>(a, t8, st);
  }
  if (N <= FPS_REG_MAX) {
    // measured (tools/bench_fps.py, one-way exchange): a round is cheapest with 8 points per thread and as few
    // CTAs per cloud as keep a CTA at <= 256 threads (2048 points): 2 CTAs up to 4096 points, 4 up to 8192 (677 vs
    // 867 ns/round with the former 8 x 256 x 4 layout), 8 beyond
    if (N <= 4096) {  // 2 CTAs x 2048 points: 653 vs 726 ns/round with 4 CTAs
      const int ppc = (ceil_div(N, 2) + 31) & ~31;
      return fps_launch<8, 2, MODEB>(a, (ceil_div(ppc, 8) + 31) & ~31, st);
    }
    if (N <= 8192) {
      const int ppc = (ceil_div(N, 4) + 31) & ~31;
      return fps_launch<8, 4, MODEB>(a, (ceil_div(ppc, 8) + 31) & ~31, st);
    }
    const int ppc = (ceil_div(N, FPS_CL) + 31) & ~31;
    return fps_launch<8, FPS_CL, MODEB>(a, (ceil_div(ppc, 8) + 31) & ~31, st);
  }
  TPG_REQUIRE(a.temp_ws != nullptr, TPG_EWORKSPACE, "fps: N=%d > %d needs a [B,N] float workspace", N, FPS_REG_MAX);
  fps_mem_kernel<MODEB><<<a.B, 1024, 0, st>>>(a);
  TPG_CHECK_LAUNCH("fps_mem_kernel");
  return TPG_OK;
}

}  // namespace tpg

using namespace tpg;

TPG_API size_t tpg_fps_workspace_bytes(int B, int N) {
  return N > tpg::FPS_REG_MAX ? sizeof(float) * (size_t)B * (size_t)N : 0;
}

TPG_API int tpg_fps_f32(const float* xyz, int B, int N, int npoint, int32_t* idx, void* workspace,
                        size_t workspace_bytes, tpg_stream_t stream) {
  TPG_REQUIRE(B >= 0 && N >= 1 && npoint >= 0, TPG_EINVAL, "fps: bad size B=%d N=%d npoint=%d", B, N, npoint);
  if (B == 0 || npoint == 0) return TPG_OK;
  TPG_REQUIRE(xyz && idx, TPG_EINVAL, "fps: null pointer");
  TPG_REQUIRE(workspace_bytes >= tpg_fps_workspace_bytes(B, N), TPG_EWORKSPACE, "fps: workspace too small");
  FpsArgs a{xyz, B, N, 3, npoint, nullptr, idx, nullptr, reinterpret_cast<float*>(workspace), 0};
  return fps_dispatch<false>(a, as_stream(stream));
}

// tuning hook (tools/bench_fps.py; not part of the ABI): pointnet2 mode with an explicit variant
TPG_API int tpg_debug_fps_variant(const float* xyz, int B, int N, int npoint, int32_t* idx, int ppt, int cl,
                                  int threads, int flags, tpg_stream_t stream) {
  FpsArgs a{xyz, B, N, 3, npoint, nullptr, idx, nullptr, nullptr, flags};
  cudaStream_t st = as_stream(stream);
  // flags bit 1: cluster variants with the per-round cluster.sync() (the pre-one-way exchange), for comparison
#define V(P, C) if (ppt == P && cl == C) return (flags & 2) ? fps_launch<P, C, false, false>(a, threads, st) : fps_launch<P, C, false>(a, threads, st)
  V(1, 1); V(2, 1); V(4, 1); V(8, 1); V(16, 1);
  V(1, 8); V(2, 8); V(4, 8); V(8, 8);
  V(2, 4); V(4, 4); V(8, 4); V(16, 4); V(4, 2); V(8, 2); V(16, 2); V(16, 8);
#undef V
  set_error("fps variant ppt=%d cl=%d not built", ppt, cl);
  return TPG_EUNSUPPORTED;
}

TPG_API int tpg_fps_start_f32(const float* pts, int B, int N, int D, int k, const int64_t* start,
                              int64_t* idx, float* dist_rows, void* workspace, size_t workspace_bytes,
                              tpg_stream_t stream) {
  TPG_REQUIRE(B >= 0 && N >= 1 && k >= 0, TPG_EINVAL, "fps_start: bad size");
  TPG_REQUIRE(D >= 1 && D <= 3, TPG_EUNSUPPORTED, "fps_start: D=%d outside [1,3]", D);
  if (B == 0 || k == 0) return TPG_OK;
  TPG_REQUIRE(pts && start && idx, TPG_EINVAL, "fps_start: null pointer");
  TPG_REQUIRE(workspace_bytes >= tpg_fps_workspace_bytes(B, N), TPG_EWORKSPACE, "fps_start: workspace too small");
  FpsArgs a{pts, B, N, D, k, start, idx, dist_rows, reinterpret_cast<float*>(workspace), 0};
  return fps_dispatch<true>(a, as_stream(stream));
}
