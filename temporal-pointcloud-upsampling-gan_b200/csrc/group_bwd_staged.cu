// Staged grouping backward (the gradient of pointnet2 `grouping_operation` / pytorch3d `knn_gather`,
// reference: models/pointnet2 grouping backward via pointnet2_ops, SURVEY.md §8 row "group bwd").
//
//   grad_f[b, c, n] = sum over l with idx[b, l] == n of grad_out[b, c, l]      (l = m * k + kk, ascending)
//
// The plain CSR kernel (group.cu) gathers 4-byte values straight from HBM/L2: every 32-byte sector of
// grad_out is touched by up to 8 different warps and every item index is a dependent L2 access, so the op
// runs at ~15 % of the HBM roofline.  Here a persistent CTA per SM keeps everything it indexes in shared memory:
//
//   * the CSR item list of the current batch element lives in shared memory as 16-bit row positions
//     (converted once per batch element with coalesced loads; a CTA owns a contiguous range of tiles, so
//     consecutive tiles share it);
//   * rows of grad_out (one per (b, c)) are cut into chunks of Lc floats; TC rows x one chunk form a stage; a
//     ring of 2-3 stages is filled by 1-D TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx) issued by
//     warp 0 ahead of the consumers -> grad_out is read from HBM exactly once, fully coalesced;
//   * every thread owns source points n (NPT of them) and walks their item lists with a cursor; items are
//     ascending in l, so the items falling into the staged chunk are a contiguous run of the list.  Values
//     are picked from shared memory and added in list order, which is exactly the summation order of the
//     oracle (oracle/tpg_oracle.c orc_group_bwd) and of group.cu: results are bit-identical;
//   * when N < 1024 the CTA is split into G = 1024 / roundup32(N) thread groups that work on different
//     channels of the same stage, and every thread serves TCG channels per item.
//
// Algorithmic bytes per call: 4 * B * (C*L + L + C*N)  (DESIGN.md §kernels).
#include <cuda_runtime.h>

#include <climits>
#include <cstdint>

#include "common.cuh"
#include "internal.cuh"

namespace tpg {
namespace {

constexpr int SB_THREADS = 1024;
constexpr int SB_MAX_STAGES = 3;
constexpr int SB_SMEM_BYTES = 224 * 1024;  // items + ring
constexpr int SB_MAX_TC = 32;

struct StagedArgs {
  const float* go;       // [B, C, L]
  const int32_t* off;    // [B, N + 1]
  const int32_t* items;  // [B, L], ascending inside a segment
  float* gf;             // [B, C, N]
  int B, C, N, L;        // L = positions of the segment being processed
  int Ltot, seg, NSEG;   // row length of grad_out, segment index, segments per cloud (1: the whole row)
  int init;              // 1: accumulators start from grad_f (the partial sums of the segments before this one)
  int G, NPs;            // thread groups per CTA and their stride in threads
  int TC, Lc, nch;       // channels per stage, chunk length (floats), chunks per row
  int S, items_bytes;    // ring stages; bytes reserved in front of the ring for the 16-bit item list
  int ctiles, ntiles;    // channel tiles per batch element, tiles in total
};

__device__ __forceinline__ uint32_t sb_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sb_mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void sb_mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t spin = 0; spin < (1u << 26); ++spin) {  // bounded: a protocol bug must trap, not hang the GPU
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
  }
  __trap();
}
__device__ __forceinline__ void sb_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sb_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void sb_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

template <int NPT, int TCG>
__global__ void __launch_bounds__(SB_THREADS, 1) group_bwd_staged_kernel(const StagedArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_s[SB_MAX_STAGES], empty_s[SB_MAX_STAGES];
  unsigned short* items_s = reinterpret_cast<unsigned short*>(smem_raw);
  float* ring = reinterpret_cast<float*>(smem_raw + a.items_bytes);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int stage_floats = a.TC * a.Lc;
  const int S = a.S;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      sb_mbar_init(sb_smem_u32(&full_s[s]), 1);
      sb_mbar_init(sb_smem_u32(&empty_s[s]), SB_THREADS / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  // a contiguous range of tiles (b-major): consecutive tiles share the batch element and its item list
  const int t0 = (int)((long long)blockIdx.x * a.ntiles / gridDim.x);
  const int t1 = (int)((long long)(blockIdx.x + 1) * a.ntiles / gridDim.x);
  const int my_chunks = (t1 - t0) * a.nch;

  // warp 0 stages chunk q of this CTA's chunk sequence (tile-major, chunk-minor)
  auto issue = [&](int q) {
    const int tile = t0 + q / a.nch, j = q % a.nch;
    const int b = tile / a.ctiles, c0 = (tile % a.ctiles) * a.TC;
    const int vc = min(a.TC, a.C - c0);
    const int len = min(a.Lc, a.L - j * a.Lc);
    const int st = q % S;
    const uint32_t bar = sb_smem_u32(&full_s[st]);
    if (lane == 0) sb_mbar_expect_tx(bar, (uint32_t)vc * (uint32_t)len * 4u);
    __syncwarp();
    if (lane < vc)
      sb_bulk_load(sb_smem_u32(ring + (size_t)st * stage_floats + (size_t)lane * a.Lc),
                   a.go + ((size_t)b * a.C + c0 + lane) * a.Ltot + (size_t)a.seg * a.L + (size_t)j * a.Lc,
                   (uint32_t)len * 4u, bar);
  };
  if (warp == 0)
    for (int q = 0; q < S - 1 && q < my_chunks; ++q) issue(q);

  // thread -> (channel group g, source points nl + p * 1024)
  const int g = tid / a.NPs, nl = tid - g * a.NPs;
  const bool active = g < a.G;
  const int ch0 = g * TCG;  // first channel (inside the stage) this thread serves

  int q = 0, items_b = -1;
  for (int tile = t0; tile < t1; ++tile) {
    const int b = tile / a.ctiles, c0 = (tile % a.ctiles) * a.TC;
    if (b != items_b) {  // (uniform over the CTA) 32-bit CSR items -> 16-bit row positions in shared memory
      __syncthreads();   // nobody still walks the previous list
      const int4* src = reinterpret_cast<const int4*>(a.items + ((size_t)b * a.NSEG + a.seg) * a.L);
      uint2* dst = reinterpret_cast<uint2*>(items_s);
      for (int e = tid; e < (a.L >> 2); e += SB_THREADS) {
        const int4 v = __ldg(src + e);
        dst[e] = make_uint2((unsigned)v.x | ((unsigned)v.y << 16), (unsigned)v.z | ((unsigned)v.w << 16));
      }
      items_b = b;
      __syncthreads();
    }
    const int32_t* ob = a.off + ((size_t)b * a.NSEG + a.seg) * (a.N + 1);
    int cur[NPT], end[NPT], nxt[NPT];
    float acc[NPT][TCG];
#pragma unroll
    for (int p = 0; p < NPT; ++p) {
      const int n = nl + p * SB_THREADS;
      cur[p] = end[p] = 0;
      if (active && n < a.N) {
        cur[p] = __ldg(ob + n);
        end[p] = __ldg(ob + n + 1);
      }
      nxt[p] = cur[p] < end[p] ? (int)items_s[cur[p]] : INT_MAX;
#pragma unroll
      for (int c = 0; c < TCG; ++c) {
        acc[p][c] = 0.0f;
        // later segments continue the sequential sum of the earlier ones: same order as one pass over the row
        if (a.init && active && n < a.N && c0 + ch0 + c < a.C) acc[p][c] = a.gf[((size_t)b * a.C + c0 + ch0 + c) * a.N + n];
      }
    }
    for (int j = 0; j < a.nch; ++j, ++q) {
      if (warp == 0 && q + S - 1 < my_chunks) {
        // the stage of chunk q-1 is free once every warp has left it
        if (q >= 1) sb_mbar_wait(sb_smem_u32(&empty_s[(q - 1) % S]), (uint32_t)(((q - 1) / S) & 1));
        issue(q + S - 1);
      }
      const int st = q % S;
      sb_mbar_wait(sb_smem_u32(&full_s[st]), (uint32_t)((q / S) & 1));
      const float* sp = ring + (size_t)st * stage_floats + (size_t)ch0 * a.Lc;
      const int lo = j * a.Lc;
      const unsigned len = (unsigned)min(a.Lc, a.L - lo);
#pragma unroll
      for (int p = 0; p < NPT; ++p) {
        while ((unsigned)(nxt[p] - lo) < len) {
          const int rel = nxt[p] - lo;
          ++cur[p];
          // (reads one entry past the list at its end: still inside items_s, the value is discarded)
          const int peek = (int)items_s[cur[p]];
#pragma unroll
          for (int c = 0; c < TCG; ++c) acc[p][c] = __fadd_rn(acc[p][c], sp[c * a.Lc + rel]);
          nxt[p] = cur[p] < end[p] ? peek : INT_MAX;
        }
      }
      __syncwarp();
      if (lane == 0) sb_mbar_arrive(sb_smem_u32(&empty_s[st]));
    }
    if (active) {
#pragma unroll
      for (int p = 0; p < NPT; ++p) {
        const int n = nl + p * SB_THREADS;
        if (n < a.N) {
#pragma unroll
          for (int c = 0; c < TCG; ++c)
            if (c0 + ch0 + c < a.C) a.gf[((size_t)b * a.C + c0 + ch0 + c) * a.N + n] = acc[p][c];
        }
      }
    }
  }
}

template <int NPT, int TCG>
int launch_staged(const StagedArgs& a, int grid, size_t smem, cudaStream_t st) {
  auto kern = group_bwd_staged_kernel<NPT, TCG>;
  // per call: the attribute is per device, and a process may drive several
  TPG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SB_SMEM_BYTES));
  kern<<<grid, SB_THREADS, smem, st>>>(a);
  TPG_CHECK_LAUNCH("group_bwd_staged_kernel");
  return TPG_OK;
}

}  // namespace

// segments the staged kernel needs for rows of L positions: 1 (whole row) up to L = 65536 (16-bit row positions in
// shared memory), else the smallest S with L % (4 S) == 0 and L / S <= 65536; 0 = not eligible
int group_bwd_staged_segments(int B, int C, int N, int L) {
  if (N < 1 || N > 8 * SB_THREADS || (L & 3) || L < 1024) return 0;
  if ((long long)B * C * L < (1LL << 18)) return 0;  // tiny: one launch of the plain kernel is cheaper
  if (L <= 65536) return N <= 4 * SB_THREADS ? 1 : 0;  // (N > 4096 with short rows: the plain kernel)
  for (int S = 2; S <= 16; ++S)
    if (L % (4 * S) == 0 && L / S <= 65536) return S;
  return 0;
}

bool group_bwd_staged_eligible(const float* go, const int32_t* items, int B, int C, int N, int L) {
  if (N < 1 || N > 4 * SB_THREADS || (L & 3) || L < 1024 || L > 65536) return false;  // 16-bit row positions
  if (((reinterpret_cast<uintptr_t>(go) | reinterpret_cast<uintptr_t>(items)) & 15) != 0) return false;
  if ((long long)B * C * L < (1LL << 18)) return false;  // tiny: one launch of the plain kernel is cheaper
  return true;
}

// force: 0 = heuristic, else 10 * stages + channels per thread
static int group_bwd_staged_seg(const float* go, const int32_t* off, const int32_t* items, int B, int C, int N, int L,
                                int Ltot, int seg, int S_seg, float* gf, int force, cudaStream_t st);

int group_bwd_staged(const float* go, const int32_t* off, const int32_t* items, int B, int C, int N, int L, float* gf,
                     int force, cudaStream_t st) {
  return group_bwd_staged_seg(go, off, items, B, C, N, L, L, 0, 1, gf, force, st);
}

// rows longer than 65536 positions: S launches, one per segment of L / S positions, each continuing the sums of the
// previous one (off / items: the inverse index of idx viewed as [B*S, L/S], positions relative to the segment)
int group_bwd_staged_segmented(const float* go, const int32_t* off, const int32_t* items, int B, int C, int N, int L,
                               int S, float* gf, cudaStream_t st) {
  TPG_REQUIRE(S >= 1 && L % (4 * S) == 0 && L / S <= 65536 && N <= 8 * SB_THREADS, TPG_EUNSUPPORTED,
              "group_bwd_segmented: unsupported (N=%d L=%d S=%d)", N, L, S);
  TPG_REQUIRE(((reinterpret_cast<uintptr_t>(go) | reinterpret_cast<uintptr_t>(items)) & 15) == 0, TPG_EUNSUPPORTED,
              "group_bwd_segmented: grad_out / items must be 16-byte aligned");
  for (int s = 0; s < S; ++s) {
    const int rc = group_bwd_staged_seg(go, off, items, B, C, N, L / S, L, s, S, gf, 0, st);
    if (rc) return rc;
  }
  return TPG_OK;
}

static int group_bwd_staged_seg(const float* go, const int32_t* off, const int32_t* items, int B, int C, int N, int L,
                                int Ltot, int seg, int S_seg, float* gf, int force, cudaStream_t st) {
  StagedArgs a{};
  a.go = go; a.off = off; a.items = items; a.gf = gf;
  a.B = B; a.C = C; a.N = N; a.L = L;
  a.Ltot = Ltot; a.seg = seg; a.NSEG = S_seg; a.init = seg > 0 ? 1 : 0;
  const int NPT = N <= SB_THREADS ? 1 : (N <= 2 * SB_THREADS ? 2 : (N <= 4 * SB_THREADS ? 4 : 8));
  a.NPs = N >= SB_THREADS ? SB_THREADS : ((N + 31) & ~31);
  a.G = SB_THREADS / a.NPs;
  a.items_bytes = (int)align_up((size_t)2 * L + 2, 128);  // + the one-past-the-end peek
  const int ring_floats = (SB_SMEM_BYTES - a.items_bytes) / 4;
  // channels per thread (one item walk serves TCG channels): as many as still leave >= min_tiles tiles and chunks
  // of >= 1536 floats (or whole rows; measured on the step's shapes, tools/bench_group_bwd.py: shorter chunks cost
  // less than walking the item list twice as often).  min_tiles is half a wave: the step replays many groupings concurrently,
  // so SM-time matters more than the latency of one call.
  int TCG = 1;
  const int max_tcg = NPT <= 2 ? 4 : (NPT == 4 ? 2 : 1);  // NPT = 8: one channel per item walk (registers)
  const int min_tiles = num_sms() / 2;
  for (int t = max_tcg; t >= 2; t >>= 1) {
    const long long TC = (long long)a.G * t;
    if (TC > SB_MAX_TC) continue;
    const long long lc = ring_floats / (2 * TC);
    if (lc >= min(L, 1536) && (long long)B * ((C + TC - 1) / TC) >= min_tiles) { TCG = t; break; }
  }
  int S = 0;
  if (force > 0) {
    S = force / 10;
    const int t = force % 10;
    if (t == 1 || ((t == 2 || t == 4) && t <= max_tcg)) TCG = t;
  }
  a.TC = a.G * TCG;
  if (a.TC > SB_MAX_TC) { a.G = SB_MAX_TC / TCG; a.TC = a.G * TCG; }
  auto chunks = [&](int stages) { const int lc = (ring_floats / (stages * a.TC)) & ~3; return lc < 4 ? INT_MAX : (L + lc - 1) / lc; };
  if (S != 2 && S != 3) S = chunks(3) <= chunks(2) ? 3 : 2;  // deeper ring unless it costs extra chunks
  a.S = S;
  a.nch = chunks(S);
  TPG_REQUIRE(a.nch != INT_MAX, TPG_EUNSUPPORTED, "group_bwd_staged: row does not fit");
  a.Lc = (((L + a.nch - 1) / a.nch) + 3) & ~3;
  a.nch = (L + a.Lc - 1) / a.Lc;
  a.ctiles = (C + a.TC - 1) / a.TC;
  a.ntiles = B * a.ctiles;
  const int grid = min(a.ntiles, num_sms());
  const size_t smem = (size_t)a.items_bytes + (size_t)S * a.TC * a.Lc * sizeof(float);
  switch (NPT * 10 + TCG) {
    case 11: return launch_staged<1, 1>(a, grid, smem, st);
    case 12: return launch_staged<1, 2>(a, grid, smem, st);
    case 14: return launch_staged<1, 4>(a, grid, smem, st);
    case 21: return launch_staged<2, 1>(a, grid, smem, st);
    case 22: return launch_staged<2, 2>(a, grid, smem, st);
    case 24: return launch_staged<2, 4>(a, grid, smem, st);
    case 41: return launch_staged<4, 1>(a, grid, smem, st);
    case 42: return launch_staged<4, 2>(a, grid, smem, st);
    case 81: return launch_staged<8, 1>(a, grid, smem, st);
  }
  set_error("group_bwd_staged: no variant for NPT=%d TCG=%d", NPT, TCG);
  return TPG_EUNSUPPORTED;
}

}  // namespace tpg
