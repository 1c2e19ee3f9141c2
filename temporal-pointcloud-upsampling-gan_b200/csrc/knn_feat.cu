// knn_feat.cu — K2: feature-space kNN (D = 32 / 64) on the 5th-gen tensor cores.
//
// Replaces pytorch3d knn_points on the generator's dynamic-graph features
// (gcn_lib/pointnet/gcn.py:200-203,258: D = 32/64, K = 4..20, P = 2048).
//
// The distance matrix is a genuine dense contraction, so it goes to tcgen05:
//   e[j][i] = |y_j|^2 - 2 <y_j, x_i>        (the |x_i|^2 term is constant per query)
// Operands are prepared by a pre-pass (feat_split_kernel): every cloud is CENTRED on its mean
// (distances are translation-invariant; real network features carry a large common offset) and
// every centred value is split into two bf16 terms, v ~ b1 + b2 (16 significand bits), stored as
// rows [b1(0..D-1) | b2(0..D-1)] — the same bytes per row as the fp32 input.  <y,x> is then the
// sum of the bf16 x bf16 products b1b1 + b1b2 + b2b1 (b2b2 <= 2^-16 is dropped; tcgen05.mma kind::f16,
// bf16 inputs, exact products, fp32 accumulate in TMEM): ~2^-15 relative instead of the 2^-10 of a
// tf32 contraction.  (Measured on the generator's real activations, tools/dump_knn_inputs.py: a
// plain tf32 contraction sends 77-98 % of the queries to the exact fallback, centring alone
// 15-34 %, centring + split 0.0 %.)  e only SELECTS candidates; the neighbours that are
// returned are re-ranked with the canonical distance (sequential fp32, no FMA) on the ORIGINAL
// rows, so indices and distances are identical to the brute-force kernel:
//   * pass 0 over the e-matrix keeps, per query, the minima of 64 disjoint candidate groups; the K smallest
//     group minima are K distinct candidates, so their K-th smallest T bounds the K-th smallest e overall;
//   * with eps a rigorous bound on |e - (d_true - |x|^2)| + |d_canon - d_true| (feat_eps), every canonical
//     top-K neighbour has e <= T + 2 eps (proof in DESIGN.md §4 K2); pass 1 re-issues the MMAs and records,
//     as one ballot per (query, 32 candidates), WHICH candidates have e <= tau0 = T + 2.25 eps
//     (~K + 5 per query) — bit masks in global memory, no atomics, no per-hit stores;
//   * knn_feat_rank_kernel (one warp per query, many warps per SM) expands the masks, sums the canonical
//     distance of every hit and writes the K best by (d_canon, idx); it re-checks completeness on the
//     canonical distances themselves (K-th smallest < tau0 + |x|^2 - eps) and sends a query with more than
//     64 hits (near-duplicate features, huge norms) to the exact SIMT fallback kernel.
//
// Kernel anatomy (knn_feat_tc_kernel: one CTA = 128 queries of one cloud, one wave of CTAs; 16 epilogue warps
// + one MMA-issuer warp + one TMA-producer warp):
//   MMA shape M=128 (candidates) x N=128 (queries) x K=16 (bf16) per instruction, cta_group::1.
//   Candidates are the M operand on purpose: TMEM lane == candidate, so a thread owns one
//   candidate row of the accumulator tile and sweeps its queries without any cross-lane
//   operation.  smem: 4-stage ring of candidate tiles (128 x 2D bf16, K-major, 128B-swizzled,
//   filled by TMA: cp.async.bulk.tensor.2d, one box per 32-float K-slab, complete_tx on a
//   per-stage mbarrier; UMMA descriptors built by hand) + the query tile; TMEM: 4 x 128 columns
//   (three MMAs in flight behind the tile being drained).  mbarrier pipeline, no CTA barrier in the loop:
//     full[stage]   TMA -> MMA issuer            tile landed
//     done[buf]     tcgen05.commit -> epilogue, producer   accumulator ready, smem stage free
//     tfree[buf]    16 epilogue warps -> MMA issuer         accumulator drained (tcgen05.ld complete)
//   issuer warp (warp-uniform control flow, one elected lane): for every tile u: wait full / tfree, issue
//     3 x D/16 MMAs back to back from uniform registers, commit (~91 cycles per MMA: the tensor pipe's rate);
//   producer warp (one elected lane): wait done(u), TMA tile u+4 into the stage MMA(u) read;
//   epilogue warps: wait done(u); tcgen05.ld -> registers; per-lane min (pass 0) / one ballot per query
//   column (pass 1); arrive tfree(u).
#include "common.cuh"
#include "internal.cuh"

#include <cuda_bf16.h>
#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)
#include <stdlib.h>

namespace tpg {

constexpr int FT_THREADS = 512;   // 16 epilogue warps (4 per TMEM lane quarter)
constexpr int FT_BLOCK = FT_THREADS + 64;  // + one warp that only issues the MMAs + one that only issues the TMA loads
constexpr int FT_TM = 128;   // candidates per tile (UMMA M)
constexpr int FT_NQ = 128;   // queries per CTA (UMMA N): 8 clouds x 16 CTAs = 128 CTAs, one wave on 148 SMs
constexpr int FT_QW = 32;    // queries (accumulator columns) per warp
constexpr int FT_NBUF = 4;             // TMEM accumulator buffers: MMA(u+1..u+3) in flight while tile u is consumed
constexpr int FT_TMEM_COLS = FT_NBUF * FT_NQ;  // 512 columns: all of TMEM (one CTA per SM)
constexpr int FT_MAX_K = 24;
constexpr int FT_NMAX_PARTS = 32;  // per-cloud partial maxima of the candidate norms (one per CTA of the split launch)
constexpr int FT_STAGES = 4;       // candidate-tile ring: tiles u+1..u+3 feed the MMAs in flight, u+4 is loading

struct FeatArgs {
  const float* p1;
  const float* p2;
  const int64_t* len1;
  const int64_t* len2;
  int B, P1, P2, D, K;
  const float* nrm1;       // [B,P1] squared norms of the (centred, split) queries
  const float* nrm2;       // [B,P2] squared norms of the (centred, split) candidates
  const unsigned* nmax2;   // [B][FT_NMAX_PARTS] partial maxima of the squared candidate norms (float bits)
  float* dists;            // [B,P1,K]
  int64_t* idx;            // [B,P1,K]
  int* fb_count;           // [1]
  int* fb_list;            // [B*P1]
  unsigned* masks;         // [B][P1][4*T] hit masks of pass 1: word (t*4 + quarter) bit l = candidate t*128 + quarter*32 + l
  float* tau0;             // [B,P1] admission bound of every query
  int T;                   // candidate tiles per cloud = ceil(P2 / 128)
  long long* dbg;          // [ctas][8] phase timestamps (tools/bench_knn_feat.py)
  const int32_t* skip;     // nonzero -> every kernel of the call returns at once (results come from a memoised call)
};

struct FeatMaps {  // TMA descriptors over the split operands: [B*P, 2D] bf16 moved as [B*P, D] 4-byte words,
                   // box 32 words x 128 rows, 128B swizzle
  CUtensorMap cand, query;
};

// ---- PTX wrappers -------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  // bounded spin: a descriptor bug must surface as a trap, never as a hung GPU
  for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
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

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
// one 32-float x 128-row box of a [rows, D] fp32 matrix -> 128B-swizzled K-major slab (16 KB)
__device__ __forceinline__ void tma_load_slab(uint32_t dst, const CUtensorMap* map, int col, int row, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(dst),
      "l"(map), "r"(col), "r"(row), "r"(bar)
      : "memory");
}

// barrier among the 16 epilogue warps only (the MMA-issuer warp never joins it)
__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, 512;\n" ::: "memory"); }

// K-major, 128B-swizzled shared-memory matrix descriptor (sm_100 "version 1"):
//   start address >> 4 | LBO(=1, unused for swizzled K-major) << 16 | SBO(1024 B) >> 4 << 32
//   | version 1 << 46 | layout SWIZZLE_128B (2) << 61
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}
// instruction descriptor: D=f32 (1<<4), A=B=bf16 (1<<7, 1<<10), K-major both, N>>3 at 17, M>>4 at 24
constexpr uint32_t FT_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(FT_NQ >> 3) << 17) |
                              ((uint32_t)(FT_TM >> 4) << 24);

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(FT_IDESC), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// rigorous bound on |e - (d_true - |x|^2)| + |d_canon - d_true|, with xs, ys the split
// values (b1 + b2) of the centred rows x_c, y_c, xn = |xs|, yn = max |ys|, s = xn + yn:
//   contraction: three of the four term pairings are issued (b1b1 + b1b2 + b2b1); the dropped b2*b2 is at most
//     2^-16 xn yn (2^-15 in e).  The 3D bf16 products are exact; fp32 accumulation inside the tensor core, taken
//     as one truncation (2^-23) per product: <= 3D 2^-23 xn yn, twice that in e;
//   split: |xs - x_c| <= 2^-16 |x_c| per coordinate (two round-to-nearest bf16 steps), so
//     |d(xs,ys) - d(x_c,y_c)| <= 2 s * 2^-16 s (+ second order) = 2^-15 s^2;
//   centring (relative 2^-24 per coordinate), fp32 roundings of the norms, of e and of the canonical sum:
//     (2D+32) 2^-24 s^2;  x1.25 safety on everything.
__device__ __forceinline__ float feat_eps(float nq, float nmax, int D) {
  const float xn = sqrtf(nq) * 1.001f, yn = sqrtf(nmax) * 1.001f;
  const float s = xn + yn;
  return 1.25f * (((float)(6 * D) * 1.1920929e-7f + 3.0517578e-5f) * xn * yn +
                  (3.0517578e-5f + (float)(2 * D + 32) * 5.9604645e-8f) * s * s);
}

// same bound as feat_eps with MUFU square roots (2 ulp; the 1.001 factors cover them)
__device__ __forceinline__ float feat_eps_fast(float nq, float nmax, int D) {
  const float xn = __frcp_rn(rsqrtf(nq)) * 1.001f, yn = __frcp_rn(rsqrtf(nmax)) * 1.001f;
  const float s = xn + yn;
  return 1.25f * (((float)(6 * D) * 1.1920929e-7f + 3.0517578e-5f) * xn * yn +
                  (3.0517578e-5f + (float)(2 * D + 32) * 5.9604645e-8f) * s * s);
}

// k-th smallest (1-based) of the 64 unsigned keys a warp holds two per lane (element 2 * lane + s), for NQ
// independent key sets at once: a bitonic sorting network — 21 compare-exchange stages, 15 of them one SHFL + one
// predicated min/max per key, 6 inside the lane.  ~110 instructions per set instead of the ~340 of the 16-step
// radix select it replaces (which bounded the kernel's admission-bound phase at 10.7 K cycles), and the result is
// the exact k-th smallest, not the upper end of a radix bucket.  Keys of absent values must be 0xffffffff.
template <int NQ>
__device__ __forceinline__ void warp_kth_smallest64_multi(unsigned (&v)[NQ][2], int k, unsigned (&out)[NQ]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int kk = 2; kk <= 64; kk <<= 1) {
    const bool asc = ((2 * lane) & kk) == 0;  // direction of this element's block (kk == 64: ascending everywhere)
#pragma unroll
    for (int j = kk >> 1; j >= 1; j >>= 1) {
      if (j == 1) {  // partner = the lane's other key
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const unsigned lo = min(v[q][0], v[q][1]), hi = max(v[q][0], v[q][1]);
          v[q][0] = asc ? lo : hi;
          v[q][1] = asc ? hi : lo;
        }
      } else {       // partner = the same slot of lane ^ (j / 2)
        const int lj = j >> 1;
        const bool take_max = ((lane & lj) != 0) == asc;  // the higher element of an ascending pair keeps the maximum
#pragma unroll
        for (int q = 0; q < NQ; ++q)
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2) {
            const unsigned p = __shfl_xor_sync(FULL, v[q][s2], lj);
            v[q][s2] = take_max ? max(v[q][s2], p) : min(v[q][s2], p);
          }
      }
    }
  }
  const int e = k - 1;
#pragma unroll
  for (int q = 0; q < NQ; ++q) out[q] = __shfl_sync(FULL, (e & 1) ? v[q][1] : v[q][0], e >> 1);
}
__device__ __forceinline__ unsigned ordered_key(float f) {  // unsigned order == float order
  const int bits = __float_as_int(f);
  return (unsigned)(bits ^ ((bits >> 31) & 0x7fffffff)) ^ 0x80000000u;
}
__device__ __forceinline__ float ordered_key_inv(unsigned uk) {
  const int key = (int)(uk ^ 0x80000000u);
  int bits = key ^ ((key >> 31) & 0x7fffffff);
  if ((bits & 0x7fffffff) > 0x7f800000) bits = 0x7f800000;  // bucket of +inf
  return __int_as_float(bits);
}

// ---- pre-pass: centre, centred bf16 x 2 split, squared norms, per-cloud max norm ------------------------------
// ONE kernel.  The centre is the mean of FT_CENTRE_ROWS rows sampled evenly over the candidate cloud, recomputed by
// every CTA (128 rows from L2, fixed summation order): ANY common shift of both operands is correct — the centre
// only sets the size of the norms, hence of the margins — and a sampled mean removes the large common offset of
// real network features as well as the exact one, without a second kernel in front of this one.
constexpr int FT_CENTRE_ROWS = 128;

// One float4 chunk per thread (LPR = D/4 lanes per row): v = x - centre; b1 = bf16_rn(v); b2 = bf16_rn(v - b1);
// out row = [b1(0..D-1) | b2(0..D-1)] (2D bf16 = the bytes of the fp32 row); nrm = sum (b1 + b2)^2.
// grid.x <= FT_NMAX_PARTS CTAs per cloud, each looping over its row groups; CTA x owns nmax_part[b][x].
__global__ void __launch_bounds__(256) feat_split_kernel(const float* __restrict__ p, const float* __restrict__ pc,
                                                         const int64_t* __restrict__ len_c, int Pc, int P, int D,
                                                         uint2* __restrict__ split, float* __restrict__ nrm,
                                                         unsigned* __restrict__ nmax_part, const int32_t* __restrict__ skip) {
  __shared__ float red_s[256];
  __shared__ float mean_s[64];
  if (skip && *skip) return;
  const int b = blockIdx.y, tid = threadIdx.x;
  const int n = len_c ? min((int)len_c[b], Pc) : Pc;  // valid rows of the candidate cloud
  {
    // all of a thread's sample rows are requested before the first is used (the kernel is pure latency)
    const int col = tid % D, sub = tid / D, nsub = 256 / D;
    constexpr int PER = FT_CENTRE_ROWS / 4;  // >= rows per thread (nsub = 8 or 4)
    float v[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const int i = sub + u * nsub;
      v[u] = 0.0f;
      if (n > 0 && i < FT_CENTRE_ROWS) v[u] = __ldg(pc + ((size_t)b * Pc + (int)(((long long)i * n) / FT_CENTRE_ROWS)) * D + col);
    }
    float acc = 0.0f;
#pragma unroll
    for (int u = 0; u < PER; ++u) acc += v[u];
    red_s[tid] = acc;
    __syncthreads();
    if (tid < D) {
      float t = 0.0f;
      for (int u = 0; u < nsub; ++u) t += red_s[u * D + tid];
      mean_s[tid] = t * (1.0f / FT_CENTRE_ROWS);
    }
    __syncthreads();
  }
  const int lpr = D >> 2, rows_per_it = 256 / lpr, ch = tid % lpr;
  const int iters = (P + (int)gridDim.x * rows_per_it - 1) / ((int)gridDim.x * rows_per_it);
  unsigned mloc = 0u;
  const float m0 = mean_s[4 * ch], m1 = mean_s[4 * ch + 1], m2 = mean_s[4 * ch + 2], m3 = mean_s[4 * ch + 3];
  for (int it0 = 0; it0 < iters; it0 += 4) {
    float4 rv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {  // up to 4 row groups in flight
      const int row = (blockIdx.x * iters + it0 + u) * rows_per_it + tid / lpr;
      if (it0 + u < iters && row < P) rv[u] = __ldg(reinterpret_cast<const float4*>(p + ((size_t)b * P + row) * D) + ch);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (it0 + u >= iters) break;  // (uniform)
      const int row = (blockIdx.x * iters + it0 + u) * rows_per_it + tid / lpr;
      float s = 0.0f;
      if (row < P) {
        const float c[4] = {rv[u].x - m0, rv[u].y - m1, rv[u].z - m2, rv[u].w - m3};
        unsigned short h1[4], h2[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const __nv_bfloat16 b1 = __float2bfloat16_rn(c[k]);
          const float f1 = __bfloat162float(b1);
          const __nv_bfloat16 b2 = __float2bfloat16_rn(c[k] - f1);
          const float xs = f1 + __bfloat162float(b2);  // exact in fp32 (<= 17 significant bits)
          s = fmaf(xs, xs, s);
          h1[k] = __bfloat16_as_ushort(b1);
          h2[k] = __bfloat16_as_ushort(b2);
        }
        uint2* orow = split + ((size_t)b * P + row) * (size_t)(D >> 1);  // row = D/2 uint2 (2D bf16)
        orow[ch] = make_uint2((unsigned)h1[0] | ((unsigned)h1[1] << 16), (unsigned)h1[2] | ((unsigned)h1[3] << 16));
        orow[lpr + ch] = make_uint2((unsigned)h2[0] | ((unsigned)h2[1] << 16), (unsigned)h2[2] | ((unsigned)h2[3] << 16));
      }
      // sum over the row's lanes (lpr = 8 or 16 consecutive lanes)
      for (int o = lpr >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
      if (row < P && ch == 0) nrm[(size_t)b * P + row] = s;
      // rows beyond the cloud's length never become candidates; non-finite norms (inf / NaN inputs) must poison
      // the bound, not vanish in the integer max: NaN -> +inf
      const float sm = (row < n) ? ((s == s) ? s : __int_as_float(0x7f800000)) : 0.0f;
      mloc = max(mloc, __float_as_uint(sm));  // s >= 0: bit order == value order
    }
  }
  if (nmax_part) {
    const unsigned m = __reduce_max_sync(FULL, mloc);
    __syncthreads();  // red_s is free again
    if ((tid & 31) == 0) reinterpret_cast<unsigned*>(red_s)[tid >> 5] = m;
    __syncthreads();
    if (tid == 0) {
      unsigned t = 0u;
      for (int w = 0; w < 8; ++w) t = max(t, reinterpret_cast<unsigned*>(red_s)[w]);
      nmax_part[(size_t)b * FT_NMAX_PARTS + blockIdx.x] = t;
    }
    // slots of CTAs that do not exist
    if (blockIdx.x == 0 && tid >= (int)gridDim.x && tid < FT_NMAX_PARTS) nmax_part[(size_t)b * FT_NMAX_PARTS + tid] = 0u;
  }
}

// ---- main tcgen05 kernel -----------------------------------------------------------------------
// 16 warps: warp w reads TMEM lane quarter (w & 3) — candidates 32(w&3)..+31 of every tile —
// and owns the column group (w >> 2): 32 of the CTA's 128 queries (accumulator columns).
// The kernel is bound by the latency of its 2*T dependent tile iterations, so the CTA is made
// as wide as TMEM/smem allow and the whole problem runs as ONE wave of CTAs.
//
// Two passes over the e-matrix (the MMAs are simply issued twice: the tiles come from L2, and re-issuing costs less
// than keeping 1 MB of accumulators per CTA).  Cross-lane work is one ballot per query column in pass 1 only:
//   pass 0  every lane keeps, per query, the running minimum of the e-values of the candidates it
//           sees (one FMNMX per value): 128 disjoint candidate groups per query (4 lane quarters
//           x 32 lanes), merged pairwise to 64.  The K smallest group minima are K distinct candidates, so the
//           K-th smallest of them (bitonic sort of the 64 merged minima in registers) bounds the K-th smallest e overall;
//           tau0 = that bound + 2.25 eps.
//   pass 1  the 32 bounds of a warp's queries sit in registers; for every query column the warp ballots
//           "e <= tau0" over its 32 candidates and publishes the 32-bit hit mask ([4T][P1] words per cloud).
// The ranking (canonical distances of the hits, K best by (d_canon, idx)) is knn_feat_rank_kernel below.

// one lane of a converged warp (elect.sync): the branch it guards stays warp-uniform for the compiler, so the
// operands of the tcgen05 / TMA instructions inside live in uniform registers
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}

template <int DD>
__global__ void __launch_bounds__(FT_BLOCK, 1) knn_feat_tc_kernel(FeatArgs a, const __grid_constant__ FeatMaps maps) {
  extern __shared__ __align__(1024) unsigned char ft_smem_raw[];
  __shared__ __align__(8) uint64_t mbar_s[FT_NBUF];       // MMA(u) done (tcgen05.commit), per TMEM buffer
  __shared__ __align__(8) uint64_t full_s[FT_STAGES];     // candidate tile landed (TMA complete_tx), per smem stage
  __shared__ __align__(8) uint64_t qfull_s;               // query tile landed
  __shared__ __align__(8) uint64_t tfree_s[FT_NBUF];      // all 16 warps have drained the TMEM buffer
  __shared__ uint32_t tmem_base_s;
  __shared__ float tau0_s[FT_NQ];
  __shared__ unsigned wmask_s[FT_THREADS / 32][32];   // pass 1: a warp's 32 hit masks on their way to lane-major order

  if (a.skip && *a.skip) return;  // (uniform) memoised call: nothing to do, nothing allocated yet
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *a.fb_count = 0;  // (replaces a memset node) counted up by the rank kernel
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(FULL, tid >> 5, 0);  // warp-uniform for the compiler too (role branches, uniform registers)
  const int quarter = warp & 3, colgrp = warp >> 2;
  const int nq0 = colgrp * FT_QW;  // first query (column) of this warp
  const int b = blockIdx.y, q0 = blockIdx.x * FT_NQ;
  constexpr int D = DD;
  const int K = a.K;
  const int n1 = a.len1 ? min((int)a.len1[b], a.P1) : a.P1;
  const int n2 = a.len2 ? min((int)a.len2[b], a.P2) : a.P2;
  const float INF = __int_as_float(0x7f800000);

  // dynamic smem, 1024-aligned: [stage0 | stage1 | queries | group values / candidate buffers]
  const uint32_t raw = smem_u32(ft_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* base_ptr = ft_smem_raw + (base - raw);
  const uint32_t stage_bytes = (uint32_t)D * 512u;   // 128 rows x D floats
  const uint32_t atomA = 128u * 128u;                // one 32-float K-slab of a stage
  const uint32_t qtile = base + (uint32_t)FT_STAGES * stage_bytes;
  const uint32_t atomB = (uint32_t)FT_NQ * 128u;
  const int T = (n2 + FT_TM - 1) / FT_TM;

  // candidate tile t -> smem stage: one TMA box per 32-float K-slab, issued by ONE thread; rows of
  // the next cloud that ride along in a ragged last tile are masked by their +inf norm
  constexpr int slabs = D >> 5;
  auto load_tile = [&](int t, int stage) {
    const uint32_t bar = smem_u32(&full_s[stage]);
    mbar_expect_tx(bar, stage_bytes);
    for (int sl = 0; sl < slabs; ++sl)
      tma_load_slab(base + (uint32_t)stage * stage_bytes + (uint32_t)sl * atomA, &maps.cand, sl * 32, b * a.P2 + t * FT_TM, bar);
  };

  long long* dbg = a.dbg ? a.dbg + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 16 : nullptr;
  if (dbg && tid == 0) dbg[0] = clock64();
  // ---- setup ----
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(FT_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 0) {
    for (int st = 0; st < FT_NBUF; ++st) {
      mbar_init(smem_u32(&mbar_s[st]), 1);
      mbar_init(smem_u32(&tfree_s[st]), FT_THREADS / 32);
    }
    for (int st = 0; st < FT_STAGES; ++st) mbar_init(smem_u32(&full_s[st]), 1);
    mbar_init(smem_u32(&qfull_s), 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  }
  const int U = 2 * T;
  unsigned char* aux = base_ptr + (uint32_t)FT_STAGES * stage_bytes + (uint32_t)D * 4u * FT_NQ;
  float* gval_s = reinterpret_cast<float*>(aux);    // pass 0 -> select: [FT_NQ][64] group minima

  __syncthreads();  // barriers initialised, TMEM allocated
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (tid == 0) {
    mbar_expect_tx(smem_u32(&qfull_s), (uint32_t)D * 4u * FT_NQ);
    for (int sl = 0; sl < slabs; ++sl)
      tma_load_slab(qtile + (uint32_t)sl * atomB, &maps.query, sl * 32, b * a.P1 + q0, smem_u32(&qfull_s));
    for (int u0 = 0; u0 < FT_STAGES && u0 < U; ++u0) load_tile(u0 % T, u0);
  }

  // MMA(u): needs tile u in smem and TMEM buffer u % 4 drained by the epilogue of u - 4.  Run by ALL lanes of the
  // issuer warp in warp-uniform control flow (descriptors stay in uniform registers: a one-lane branch made the
  // compiler move every operand through an ELECT / R2UR loop, ~33 dependent instructions and ~295 cycles per
  // MMA — the issue, not the tensor pipe, bounded both passes); one elected lane issues.
  constexpr int KH = D >> 4;  // K-steps (16 bf16 = 32 bytes) per term: a row = [b1: KH steps | b2: KH steps]
  auto koff = [](int kk) -> uint64_t { return (uint64_t)(((uint32_t)(kk >> 2) * (128u * 128u) + (uint32_t)(kk & 3) * 32u) >> 4); };
  auto issue_mma = [&](int u) {
    const int s = u % FT_NBUF;  // TMEM buffer / mbarrier
    mbar_wait(smem_u32(&full_s[u % FT_STAGES]), (uint32_t)((u / FT_STAGES) & 1));
    if (u >= FT_NBUF) mbar_wait(smem_u32(&tfree_s[s]), (uint32_t)((u / FT_NBUF - 1) & 1));
    tc_fence_after();
    const uint32_t td = tmem_base + (uint32_t)s * FT_NQ;
    // descriptors differ from their stage base only in the 14-bit start-address field (>> 4)
    const uint64_t ad0 = umma_desc_sw128(base + (uint32_t)(u % FT_STAGES) * stage_bytes);
    const uint64_t bd0 = umma_desc_sw128(qtile);
    if (elect_one()) {
      // <y,x> ~ b1b1 + b1b2 + b2b1, all accumulated into one TMEM tile; b2*b2 (<= 2^-16 |x||y|, accounted for in
      // feat_eps) is not worth a quarter of the tensor-pipe time
#pragma unroll
      for (int pr = 0; pr < 3; ++pr) {
        const int ta = pr == 2 ? 1 : 0, tb = pr == 1 ? 1 : 0;
#pragma unroll
        for (int kk = 0; kk < KH; ++kk)
          umma_bf16(td, ad0 + koff(ta * KH + kk), bd0 + koff(tb * KH + kk), (pr | kk) ? 1u : 0u);  // atomA == atomB
      }
      umma_commit(smem_u32(&mbar_s[s]));
    }
    __syncwarp();
  };
  if (warp == FT_THREADS / 32 + 1) {
    // ===== TMA-producer warp: refills a stage the moment the MMA that read it has completed =====
    for (int u = 0; u + FT_STAGES < U; ++u) {
      mbar_wait(smem_u32(&mbar_s[u % FT_NBUF]), (uint32_t)((u / FT_NBUF) & 1));
      if (elect_one()) load_tile((u + FT_STAGES) % T, u % FT_STAGES);
      __syncwarp();
    }
    __syncthreads();
    return;
  }
  if (warp == FT_THREADS / 32) {
    // ===== MMA-issuer warp: feeds the tensor pipe as fast as tiles land and accumulators drain; it does no
    //       epilogue work, so issue never waits behind this warp's own share =====
    mbar_wait(smem_u32(&qfull_s), 0u);
    if (dbg) {  // tuning: where the issuer's time goes (tools/bench_knn_feat.py)
      long long tf = 0, tt = 0, ti = 0;
      for (int u = 0; u < U; ++u) {
        const long long c0 = clock64();
        mbar_wait(smem_u32(&full_s[u % FT_STAGES]), (uint32_t)((u / FT_STAGES) & 1));
        const long long c1 = clock64();
        if (u >= FT_NBUF) mbar_wait(smem_u32(&tfree_s[u % FT_NBUF]), (uint32_t)((u / FT_NBUF - 1) & 1));
        const long long c2 = clock64();
        issue_mma(u);
        tf += c1 - c0; tt += c2 - c1; ti += clock64() - c2;
      }
      if (lane == 0) { dbg[8] = tf; dbg[9] = tt; dbg[10] = ti; }
    } else {
      for (int u = 0; u < U; ++u) issue_mma(u);
    }
    __syncthreads();  // joins the epilogue warps before the TMEM dealloc
    return;
  }
  if (dbg && tid == 0) dbg[1] = clock64();

  // pass 0: running group minima; pass 1: admission bounds of this warp's 32 queries
  float reg[FT_QW];
#pragma unroll
  for (int n = 0; n < FT_QW; ++n) reg[n] = INF;
  unsigned* wmask = &wmask_s[warp][0];
  unsigned* mrow = a.masks + (size_t)b * (size_t)(4 * T) * a.P1;   // this cloud's mask words: [P1][4T] (query-major)
  const bool qcol_ok = q0 + nq0 + lane < a.P1;                       // lane c publishes the masks of query q0 + nq0 + c
  // for the admission bound (tau0): the cloud's largest candidate norm and the norms of this warp's QPW queries
  const unsigned nmax_part = __ldg(a.nmax2 + (size_t)b * FT_NMAX_PARTS + lane);
  const float nq_lane = __ldg(a.nrm1 + (size_t)b * a.P1 + min(q0 + warp * (FT_NQ / (FT_THREADS / 32)) + (lane & 7), a.P1 - 1));

  for (int u = 0; u < U; ++u) {
    const int t = u < T ? u : u - T;
    const bool list_pass = u >= T;
    if (u == T) {
      if (dbg && tid == 0) dbg[2] = clock64();
#pragma unroll
      for (int n = 0; n < FT_QW; ++n) {  // 64 groups per query: lanes l and l^16 merge (halves the smem footprint)
        const float m2 = fminf(reg[n], __shfl_xor_sync(FULL, reg[n], 16));
        if (lane < 16) gval_s[(nq0 + n) * 64 + quarter * 16 + lane] = m2;
      }
      if (dbg && tid == 0) dbg[11] = clock64();
      epi_sync();
      if (dbg && tid == 0) dbg[12] = clock64();
      // ---- tau0 = R-th smallest of the 128 group minima of a query; the warp's queries are
      //      processed together so the stages of their sorting networks interleave
      constexpr int QPW = FT_NQ / (FT_THREADS / 32);
      // The K smallest group minima are K distinct candidates, so their K-th smallest T bounds the K-th smallest
      // e overall; every canonical top-K neighbour has e <= e_(K) + 2 eps <= T + 2 eps (eps = feat_eps: the error of
      // e plus that of the canonical sum), so tau0 = T + 2.25 eps admits a superset — about K + 5 hits per query
      // (group collisions), and the ranking kernel's completeness check then holds by construction.
      const int R = min(32, K);
      unsigned uk[QPW][2];
#pragma unroll
      for (int qq = 0; qq < QPW; ++qq)
#pragma unroll
        for (int v = 0; v < 2; ++v) uk[qq][v] = ordered_key(gval_s[(warp * QPW + qq) * 64 + v * 32 + lane]);
      unsigned bound[QPW];
      warp_kth_smallest64_multi<QPW>(uk, R, bound);
      if (dbg && tid == 0) dbg[13] = clock64();
      const unsigned nm = __reduce_max_sync(FULL, nmax_part);  // (loaded before the passes)
      float eps_q = 0.0f;
      if (lane < QPW) eps_q = feat_eps_fast(nq_lane, __uint_as_float(nm), D);
#pragma unroll
      for (int qq = 0; qq < QPW; ++qq) {
        const float tq = ordered_key_inv(bound[qq]) + 2.25f * __shfl_sync(FULL, eps_q, qq);
        if (lane == 0) {
          tau0_s[warp * QPW + qq] = tq;
          if (q0 + warp * QPW + qq < a.P1) a.tau0[(size_t)b * a.P1 + q0 + warp * QPW + qq] = tq;
        }
      }
      epi_sync();
#pragma unroll
      for (int n = 0; n < FT_QW; ++n) reg[n] = tau0_s[nq0 + n];
      if (dbg && tid == 0) dbg[3] = clock64();
    }
    const int j = t * FT_TM + quarter * 32 + lane;
    const float ncj = j < n2 ? __ldg(a.nrm2 + (size_t)b * a.P2 + j) : INF;
    mbar_wait(smem_u32(&mbar_s[u % FT_NBUF]), (uint32_t)((u / FT_NBUF) & 1));
    tc_fence_after();
    uint32_t acc[FT_QW];
    tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((u % FT_NBUF) * FT_NQ + nq0), acc);
    if (!list_pass) {
#pragma unroll
      for (int n = 0; n < FT_QW; ++n) reg[n] = fminf(reg[n], fmaf(-2.0f, __uint_as_float(acc[n]), ncj));
    } else {
      // Pass 1 records WHICH candidates lie under the bound, nothing else: one ballot per query column gives the
      // 32-bit hit mask of this warp's 32 candidates (no shared-memory atomics, no per-hit stores: the first
      // versions appended (e, idx) pairs through smem counters and spent 2700-4000 cycles per tile on the
      // ATOMS round trips).  Lane 0 parks the 32 masks in smem, lane c publishes the word of column c.
      const bool valid = j < n2;
#pragma unroll
      for (int c = 0; c < FT_QW; ++c) {
        const float e = fmaf(-2.0f, __uint_as_float(acc[c]), ncj);  // inf for padded candidates
        const unsigned m = __ballot_sync(FULL, valid && e <= reg[c]);
        if (lane == 0) wmask[c] = m;
      }
      __syncwarp();
      if (qcol_ok) mrow[(size_t)(q0 + nq0 + lane) * (size_t)(4 * T) + t * 4 + quarter] = wmask[lane];
    }
    // this warp has drained TMEM buffer u % 4 (tcgen05.wait::ld inside tmem_ld32): MMA(u+4) may overwrite it.
    // No CTA-wide barrier in the loop: warps run ahead until the next MMA-done barrier.
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&tfree_s[u % FT_NBUF]));
  }
  if (dbg && tid == 0) dbg[4] = clock64();
  __syncthreads();
  if (dbg && tid == 0) dbg[5] = clock64();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(FT_TMEM_COLS) : "memory");
  }
}

// ---- ranking kernel: one warp per query, many warps per SM (the chain below is all latency) -----------------
// The hits of a query (pass 1 masks) are expanded into a candidate list, their rows are fetched with coalesced
// 16-byte loads into a padded per-warp tile, lane r sums row r sequentially in d order (the canonical distance),
// and the K best by (d_canon, idx) are written.  Completeness: a candidate j that was NOT recorded has
// e_j > tau0, hence d_canon(j) > tau0 + |x|^2 - eps (eps bounds |e - (d_true - |x|^2)| + |d_canon - d_true|,
// feat_eps); so when the K-th smallest canonical distance among the hits lies strictly below that bound, no
// unrecorded candidate can enter (or tie into) the result.  Otherwise, or with more than FT_HCAP hits, the query
// goes to the exact fallback.
// (Tried: a CTA per 128 queries that streams the whole candidate cloud through shared memory in padded stages and
// sums the canonical distances from there, 32 (query, row) items at a time: 28 / 35 us against 23 / 27 us for this
// gather at 8 x 2048^2, D = 32 / 64 — the staging of 256-512 KB per CTA costs more than the L2 gathers it saves.)
constexpr int FT_HCAP = 64;          // hits handled per query (two rounds of 32)
constexpr int FT_RANK_WARPS = 8;

template <int DD>
__global__ void __launch_bounds__(FT_RANK_WARPS * 32, 4) knn_feat_rank_kernel(FeatArgs a) {
  constexpr int D = DD, cpr = D >> 2;
  __shared__ __align__(16) float q_s[FT_RANK_WARPS][D];
  __shared__ __align__(16) unsigned long long key_s[FT_RANK_WARPS][FT_HCAP];
  __shared__ int cand_s[FT_RANK_WARPS][FT_HCAP];
  if (a.skip && *a.skip) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y, qi = blockIdx.x * FT_RANK_WARPS + warp;
  if (qi >= a.P1) return;
  const int K = a.K;
  const int n1 = a.len1 ? min((int)a.len1[b], a.P1) : a.P1;
  float* od = a.dists + ((size_t)b * a.P1 + qi) * K;
  int64_t* oi = a.idx + ((size_t)b * a.P1 + qi) * K;
  if (qi >= n1) {  // rows beyond lengths1: zeros (pytorch3d convention)
    if (lane < K) { od[lane] = 0.0f; oi[lane] = 0; }
    return;
  }
  const float INF = __int_as_float(0x7f800000);
  float* wq = &q_s[warp][0];
  int* wcand = &cand_s[warp][0];
  unsigned long long* wkey = &key_s[warp][0];
  const float* p2b = a.p2 + (size_t)b * a.P2 * D;
  // everything the tail needs is requested up front: the warp's chain is latency, not bandwidth
  const unsigned* mq = a.masks + ((size_t)b * a.P1 + qi) * (size_t)(4 * a.T);  // the query's 4T mask words, contiguous
  const int nwords = 4 * a.T;
  unsigned mw0 = lane < nwords ? __ldg(mq + lane) : 0u;
  unsigned mw1 = lane + 32 < nwords ? __ldg(mq + lane + 32) : 0u;
  const float t0 = __ldg(a.tau0 + (size_t)b * a.P1 + qi);
  const float nq = __ldg(a.nrm1 + (size_t)b * a.P1 + qi);
  const unsigned nmax_bits = __reduce_max_sync(FULL, __ldg(a.nmax2 + (size_t)b * FT_NMAX_PARTS + lane));
  if (lane < cpr) reinterpret_cast<float4*>(wq)[lane] = __ldg(reinterpret_cast<const float4*>(a.p1 + ((size_t)b * a.P1 + qi) * D) + lane);
  // ---- candidate list from the hit masks (ascending candidate index); word w bit i = candidate 32 w + i ----
  int H;
  {
    // the first 64 words (all of them up to 2048 candidates): both counts ride one scan
    const int c0 = __popc(mw0), c1 = __popc(mw1);
    int incl = c0 | (c1 << 16);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(FULL, incl, o);
      if (lane >= o) incl += v;
    }
    const int tot = __shfl_sync(FULL, incl, 31);
    const int tot0 = tot & 0xffff;
    int pos = (incl & 0xffff) - c0;
    while (mw0) {
      const int bit = __ffs(mw0) - 1;
      mw0 &= mw0 - 1;
      if (pos < FT_HCAP) wcand[pos] = lane * 32 + bit;
      ++pos;
    }
    pos = tot0 + (incl >> 16) - c1;
    while (mw1) {
      const int bit = __ffs(mw1) - 1;
      mw1 &= mw1 - 1;
      if (pos < FT_HCAP) wcand[pos] = (lane + 32) * 32 + bit;
      ++pos;
    }
    H = tot0 + (tot >> 16);
  }
  for (int w0 = 64; w0 < nwords; w0 += 32) {  // clouds of more than 2048 candidates
    const int w = w0 + lane;
    unsigned m = w < nwords ? __ldg(mq + w) : 0u;
    const int c = __popc(m);
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(FULL, incl, o);
      if (lane >= o) incl += v;
    }
    int pos = H + incl - c;
    while (m) {
      const int bit = __ffs(m) - 1;
      m &= m - 1;
      if (pos < FT_HCAP) wcand[pos] = w * 32 + bit;
      ++pos;
    }
    H += __shfl_sync(FULL, incl, 31);
  }
  __syncwarp();
  bool ok = H <= FT_HCAP;
  // ---- canonical distances: lane r reads the row of candidate r itself (16-byte loads, all requested before the
  //      first is used) and sums it sequentially in d order; key = (d_canon bits, idx): d >= 0, so the integer
  //      order of the keys is the (d, idx) order ----
  auto canon = [&](int r) -> unsigned long long {
    const int ci = wcand[r];
    const float4* y = reinterpret_cast<const float4*>(p2b + (size_t)ci * D);
    const float4* x = reinterpret_cast<const float4*>(wq);
    float acc = 0.0f;
#pragma unroll
    for (int h = 0; h < cpr; h += 8) {
      float4 yv[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) yv[c] = __ldg(y + h + c);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 x0 = x[h + c];
        acc = sq_acc(acc, x0.x, yv[c].x);
        acc = sq_acc(acc, x0.y, yv[c].y);
        acc = sq_acc(acc, x0.z, yv[c].z);
        acc = sq_acc(acc, x0.w, yv[c].w);
      }
    }
    return ((unsigned long long)__float_as_uint(acc) << 32) | (unsigned)ci;
  };
  unsigned long long key0 = ~0ull, key1 = ~0ull;
  if (ok) {
    if (lane < H) key0 = canon(lane);
    if (H > 32 && lane + 32 < H) key1 = canon(lane + 32);  // (first test warp-uniform)
    wkey[lane] = key0;
    if (H > 32) wkey[lane + 32] = key1;
  }
  __syncwarp();
  // ---- rank by (d_canon, idx): every lane walks the key list (broadcast 16-byte reads, two keys each) ----
  int rank0 = 0, rank1 = 0;
  if (ok) {
    const int He = (H + 1) & ~1;  // the list is padded with ~0 keys up to 32 / 64 entries
    const ulonglong2* kp = reinterpret_cast<const ulonglong2*>(wkey);
    if (H <= 32) {
      for (int m2 = 0; m2 < He; m2 += 2) {
        const ulonglong2 o = kp[m2 >> 1];
        rank0 += (o.x < key0 ? 1 : 0) + (o.y < key0 ? 1 : 0);
      }
    } else {
      for (int m2 = 0; m2 < He; m2 += 2) {
        const ulonglong2 o = kp[m2 >> 1];
        rank0 += (o.x < key0 ? 1 : 0) + (o.y < key0 ? 1 : 0);
        rank1 += (o.x < key1 ? 1 : 0) + (o.y < key1 ? 1 : 0);
      }
    }
    // K-th smallest canonical distance among the hits
    const int kk = min(K, H);
    unsigned top = 0u;
    if (lane < H && rank0 < kk) top = (unsigned)(key0 >> 32);
    if (lane + 32 < H && rank1 < kk) top = max(top, (unsigned)(key1 >> 32));
    const float dk = __uint_as_float(__reduce_max_sync(FULL, top));
    if (t0 != INF) {  // tau0 == inf: every candidate of the cloud is a hit
      const float eps = feat_eps_fast(nq, __uint_as_float(nmax_bits), D);
      // 1/16 more than eps covers the roundings of this very expression
      ok = H >= K && dk < (t0 + nq) - 1.0625f * eps;
    }
  }
  if (!ok) {  // (warp-uniform) exact fallback kernel takes this query
    if (lane == 0) {
      const int pos = atomicAdd(a.fb_count, 1);
      a.fb_list[pos] = b * a.P1 + qi;
    }
    return;
  }
  if (lane < H && rank0 < K) { od[rank0] = __uint_as_float((unsigned)(key0 >> 32)); oi[rank0] = (int64_t)(unsigned)key0; }
  if (lane + 32 < H && rank1 < K) { od[rank1] = __uint_as_float((unsigned)(key1 >> 32)); oi[rank1] = (int64_t)(unsigned)key1; }
  if (lane < K && lane >= H) { od[lane] = 0.0f; oi[lane] = 0; }  // fewer than K candidates in the cloud
}

// ---- exact fallback: one CTA per flagged query (16 warps x 1/16 of the candidates, then a two-level merge) ------
template <int CH>  // CH = D / 4 float4 chunks per row (8 or 16)
__global__ void __launch_bounds__(512) knn_feat_fallback_kernel(FeatArgs a) {
  __shared__ float2 part_s[16][32];
  if (a.skip && *a.skip) return;
  const int total = *a.fb_count;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float INF = __int_as_float(0x7f800000);
  // merges the (ascending) list `v` of another slice into L; slices are merged in ascending index order, so
  // insert_tail's "equal distance -> lower index first" rule is kept
  auto merge = [&](WarpList& L, float& tau, const float2 v) {
    const int vi = __float_as_int(v.y);
    unsigned m = __ballot_sync(FULL, vi >= 0 && v.x < tau);
    while (m) {
      const int l = __ffs(m) - 1;
      m &= m - 1;
      const float dcand = __shfl_sync(FULL, v.x, l);
      const int icand = __shfl_sync(FULL, vi, l);
      if (dcand < tau) {
        L.insert_tail(dcand, icand, lane);
        tau = L.kth(a.K);
      }
    }
  };
  for (int e = blockIdx.x; e < total; e += gridDim.x) {
    const int flat = a.fb_list[e];
    const int b = flat / a.P1, qi = flat - b * a.P1;
    const int n2 = a.len2 ? min((int)a.len2[b], a.P2) : a.P2;
    const float4* xr = reinterpret_cast<const float4*>(a.p1 + ((size_t)b * a.P1 + qi) * a.D);
    const float* p2b = a.p2 + (size_t)b * a.P2 * a.D;
    float4 x[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) x[c] = __ldg(xr + c);
    WarpList L;
    L.init();
    float tau = INF;
    const int per = ((n2 + 15) / 16 + 31) & ~31;  // contiguous, 32-aligned slice per warp: ascending indices
    const int jlo = warp * per, jhi = min(n2, jlo + per);
    for (int j0 = jlo; j0 < jhi; j0 += 32) {
      const int j = j0 + lane;
      float acc = INF;
      if (j < jhi) {
        const float4* yr = reinterpret_cast<const float4*>(p2b + (size_t)j * a.D);
        float4 y[CH];  // the whole row first: one L2 round trip instead of CH dependent ones
#pragma unroll
        for (int c = 0; c < CH; ++c) y[c] = __ldg(yr + c);
        acc = 0.0f;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          acc = sq_acc(acc, x[c].x, y[c].x); acc = sq_acc(acc, x[c].y, y[c].y);
          acc = sq_acc(acc, x[c].z, y[c].z); acc = sq_acc(acc, x[c].w, y[c].w);
        }
      }
      unsigned m = __ballot_sync(FULL, j < jhi && acc < tau);
      while (m) {
        const int l = __ffs(m) - 1;
        m &= m - 1;
        const float dcand = __shfl_sync(FULL, acc, l);
        if (dcand < tau) {
          L.insert_tail(dcand, j0 + l, lane);
          tau = L.kth(a.K);
        }
      }
    }
    __syncthreads();  // previous query's merge has finished reading part_s
    part_s[warp][lane] = make_float2(L.d, __int_as_float(L.i));
    __syncthreads();
    // two-level merge: warps 0, 4, 8, 12 absorb the 3 slices after theirs, then warp 0 absorbs those three
    if ((warp & 3) == 0) {
      for (int w = warp + 1; w < warp + 4; ++w) merge(L, tau, part_s[w][lane]);
      if (warp != 0) part_s[warp][lane] = make_float2(L.d, __int_as_float(L.i));
    }
    __syncthreads();
    if (warp == 0) {
      for (int w = 4; w < 16; w += 4) merge(L, tau, part_s[w][lane]);
      if (lane < a.K) {
        const size_t o = ((size_t)b * a.P1 + qi) * a.K + lane;
        const bool found = L.i >= 0;
        a.dists[o] = found ? L.d : 0.0f;
        a.idx[o] = found ? (int64_t)L.i : 0;
      }
    }
  }
}

// ---- host side -------------------------------------------------------------------------------------
struct FeatWs {
  uint2* split1;   // [B*P1][2D bf16] centred, split queries
  uint2* split2;   // [B*P2][2D bf16] centred, split candidates
  float* nrm1;
  float* nrm2;
  unsigned* nmax2;
  int* fb_count;
  int* fb_list;
  unsigned* masks; // [B][4T][P1] pass-1 hit masks
  float* tau0;     // [B*P1]
  long long* dbg;
  size_t total;
};

static FeatWs feat_carve(void* base, int B, int P1, int P2, int D) {
  FeatWs w;
  char* p = reinterpret_cast<char*>(base);
  size_t o = 0;
  w.nmax2 = reinterpret_cast<unsigned*>(p + o); o += align_up(sizeof(unsigned) * (size_t)B * 32, 256);
  w.fb_count = reinterpret_cast<int*>(p + o);   o += 256;
  w.nrm1 = reinterpret_cast<float*>(p + o);     o += align_up(sizeof(float) * (size_t)B * P1, 256);
  w.nrm2 = reinterpret_cast<float*>(p + o);     o += align_up(sizeof(float) * (size_t)B * P2, 256);
  w.fb_list = reinterpret_cast<int*>(p + o);    o += align_up(sizeof(int) * (size_t)B * P1, 256);
  w.split1 = reinterpret_cast<uint2*>(p + o);   o += align_up(sizeof(float) * (size_t)B * P1 * D, 1024);
  w.split2 = reinterpret_cast<uint2*>(p + o);   o += align_up(sizeof(float) * (size_t)B * P2 * D, 1024);
  w.tau0 = reinterpret_cast<float*>(p + o);     o += align_up(sizeof(float) * (size_t)B * P1, 256);
  w.masks = reinterpret_cast<unsigned*>(p + o); o += align_up(sizeof(unsigned) * (size_t)B * P1 * 4 * (size_t)((P2 + FT_TM - 1) / FT_TM), 256);
  w.dbg = reinterpret_cast<long long*>(p + o);  o += align_up(sizeof(long long) * 16 * (size_t)B * (size_t)((P1 + FT_NQ - 1) / FT_NQ), 256);
  w.total = o;
  return w;
}

typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime (no link against libcuda)
static TmapEncodeFn tmap_encoder() {
  static TmapEncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<TmapEncodeFn>(p);
  }
  return fn;
}

// [rows, D] fp32 row-major -> boxes of 32 floats x 128 rows, 128B swizzle (the K-major UMMA slab layout)
static bool make_feat_map(CUtensorMap* m, const float* ptr, long long rows, int D) {
  TmapEncodeFn fn = tmap_encoder();
  if (!fn) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)D, (cuuint64_t)rows};
  const cuuint64_t gstr[1] = {(cuuint64_t)D * sizeof(float)};
  const cuuint32_t box[2] = {32, 128};
  const cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstr, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool feat_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* s = getenv("TPG_KNN_FEAT");
    v = (s && s[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

bool knn_feat_eligible(const KnnArgs& a) {
  if (!feat_enabled()) return false;
  if (a.out_mode != OUT_KNN || a.use_radius) return false;
  if (a.D != 32 && a.D != 64) return false;  // the shapes the generator uses (gcn.py:200-203,258)
  if (a.K > FT_MAX_K || a.P2 < 1024) return false;  // below that the brute-force kernel is as fast
  if ((long long)a.B * a.P1 >= (1LL << 31)) return false;
  // the hit masks of pass 1 are a P1 x P2 bit matrix per cloud: keep that workspace below 1 GiB (8 x 16384^2 fits)
  if ((long long)a.B * a.P1 * (long long)((a.P2 + FT_TM - 1) / FT_TM) * 16 > (1LL << 30)) return false;
  if ((reinterpret_cast<uintptr_t>(a.p1) | reinterpret_cast<uintptr_t>(a.p2)) & 15) return false;
  if (!tmap_encoder()) return false;  // no TMA descriptor encoder: SIMT path
  return true;
}

size_t knn_feat_fallback_count_offset(int B) {
  return (size_t)((char*)feat_carve(nullptr, B, 1, 1, 32).fb_count - (char*)nullptr);
}

size_t knn_feat_workspace_bytes(int B, int P1, int P2, int D) { return feat_carve(nullptr, B, P1, P2, D).total; }

int knn_feat_dispatch(const KnnArgs& k, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  TPG_REQUIRE(workspace && workspace_bytes >= knn_feat_workspace_bytes(k.B, k.P1, k.P2, k.D), TPG_EWORKSPACE,
              "knn: workspace too small for the tensor-core path (need %zu bytes)",
              knn_feat_workspace_bytes(k.B, k.P1, k.P2, k.D));
  TPG_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, TPG_EWORKSPACE,
              "knn: workspace must be 256-byte aligned for the tensor-core path");
  FeatWs w = feat_carve(workspace, k.B, k.P1, k.P2, k.D);
  // tuning (tools/bench_knn_feat.py): TPG_KNN_STOP=n returns after the n-th stage (1 split, 2 tcgen05, 3 rank)
  const char* stop_s = getenv("TPG_KNN_STOP");
  const int stop = stop_s ? atoi(stop_s) : 99;
  {
    // centre = sampled mean of the candidate cloud (rows below its length); both operands are shifted by it
    const int rows_per_it = 256 / (k.D >> 2);
    auto split_ctas = [&](int P) { return max(1, min(FT_NMAX_PARTS, ceil_div(P, rows_per_it))); };
    feat_split_kernel<<<dim3(split_ctas(k.P2), k.B), 256, 0, st>>>(k.p2, k.p2, k.len2, k.P2, k.P2, k.D, w.split2, w.nrm2,
                                                                  w.nmax2, k.skip);
    TPG_CHECK_LAUNCH("feat_split_kernel");
    if (k.p1 == k.p2 && k.P1 == k.P2) {
      w.nrm1 = w.nrm2;  // self search: one pre-pass
      w.split1 = w.split2;
    } else {
      feat_split_kernel<<<dim3(split_ctas(k.P1), k.B), 256, 0, st>>>(k.p1, k.p2, k.len2, k.P2, k.P1, k.D, w.split1, w.nrm1,
                                                                    nullptr, k.skip);
      TPG_CHECK_LAUNCH("feat_split_kernel");
    }
  }
  if (stop <= 1) return TPG_OK;
  FeatArgs a{k.p1, k.p2, k.len1, k.len2, k.B, k.P1, k.P2, k.D, k.K, w.nrm1, w.nrm2, w.nmax2,
             k.dists, reinterpret_cast<int64_t*>(k.idx), w.fb_count, w.fb_list, w.masks, w.tau0, ceil_div(k.P2, FT_TM),
             getenv("TPG_KNN_DBG") ? w.dbg : nullptr, k.skip};
  const size_t aux = (size_t)FT_NQ * 64 * sizeof(float);
  const size_t smem = (size_t)k.D * 512 * FT_STAGES + (size_t)k.D * 4 * FT_NQ + 1024 + aux;
  auto kern = k.D == 32 ? knn_feat_tc_kernel<32> : knn_feat_tc_kernel<64>;
  TPG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(ceil_div(k.P1, FT_NQ), k.B);
  FeatMaps maps;
  TPG_REQUIRE(make_feat_map(&maps.cand, reinterpret_cast<const float*>(w.split2), (long long)k.B * k.P2, k.D) &&
                  make_feat_map(&maps.query, reinterpret_cast<const float*>(w.split1), (long long)k.B * k.P1, k.D),
              TPG_ECUDA, "knn: cuTensorMapEncodeTiled failed");
  kern<<<grid, FT_BLOCK, smem, st>>>(a, maps);
  TPG_CHECK_LAUNCH("knn_feat_tc_kernel");
  if (stop <= 2) return TPG_OK;
  {
    const dim3 rgrid(ceil_div(k.P1, FT_RANK_WARPS), k.B);
    if (k.D == 32) knn_feat_rank_kernel<32><<<rgrid, FT_RANK_WARPS * 32, 0, st>>>(a);
    else knn_feat_rank_kernel<64><<<rgrid, FT_RANK_WARPS * 32, 0, st>>>(a);
    TPG_CHECK_LAUNCH("knn_feat_rank_kernel");
  }
  if (stop <= 3) return TPG_OK;
  if (k.D == 32) knn_feat_fallback_kernel<8><<<num_sms(), 512, 0, st>>>(a);
  else knn_feat_fallback_kernel<16><<<num_sms(), 512, 0, st>>>(a);
  TPG_CHECK_LAUNCH("knn_feat_fallback_kernel");
  return TPG_OK;
}

}  // namespace tpg
