// assemble.cu — K11/K12: conv-input assembly around the grouping.
//
// K11 tpg_group_assemble_f32 writes the concatenated [B, C_0 + C_1 + ..., M, k] input of a shared MLP in ONE
// pass, straight from the per-point tensors, instead of grouping each tensor into its own [B,C_p,M,k] buffer,
// subtracting / repeating in further passes and copying everything once more for the torch.cat:
//   * pointnet2_utils.QueryAndGroup.forward: cat([xyz[idx] - new_xyz, features[idx]])   (discriminator.py:190)
//   * FlowEmbedding.forward: cat([pos2[idx] - pos1, feat2[idx], feat1 repeated over k])  (discriminator.py:270-277)
// A part is a channel range of the output: GATHER (optionally minus a per-centre value) or BROADCAST.
//
// K12 tpg_edge_affine_fwd_f32 is the k-expanded half of EdgeConv after the algebraic restructure
// (gcn_lib/pointnet/gcn.py:206-211): with P = act(W_n f + b_n) and Q = W_e f + b_e computed per NODE,
//   node_affine(f_j) + edge_affine(f_j - f_i) = P[j] + act(Q[j] - (Q[i] - b_e)),   act = LeakyReLU(slope)
// so the two 1x1 convolutions run over N columns instead of N*k and the [B,C,N,k] tensors `grouped`,
// `edge_feat`, the two conv outputs and their activations are never materialised.  tpg_edge_affine_bwd_f32
// turns grad_out into the masked gradient g2 = grad_out * act'(Q[j] - c) (+ the per-centre sum of g2);
// the scatter parts of the backward are two ordinary grouping backwards (grad_out -> dP, g2 -> dQ).
//
// Same forward design as group.cu: a CTA owns (cloud, channel tile, range of flat (m,j) positions); the
// tile's source rows are staged in shared memory when they fit, every thread turns one int4 of indices into
// float4 stores (coalesced 16-byte stores along l; random reads hit shared memory).
#include "common.cuh"
#include "internal.cuh"

namespace tpg {

constexpr int AS_THREADS = 256;
constexpr size_t AS_SMEM_MAX = 96 * 1024;
constexpr int AS_MAX_PARTS = TPG_ASSEMBLE_MAX_PARTS;

struct AsmPart {
  const float* src;
  const float* center;
  int C, N, mode;
  int c_off;      // first output channel of the part
  int tile0;      // first blockIdx.y of the part
  int TC;         // channels per tile
  int smem;       // rows staged in shared memory
};

struct AsmArgs {
  AsmPart part[AS_MAX_PARTS];
  int nparts;
  const int32_t* idx;  // [B,L]
  int B, M, k, L, Ctot, LT;
  float* out;          // [B,Ctot,L]
};

template <bool VEC4>
__global__ void __launch_bounds__(AS_THREADS) group_assemble_kernel(const __grid_constant__ AsmArgs a) {
  extern __shared__ float rows_s[];
  const int b = blockIdx.z, tid = threadIdx.x;
  int p = 0;
#pragma unroll
  for (int q = 1; q < AS_MAX_PARTS; ++q)
    if (q < a.nparts && (int)blockIdx.y >= a.part[q].tile0) p = q;
  const AsmPart& pt = a.part[p];
  const int c0 = ((int)blockIdx.y - pt.tile0) * pt.TC;
  const int tc = min(pt.TC, pt.C - c0);
  const int l0 = blockIdx.x * a.LT, l1 = min(a.L, l0 + a.LT);
  float* ob = a.out + ((size_t)b * a.Ctot + pt.c_off + c0) * a.L;
  if (pt.mode == TPG_PART_BROADCAST) {
    const float* sb = pt.src + ((size_t)b * pt.C + c0) * a.M;
    if (VEC4) {
      for (int g = (l0 >> 2) + tid; g < (l1 >> 2); g += AS_THREADS) {
        const int l = g << 2;
        const int m0 = l / a.k, m1 = (l + 1) / a.k, m2 = (l + 2) / a.k, m3 = (l + 3) / a.k;
        for (int c = 0; c < tc; ++c) {
          const float* r = sb + (size_t)c * a.M;
          float4 v;
          v.x = __ldg(r + m0); v.y = __ldg(r + m1); v.z = __ldg(r + m2); v.w = __ldg(r + m3);
          reinterpret_cast<float4*>(ob + (size_t)c * a.L)[g] = v;
        }
      }
    } else {
      for (int l = l0 + tid; l < l1; l += AS_THREADS) {
        const int m = l / a.k;
        for (int c = 0; c < tc; ++c) ob[(size_t)c * a.L + l] = __ldg(sb + (size_t)c * a.M + m);
      }
    }
    return;
  }
  const float* fb = pt.src + ((size_t)b * pt.C + c0) * pt.N;
  if (pt.smem) {
    const int total = tc * pt.N;
    for (int e = tid; e < total; e += AS_THREADS) rows_s[e] = __ldg(fb + e);
    __syncthreads();
  }
  const float* rows = pt.smem ? rows_s : fb;
  const bool sm = pt.smem != 0;
  const int32_t* ib = a.idx + (size_t)b * a.L;
  const float* cb = pt.center ? pt.center + ((size_t)b * pt.C + c0) * a.M : nullptr;
  if (VEC4) {
    const int4* ib4 = reinterpret_cast<const int4*>(ib);
    for (int g = (l0 >> 2) + tid; g < (l1 >> 2); g += AS_THREADS) {
      const int4 ii = __ldg(ib4 + g);
      int m0 = 0, m1 = 0, m2 = 0, m3 = 0;
      if (cb) { const int l = g << 2; m0 = l / a.k; m1 = (l + 1) / a.k; m2 = (l + 2) / a.k; m3 = (l + 3) / a.k; }
#pragma unroll 4
      for (int c = 0; c < tc; ++c) {
        const float* r = rows + (size_t)c * pt.N;
        float4 v;
        if (sm) { v.x = r[ii.x]; v.y = r[ii.y]; v.z = r[ii.z]; v.w = r[ii.w]; }
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
    for (int l = l0 + tid; l < l1; l += AS_THREADS) {
      const int i = __ldg(ib + l);
      const int m = cb ? l / a.k : 0;
      for (int c = 0; c < tc; ++c) {
        float v = sm ? rows[(size_t)c * pt.N + i] : __ldg(rows + (size_t)c * pt.N + i);
        if (cb) v = __fsub_rn(v, __ldg(cb + (size_t)c * a.M + m));
        ob[(size_t)c * a.L + l] = v;
      }
    }
  }
}

// ---- K12: EdgeConv pre-activation -------------------------------------------------------------------------
struct EdgeArgs {
  const float* p;        // [B,C,N]  act(W_n f + b_n)
  const float* q;        // [B,C,N]  W_e f + b_e
  const float* center;   // [B,C,M]  q at the centre minus b_e
  const int32_t* idx;    // [B,L]
  const float* grad_out; // bwd: [B,C,L]
  float slope;
  int B, C, N, M, k, L, TC, LT, smem;
  float* out;            // fwd: [B,C,L];  bwd: g2 [B,C,L]
  float* gcenter;        // bwd: [B,C,M] = -sum_j g2 (only when LT covers whole centres; see launcher)
};

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.0f ? v : __fmul_rn(v, slope); }

// BWD = false: out = p[i] + lrelu(q[i] - c);  BWD = true: out = grad_out * lrelu'(q[i] - c)
template <bool BWD, bool VEC4>
__global__ void __launch_bounds__(AS_THREADS) edge_affine_kernel(const EdgeArgs a) {
  extern __shared__ float rows_s[];  // [TC][N] q rows, then (fwd) [TC][N] p rows
  const int b = blockIdx.z, c0 = blockIdx.y * a.TC, tid = threadIdx.x;
  const int tc = min(a.TC, a.C - c0);
  const float* qb = a.q + ((size_t)b * a.C + c0) * a.N;
  const float* pb = BWD ? nullptr : a.p + ((size_t)b * a.C + c0) * a.N;
  const int total = tc * a.N;
  if (a.smem) {
    for (int e = tid; e < total; e += AS_THREADS) rows_s[e] = __ldg(qb + e);
    if (!BWD)
      for (int e = tid; e < total; e += AS_THREADS) rows_s[a.TC * a.N + e] = __ldg(pb + e);
    __syncthreads();
  }
  const bool sm = a.smem != 0;
  const float* qrows = sm ? rows_s : qb;
  const float* prows = BWD ? nullptr : (sm ? rows_s + a.TC * a.N : pb);
  const int l0 = blockIdx.x * a.LT, l1 = min(a.L, l0 + a.LT);
  const int32_t* ib = a.idx + (size_t)b * a.L;
  const float* cb = a.center + ((size_t)b * a.C + c0) * a.M;
  float* ob = a.out + ((size_t)b * a.C + c0) * a.L;
  const float* gb = BWD ? a.grad_out + ((size_t)b * a.C + c0) * a.L : nullptr;
  auto one = [&](int c, int i, int m, int l) -> float {
    const float qv = sm ? qrows[(size_t)c * a.N + i] : __ldg(qrows + (size_t)c * a.N + i);
    const float pre = __fsub_rn(qv, __ldg(cb + (size_t)c * a.M + m));
    if (BWD) {
      const float g = __ldg(gb + (size_t)c * a.L + l);
      return pre > 0.0f ? g : __fmul_rn(g, a.slope);
    }
    const float pv = sm ? prows[(size_t)c * a.N + i] : __ldg(prows + (size_t)c * a.N + i);
    return __fadd_rn(pv, lrelu(pre, a.slope));
  };
  if (VEC4) {
    const int4* ib4 = reinterpret_cast<const int4*>(ib);
    for (int g = (l0 >> 2) + tid; g < (l1 >> 2); g += AS_THREADS) {
      const int4 ii = __ldg(ib4 + g);
      const int l = g << 2;
      const int m0 = l / a.k, m1 = (l + 1) / a.k, m2 = (l + 2) / a.k, m3 = (l + 3) / a.k;
#pragma unroll 2
      for (int c = 0; c < tc; ++c) {
        float4 v;
        v.x = one(c, ii.x, m0, l); v.y = one(c, ii.y, m1, l + 1);
        v.z = one(c, ii.z, m2, l + 2); v.w = one(c, ii.w, m3, l + 3);
        reinterpret_cast<float4*>(ob + (size_t)c * a.L)[g] = v;
      }
    }
  } else {
    for (int l = l0 + tid; l < l1; l += AS_THREADS) {
      const int i = __ldg(ib + l);
      const int m = l / a.k;
      for (int c = 0; c < tc; ++c) ob[(size_t)c * a.L + l] = one(c, i, m, l);
    }
  }
}

// gcenter[b,c,m] = -sum_j g2[b,c,m,j], sequential in j (deterministic); one thread per (b,c,m)
__global__ void __launch_bounds__(256) neg_rowsum_kernel(const float* __restrict__ g2, long long rows, int k,
                                                         float* __restrict__ out) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float* s = g2 + r * k;
  float acc = 0.0f;
  for (int j = 0; j < k; ++j) acc = __fadd_rn(acc, __ldg(s + j));
  out[r] = -acc;
}

// channels per tile / positions per tile: enough CTAs for ~3 per SM, rows in shared memory when they fit
static void asm_tiles(int B, int C, int N, int L, int rows_per_channel, int& TC, int& LT, int& smem) {
  const int target = 3 * num_sms();
  const size_t row_bytes = (size_t)N * sizeof(float) * rows_per_channel;
  const int max_tc = row_bytes ? (int)(AS_SMEM_MAX / row_bytes) : 8;
  smem = max_tc >= 1 ? 1 : 0;
  TC = 1;
  for (int t = 8; t >= 1; t >>= 1) {
    if (t > C && t > 1) continue;
    if (smem && t > max_tc) continue;
    TC = t;
    if ((long long)B * ceil_div(C, t) >= target) break;
  }
  int lt = max(1, ceil_div(target, B * ceil_div(C, TC)));
  // staging the rows must stay small next to the tile's stores: at least 4*N positions per tile
  lt = min(lt, max(1, L / max(1024, 4 * N)));
  LT = ceil_div(ceil_div(L, lt), 1024) * 1024;
}

}  // namespace tpg

using namespace tpg;

TPG_API int tpg_group_assemble_f32(const tpg_assemble_part* parts, int nparts, const int32_t* idx, int B, int M,
                                   int k, float* out, tpg_stream_t stream) {
  TPG_REQUIRE(parts && nparts >= 1 && nparts <= AS_MAX_PARTS, TPG_EINVAL, "group_assemble: 1..%d parts", AS_MAX_PARTS);
  TPG_REQUIRE(B >= 0 && M >= 0 && k >= 0, TPG_EINVAL, "group_assemble: negative size");
  const long long L64 = (long long)M * k;
  TPG_REQUIRE(L64 < (1LL << 31), TPG_EUNSUPPORTED, "group_assemble: M*k too large");
  TPG_REQUIRE(B <= 65535, TPG_EUNSUPPORTED, "group_assemble: B > 65535");
  AsmArgs a{};
  a.nparts = nparts; a.idx = idx; a.B = B; a.M = M; a.k = k; a.L = (int)L64; a.out = out;
  int ctot = 0;
  bool any_gather = false;
  for (int p = 0; p < nparts; ++p) {
    TPG_REQUIRE(parts[p].C >= 1, TPG_EINVAL, "group_assemble: part %d has no channels", p);
    TPG_REQUIRE(parts[p].mode == TPG_PART_GATHER || parts[p].mode == TPG_PART_BROADCAST, TPG_EINVAL,
                "group_assemble: part %d: unknown mode %d", p, parts[p].mode);
    ctot += parts[p].C;
    any_gather = any_gather || parts[p].mode == TPG_PART_GATHER;
  }
  if (B == 0 || L64 == 0) return TPG_OK;
  TPG_REQUIRE(out && (!any_gather || idx), TPG_EINVAL, "group_assemble: null pointer");
  a.Ctot = ctot;
  // one LT for the whole launch: the tile shape of the widest gather part (the dominant store volume)
  int lt_all = 0, tiles = 0, coff = 0;
  size_t smem_max = 0;
  for (int p = 0; p < nparts; ++p) {
    AsmPart& pt = a.part[p];
    pt.src = parts[p].src; pt.center = parts[p].center; pt.C = parts[p].C; pt.mode = parts[p].mode;
    pt.N = parts[p].mode == TPG_PART_GATHER ? parts[p].N : 0;
    TPG_REQUIRE(pt.src, TPG_EINVAL, "group_assemble: part %d: null source", p);
    TPG_REQUIRE(pt.mode != TPG_PART_GATHER || pt.N >= 1, TPG_EINVAL, "group_assemble: part %d: empty source cloud", p);
    TPG_REQUIRE(pt.mode != TPG_PART_BROADCAST || !pt.center, TPG_EINVAL, "group_assemble: part %d: broadcast takes no center", p);
    int LT;
    asm_tiles(B, pt.C, pt.N, a.L, pt.mode == TPG_PART_GATHER ? 1 : 0, pt.TC, LT, pt.smem);
    if (pt.mode == TPG_PART_BROADCAST) pt.smem = 0;
    lt_all = max(lt_all, LT);
    pt.c_off = coff; pt.tile0 = tiles;
    coff += pt.C; tiles += ceil_div(pt.C, pt.TC);
    if (pt.smem) smem_max = max(smem_max, (size_t)pt.TC * pt.N * sizeof(float));
  }
  a.LT = lt_all;
  TPG_REQUIRE(tiles <= 65535, TPG_EUNSUPPORTED, "group_assemble: too many channels");
  const bool vec4 = (a.L & 3) == 0 && ((reinterpret_cast<uintptr_t>(idx) & 15) == 0) &&
                    ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  dim3 grid(ceil_div(a.L, a.LT), tiles, B);
  cudaStream_t st = as_stream(stream);
  auto kern = vec4 ? group_assemble_kernel<true> : group_assemble_kernel<false>;
  if (smem_max > 48 * 1024)
    TPG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
  kern<<<grid, AS_THREADS, smem_max, st>>>(a);
  TPG_CHECK_LAUNCH("group_assemble_kernel");
  return TPG_OK;
}

static int edge_args(EdgeArgs& a, const char* who, const float* q, const float* center, const int32_t* idx, int B, int C,
                     int N, int M, int k, float* out, int rows_per_channel) {
  TPG_REQUIRE(B >= 0 && C >= 0 && N >= 0 && M >= 0 && k >= 0, TPG_EINVAL, "%s: negative size", who);
  const long long L64 = (long long)M * k;
  TPG_REQUIRE(L64 < (1LL << 31), TPG_EUNSUPPORTED, "%s: M*k too large", who);
  TPG_REQUIRE(B <= 65535, TPG_EUNSUPPORTED, "%s: B > 65535", who);
  a.L = (int)L64;
  if (B == 0 || C == 0 || L64 == 0) return 1;
  TPG_REQUIRE(N >= 1, TPG_EINVAL, "%s: empty source cloud", who);
  TPG_REQUIRE(q && center && idx && out, TPG_EINVAL, "%s: null pointer", who);
  asm_tiles(B, C, N, a.L, rows_per_channel, a.TC, a.LT, a.smem);
  TPG_REQUIRE(ceil_div(C, a.TC) <= 65535, TPG_EUNSUPPORTED, "%s: C too large", who);
  return TPG_OK;
}

TPG_API int tpg_edge_affine_fwd_f32(const float* p, const float* q, const float* center, const int32_t* idx,
                                    float slope, int B, int C, int N, int M, int k, float* out, tpg_stream_t stream) {
  EdgeArgs a{p, q, center, idx, nullptr, slope, B, C, N, M, k, 0, 1, 1024, 0, out, nullptr};
  const int rc = edge_args(a, "edge_affine_fwd", q, center, idx, B, C, N, M, k, out, 2);
  if (rc != TPG_OK) return rc < 0 ? rc : TPG_OK;
  TPG_REQUIRE(p, TPG_EINVAL, "edge_affine_fwd: null pointer");
  const bool vec4 = (a.L & 3) == 0 && ((reinterpret_cast<uintptr_t>(idx) & 15) == 0) &&
                    ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  dim3 grid(ceil_div(a.L, a.LT), ceil_div(C, a.TC), B);
  const size_t sm = a.smem ? (size_t)2 * a.TC * N * sizeof(float) : 0;
  auto kern = vec4 ? edge_affine_kernel<false, true> : edge_affine_kernel<false, false>;
  if (sm > 48 * 1024) TPG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  kern<<<grid, AS_THREADS, sm, as_stream(stream)>>>(a);
  TPG_CHECK_LAUNCH("edge_affine_kernel<fwd>");
  return TPG_OK;
}

TPG_API int tpg_edge_affine_bwd_f32(const float* grad_out, const float* q, const float* center, const int32_t* idx,
                                    float slope, int B, int C, int N, int M, int k, float* g2, float* grad_center,
                                    tpg_stream_t stream) {
  EdgeArgs a{nullptr, q, center, idx, grad_out, slope, B, C, N, M, k, 0, 1, 1024, 0, g2, grad_center};
  const int rc = edge_args(a, "edge_affine_bwd", q, center, idx, B, C, N, M, k, g2, 1);
  if (rc != TPG_OK) return rc < 0 ? rc : TPG_OK;
  TPG_REQUIRE(grad_out, TPG_EINVAL, "edge_affine_bwd: null pointer");
  const bool vec4 = (a.L & 3) == 0 && ((reinterpret_cast<uintptr_t>(idx) & 15) == 0) &&
                    ((reinterpret_cast<uintptr_t>(g2) & 15) == 0) && ((reinterpret_cast<uintptr_t>(grad_out) & 15) == 0);
  dim3 grid(ceil_div(a.L, a.LT), ceil_div(C, a.TC), B);
  const size_t sm = a.smem ? (size_t)a.TC * N * sizeof(float) : 0;
  cudaStream_t st = as_stream(stream);
  auto kern = vec4 ? edge_affine_kernel<true, true> : edge_affine_kernel<true, false>;
  if (sm > 48 * 1024) TPG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  kern<<<grid, AS_THREADS, sm, st>>>(a);
  TPG_CHECK_LAUNCH("edge_affine_kernel<bwd>");
  if (grad_center) {
    const long long rows = (long long)B * C * M;
    neg_rowsum_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(g2, rows, k, grad_center);
    TPG_CHECK_LAUNCH("neg_rowsum_kernel");
  }
  return TPG_OK;
}
