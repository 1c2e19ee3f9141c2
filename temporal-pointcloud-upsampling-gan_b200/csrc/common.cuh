// common.cuh — shared helpers for libtpugan_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/tpugan_b200.h"

#ifndef TPG_API
#define TPG_API extern "C" __attribute__((visibility("default")))
#endif

namespace tpg {

constexpr unsigned FULL = 0xffffffffu;
constexpr int kNumSMsB200 = 148;

// ---- error / bookkeeping (api.cu) ----------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int num_sms();

#define TPG_REQUIRE(cond, code, ...)   \
  do {                                 \
    if (!(cond)) {                     \
      ::tpg::set_error(__VA_ARGS__);   \
      return (code);                   \
    }                                  \
  } while (0)

#define TPG_CHECK_LAUNCH(name)                                                   \
  do {                                                                           \
    cudaError_t e__ = cudaGetLastError();                                        \
    if (e__ != cudaSuccess) {                                                    \
      ::tpg::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));  \
      return TPG_ECUDA;                                                          \
    }                                                                            \
    ::tpg::count_launch();                                                       \
  } while (0)

#define TPG_CUDA(call)                                                            \
  do {                                                                            \
    cudaError_t e__ = (call);                                                     \
    if (e__ != cudaSuccess) {                                                     \
      ::tpg::set_error("%s failed: %s", #call, cudaGetErrorString(e__));          \
      return TPG_ECUDA;                                                           \
    }                                                                             \
  } while (0)

static inline cudaStream_t as_stream(tpg_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- canonical arithmetic: never contracted into FMA ------------------------
// d2 = ((dx*dx) + dy*dy) + dz*dz with separate rounding of every product and sum
// (the order pytorch3d's CPU loop uses).
__device__ __forceinline__ float sq_acc(float acc, float a, float b) {
  float diff = __fsub_rn(a, b);
  return __fadd_rn(acc, __fmul_rn(diff, diff));
}
__device__ __forceinline__ float sqdist3(float ax, float ay, float az, float bx, float by, float bz) {
  float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// ---- warp-resident sorted list (one element per lane, rank == lane) ----------
// Holds the best (d, i) pairs seen so far in ascending (d, i) order.  `insert_tail`
// is for candidates that arrive in ascending index order (brute-force scans): a new
// element goes after every stored element with d <= dc, so equal distances keep the
// lower index first.  `insert_key` takes candidates in any order (grid walks) and
// compares the full (d, i) key.
struct WarpList {
  float d;  // this lane's element
  int i;
  __device__ __forceinline__ void init() { d = __int_as_float(0x7f800000); i = -1; }
  __device__ __forceinline__ void insert_tail(float dc, int ic, int lane) {
    unsigned le = __ballot_sync(FULL, d <= dc);
    int pos = __popc(le);
    float ud = __shfl_up_sync(FULL, d, 1);
    int ui = __shfl_up_sync(FULL, i, 1);
    if (lane == pos) { d = dc; i = ic; }
    else if (lane > pos) { d = ud; i = ui; }
  }
  __device__ __forceinline__ void insert_key(float dc, int ic, int lane) {
    unsigned le = __ballot_sync(FULL, d < dc || (d == dc && i < ic && i >= 0));
    int pos = __popc(le);
    float ud = __shfl_up_sync(FULL, d, 1);
    int ui = __shfl_up_sync(FULL, i, 1);
    if (lane == pos) { d = dc; i = ic; }
    else if (lane > pos) { d = ud; i = ui; }
  }
  __device__ __forceinline__ float kth(int K) const { return __shfl_sync(FULL, d, K - 1); }
};

}  // namespace tpg
