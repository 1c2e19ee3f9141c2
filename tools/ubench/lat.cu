// latency micro-benchmarks for the FPS round (dependent chains), sm_100a.  nvcc -arch=sm_100a lat.cu -o lat
#include <cstdio>
#include <cuda_runtime.h>
#define FULL 0xffffffffu
__global__ void k(long long* out, int iters, unsigned seed) {
  __shared__ unsigned sm[64];
  unsigned v = seed + threadIdx.x * 2654435761u;
  long long t0, t1;
  // REDUX.MAX chain
  t0 = clock64();
  for (int i = 0; i < iters; ++i) v = __reduce_max_sync(FULL, v) ^ (threadIdx.x + i);
  t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  // SHFL chain
  t0 = clock64();
  for (int i = 0; i < iters; ++i) v = __shfl_xor_sync(FULL, v, 1) + i;
  t1 = clock64();
  if (threadIdx.x == 0) out[1] = t1 - t0;
  // VOTE chain
  t0 = clock64();
  for (int i = 0; i < iters; ++i) v = __ballot_sync(FULL, v & 1) + threadIdx.x + i;
  t1 = clock64();
  if (threadIdx.x == 0) out[2] = t1 - t0;
  // STS + BAR + LDS chain
  t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if ((threadIdx.x & 31) == 0) sm[(i & 1) * 32 + (threadIdx.x >> 5)] = v;
    __syncthreads();
    v += sm[(i & 1) * 32 + (threadIdx.x & 31) % (blockDim.x >> 5)];
  }
  t1 = clock64();
  if (threadIdx.x == 0) out[3] = t1 - t0;
  // FP chain: sub mul add add min
  float f = __uint_as_float((v & 0x7fffff) | 0x3f000000);
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { float d = f - 0.5f; f = fminf(__fadd_rn(__fadd_rn(__fmul_rn(d, d), 0.1f), 0.2f), f + 1.0f); }
  t1 = clock64();
  if (threadIdx.x == 0) out[4] = t1 - t0;
  // LDS chain (pointer chase)
  sm[threadIdx.x & 63] = (threadIdx.x + 1) & 63;
  __syncthreads();
  unsigned p = threadIdx.x & 63;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) p = sm[p];
  t1 = clock64();
  if (threadIdx.x == 0) out[5] = t1 - t0;
  // REDUX.MIN after compare (the pair)
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { unsigned m = __reduce_max_sync(FULL, v); v = __reduce_min_sync(FULL, v == m ? threadIdx.x : 0xffffffffu) + i + v; }
  t1 = clock64();
  if (threadIdx.x == 0) out[6] = t1 - t0;
  out[7 + threadIdx.x % 2] = v + p + (unsigned)f;
}
int main() {
  long long* d; cudaMalloc(&d, 16 * sizeof(long long));
  const char* names[] = {"REDUX.MAX", "SHFL", "VOTE", "STS+BAR+LDS", "FP chain (5 dep ops)", "LDS", "REDUX pair"};
  for (int threads : {32, 128, 512, 1024}) {
    k<<<1, threads>>>(d, 1000, 1); cudaDeviceSynchronize();
    k<<<1, threads>>>(d, 1000, 2); cudaDeviceSynchronize();
    long long h[16]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("threads=%d:", threads);
    for (int i = 0; i < 7; ++i) printf("  %s=%.1f", names[i], h[i] / 1000.0);
    printf(" cycles/iter\n");
  }
  return 0;
}
