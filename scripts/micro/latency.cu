// Dependent-chain latencies on this GPU (cycles): the numbers behind the per-pivot critical-path estimates.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o latency latency.cu
#include <cstdio>
#include <cuda_runtime.h>
#define N 256
__global__ void k(double *out, long long *cyc, double a, double b, int mode) {
  __shared__ double sm[64];
  __shared__ int cnt;
  if (threadIdx.x < 64) sm[threadIdx.x] = a + threadIdx.x;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  double x = a + threadIdx.x * 1e-3;
  unsigned u = threadIdx.x * 2654435761u;
  long long t0 = clock64();
  if (mode == 0) {
#pragma unroll 16
    for (int i = 0; i < N; i++) x = __dmul_rn(x, b);
  } else if (mode == 1) {
#pragma unroll 16
    for (int i = 0; i < N; i++) x = __dadd_rn(x, b);
  } else if (mode == 2) {
#pragma unroll 16
    for (int i = 0; i < N; i++) x = __fma_rn(x, b, a);
  } else if (mode == 3) {
#pragma unroll 4
    for (int i = 0; i < N; i++) x = __ddiv_rn(a, x);
  } else if (mode == 4) {
#pragma unroll 16
    for (int i = 0; i < N; i++) u = __reduce_max_sync(0xffffffffu, u + i);
  } else if (mode == 5) {
#pragma unroll 16
    for (int i = 0; i < N; i++) u = __shfl_xor_sync(0xffffffffu, u + i, 1);
  } else if (mode == 6) {
#pragma unroll 16
    for (int i = 0; i < N; i++) { __syncthreads(); }
  } else if (mode == 7) {
#pragma unroll 16
    for (int i = 0; i < N; i++) u = __syncthreads_or(u == 12345u + i);
  } else if (mode == 8) {
    int idx = threadIdx.x & 63;
#pragma unroll 16
    for (int i = 0; i < N; i++) idx = (int)sm[idx & 63] & 63;
    u = idx;
  } else if (mode == 9) {
#pragma unroll 16
    for (int i = 0; i < N; i++) u = atomicAdd(&cnt, (int)(u & 1));
  } else if (mode == 10) {
    // STS -> BAR -> LDS round trip
    for (int i = 0; i < N; i++) { sm[threadIdx.x & 63] = x; __syncthreads(); x = sm[(threadIdx.x + 1) & 63] + 1.0; __syncthreads(); }
  } else if (mode == 11) {
#pragma unroll 16
    for (int i = 0; i < N; i++) u = __ballot_sync(0xffffffffu, (u + i) & 1);
  } else if (mode == 12) {  // independent DMUL/DADD pairs: issue rate
    double y0 = x, y1 = x + 1, y2 = x + 2, y3 = x + 3, y4 = x + 4, y5 = x + 5, y6 = x + 6, y7 = x + 7;
#pragma unroll 4
    for (int i = 0; i < N / 8; i++) {
      y0 = __dmul_rn(y0, b); y1 = __dmul_rn(y1, b); y2 = __dmul_rn(y2, b); y3 = __dmul_rn(y3, b);
      y4 = __dmul_rn(y4, b); y5 = __dmul_rn(y5, b); y6 = __dmul_rn(y6, b); y7 = __dmul_rn(y7, b);
    }
    x = y0 + y1 + y2 + y3 + y4 + y5 + y6 + y7;
  } else if (mode == 13) {  // integer IMAD chain
#pragma unroll 16
    for (int i = 0; i < N; i++) u = u * 3u + 7u;
  } else if (mode == 14) {  // double compare + select chain
#pragma unroll 16
    for (int i = 0; i < N; i++) x = (x > b) ? x - 1.0 : x + a;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[mode] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = x + u;
}
int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, 1024 * 8); cudaMallocManaged(&cyc, 16 * 8);
  const char *names[] = {"DMUL chain", "DADD chain", "DFMA chain", "ddiv_rn chain", "REDUX.MAX chain", "SHFL chain", "BAR.SYNC", "BAR.RED.OR", "LDS.64 pointer chase", "ATOMS add (same addr)", "STS+BAR+LDS+BAR", "VOTE.ballot chain", "8 indep DMUL (per op)", "IMAD chain", "DSETP+select chain"};
  for (int threads : {32, 256}) {
    for (int mode = 0; mode < 15; mode++) {
      for (int rep = 0; rep < 2; rep++) { k<<<1, threads>>>(out, cyc, 1.0000001, 0.9999999, mode); cudaDeviceSynchronize(); }
      printf("threads=%4d  %-26s %8.1f cycles/op\n", threads, names[mode], (double)cyc[mode] / N);
    }
  }
  return 0;
}
